"""CPU model of the phase-1 scheme of the engine (partner search with truncated top-k lists, cluster
search against the state at batch start, the resolver with its in-batch repair tables and its speculative WINDOWS,
batches prepared ahead from a stale state) against the reference's sequential loop
(LimitedGreedySequenceClusterer.java:77-120) on random instances.

What is modelled is the LOGIC of hmk_p1_resolve_kernel and of Engine::phase1 / stage_partner_search, not their CUDA:
the same inputs (best-kb partner lists + "more existed" flag, static cluster candidates with their best, the dense
tables ib / pd, the masks ibm / ibm2), the same trust conditions for a window lane, the same sequential step, the same
statuses (RESTART on an exhausted truncated list, DONE, the null-object exception).  Pair scores are arbitrary integer
tables here (asymmetric, heavy ties), so the order-dependent parts are stressed much harder than real peptides do.
The claim tested: for every batch size, list length, window width and staleness of the prepared batches the model
ends phase 1 in exactly the reference's state (clusters with their member order, orphans, counters, exception step)."""
import numpy as np
import pytest

JMIN, JMAX = -(2 ** 31), 2 ** 31 - 1
CONTINUE, DONE, RESTART, NPE = 0, 1, 2, 3


def _wadd(a, b):
    return (a + b + 2 ** 31) % 2 ** 32 - 2 ** 31


def _better(cand, best):
    """NearestClusterRunner order (ClinkageSequenceClusterer.java:258-293): score desc, size desc, id asc"""
    return best is None or (-cand[0], -cand[1], cand[2]) < (-best[0], -best[1], best[2])


# ------------------------------------------------------------------------------------------------ the reference
def sequential(S, ab, T, K):
    n = len(ab)
    in_list = [True] * n
    clusters = []                                  # dict(fid, size, members)
    orphans = []
    processed = pos = steps = 0
    status, npe_step = 0, -1
    while processed < sum(in_list) and len(clusters) < K:
        while not in_list[pos]:
            pos += 1
        q = pos
        after = [i for i in range(q + 1, n) if in_list[i]]
        if not clusters:
            A = "empty"
        else:
            A = None
            for ci, c in enumerate(clusters):
                sc = min(S[m][q] for m in c["members"])
                if sc >= T and _better((sc, c["size"], c["fid"], ci), A):
                    A = (sc, c["size"], c["fid"], ci)
        if not after:
            Bn = "empty"
        else:
            Bn = None
            for i in after:
                if S[i][q] >= T and (Bn is None or (-S[i][q], -ab[i], i) < (-Bn[0], -ab[Bn[1]], Bn[1])):
                    Bn = (S[i][q], i)
        a_obj, b_obj = A is not None, Bn is not None
        a_score = JMIN if A == "empty" else (A[0] if a_obj else None)
        b_score = JMIN if Bn == "empty" else (Bn[0] if b_obj else None)
        action = "orphan"
        if a_obj:
            action = ("join" if a_score >= b_score else "create") if b_obj else "join"
        elif b_obj:
            action = "create"
        if (action == "join" and A == "empty") or (action == "create" and Bn == "empty"):
            status, npe_step = NPE, steps
            break
        if action == "join":
            c = clusters[A[3]]
            c["members"].append(q)
            c["size"] = _wadd(c["size"], ab[q])
        elif action == "create":
            p = Bn[1]
            clusters.append({"fid": q, "size": _wadd(ab[q], ab[p]), "members": [q, p]})
            in_list[p] = False
        else:
            orphans.append(q)
        steps += 1
        processed += 1
        pos += 1
    return {"status": status, "npe_step": npe_step, "steps": steps, "orphans": orphans,
            "clusters": [(c["fid"], c["size"], list(c["members"])) for c in clusters]}


# ------------------------------------------------------------------------------------------------ the engine's scheme
class Model:
    def __init__(self, S, ab, T, K, batch, kb, win, gate, depth, rng, join_lanes=False, ranks=1):
        self.S, self.ab, self.T, self.K = S, ab, T, K
        self.ranks = ranks                         # GPUs: each scores a stripe of the later sequences (section 6)
        # join_lanes: NOT in the kernel -- the extension DESIGN.md section 10 proposes (window lanes that join their best
        # pre-batch cluster), kept here so that its trust conditions stay checked against the sequential loop
        self.join_lanes = join_lanes
        self.n = len(ab)
        # a window never reaches the LAST unprocessed sequence (whose empty sub-list is the null-object case): every lane
        # consumes at most two sequences, so more than 2 x lanes must be left (the kernel: 32 lanes, gate 96)
        self.B, self.kb, self.win, self.gate, self.depth, self.rng = batch, kb, win, max(gate, 2 * win), depth, rng
        self.slot = [-1] * self.n                  # -1 = still a singleton
        self.clusters = []                         # dict(fid, size, members)
        self.orphans = []
        self.cur, self.unproc = 0, self.n
        self.steps = 0
        self.status, self.npe_step = CONTINUE, -1
        self.stats = {"windows": 0, "window_steps": 0, "sequential": 0, "restarts": 0, "prepared": 0, "window_joins": 0}
        # phase-2 hit reuse (Engine::phase2, hmk_member_check mode 2): every qualifying hit of every partner search, tagged
        # with the batch that produced it, and per sequence the batch in which it was RESOLVED as a query
        self.batch_id = 0
        self.xhits = []
        self.qbatch = [-1] * self.n

    # ---- stage_partner_search: what does not depend on the clustering state (or tolerates a stale one)
    def select(self, slot_seen, first, want):
        ids = [i for i in range(first, self.n) if slot_seen[i] < 0][:want]
        return ids

    def partner_lists(self, qid, snapshots, tag):
        """per rank: the best kb hits of ITS stripe of the later sequences, against ITS (possibly stale) view of who is
        still a singleton; then the exchange: every rank merges all lists -> one list per query, identical everywhere"""
        S, ab, T, kb, R = self.S, self.ab, self.T, self.kb, len(snapshots)
        lo0 = qid[0] + 1
        bounds = [lo0 + (self.n - lo0) * r // R for r in range(R + 1)]          # contiguous stripes of [first query + 1, n)
        lists = []
        for q in qid:
            merged, ovf = [], False
            for r in range(R):
                hits = sorted((-S[i][q], -ab[i], i) for i in range(max(q + 1, bounds[r]), bounds[r + 1])
                              if snapshots[r][i] < 0 and S[i][q] >= T)
                self.xhits += [(q, h[2], -h[0], tag) for h in hits]      # before the top-k cut
                merged += hits[:kb]
                ovf = ovf or len(hits) > kb
            merged.sort()
            ovf = ovf or len(merged) > kb                                # hmk_topk_merge: "more existed than the list holds"
            lists.append({"id": [h[2] for h in merged[:kb]], "score": [-h[0] for h in merged[:kb]], "ovf": ovf})
        return lists

    def stage(self, qid, snapshots):
        S, T = self.S, self.T
        self.batch_id += 1
        lists = self.partner_lists(qid, snapshots, self.batch_id)
        nq = len(qid)
        ib = [[S[qid[b2]][qid[b]] for b2 in range(nq)] for b in range(nq)]
        pd = [[[S[p][qid[b]] for p in lists[b2]["id"]] for b2 in range(nq)] for b in range(nq)]
        ibm, ibm2 = [], []
        for b in range(nq):
            thr = lists[b]["score"][-1] if lists[b]["id"] else JMAX
            ibm.append({b2 for b2 in range(b) if ib[b][b2] >= T})
            ibm2.append({b2 for b2 in ibm[b] if ib[b][b2] >= thr})
        return {"qid": qid, "lists": lists, "ib": ib, "pd": pd, "ibm": ibm, "ibm2": ibm2, "batch_id": self.batch_id}

    # ---- cluster search at the start of a batch's resolution: founder filter + member check + prepare_candidates
    def static_candidates(self, qid):
        S, T = self.S, self.T
        out = []
        for q in qid:
            cands, best = [], None
            for ci, c in enumerate(self.clusters):
                if S[c["fid"]][q] < T:
                    continue
                sc = min(S[m][q] for m in c["members"])
                if sc >= T:
                    cands.append((ci, sc, c["size"], c["fid"]))
                    if _better((sc, c["size"], c["fid"], ci), best):
                        best = (sc, c["size"], c["fid"], ci)
            out.append((cands, best))
        return out

    # ---- hmk_p1_resolve_kernel.  `hook(b)` is called before every step / window: the engine's side stream works then
    def resolve(self, bt, hook):
        S_, ab, T, K = self.S, self.ab, self.T, self.K
        qid, lists, ib, pd, ibm, ibm2 = bt["qid"], bt["lists"], bt["ib"], bt["pd"], bt["ibm"], bt["ibm2"]
        nq = len(qid)
        static = self.static_candidates(qid)
        s_best = [st[1] for st in static]
        consumed = set()
        for b in range(nq):                        # whatever stopped being a singleton since the lists were made
            for i in [qid[b]] + lists[b]["id"]:
                if self.slot[i] >= 0:
                    consumed.add(i)
        touched = {}                               # cluster index -> row
        fmask, f_slot, f_pick = set(), {}, {}
        dirty = set()
        status = CONTINUE
        penalty = 0

        def eval_touched(row, b, cl):
            if row["mask"] - ibm[b]:
                return None
            mn = cl
            if row["fb"] >= 0:
                s = pd[b][row["fb"]][f_pick[row["fb"]]]
                if s < T:
                    return None
                mn = min(mn, s)
            for b2 in row["mask"]:
                mn = min(mn, ib[b][b2])
            return mn

        def new_pair(b, pick):                     # {q, partner} (:99-101, 108-110)
            q, p = qid[b], lists[b]["id"][pick]
            assert self.slot[q] < 0 and self.slot[p] < 0 and p > q
            ci = len(self.clusters)
            self.clusters.append({"fid": q, "size": _wadd(ab[q], ab[p]), "members": [q, p]})
            self.slot[q] = self.slot[p] = ci
            consumed.add(p)
            touched[ci] = {"fb": b, "mask": {b}}
            fmask.add(b)
            f_slot[b], f_pick[b] = ci, pick

        b = 0
        while b < nq:
            hook(b)
            ncl = len(self.clusters)
            if penalty == 0 and ncl > 0 and ncl < K and self.unproc > self.gate and self.win > 0:
                W = min(self.win, nq - b)
                kind, bid, bscore, bpick = [3] * W, [-1] * W, [JMIN] * W, [0] * W
                jt = [-1] * W                      # join lanes (kind 4): the cluster they join

                def can_join(bi):                  # its static best is still THE best pre-batch cluster, untouched
                    sb_ = s_best[bi]
                    return self.join_lanes and sb_ is not None and bi not in dirty and sb_[3] not in touched
                for l in range(W):
                    bi = b + l
                    q = qid[bi]
                    if q in consumed:
                        kind[l] = 0
                        continue
                    pick = next((j for j, p in enumerate(lists[bi]["id"]) if p not in consumed), -1)
                    if pick < 0 and lists[bi]["ovf"]:
                        continue
                    sb = s_best[bi]
                    if pick >= 0:
                        bpick[l], bid[l], bscore[l] = pick, lists[bi]["id"][pick], lists[bi]["score"][pick]
                        if not (sb is not None and sb[0] >= bscore[l]):
                            kind[l] = 1
                        elif can_join(bi):         # a_score >= b_score whichever listed partner is still alive
                            kind[l], jt[l] = 4, sb[3]
                    elif sb is None:
                        kind[l] = 2
                    elif can_join(bi):             # no partner at all (and the list was complete)
                        kind[l], jt[l] = 4, sb[3]
                cm0 = [k == 1 for k in kind]
                for l in range(W):                 # clusters born in this batch, incl. by earlier lanes of the window
                    if kind[l] not in (1, 2, 4):
                        continue
                    bi = b + l
                    founders = fmask | {b + j for j in range(l) if cm0[j]}
                    hm = ibm2[bi] if kind[l] == 1 else ibm[bi]
                    need = max(bscore[l], T) if kind[l] == 1 else (s_best[bi][0] if kind[l] == 4 else T)
                    for b2 in founders & hm:
                        pk = bpick[b2 - b] if b2 >= b else f_pick[b2]
                        if min(ib[bi][b2], pd[bi][b2][pk]) >= need:
                            kind[l] = 3
                            break
                bad = [k == 3 for k in kind]
                jt0 = list(jt)
                for l in range(W):                 # a join lane needs its cluster as it was: no earlier lane may join it,
                    if kind[l] == 4:               # nor grow a cluster that ties with it on the score (size decides ties)
                        ties = {cc[0] for cc in static[b + l][0] if cc[1] == s_best[b + l][0]}
                        if any(jt0[j] in ties for j in range(l)):
                            bad[l] = True
                for l in range(W):
                    q = qid[b + l]
                    taken = any(kind[j] == 1 and bid[j] == q for j in range(l))
                    same = kind[l] == 1 and any(kind[j] == 1 and bid[j] == bid[l] for j in range(l))
                    if kind[l] != 0 and (taken or same):
                        bad[l] = True
                    if ncl + sum(cm0[:l]) >= K:
                        bad[l] = True
                Pn = next((l for l in range(W) if bad[l]), W)
                self.stats["windows"] += 1
                self.stats["window_steps"] += Pn
                for l in range(Pn):
                    bi = b + l
                    if kind[l] == 1:
                        new_pair(bi, bpick[l])
                        self.unproc -= 2
                    elif kind[l] == 2:
                        self.orphans.append(qid[bi])
                        self.unproc -= 1
                    elif kind[l] == 4:
                        ci = jt[l]
                        assert ci not in touched
                        touched[ci] = {"fb": -1, "mask": {bi}}
                        for b3 in range(bi + 1, nq):
                            if any(cc[0] == ci for cc in static[b3][0]):
                                dirty.add(b3)
                        c = self.clusters[ci]
                        c["members"].append(qid[bi])
                        c["size"] = _wadd(c["size"], ab[qid[bi]])
                        self.slot[qid[bi]] = ci
                        self.unproc -= 1
                        self.stats["window_joins"] += 1
                    if kind[l] != 0:
                        self.steps += 1
                        self.cur = qid[bi] + 1
                        self.qbatch[qid[bi]] = bt["batch_id"]
                b += Pn
                penalty = 0 if Pn >= min(4, self.win) else 4
                if Pn == W:
                    continue
            elif penalty > 0:
                penalty -= 1
            # ---- one step in reference order
            self.stats["sequential"] += 1
            q = qid[b]
            ncl = len(self.clusters)
            if ncl >= K:
                status, self.cur = DONE, q
                break
            if q in consumed:
                b += 1
                continue
            bkind, bscore1, pick = 0, JMIN, -1
            if self.unproc - 1 == 0:
                bkind = 2
            else:
                pick = next((j for j, p in enumerate(lists[b]["id"]) if p not in consumed), -1)
                if pick >= 0:
                    bkind, bscore1 = 1, lists[b]["score"][pick]
                elif lists[b]["ovf"]:
                    status, self.cur = RESTART, q
                    self.stats["restarts"] += 1
                    break
            akind, best = 0, None
            if ncl == 0:
                akind = 2
            else:
                cands, sb = static[b]
                todo = []
                if b not in dirty:                 # none of its pre-batch candidates changed: the staged best is final
                    if sb is not None:
                        best = sb
                else:
                    for ci, sc, sz, fd in cands:
                        if ci in touched:
                            todo.append((ci, sc))
                        elif _better((sc, sz, fd, ci), best):
                            best = (sc, sz, fd, ci)
                for b2 in fmask & ibm[b]:
                    todo.append((f_slot[b2], JMAX))
                for ci, cl in todo:
                    v = eval_touched(touched[ci], b, cl)
                    if v is not None:
                        c = self.clusters[ci]
                        if _better((v, c["size"], c["fid"], ci), best):
                            best = (v, c["size"], c["fid"], ci)
                if best is not None:
                    akind = 1
            ascore = best[0] if akind == 1 else JMIN
            join = create = False
            if akind != 0:
                if bkind != 0:
                    join = ascore >= bscore1
                    create = not join
                else:
                    join = True
            elif bkind != 0:
                create = True
            if (join and akind == 2) or (create and bkind == 2):
                status, self.npe_step, self.cur = NPE, self.steps, q
                break
            if join:
                ci = best[3]
                if ci not in touched:
                    touched[ci] = {"fb": -1, "mask": set()}
                    for b3 in range(b + 1, nq):    # later queries that list this cluster lose their staged best
                        if any(cc[0] == ci for cc in static[b3][0]):
                            dirty.add(b3)
                touched[ci]["mask"].add(b)
                c = self.clusters[ci]
                c["members"].append(q)
                c["size"] = _wadd(c["size"], ab[q])
                self.slot[q] = ci
            elif create:
                new_pair(b, pick)
                self.unproc -= 1
            else:
                self.orphans.append(q)
            self.steps += 1
            self.unproc -= 1
            self.cur = q + 1
            self.qbatch[q] = bt["batch_id"]
            b += 1
        if status == CONTINUE and (len(self.clusters) >= K or self.unproc <= 0):
            status = DONE
        self.status = status
        return status

    # ---- Engine::phase1: the ring of prepared batches
    def run(self):
        """The look-ahead is issued right behind the resolver of the current batch and runs on the side stream: rank 0's
        query selection and every rank's partner search see SOME state between that moment and the start of the prepared
        batch's own resolution -- possibly one resolution later, and a different one on every rank (the ranks' resolvers
        are replicas, but nothing keeps them in step).  Modelled with a global tick per resolver step / window and a random
        tick per event, in order per rank (one side stream each); rank 0's selection is what every rank uses."""
        R = self.ranks
        ring = []                                  # prepared batches behind the current one, in issue order
        self.tick = 0

        def advance(force=None):
            for nb in ring if force is None else [force]:
                if nb["sel"] is None and (nb is force or self.tick >= nb["sel_at"]):
                    nb["sel"] = list(self.slot)
                for r in range(R):
                    if nb["seen"][r] is None and (nb is force or self.tick >= nb["search_at"][r]):
                        nb["seen"][r] = list(self.slot)

        def hook(b):
            self.tick += 1
            advance()

        while len(self.clusters) < self.K and self.unproc > 0:
            if ring:
                cb = ring.pop(0)
                advance(force=cb)                  # the main stream waits for the batch's `ready` event
                qid = self.select(cb["sel"], cb["after"]["qid"][-1] + 1, self.B)
                assert len(qid) == self.B          # the margin of Engine::phase1 guarantees a full batch
                cb.update(self.stage(qid, cb["seen"]))
                self.stats["prepared"] += 1
            else:
                qid = self.select(self.slot, self.cur, min(self.B, self.unproc))
                cb = self.stage(qid, [self.slot] * R)
            nq = len(cb["qid"])
            prev = cb
            for d in range(1, self.depth + 1):
                if len(ring) >= d:
                    prev = ring[d - 1]
                    continue
                if (prev["nq"] if "nq" in prev else len(prev["qid"])) != self.B or self.unproc < nq + (2 + d) * self.B:
                    break
                horizon = self.tick + 1 + int(self.rng.integers(0, 2 * self.B))
                last = ring[-1] if ring else None
                sel_at = int(self.rng.integers(self.tick, horizon + 1))
                if last is not None:
                    sel_at = max(sel_at, last["sel_at"])
                search_at = []
                for r in range(R):
                    t = int(self.rng.integers(self.tick, horizon + 1))
                    if last is not None:
                        t = max(t, last["search_at"][r])
                    search_at.append(t)
                nb = {"after": prev, "nq": self.B, "qid": None, "sel": None, "seen": [None] * R, "sel_at": sel_at, "search_at": search_at}
                ring.append(nb)
                prev = nb
            st = self.resolve(cb, hook)
            if st != CONTINUE:
                ring = []
            if st in (NPE, DONE):
                break
        return {"status": NPE if self.status == NPE else 0, "npe_step": self.npe_step, "steps": self.steps, "orphans": self.orphans,
                "clusters": [(c["fid"], c["size"], list(c["members"])) for c in self.clusters]}

    # ---- phase 2's candidate pairs (singleton q, cluster c, min score over the phase-1 members) ...
    def pairs_from_kept_hits(self):
        """... from the partner-search hits of phase 1: a hit counts if its batch is the one that resolved its query and it
        links a founder with a sequence that is still a singleton; then the member check (hmk_member_check, mode 2).
        (The second role assignment -- the QUERY of the hit is the singleton -- is kept as in the kernel but can never
        fire: a resolved query that is still a singleton was an orphan, i.e. had no live hit at all.)"""
        S, T = self.S, self.T
        out = []
        for x, y, sc, tag in self.xhits:
            if self.qbatch[x] != tag:
                continue
            sx, sy = self.slot[x], self.slot[y]
            if sx >= 0 and sy < 0 and self.clusters[sx]["fid"] == x:
                c, q = sx, y
            elif sy >= 0 and sx < 0 and self.clusters[sy]["fid"] == y:
                c, q = sy, x
            else:
                continue
            cl = min([sc] + [S[m][q] for m in self.clusters[c]["members"][1:]])
            if cl >= T:
                out.append((q, c, cl))
        return out

    def pairs_direct(self):
        """... by the separate founder pass: every cluster against every remaining singleton"""
        S, T = self.S, self.T
        out = []
        for c, cl_ in enumerate(self.clusters):
            for q in range(self.n):
                if self.slot[q] < 0:
                    cl = min(S[m][q] for m in cl_["members"])
                    if cl >= T:
                        out.append((q, c, cl))
        return out


def _instance(rng, n, family, tie_heavy, asym):
    """pair scores with family structure: sequences of one family score high against each other"""
    T = 20
    fam = rng.integers(0, max(1, n // family), size=n)
    lo, hi = (T - 3, T + 4) if tie_heavy else (T - 15, T + 30)
    S = rng.integers(lo, hi, size=(n, n))
    S = np.where(fam[:, None] == fam[None, :], S, S - (8 if tie_heavy else 40))
    if not asym:
        S = np.minimum(S, S.T)
    ab = np.sort(rng.integers(1, 4 if tie_heavy else 60, size=n))[::-1]      # clustering order: abundance descending
    return [[int(v) for v in row] for row in S], [int(a) for a in ab], T


CONFIGS = [  # batch, kb, window lanes, window gate (unprocessed sequences), look-ahead depth
    (1, 1, 0, 0, 0), (4, 1, 0, 0, 0), (7, 2, 4, 0, 0), (16, 3, 8, 0, 1), (16, 8, 32, 0, 2), (32, 2, 32, 96, 2), (64, 8, 32, 0, 1),
    (5, 1, 4, 0, 2), (12, 32, 4, 0, 2),
]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("tie_heavy", [False, True])
def test_phase1_scheme_reaches_the_sequential_state(cfg, tie_heavy):
    batch, kb, win, gate, depth = cfg
    rng = np.random.default_rng(99 + 7 * batch + kb + (1000 if tie_heavy else 0))
    seen = {"windows": 0, "window_steps": 0, "sequential": 0, "restarts": 0, "prepared": 0}
    npairs = 0
    for trial in range(14):
        n = int(rng.integers(2, 140))
        S, ab, T = _instance(rng, n, int(rng.choice([2, 5, 12])), tie_heavy, asym=bool(trial % 2))
        K = int(rng.choice([1, 3, max(1, n // 8), n]))
        want = sequential(S, ab, T, K)
        m = Model(S, ab, T, K, batch, kb, win, gate, depth, rng)
        got = m.run()
        assert got == want, (cfg, tie_heavy, trial, n, K)
        if not trial % 2 and got["status"] == 0:       # symmetric scores: phase 2 can start from the kept phase-1 hits
            kept = m.pairs_from_kept_hits()
            assert len(kept) == len(set(kept)) and sorted(kept) == sorted(m.pairs_direct()), (cfg, tie_heavy, trial)
            npairs += len(kept)
        for k in seen:
            seen[k] += m.stats[k]
    if win > 0:
        assert seen["windows"] > 0 and seen["window_steps"] > 0        # the speculative path really ran ...
    assert seen["sequential"] > 0                                      # ... and so did the sequential one
    assert npairs > 0                                    # ... and the kept hits produced candidate pairs
    if depth > 0 and batch <= 16:
        assert seen["prepared"] > 0                                    # batches prepared from a stale state were used


def test_phase1_scheme_statuses():
    """the null-object quirks and the truncated-list restart (LimitedGreedySequenceClusterer.java:94-114)"""
    T = 20
    # one sequence: no cluster, empty sub-list -> both "nearest" objects exist with MIN_VALUE -> join on null -> exception
    want = sequential([[25]], [1], T, 5)
    assert want["status"] == NPE and want["npe_step"] == 0
    assert Model([[25]], [1], T, 5, 4, 2, 4, 0, 1, np.random.default_rng(0)).run() == want
    # three mutually similar sequences and one loner: pair {0, 1}; 2 joins it; 3 is last: empty sub-list -> MIN_VALUE
    # object, no valid cluster -> "create" with a null cluster -> exception at step 2
    S = [[30, 30, 30, 0], [30, 30, 30, 0], [30, 30, 30, 0], [0, 0, 0, 30]]
    want = sequential(S, [3, 2, 1, 1], T, 5)
    assert want["clusters"] == [(0, 6, [0, 1, 2])] and want["status"] == NPE and want["npe_step"] == 2
    for batch, kb in ((1, 1), (4, 1), (4, 8)):
        assert Model(S, [3, 2, 1, 1], T, 5, batch, kb, 4, 0, 1, np.random.default_rng(0)).run() == want
    # a one-entry list whose only candidate was taken by an earlier query of the batch while more hits existed: the batch
    # restarts at that query (HMK_P1_RESTART) and still ends in the reference's state
    rng = np.random.default_rng(4)
    restarts = 0
    for trial in range(6):
        S, ab, T = _instance(rng, 80, 20, bool(trial % 2), False)
        m = Model(S, ab, T, 80, 32, 1, 0, 0, 0, rng)
        assert m.run() == sequential(S, ab, T, 80)
        restarts += m.stats["restarts"]
    assert restarts > 0


def test_phase1_window_sees_the_pairs_of_earlier_lanes():
    """A pair created by an EARLIER lane of the same window can attract a later lane's query only through an exact tie (its
    partner would otherwise head the later lane's own list and collide there).  Hand-built: {0, 1} exists; lane q2 takes
    partner 5; lane q3 lists 4 and 5 with equal scores (4 first: smaller id) -- and {2, 5} scores 25 against q3, which
    equals its own pair's score, so the reference JOINS (a_score >= b_score, LimitedGreedySequenceClusterer.java:96-98)."""
    n, T = 12, 20
    S = [[0] * n for _ in range(n)]
    for i in range(n):
        S[i][i] = 40
    S[1][0] = S[0][1] = 40
    S[5][2] = S[2][5] = 30
    S[4][3] = S[3][4] = 25
    S[5][3] = S[3][5] = 25
    S[2][3] = S[3][2] = 27
    ab = [5] * n
    want = sequential(S, ab, T, n)
    assert want["clusters"][:2] == [(0, 10, [0, 1]), (2, 15, [2, 5, 3])] and 4 in want["orphans"]
    for win in (0, 4, 32):
        m = Model(S, ab, T, n, 8, 4, win, 0, 0, np.random.default_rng(0))
        if win == 4:
            m.gate = 8
        assert m.run() == want, win
    m = Model(S, ab, T, n, 8, 4, 4, 0, 0, np.random.default_rng(0))
    m.run()
    assert m.stats["windows"] > 0


def test_phase2_pairs_from_kept_hits_survive_restarts_and_discarded_batches():
    """The kept phase-1 hits are only usable because of the batch tag: a batch that restarts (or a prepared batch that is
    thrown away) leaves hits of queries that are searched AGAIN later -- without the tag the same candidate pair would be
    listed twice.  One-entry lists on family-rich data restart often; K well below n leaves many singletons for phase 2."""
    rng = np.random.default_rng(11)
    restarts = pairs = dropped = 0
    for trial in range(30):
        n = int(rng.integers(40, 140))
        S, ab, T = _instance(rng, n, int(rng.choice([12, 25])), bool(trial % 3 == 0), False)
        K = max(2, n // int(rng.choice([4, 6, 10])))
        m = Model(S, ab, T, K, int(rng.choice([8, 16, 32])), 1, int(rng.choice([0, 4])), 0, int(rng.choice([0, 1, 2])), rng)
        got = m.run()
        assert got == sequential(S, ab, T, K)
        if got["status"] != 0:
            continue
        kept = m.pairs_from_kept_hits()
        assert len(kept) == len(set(kept)) and sorted(kept) == sorted(m.pairs_direct()), trial
        restarts += m.stats["restarts"]
        pairs += len(kept)
        dropped += sum(1 for x, y, sc, tag in m.xhits if m.qbatch[x] >= 0 and m.qbatch[x] != tag)
    assert restarts > 0 and pairs > 0 and dropped > 0


@pytest.mark.parametrize("tie_heavy", [False, True])
def test_phase1_window_join_lanes_extension(tie_heavy):
    """NOT built: the next step for the resolver proposed in DESIGN.md section 10.  A window lane may also JOIN when its
    staged best pre-batch cluster is still the true arg-max (none of its candidates touched in this batch), no earlier
    lane joins that cluster or grows one that ties with it, no cluster born in this batch can reach its score, and the
    partner -- whichever is still alive -- scores no higher.  Exactness of these conditions, same instances as above."""
    rng = np.random.default_rng(4242 + tie_heavy)
    joins = 0
    for trial in range(60):
        n = int(rng.integers(20, 160))
        # big families and no early stop: most queries join a cluster that already exists
        S, ab, T = _instance(rng, n, int(rng.choice([12, 30, 60])), tie_heavy, asym=bool(trial % 2))
        if trial % 3 == 0:                     # ... and families that agree internally: long runs of joins to one cluster
            S = [[v + 12 if v >= T - 3 else v for v in row] for row in S]
        K = int(rng.choice([max(2, n // 8), n, n]))
        want = sequential(S, ab, T, K)
        m = Model(S, ab, T, K, int(rng.choice([8, 16, 64])), int(rng.choice([1, 3, 8])), int(rng.choice([4, 8, 32])), 0,
                  int(rng.choice([0, 1, 2])), rng, join_lanes=True)
        assert m.run() == want, (tie_heavy, trial)
        joins += m.stats["window_joins"]
    assert joins > 30


@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_phase1_scheme_on_several_ranks(ranks):
    """every rank scores its stripe of the later sequences against ITS OWN stale view of the singletons
    (the ranks' resolvers are replicas but not in step), keeps its best kb hits per query, and all ranks merge all lists.
    With rank 0's query selection shared, the merged lists lead the replicated resolver to the sequential result -- and the
    union of the ranks' kept hits gives phase 2 the candidate pairs of the founder pass."""
    rng = np.random.default_rng(77 + ranks)
    prepared = pairs = 0
    for trial in range(24):
        n = int(rng.integers(30, 150))
        S, ab, T = _instance(rng, n, int(rng.choice([5, 12, 30])), bool(trial % 3 == 0), asym=bool(trial % 2))
        K = int(rng.choice([3, max(1, n // 8), n]))
        want = sequential(S, ab, T, K)
        m = Model(S, ab, T, K, int(rng.choice([4, 8, 16])), int(rng.choice([1, 2, 8])), int(rng.choice([0, 4, 32])), 0,
                  int(rng.choice([1, 2])), rng, ranks=ranks)
        got = m.run()
        assert got == want, (ranks, trial)
        prepared += m.stats["prepared"]
        if not trial % 2 and got["status"] == 0:
            kept = m.pairs_from_kept_hits()
            assert len(kept) == len(set(kept)) and sorted(kept) == sorted(m.pairs_direct()), (ranks, trial)
            pairs += len(kept)
    assert prepared > 0 and pairs > 0
