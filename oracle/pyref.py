"""Second, independently written restatement of the reference path (numpy).

TEST INFRASTRUCTURE ONLY (same rules as oracle/hammock_oracle.h).  It exists because the
reference (Java, no JVM here) cannot be executed: two restatements written separately from
the Java source and compared on every fixture are the only defence against a mis-reading.
Structured differently from the C oracle on purpose: scores are computed per query against
whole candidate arrays, clusters are plain Python lists, and the nearest search is a sort
key instead of an incremental compare.  PARITY UNPINNED.

Citations: /root/reference/src/cz/krejciadam/hammock/<file>:<line>.
"""
from __future__ import annotations

import numpy as np

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"  # UniqueSequence.java:23-26
JMIN = -(2 ** 31)


def _i32(x):
    """Java int wrap-around."""
    return ((np.asarray(x, dtype=np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31)


class PyRef:
    def __init__(self, seqs, abundance, matrix, threshold, max_shift, shift_penalty, max_clusters):
        """seqs: list of uint8 code arrays ALREADY in clustering order."""
        self.seqs = [np.asarray(s, dtype=np.int64) for s in seqs]
        self.n = len(seqs)
        self.ab = [int(a) for a in abundance]
        self.M = np.asarray(matrix, dtype=np.int64).reshape(24, 24)
        self.T, self.X, self.P, self.K = int(threshold), int(max_shift), int(shift_penalty), int(max_clusters)
        self.lens = np.array([len(s) for s in self.seqs], dtype=np.int64)
        self.by_len = {}
        for L in np.unique(self.lens):
            idx = np.nonzero(self.lens == L)[0]
            self.by_len[int(L)] = (idx, np.stack([self.seqs[i] for i in idx]) if L > 0 else np.zeros((len(idx), 0), np.int64))
        self.row_of = np.zeros(self.n, dtype=np.int64)
        for L, (idx, _) in self.by_len.items():
            self.row_of[idx] = np.arange(len(idx))
        self.pairs = 0

    # ShiftedScorer.java:48-95, evaluated for many first arguments ("members") at once
    def scores(self, members, q):
        """S(member, query) for an int array of member ids -> int64 array (Java-int valued)."""
        members = np.asarray(members, dtype=np.int64)
        out = np.empty(len(members), dtype=np.int64)
        qs = self.seqs[q]
        lq = len(qs)
        X, P, M = self.X, self.P, self.M
        self.pairs += len(members)
        for L in np.unique(self.lens[members]):
            L = int(L)
            sel = np.nonzero(self.lens[members] == L)[0]
            rows = self.by_len[L][1][self.row_of[members[sel]]]  # (m, L)
            if L >= lq:   # :51-53  member is "longer" (ties: second argument = query is shorter)
                ls, ll, member_is_longer = lq, L, True
            else:         # :54-57
                ls, ll, member_is_longer = L, lq, False
            if X >= ls:   # :59-62
                raise ValueError("Shift too big")
            d = ll - ls
            best = np.full(len(sel), JMIN, dtype=np.int64)
            for k in range(-X, X + d + 1):                                   # :67
                if k <= 0:
                    cnt = ls + k
                    s_sl, l_sl = slice(-k, -k + cnt), slice(0, cnt)          # :70-72
                else:
                    cnt = min(ls, ll - k)
                    s_sl, l_sl = slice(0, cnt), slice(k, k + cnt)            # :74-76
                if member_is_longer:
                    sc = M[qs[s_sl][None, :], rows[:, l_sl]].sum(axis=1)     # M[shorter][longer] :110-112
                else:
                    sc = M[rows[:, s_sl], qs[l_sl][None, :]].sum(axis=1)
                sc = sc + d * P                                              # :79
                if k < 0:
                    sc = sc + (-k) * 2 * P                                   # :80-82
                if k > d:
                    sc = sc + (k - d) * 2 * P                                # :83-85
                sc = _i32(sc)
                best = np.where(sc > best, sc, best)                         # :86-89 strict >
            out[sel] = best
        return out

    # ClinkageClusterScorer.java:30-49 (dense evaluation; the early exit only replaces
    # values < T by "invalid", so min-over-all-members is equivalent)
    def _cl(self, members, q):
        s = self.scores(members, q)
        return int(s.min()) if (s >= self.T).all() else JMIN + 1

    # findNearestClusterParallel + NearestClusterRunner (ClinkageSequenceClusterer.java:137-177, 258-293)
    def _nearest_cluster(self, clusters, q):
        """clusters: list of dict(id, size, members).  -> ('empty',) | None | (score, cluster)"""
        if not clusters:
            return ("empty",)                                                # :138-140
        cand = []
        for c in clusters:
            sc = self._cl(c["members"], q)
            cand.append((-sc, -c["size"], c["id"], c))
        cand.sort(key=lambda t: t[:3])
        sc = -cand[0][0]
        if sc < JMIN + 42:                                                   # :151,159-161
            return None
        return (sc, cand[0][3])

    def _nearest_single(self, alive_after, q):
        if len(alive_after) == 0:
            return ("empty",)
        s = self.scores(alive_after, q)
        ok = s >= self.T
        if not ok.any():
            return None
        ids = np.asarray(alive_after)[ok]
        sc = s[ok]
        sizes = np.array([self.ab[i] for i in ids], dtype=np.int64)
        order = np.lexsort((ids, -sizes, -sc))                               # score desc, size desc, id asc
        j = order[0]
        return (int(sc[j]), int(ids[j]))

    def run(self):
        n, K = self.n, self.K
        in_list = np.ones(n, dtype=bool)          # initialList membership
        clusters = []                             # actualClusters (creation order)
        orphans = []                              # actualSequences
        processed = 0
        pos = 0
        steps = joins = created = 0
        status, npe_step = 0, -1
        # firstPhase, LimitedGreedySequenceClusterer.java:77-120
        while processed < int(in_list.sum()) and len(clusters) < K:           # :90
            while not in_list[pos]:
                pos += 1
            q = pos
            after = np.nonzero(in_list[q + 1:])[0] + q + 1
            A = self._nearest_cluster(clusters, q)                            # :92
            B = self._nearest_single(after, q)                                # :93
            a_obj = A is not None
            b_obj = B is not None
            a_score = JMIN if A == ("empty",) else (A[0] if a_obj else None)
            b_score = JMIN if B == ("empty",) else (B[0] if b_obj else None)
            action = "orphan"
            if a_obj:
                if b_obj:
                    action = "join" if a_score >= b_score else "create"       # :96-102
                else:
                    action = "join"                                           # :104
            elif b_obj:
                action = "create"                                             # :108-110
            if (action == "join" and A == ("empty",)) or (action == "create" and B == ("empty",)):
                status, npe_step = 2, steps                                   # NullPointerException
                break
            if action == "join":
                c = A[1]
                c["members"].append(q)
                c["size"] = int(_i32(c["size"] + self.ab[q]))
                joins += 1
            elif action == "create":
                p = B[1]
                clusters.append({"id": q, "size": int(_i32(self.ab[q] + self.ab[p])), "members": [q, p]})
                in_list[p] = False
                created += 1
            else:
                orphans.append(q)
            steps += 1
            processed += 1
            pos += 1
        out = {"status": status, "npe_step": npe_step, "p1_steps": steps, "p1_joins": joins,
               "p1_new_clusters": created, "p1_orphans": len(orphans)}
        if status:
            return out
        rest = [int(i) for i in np.nonzero(in_list)[0] if i >= pos]
        singles = orphans + rest                                              # :117-119, :43-51
        remaining = []
        assigned = 0
        for q in singles:                                                     # :59-66
            A = self._nearest_cluster(clusters, q)
            if A is not None and A != ("empty",) and A[0] >= self.T:
                A[1]["members"].append(q)
                A[1]["size"] = int(_i32(A[1]["size"] + self.ab[q]))
                assigned += 1
            else:
                remaining.append(q)
        cid = np.arange(n, dtype=np.int32)
        rank = np.zeros(n, dtype=np.int32)
        for c in clusters:
            for r, m in enumerate(c["members"]):
                cid[m] = c["id"]
                rank[m] = r
        out.update({"cluster_id": cid, "member_rank": rank,
                    "result_order": np.array([c["id"] for c in clusters] + remaining, dtype=np.int32),
                    "n_multi": len(clusters), "p2_queries": len(singles), "p2_assigned": assigned})
        return out


def sort_order_size(strings, abundance):
    """UniqueSequence.java:176-181, 238-261: stable, abundance desc then string desc."""
    idx = list(range(len(strings)))
    # Python's sort is stable; reverse=True keeps stability semantics of reverseOrder() for
    # distinct keys, and equal keys (identical upper-case strings) must keep input order:
    idx.sort(key=lambda i: (abundance[i], strings[i].upper()), reverse=True)
    # reverse=True reverses the order of EQUAL elements relative to a reversed comparator;
    # restore input order inside runs of equal keys.
    out, i = [], 0
    while i < len(idx):
        j = i
        while j + 1 < len(idx) and (abundance[idx[j + 1]], strings[idx[j + 1]].upper()) == (abundance[idx[i]], strings[idx[i]].upper()):
            j += 1
        out.extend(sorted(idx[i:j + 1]))
        i = j + 1
    return np.array(out, dtype=np.int32)
