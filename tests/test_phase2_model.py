"""CPU model of the phase-2 scheme of the engine (hmk_p2_window) against the reference's sequential
loop (LimitedGreedySequenceClusterer.java:59-66) on random instances.

What is modelled -- the LOGIC of the kernel, not its CUDA: windows of consecutive queries; the base pass that folds the
first final phase-2 member of a cluster into every candidate pair; fixed-point iterations in which (R) only the clusters
whose joiner set changed rebuild their list and put the candidate queries BEHIND the first change on the work list (once,
generation stamp) and (D) only the queries on the work list are re-decided; lazy evaluation (a candidate whose static
score is below the best valid score found so far is skipped; a pair once refuted by a FINAL member stays refuted); the
commit of a converged window.  The claim tested: this reaches exactly the assignment of the sequential loop, for every
window size, including ties on the score (size, then founder id decide) and sizes that change with the joiners, within
the `window + 2` iterations the host allows (hmk_engine.cu: P.max_iters).  Kernel counterparts: the base pass and the R / D
phases of hmk_p2_window, the chunk loop of hmk_p2_decide_query (hmk_kernels.cuh)."""
import numpy as np
import pytest

JMIN = -(2 ** 31)


def _instance(rng, ns, ncl, density, T, tie_heavy):
    """queries 0..ns-1 (already in order), clusters 0..ncl-1 with founder ids, sizes; candidate pairs with static scores;
    a symmetric table of query-query pair scores"""
    hi = T + (3 if tie_heavy else 25)
    pair = rng.integers(T - (4 if tie_heavy else 12), hi, size=(ns, ns))
    pair = np.minimum(pair, pair.T)
    ab = rng.integers(1, 4 if tie_heavy else 50, size=ns)
    c_size = rng.integers(2, 6 if tie_heavy else 200, size=ncl)
    c_fid = rng.permutation(10 * ncl)[:ncl]
    cand = []
    for q in range(ns):
        cs = np.nonzero(rng.random(ncl) < density)[0]
        cand.append([(int(c), int(rng.integers(T, hi))) for c in cs])        # static score >= T
    return pair, ab, c_size, c_fid, cand


def _key(score, size, fid):
    return (-score, -size, fid)          # smaller tuple == preferred (score desc, size desc, id asc)


def sequential(pair, ab, c_size, c_fid, cand, T):
    """the reference: every query in order against the clusters as they are NOW"""
    ns = len(cand)
    members = {}
    size = c_size.copy()
    A = np.full(ns, -1)
    for q in range(ns):
        best = None
        for c, st in cand[q]:
            js = members.get(c, [])
            sc = min([st] + [int(pair[j, q]) for j in js])
            if sc < T:
                continue
            k = _key(sc, int(size[c]), int(c_fid[c]))
            if best is None or k < best[0]:
                best = (k, c)
        if best is not None:
            c = best[1]
            A[q] = c
            members.setdefault(c, []).append(q)
            size[c] += ab[q]
    return A


def engine_model(pair, ab, c_size, c_fid, cand, T, W, chunk=4):
    ns, ncl = len(cand), len(c_size)
    size = c_size.copy()                     # final sizes (committed windows)
    final = {c: [] for c in range(ncl)}      # final phase-2 members, join order == query order
    by_cluster = {c: [] for c in range(ncl)}
    for q in range(ns):
        for c, _ in cand[q]:
            by_cluster[c].append(q)          # ascending
    A = np.full(ns, -1)
    base = [dict() for _ in range(ns)]       # per pair: static score with member 0 folded in, JMIN = refuted for good
    stamp = np.full(ns, -1)
    gen = 0
    stats = {"evaluations": 0, "iterations": 0}
    for qa in range(0, ns, W):
        qb = min(ns, qa + W)
        # ---- base pass
        for q in range(qa, qb):
            for c, st in cand[q]:
                cl = st
                if final[c]:
                    s = int(pair[final[c][0], q])
                    cl = JMIN if s < T else min(cl, s)
                base[q][c] = cl
        tent = {c: [] for c in range(ncl)}   # tentative joiners of the window, in query order
        dirty = {}                           # cluster -> smallest query whose assignment to / from it changed
        work = list(range(qa, qb))
        t = 0
        while True:
            if t > 0:
                # ---- R: only the changed clusters rebuild their lists and mark the queries behind the change
                work = []
                for c, dpos in dirty.items():
                    tent[c] = [q for q in by_cluster[c] if qa <= q < qb and A[q] == c]
                    for q in by_cluster[c]:
                        if qa <= q < qb and q > dpos and stamp[q] < gen + t:
                            stamp[q] = gen + t
                            work.append(q)
                dirty = {}
            # ---- D: decide the queries on the work list against the lists as of the last R
            new_dirty = {}
            changed = False
            decisions = {}
            for q in work:
                stats["evaluations"] += 1
                best = None                  # (key, cluster)
                cs = cand[q]
                for k0 in range(0, len(cs), chunk):          # chunks of `chunk` candidates, like the warp's 32
                    part = cs[k0:k0 + chunk]
                    alive = [(c, st) for c, st in part if base[q][c] != JMIN and not (best is not None and st < -best[0][0])]
                    # candidates without further members are decided at once
                    todo = []
                    for c, st in alive:
                        nd, start = len(final[c]), (1 if final[c] else 0)
                        if nd + len(tent[c]) > start:
                            todo.append((c, st))
                        else:
                            k = _key(base[q][c], int(size[c]), int(c_fid[c]))
                            if best is None or k < best[0]:
                                best = (k, c)
                    # the others best static score first, skipped once they cannot reach the best any more
                    for c, st in sorted(todo, key=lambda x: -x[1]):
                        if best is not None and st < -best[0][0]:
                            break
                        ok, mn, sz = True, base[q][c], int(size[c])
                        for j in final[c][1:]:
                            s = int(pair[j, q])
                            if s < T:
                                ok = False
                                base[q][c] = JMIN            # refuted by a final member: for good
                                break
                            mn = min(mn, s)
                        if ok:
                            for j in tent[c]:
                                if j >= q:
                                    break                     # joiners at or behind q do not count
                                s = int(pair[j, q])
                                if s < T:
                                    ok = False
                                    break
                                mn = min(mn, s)
                                sz += int(ab[j])
                        if ok:
                            k = _key(mn, sz, int(c_fid[c]))
                            if best is None or k < best[0]:
                                best = (k, c)
                decisions[q] = -1 if best is None else best[1]
            for q, new in decisions.items():                  # (all decisions of an iteration see the same lists)
                old = A[q]
                if new != old:
                    changed = True
                    A[q] = new
                    for c in (old, new):
                        if c >= 0:
                            new_dirty[c] = min(new_dirty.get(c, q), q)
            stats["iterations"] += 1
            dirty = new_dirty
            t += 1
            if not changed:
                break
            assert t <= (qb - qa) + 2, "no fixed point within the guaranteed number of iterations"
        gen += t + 1
        # ---- commit: the lists as of the last R are the final joiners (nothing changed since)
        for c in range(ncl):
            for q in tent[c]:
                final[c].append(q)
                size[c] += ab[q]
    return A, stats


@pytest.mark.parametrize("tie_heavy", [False, True])
@pytest.mark.parametrize("W", [1, 3, 16, 64, 10 ** 6])
def test_phase2_scheme_reaches_the_sequential_assignment(W, tie_heavy):
    rng = np.random.default_rng(1234 + W + (7 if tie_heavy else 0))
    for trial in range(12):
        ns = int(rng.integers(5, 120))
        ncl = int(rng.integers(1, 12))
        T = 20
        inst = _instance(rng, ns, ncl, float(rng.choice([0.1, 0.4, 0.9])), T, tie_heavy)
        want = sequential(*inst, T)
        got, stats = engine_model(*inst, T, W, chunk=int(rng.choice([1, 2, 4, 32])))
        assert (got == want).all(), (W, tie_heavy, trial, np.nonzero(got != want)[0][:5])
        assert stats["iterations"] >= 1


def test_phase2_scheme_work_is_bounded():
    """late iterations re-decide few queries: the work lists shrink (this is what the cluster-driven marking buys)"""
    rng = np.random.default_rng(5)
    inst = _instance(rng, 400, 10, 0.5, 20, False)
    want = sequential(*inst, 20)
    got, stats = engine_model(*inst, 20, 128)
    assert (got == want).all()
    assert stats["evaluations"] < 400 * stats["iterations"] / 2
