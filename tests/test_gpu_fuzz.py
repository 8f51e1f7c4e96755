"""Randomised differential test (GPU vs oracle): small inputs, random parameters -- thresholds around and far
from the automatic value, shifts up to the limit, penalties of both signs, tiny / huge cluster limits, every
bundled matrix, uniform / mixed / long lengths, sorted and unsorted abundances, tiny batches and candidate
lists (restarts, grows, look-ahead aborts).  Every status code and every assignment must agree."""
import os

import numpy as np
import pytest

import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _case(rng, mats, names):
    kind = rng.integers(0, 6)
    if kind == 0:
        lo = hi = int(rng.integers(5, 13))
    elif kind == 1:
        lo = int(rng.integers(4, 10)); hi = int(rng.integers(lo + 1, 13))
    elif kind == 2:
        lo = hi = int(rng.integers(13, 31))
    elif kind == 3:
        lo = int(rng.integers(13, 20)); hi = lo + int(rng.integers(1, 6))
    elif kind == 4:
        lo = int(rng.integers(5, 12)); hi = int(rng.integers(14, 40))
    else:
        lo = hi = 12
    n = int(rng.integers(2, 500))
    d = synth.generate(n, lo, hi, seed=int(rng.integers(1, 1 << 30)), top_abundance=int(rng.choice([1, 3, 50, 100000])))
    if rng.random() < 0.25:
        d["abundance"] = np.ascontiguousarray(rng.permutation(d["abundance"]))
    T0, X0, K0 = synth.default_params(d["lengths"])
    T = int(T0 + rng.choice([0, 0, -3, 4, -12, 15, 40, -60]))
    X = int(rng.choice([X0, X0, 0, 1, lo - 1, lo, X0 + 1]))
    P = int(rng.choice([0, 0, 0, -1, -2, -5, 1]))
    K = int(rng.choice([K0, K0, 0, 1, 2, n // 3, n, 2 * n]))
    m = names[int(rng.integers(0, len(names)))]
    opts = {}
    if rng.random() < 0.5:
        opts = {"batch": int(rng.choice([1, 2, 5, 17, 64])), "kb": int(rng.choice([1, 2, 8])), "capq": int(rng.choice([1, 4, 256])),
                "p2_window": int(rng.choice([1, 7, 64, 65536])), "lookahead": int(rng.integers(0, 2)) + n % 2}   # (0..2; derived from n so that the random stream stays as it was)
    if rng.random() < 0.4:      # the older code paths stay covered: exact packed kernel, separate phase-2 founder pass
        opts = dict(opts, filter=int(rng.integers(0, 2)), reuse=int(rng.integers(0, 2)))
    M = mats[m]
    if rng.random() < 0.12:     # asymmetric matrix: S(a, b) != S(b, a) for equal lengths, phase-2 hit reuse must switch off
        M = M.copy()
        for _ in range(int(rng.integers(1, 6))):
            i, j = (int(v) for v in rng.integers(0, 20, size=2))
            if i != j:
                M[i, j] += int(rng.integers(1, 4))
        m = m + "+asym"
    return d, M, T, max(X, 0), P, K, opts, (n, lo, hi, m)


def test_fuzz_against_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    mats = {k: z[k] for k in z.files}
    names = sorted(mats)
    asym = mats["blosum62"].copy()
    asym[np.triu_indices(24, 1)] -= 1          # orientation M[shorter][longer] matters
    mats["asym"] = asym
    names.append("asym")
    big = mats["blosum62"] * 700               # does not fit 8/16-bit lanes: generic kernel, int32 sums
    mats["big"] = big
    names.append("big")
    rng = np.random.default_rng(20260101)
    seen = {"status": set(), "path": set()}
    for it in range(220):
        d, M, T, X, P, K, opts, tag = _case(rng, mats, names)
        if tag[3] == "big":
            T *= 700
        R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K, nthreads=1)
        ctx = hb.GreedyContext(0, **opts)
        try:
            ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
            rc, _ = ctx.run_status()
            st = ctx.stats()
            what = f"iter {it} {tag} T={T} X={X} P={P} K={K} opts={opts} path={st['fast_path']}"
            assert rc == R.status, what
            if rc == 2:
                assert st["error_step"] == R.counters["npe_step"], what
            if rc == 0:
                G = ctx.download()
                assert (G.cluster_id == R.cluster_id).all(), what
                assert (G.member_rank == R.member_rank).all(), what
                assert len(G.result_order) == len(R.result_order) and (G.result_order == R.result_order).all(), what
                assert G.n_multi == R.n_multi, what
                assert st["p1_steps"] == R.counters["p1_steps"] and st["p2_assigned"] == R.counters["p2_assigned"], what
            seen["status"].add(rc)
            seen["path"].add(st["fast_path"])
        finally:
            ctx.close()
    assert seen["status"] >= {0, 1, 2} and seen["path"] == {0, 1, 2}, seen


@pytest.mark.timeout(300, method="thread")
def test_fuzz_midsize_default_options(golden_dir):
    """A few thousand to 20 k sequences with the DEFAULT tuning (batches of 448, look-ahead on a side stream,
    parallel resolver windows, phase-2 hit reuse): the paths the small cases above hardly reach.  Each case runs
    twice on the same context -- the look-ahead races with the resolver by design, the result must not depend on it."""
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    mats = {k: z[k] for k in z.files}
    rng = np.random.default_rng(77)
    ncpu = os.cpu_count() or 1
    for it in range(10):
        n = int(rng.integers(3000, 20000))
        kind = int(rng.integers(0, 4))
        lo, hi = [(12, 12), (9, 9), (7, 12), (16, 16)][kind]
        m = str(rng.choice(["blosum62", "blosum62", "blosum45", "pam250"]))
        d = synth.generate(n, lo, hi, seed=int(rng.integers(1, 1 << 30)), top_abundance=int(rng.choice([50, 100000])))
        T0, X0, K0 = synth.default_params(d["lengths"])
        T = int(T0 + rng.choice([0, 0, -4, 5]))
        K = int(rng.choice([K0, K0, n // 10, n // 3, n]))
        P = int(rng.choice([0, 0, -1]))
        opts = {} if it % 3 else {"kb": int(rng.choice([2, 4]))}        # short candidate lists: restarts with the look-ahead on
        R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K, nthreads=ncpu)
        ctx = hb.GreedyContext(0, **opts)
        try:
            ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K)
            for rep in range(2):
                rc, _ = ctx.run_status()
                st = ctx.stats()
                what = f"iter {it} rep {rep} n={n} len={lo}-{hi} {m} T={T} X={X0} P={P} K={K} opts={opts} path={st['fast_path']}"
                assert rc == R.status, what
                if rc == 0:
                    G = ctx.download()
                    assert (G.cluster_id == R.cluster_id).all(), what
                    assert (G.member_rank == R.member_rank).all(), what
                    assert (G.result_order == R.result_order).all() and G.n_multi == R.n_multi, what
                    assert st["p1_steps"] == R.counters["p1_steps"] and st["p2_assigned"] == R.counters["p2_assigned"], what
        finally:
            ctx.close()
