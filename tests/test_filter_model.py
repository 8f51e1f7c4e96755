"""CPU model of the upper-bound filter of hmk_bulk_filter (hammock_b200/csrc/hmk_kernels.cuh, hmk_build_profiles):
the exactness argument of DESIGN.md section 5 checked with numpy on random pairs, for every bundled matrix.

For one length L and max shift X (2X+1 diagonals = "lanes"), an exact lane sums, over the L positions of the
database sequence, `M[..] + bias` where the diagonal has a cell and 0 elsewhere, plus a constant folded into
position 0 so that  score >= T  <=>  lane value >= 128.  A filter byte covers lanes 2g and 2g+1 and sums, per
position, the MAXIMUM of the two lanes' entries -- with out-of-diagonal positions counted as `bias` (score 0) and
the constant reduced by (L - cells) * bias to match.  Claims: (1) every filter byte >= both exact lanes it covers,
so no pair with score >= T is filtered out; (2) bytes stay <= 255 whenever the engine's range check passes."""
import os

import numpy as np
import pytest

from hammock_b200 import synth
from oracle import oracle as O


def _tables(q, M, L, X, T, P):
    """exact[lam, j, r] and filt[g, j, r] for profile sequence q (length L), thread-side residue r at position j"""
    bias = int(max(0, -M.min()))
    nl = 2 * X + 1
    exact = np.zeros((nl, L, 24), np.int64)
    filt_lane = np.zeros((nl, L, 24), np.int64)
    for lam in range(nl):
        k = lam - X
        cells = L - abs(k)
        for j in range(L):
            pi = j + k                      # equal lengths: the SECOND argument is the "shorter" one (hmk_pair_score)
            if 0 <= pi < L:
                exact[lam, j, :] = M[:, q[pi]] + bias
                filt_lane[lam, j, :] = M[:, q[pi]] + bias
            else:
                filt_lane[lam, j, :] = bias
        pen = 2 * P * abs(k)
        exact[lam, 0, :] += pen + 128 - T - cells * bias
        filt_lane[lam, 0, :] += pen + 128 - T - L * bias
    ng = (nl + 1) // 2
    filt = np.zeros((ng, L, 24), np.int64)
    for g in range(ng):
        l0, l1 = 2 * g, min(2 * g + 1, nl - 1)
        filt[g] = np.maximum(np.maximum(filt_lane[l0], filt_lane[l1]), 0)
    return exact, filt, bias


@pytest.mark.parametrize("L,X,P", [(12, 3, 0), (12, 3, -1), (9, 2, 0), (10, 3, -2), (7, 2, 0)])
def test_filter_bounds_every_exact_lane(golden_dir, L, X, P):
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    rng = np.random.default_rng(L * 100 + X)
    for name in z.files:
        M = z[name].astype(np.int64)
        T = int(round(1.7 * L))
        bias = int(max(0, -M.min()))
        nl = 2 * X + 1
        # the engine's range checks (choose_scheme): u8 lanes for the exact words, and for the filter bytes
        ok_exact = all(0 <= 2 * P * abs(k) + 128 - T - (L - abs(k)) * bias
                       and 2 * P * abs(k) + 128 - T - (L - abs(k)) * bias + (L - abs(k)) * (int(M.max()) + bias) <= 255
                       for k in range(-X, X + 1))
        worst = max(max(0, 2 * P * abs(k) + 128 - T - L * bias) for k in range(-X, X + 1))
        ok_filter = ok_exact and nl > 4 and worst + L * (int(M.max()) + bias) <= 255
        if not ok_filter:
            continue
        for _ in range(6):
            q = rng.integers(0, 20, size=L)
            exact, filt, _ = _tables(q, M, L, X, T, P)
            D = rng.integers(0, 24, size=(400, L))
            # make some of the database sequences close to q so that real hits occur
            D[:100] = q
            for i in range(100):
                pos = rng.integers(0, L, size=int(rng.integers(0, 4)))
                D[i, pos] = rng.integers(0, 20, size=len(pos))
            jj = np.arange(L)
            ex = exact[:, jj[None, :], D].sum(axis=2)          # [lane, item]
            fl = filt[:, jj[None, :], D].sum(axis=2)           # [group, item]
            assert fl.max() <= 255 and fl.min() >= 0 and ex.max() <= 255, name
            for lam in range(nl):
                assert (fl[lam // 2] >= ex[lam]).all(), (name, lam)
            # the lanes mean what they should: lane >= 128  <=>  that diagonal's score >= T
            best = ex.max(axis=0)
            for i in range(0, 400, 37):
                s = O.score_with_shift(q.astype(np.uint8), D[i].astype(np.uint8), M.astype(np.int32), X, P)[0]
                assert (s >= T) == (best[i] >= 128), (name, i, s, best[i])
                if s >= T:
                    assert best[i] - 128 + T == s


def test_score_symmetry_claim_behind_the_phase2_hit_reuse(golden_dir):
    """S(a, b) == S(b, a) for every pair iff the matrix is symmetric (different lengths: the
    shorter/longer roles do not depend on the argument order; equal lengths: the matrix gets transposed)."""
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    rng = np.random.default_rng(3)
    for name in ("blosum62", "pam250", "mcla71", "blosum30"):
        M = z[name].astype(np.int32)
        assert (M == M.T).all(), name                 # all bundled matrices are symmetric
        for _ in range(300):
            la, lb = int(rng.integers(4, 31)), int(rng.integers(4, 31))
            a = rng.integers(0, 24, size=la).astype(np.uint8)
            b = rng.integers(0, 24, size=lb).astype(np.uint8)
            X = int(rng.integers(0, min(la, lb)))
            P = int(rng.choice([0, -1, -3, 2]))
            assert O.score_with_shift(a, b, M, X, P)[0] == O.score_with_shift(b, a, M, X, P)[0], (name, la, lb, X, P)
    # ... and an asymmetric matrix breaks it for equal lengths only
    A = z["blosum62"].astype(np.int32).copy()
    A[np.triu_indices(24, 1)] -= 3
    diff_equal = diff_unequal = 0
    for _ in range(300):
        la = int(rng.integers(5, 13))
        a = rng.integers(0, 20, size=la).astype(np.uint8)
        b = rng.integers(0, 20, size=la).astype(np.uint8)
        c = rng.integers(0, 20, size=la + 2).astype(np.uint8)
        diff_equal += O.score_with_shift(a, b, A, 2, 0)[0] != O.score_with_shift(b, a, A, 2, 0)[0]
        diff_unequal += O.score_with_shift(a, c, A, 2, 0)[0] != O.score_with_shift(c, a, A, 2, 0)[0]
    assert diff_equal > 0 and diff_unequal == 0
