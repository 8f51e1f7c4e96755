"""CPU tests: the oracle against every pin that exists for this path (all survey-derived; the
reference has no tests -- parity unpinned), the second restatement, and the committed goldens."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle.pyref import PyRef
from tests import kats


@pytest.fixture(scope="module")
def mats(golden_dir):
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    return {k: z[k] for k in z.files}


def test_scorer_kats(mats):
    for a, b, X, P, m, score, shift in kats.SCORER_KATS:
        got = O.score_with_shift(a, b, mats[m], X, P)
        assert got[0] == score, (a, b, got)
        if shift is not None:
            assert got[1] == shift, (a, b, got)


def test_cells_and_shifts():
    for (l1, l2, X), cells, shifts in kats.CELL_KATS:
        assert O.pair_cells(l1, l2, X) == cells
        assert O.pair_shifts(l1, l2, X) == shifts
        m, n = min(l1, l2), max(l1, l2)
        assert cells == (n - m + 1) * m + 2 * X * m - X * (X + 1)


def test_shift_too_big(blosum62):
    with pytest.raises(O.OracleError) as e:
        O.score_with_shift("ACD", "ACDEFG", blosum62, 3, 0)
    assert e.value.status == O.ERR_SHIFT_TOO_BIG


def test_matrix_facts(mats, golden_dir, blosum62):
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    assert len(mats) == 17 and "gonnet250" not in mats            # gonnet250.txt is rejected by the loader
    assert meta["matrix_rejected"] == {"gonnet250.txt": O.ERR_FILE_FORMAT}
    for k, v in mats.items():
        assert v.shape == (24, 24) and (v == v.T).all() and v.min() >= -10 and v.max() <= 20, k
    assert (mats["blosum62"] == blosum62).all()


def test_matrix_loader_quirks(tmp_path, blosum62):
    rows = ["#c", "   A R"] + [f"{O.ALPHABET[i]} " + " ".join(str(int(v)) for v in blosum62[i]) for i in range(24)]
    p = tmp_path / "m.txt"
    p.write_text("\n".join(rows) + "\n")
    assert (O.load_matrix(str(p)) == blosum62).all()
    # rows are taken in FILE order, labels ignored (the header check in the reference is dead code)
    sw = rows[:2] + [rows[3], rows[2]] + rows[4:]
    p.write_text("\n".join(sw) + "\n")
    M = O.load_matrix(str(p))
    assert (M[0] == blosum62[1]).all() and (M[1] == blosum62[0]).all()
    p.write_text("\n".join(rows[:10]) + "\n")                     # fewer rows: zeros
    assert (O.load_matrix(str(p))[8:] == 0).all()
    for bad in (rows + [rows[2]], rows[:5] + [""] + rows[5:], rows[:5] + ["A 1 2 3"] + rows[5:]):
        p.write_text("\n".join(bad) + "\n")
        with pytest.raises(O.OracleError) as e:
            O.load_matrix(str(p))
        assert e.value.status == O.ERR_FILE_FORMAT


def _micro(blosum62):
    strs = [s for s, _ in kats.MICRO]
    ab = np.array([a for _, a in kats.MICRO], dtype=np.int32)
    res, offs = O.pack(strs)
    perm = O.sort_order_size(res, offs, ab)
    ordered = [strs[i] for i in perm]
    r2, o2 = O.pack(ordered)
    return ordered, r2, o2, ab[perm]


def test_micro_fixture(blosum62):
    ordered, r2, o2, a2 = _micro(blosum62)
    assert ordered == kats.MICRO_ORDER
    for K, (clusters, singles) in kats.MICRO_EXPECT.items():
        R = O.greedy_cluster(r2, o2, a2, blosum62, 24, 2, 0, K)
        assert R.status == 0
        assert kats.result_to_lists(R.cluster_id, R.member_rank, R.result_order, R.n_multi) == (clusters, singles)
    R = O.greedy_cluster(r2, o2, a2, blosum62, 24, 2, 0, 0)      # K = 0: phase 1 skipped, no NPE
    assert R.status == 0 and R.n_multi == 0 and (R.cluster_id == np.arange(14)).all()


def test_null_cluster_quirk(blosum62):
    # first query has no partner while no cluster exists -> NullPointerException in the reference
    strs = ["WWWWWWWWWW", "AAAAAAAAAA", "CCCCCCCCCC"]
    res, offs = O.pack(strs)
    R = O.greedy_cluster(res, offs, np.array([3, 2, 1], np.int32), blosum62, 24, 2, 0, 2)
    assert R.status == O.ERR_NULL_CLUSTER and R.counters["npe_step"] == 0
    # one sequence, K >= 1: both collections empty -> NPE as well
    res, offs = O.pack(["WWWWWWWWWW"])
    R = O.greedy_cluster(res, offs, np.array([1], np.int32), blosum62, 24, 2, 0, 1)
    assert R.status == O.ERR_NULL_CLUSTER


def test_defaults_and_order():
    offs = np.arange(0, 12 * 101, 12, dtype=np.int32)
    assert O.default_params(offs) == (20, 3, 3)                    # round(2.5) = 3 (Math.round half up)
    offs = np.concatenate([[0], np.cumsum([7, 8, 9, 30, 30, 30])]).astype(np.int32)
    T, X, K = O.default_params(offs)
    assert (T, X, K) == (32, 5, 0) and O.check_max_shift(offs, 9) == 6


@pytest.mark.parametrize("name", ["musi", "antibodies"])
def test_golden_structure(golden_dir, name):
    """Structural pins of SURVEY.md section 6 (survey-derived)."""
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))[name]
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    expect = {"musi": (2457, 61, 67, 6, 2329, 574, 169039, 1121022, 702, [23, 21, 20, 20, 19]),
              "antibodies": (74041, 1851, 1923, 72, 70267, 7196, 130649021, 605487235, 10970, [60, 58, 49, 42, 39])}[name]
    c = meta["counters"]
    sizes = np.bincount(z["cluster_id"])
    multi = np.sort(sizes[sizes > 1])[::-1]
    got = (meta["n"], int(z["n_multi"]), c["p1_steps"], c["p1_joins"], c["p2_queries"], c["p2_assigned"],
           c["p2_pairs_early"], c["p2_pairs_dense"], int(multi.sum()), multi[:5].tolist())
    assert got == expect


def test_oracle_reproduces_musi_golden(golden_dir, blosum62):
    z = np.load(os.path.join(golden_dir, "musi.npz"))
    T, X, P, K = (int(v) for v in z["params"])
    for nt in (1, 4):
        R = O.greedy_cluster(z["residues"], z["offsets"], z["abundance"], blosum62, T, X, P, K, nthreads=nt)
        assert R.status == 0 and (R.cluster_id == z["cluster_id"]).all() and (R.member_rank == z["member_rank"]).all()
        assert (R.result_order == z["result_order"]).all()


def test_pyref_agrees_on_random_inputs(blosum62, mats):
    """differential test of the two restatements: mixed lengths, penalties, other matrices, odd K"""
    from hammock_b200 import synth
    cases = [(300, 12, 12, 0, "blosum62", None), (300, 7, 12, -1, "blosum62", None), (250, 9, 9, 0, "pam250", 3),
             (200, 7, 16, -2, "blosum45", 40), (120, 12, 12, 0, "blosum62", 200)]
    for n, lo, hi, P, m, K in cases:
        d = synth.generate(n, lo, hi, seed=n + hi)
        T, X, K0 = synth.default_params(d["lengths"])
        K = K0 if K is None else K
        R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K)
        seqs = [d["residues"][d["offsets"][i]:d["offsets"][i + 1]] for i in range(n)]
        Pr = PyRef(seqs, d["abundance"], mats[m], T, X, P, K).run()
        assert R.status == Pr["status"]
        if R.status == 0:
            assert (R.cluster_id == Pr["cluster_id"]).all() and (R.member_rank == Pr["member_rank"]).all()
            assert (R.result_order == Pr["result_order"]).all()
            assert R.counters["p1_steps"] == Pr["p1_steps"] and R.counters["p2_assigned"] == Pr["p2_assigned"]
