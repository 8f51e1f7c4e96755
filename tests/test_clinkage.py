"""SURVEY.md 8(f) N1 -- the exact complete-linkage (nearest-neighbour chain) clusterer.
CPU: the C oracle against the second restatement and against hand-checkable facts of the java.util.HashSet emulation.
GPU (-m gpu): hmk_clinkage_cluster through the C ABI against the oracle and the committed MUSI golden."""
import os

import numpy as np
import pytest

import hammock_b200 as hb
from hammock_b200 import _lib, synth
from oracle import oracle as O, pyref_clinkage as PC


def _input_order_case(n, lo, hi, seed):
    """a synthetic set in a shuffled (input) order: runClinkageClustering does not sort (Hammock.java:449-462)"""
    d = synth.generate(n, lo, hi, seed=seed)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    res, offs = d["residues"], d["offsets"]
    seqs = [res[offs[i]:offs[i + 1]] for i in perm]
    o = np.zeros(n + 1, np.int32)
    o[1:] = np.cumsum([len(s) for s in seqs])
    T, X, _ = synth.default_params(d["lengths"])
    return {"residues": np.concatenate(seqs), "offsets": o, "abundance": np.ascontiguousarray(d["abundance"][perm])}, seqs, T, X


def test_java_hashset_emulation_known_order():
    """OpenJDK 8+ HashMap facts, checkable by hand: hash = 553 + id, default capacity 16, resize above 12 entries"""
    s = PC.JavaHashSet()
    for cid in (1, 2, 3):                      # hashes 554, 555, 556 -> bins 10, 11, 12 of 16
        s.add(cid)
    assert list(s) == [1, 2, 3] and len(s.table) == 16
    s = PC.JavaHashSet()
    for cid in (7, 6, 5):                      # hashes 560, 559, 558 -> bins 0, 15, 14: iteration order is by bin
        s.add(cid)
    assert list(s) == [7, 5, 6]
    s = PC.JavaHashSet()
    for cid in (23, 7):                        # 576 and 560 share bin 0 of 16: insertion order inside the bin
        s.add(cid)
    assert list(s) == [23, 7]
    for cid in range(100, 111):                # 13 entries > 12: table doubles, 576 -> bin 0, 560 -> bin 16 of 32
        s.add(cid)
    assert len(s.table) == 32
    order = list(s)
    assert order.index(23) < order.index(7) and order[0] == 23
    assert sorted(order) == sorted([23, 7] + list(range(100, 111)))
    assert order == [23, 100, 101, 102, 7, 103, 104, 105, 106, 107, 108, 109, 110]   # bin 16 of 32: 7 (older) before 103
    s.remove(23)
    assert s.first() == 100                    # 553 + 100 = 653 = 20 * 32 + 13


def test_java_hashset_first_after_removals():
    s = PC.JavaHashSet()
    for cid in range(1, 41):
        s.add(cid)
    assert len(s.table) == 64                  # 40 > 0.75 * 32
    # bins: (553 + id) & 63 -> id 23 lands in bin 0
    assert s.first() == 23
    s.remove(23)
    assert s.first() == 24
    s.add(87)                                  # 640 & 63 = 0
    assert s.first() == 87


@pytest.mark.parametrize("case", [(2, 12, 12, 1), (3, 12, 12, 2), (40, 12, 12, 3), (250, 7, 12, 4), (300, 12, 12, 5), (200, 9, 9, 6)])
def test_clinkage_oracle_vs_second_restatement(blosum62, case):
    n, lo, hi, seed = case
    d, seqs, T, X = _input_order_case(n, lo, hi, seed)
    for T2, P in ((T, 0), (T - 6, -1), (T + 8, 0)):
        R = O.clinkage_cluster(d["residues"], d["offsets"], d["abundance"], blosum62, T2, X, P)
        assert R.status == 0
        c, r, o = PC.clinkage_cluster(seqs, d["abundance"], blosum62, T2, X, P)
        assert (c == R.cluster_id).all() and (r == R.member_rank).all() and (o == R.result_order).all()
        # structure: ids, ranks, complete linkage
        assert sorted(set(R.cluster_id.tolist())) == sorted(R.result_order.tolist())
        for cid in R.result_order:
            mem = np.nonzero(R.cluster_id == cid)[0]
            assert sorted(R.member_rank[mem].tolist()) == list(range(len(mem)))
            assert (len(mem) == 1) == (cid <= n)


def test_clinkage_oracle_status_codes(blosum62):
    res, offs = O.pack(["WVTAPRSLPVLP", "WVT"])
    ab = np.array([1, 1], np.int32)
    assert O.clinkage_cluster(res, offs, ab, blosum62, 20, 3, 0).status == O.ERR_SHIFT_TOO_BIG
    asym = blosum62.copy()
    asym[0, 1] += 1
    res, offs = O.pack(["WVTAPRSLPVLP", "WVTAPRSLPVLA"])
    assert O.clinkage_cluster(res, offs, ab, asym, 20, 3, 0).status == O.ERR_ASYMMETRIC
    assert O.clinkage_cluster(np.zeros(0, np.uint8), np.zeros(1, np.int32), np.zeros(0, np.int32), blosum62, 20, 3, 0).status == O.ERR_EMPTY
    R = O.clinkage_cluster(res, offs, ab, blosum62, 20, 3, 0)          # two near-identical 12-mers merge into cluster n + 2
    assert R.status == 0 and R.result_order.tolist() == [4] and R.cluster_id.tolist() == [4, 4]
    R = O.clinkage_cluster(res[:12], offs[:2], ab[:1], blosum62, 20, 3, 0)   # a single sequence: one ready cluster, no scoring
    assert R.status == 0 and R.result_order.tolist() == [1]


def test_musi_clinkage_golden_is_the_oracle(golden_dir, blosum62):
    z = np.load(os.path.join(golden_dir, "musi_clinkage.npz"))
    T, X, P = (int(v) for v in z["params"])
    assert (T, X, P) == (20, 3, 0) and len(z["abundance"]) == 2457
    R = O.clinkage_cluster(z["residues"], z["offsets"], z["abundance"], blosum62, T, X, P)
    assert R.status == 0 and (R.cluster_id == z["cluster_id"]).all() and (R.member_rank == z["member_rank"]).all()
    assert (R.result_order == z["result_order"]).all() and R.nearest_searches == int(z["nearest_searches"])


# ------------------------------------------------------------------------------------------------ GPU
def _gpu(d, M, T, X, P):
    return hb.clinkage_cluster_arrays(d["residues"], d["offsets"], d["abundance"], M, T, X, P)


@pytest.mark.gpu
def test_musi_clinkage_golden_gpu(golden_dir, blosum62):
    """examples/MUSI through the clusterer Hammock uses for it by default (2 457 <= 10 000 sequences)"""
    z = np.load(os.path.join(golden_dir, "musi_clinkage.npz"))
    T, X, P = (int(v) for v in z["params"])
    rc, G, err = _gpu(z, blosum62, T, X, P)
    assert rc == 0, err
    assert (G.cluster_id == z["cluster_id"]).all() and (G.member_rank == z["member_rank"]).all()
    assert (G.result_order == z["result_order"]).all()
    assert G.n_multi == int((np.bincount(z["cluster_id"]) > 1).sum())
    # the reference-shaped interface: UniqueSequence list in, List<Cluster> out, in the reference's list order
    strs = synth.to_strings(z["residues"], z["offsets"])
    seqs = [hb.UniqueSequence(s, {"x": int(a)}) for s, a in zip(strs, z["abundance"])]
    clusters = hb.run_clinkage_clustering(seqs, blosum62)
    assert [c.get_id() for c in clusters] == z["result_order"].tolist()
    big = max(clusters, key=lambda c: c.get_unique_size())
    mem = np.nonzero(z["cluster_id"] == big.get_id())[0]
    mem = mem[np.argsort(z["member_rank"][mem])]
    assert [s.get_sequence_string() for s in big.get_sequences()] == [strs[m] for m in mem]


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 12, 12, 1), (3, 12, 12, 2), (40, 12, 12, 3), (700, 7, 12, 4), (1500, 12, 12, 5), (900, 9, 9, 6),
                                  (600, 16, 16, 7), (400, 7, 30, 8), (3000, 12, 12, 9),
                                  (12, 12, 12, 10), (13, 12, 12, 11), (25, 9, 12, 12), (49, 12, 12, 13)])   # HashMap growth points
def test_clinkage_gpu_vs_oracle(golden_dir, blosum62, case):
    n, lo, hi, seed = case
    z = np.load(os.path.join(golden_dir, "matrices.npz"))
    d, _, T, X = _input_order_case(n, lo, hi, seed)
    for M, T2, P in ((blosum62, T, 0), (z["pam250"], T + 3, -1), (z["blosum100"], T - 5, 0)):
        R = O.clinkage_cluster(d["residues"], d["offsets"], d["abundance"], M, T2, X, P)
        rc, G, err = _gpu(d, M, T2, X, P)
        assert rc == R.status == 0, (rc, R.status, err)
        assert (G.cluster_id == R.cluster_id).all() and (G.member_rank == R.member_rank).all(), case
        assert len(G.result_order) == len(R.result_order) and (G.result_order == R.result_order).all(), case


@pytest.mark.gpu
def test_clinkage_gpu_status_codes(blosum62):
    res, offs = O.pack(["WVTAPRSLPVLP", "WVT"])
    d = {"residues": res, "offsets": offs, "abundance": np.array([1, 1], np.int32)}
    assert _gpu(d, blosum62, 20, 3, 0)[0] == _lib.STATUS_SHIFT_TOO_BIG
    with pytest.raises(hb.DataException):
        hb.ClinkageSequenceClusterer(hb.ShiftedScorer(blosum62, 0, 3), 20).cluster([hb.UniqueSequence("WVTAPRSLPVLP"), hb.UniqueSequence("WVT")])
    asym = blosum62.copy()
    asym[0, 1] += 1
    res, offs = O.pack(["WVTAPRSLPVLP", "WVTAPRSLPVLA"])
    d = {"residues": res, "offsets": offs, "abundance": np.array([1, 1], np.int32)}
    assert _gpu(d, asym, 20, 3, 0)[0] == _lib.STATUS_UNSUPPORTED
    with pytest.raises(hb.UnsupportedInput):
        hb.ClinkageSequenceClusterer(hb.ShiftedScorer(asym, 0, 3), 20).cluster([hb.UniqueSequence("WVTAPRSLPVLP"), hb.UniqueSequence("WVTAPRSLPVLA")])
    rc, G, _ = _gpu(d, blosum62, 20, 3, 0)
    assert rc == 0 and G.result_order.tolist() == [4] and G.cluster_id.tolist() == [4, 4] and G.n_multi == 1
    empty = {"residues": np.zeros(0, np.uint8), "offsets": np.zeros(1, np.int32), "abundance": np.zeros(0, np.int32)}
    assert _gpu(empty, blosum62, 20, 3, 0)[0] == _lib.STATUS_BAD_ARG
    bad = {"residues": np.array([1, 2, 24, 3], np.uint8), "offsets": np.array([0, 2, 4], np.int32), "abundance": np.array([1, 1], np.int32)}
    assert _gpu(bad, blosum62, 20, 1, 0)[0] == _lib.STATUS_BAD_RESIDUE


# ------------------------------------------------------------------------------------------------ CPU model of the chain kernel
def _chain_kernel_model(pair, ab, T):
    """The LOGIC of hmk_clinkage_chain (hmk_kernels.cuh), not its CUDA -- where it differs in structure from both oracles:
    a thresholded score matrix whose rows are merged by element-wise minimum INTO THE SLOT OF THE STACK TOP, the active set
    as bins of a table that has its FINAL capacity from the start (the claim: HashMap's order-preserving resize makes that
    equivalent to growing it), removal by unlinking, and the ready set rebuilt with real resizes at the end."""
    n = len(ab)
    BELOW = -(2 ** 31) + 1
    D = [[(int(pair[a][b]) if pair[a][b] >= T else BELOW) for b in range(n)] for a in range(n)]
    hsh = lambda cid: ((553 + cid) & 0xFFFFFFFF) ^ (((553 + cid) & 0xFFFFFFFF) >> 16)
    acap = 16
    while n > int(acap * 0.75):
        acap *= 2
    bins = [[] for _ in range(acap)]
    id_of = [i + 1 for i in range(n)]
    slot_of = {i + 1: i for i in range(n)}
    size_of = [int(a) for a in ab]
    members = [[i] for i in range(n)]
    for cid in range(1, n + 1):
        bins[hsh(cid) & (acap - 1)].append(cid)
    nactive, cur_id, stack, ready = n, n + 1, [], []
    while True:
        if not stack:
            if nactive <= 1:
                break
            stack.append(next(b for b in bins if b)[0])
        top = stack[-1]
        ts = slot_of[top]
        best = None                                    # (score, size, id) under score desc, size desc, id asc
        for k in range(n):
            if id_of[k] < 0 or k == ts:
                continue
            cand = (D[ts][k], size_of[k], id_of[k])
            if best is None or (-cand[0], -cand[1], cand[2]) < (-best[0], -best[1], best[2]):
                best = cand
        bscore = best[0] if best else -(2 ** 31)
        if bscore < T:
            stack.pop()
            ready.append(top)
            bins[hsh(top) & (acap - 1)].remove(top)
            nactive -= 1
            id_of[ts] = -1
        elif len(stack) > 1 and stack[-2] == best[2]:
            bs = slot_of[best[2]]
            for k in range(n):
                m = min(D[ts][k], D[bs][k])
                D[ts][k] = m
                D[k][ts] = m
            cur_id += 1
            stack.pop(); stack.pop()
            for cid in (top, best[2]):
                bins[hsh(cid) & (acap - 1)].remove(cid)
            members[ts] = members[ts] + members[bs]
            size_of[ts] = (size_of[ts] + size_of[bs] + 2 ** 31) % 2 ** 32 - 2 ** 31
            id_of[ts], id_of[bs] = cur_id, -1
            slot_of[cur_id] = ts
            bins[hsh(cur_id) & (acap - 1)].append(cur_id)
            nactive -= 1
        else:
            stack.append(best[2])
    ready.append(next(b for b in bins if b)[0])
    rs = PC.JavaHashSet()                              # the ready set grows like a real HashSet (the kernel does the same)
    for cid in ready:
        rs.add(cid)
    order = list(rs)
    cluster_id, member_rank = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for cid in order:
        for r, m in enumerate(members[slot_of[cid]]):
            cluster_id[m], member_rank[m] = cid, r
    return cluster_id, member_rank, np.array(order, np.int32)


@pytest.mark.parametrize("n", [1, 2, 3, 11, 12, 13, 14, 24, 25, 26, 48, 49, 50, 96, 97, 130])
def test_chain_kernel_scheme_model(blosum62, n):
    """sizes around the HashMap growth points (12, 24, 48, 96 entries): the final-capacity table of the kernel against the
    second restatement, which grows a real node table"""
    from oracle.pyref import PyRef
    for seed in range(3):
        d, seqs, T, X = _input_order_case(n, 9, 12, 100 * n + seed)
        ref = PyRef(seqs, d["abundance"], blosum62, T, X, 0, 0)
        pair = np.empty((n, n), dtype=np.int64)
        for q in range(n):
            pair[:, q] = ref.scores(np.arange(n), q)
        for T2 in (T, T - 7):
            want = PC.clinkage_cluster(seqs, d["abundance"], blosum62, T2, X, 0)
            got = _chain_kernel_model(pair, d["abundance"], T2)
            for g, w in zip(got, want):
                assert (g == w).all(), (n, seed, T2)
