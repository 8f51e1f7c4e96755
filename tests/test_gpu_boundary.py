"""GPU tests of the C-ABI boundary as a long-lived host (one JVM) would use it: argument validation, several
contexts from several threads, several devices in one process, and the one-process multi-GPU entry
hmk_greedy_cluster_multi (SURVEY.md 8b `n_gpus`).  The multi-device tests skip on a one-GPU box."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

import hammock_b200 as hb
from hammock_b200 import _lib, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _oracle(d, M, T, X, P, K):
    return O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K, nthreads=os.cpu_count() or 1)


def _same(R, G):
    return ((R.cluster_id == G.cluster_id).all() and (R.member_rank == G.member_rank).all()
            and len(R.result_order) == len(G.result_order) and (R.result_order == G.result_order).all()
            and R.n_multi == G.n_multi)


def _workload(n, lo=12, hi=12, seed=5):
    d = synth.generate(n, lo, hi, seed=seed)
    return d, synth.default_params(d["lengths"])


def test_option_ranges_are_checked():
    L = _lib.load()
    ctx = hb.GreedyContext(0)
    try:
        for name, bad in (("qt", 100000), ("qt", -1), ("p2_chunk", -5), ("p2_window", 0), ("hit_cap", -1), ("hit_cap", 1 << 40),
                          ("waves", 0), ("waves", 1000), ("kb", 0), ("kb", 33), ("batch", 513), ("lookahead", 3), ("capq", 0),
                          ("xhit_cap", -1), ("p2_spec", 3), ("no_such_option", 1)):
            assert L.hmk_set_option(ctx._h, name.encode(), bad) == _lib.STATUS_BAD_ARG, (name, bad)
        for name, ok in (("qt", 0), ("qt", 16), ("kb", 32), ("batch", 512), ("lookahead", 2), ("p2_window", 1), ("hit_cap", 1024)):
            assert L.hmk_set_option(ctx._h, name.encode(), ok) == _lib.STATUS_OK, (name, ok)
    finally:
        ctx.close()


def test_bad_arguments_return_bad_arg_not_a_crash(blosum62):
    L = _lib.load()
    ctx = hb.GreedyContext(0)
    err = C.create_string_buffer(256)
    M = np.ascontiguousarray(blosum62, dtype=np.int32).reshape(-1)
    i32, u8 = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    try:
        # n == 0 with every array NULL is a valid (empty) input
        gin = _lib.GreedyIn(0, None, None, None, M.ctypes.data_as(i32), 20, 3, 0, 5)
        assert L.hmk_upload(ctx._h, C.byref(gin), err, 256) == _lib.STATUS_OK
        assert L.hmk_run(ctx._h, err, 256) == _lib.STATUS_OK
        out = _lib.GreedyOut(None, None, None, 7, 7, 0)
        assert L.hmk_download(ctx._h, C.byref(out), err, 256) == _lib.STATUS_OK and out.n_result == 0
        res = np.zeros(24, np.uint8)
        ab = np.ones(2, np.int32)
        for offs in ([1, 12, 24], [0, 12, 6]):        # offsets[0] != 0, not monotone
            o = np.array(offs, np.int32)
            gin = _lib.GreedyIn(2, res.ctypes.data_as(u8), o.ctypes.data_as(i32), ab.ctypes.data_as(i32), M.ctypes.data_as(i32), 20, 3, 0, 1)
            assert L.hmk_upload(ctx._h, C.byref(gin), err, 256) == _lib.STATUS_BAD_ARG, offs
            assert L.hmk_run(ctx._h, err, 256) == _lib.STATUS_BAD_ARG           # nothing is uploaded after a rejected upload
        gin = _lib.GreedyIn(2, None, None, None, M.ctypes.data_as(i32), 20, 3, 0, 1)
        assert L.hmk_upload(ctx._h, C.byref(gin), err, 256) == _lib.STATUS_BAD_ARG
        o = np.array([0, 12, 24], np.int32)
        gin = _lib.GreedyIn(2, res.ctypes.data_as(u8), o.ctypes.data_as(i32), ab.ctypes.data_as(i32), None, 20, 3, 0, 1)
        assert L.hmk_upload(ctx._h, C.byref(gin), err, 256) == _lib.STATUS_BAD_ARG
        # a good upload + run, then a download into NULL arrays
        gin = _lib.GreedyIn(2, res.ctypes.data_as(u8), o.ctypes.data_as(i32), ab.ctypes.data_as(i32), M.ctypes.data_as(i32), 20, 3, 0, 1)
        assert L.hmk_upload(ctx._h, C.byref(gin), err, 256) == _lib.STATUS_OK
        assert L.hmk_run(ctx._h, err, 256) == _lib.STATUS_OK
        out = _lib.GreedyOut(None, None, None, 0, 0, 0)
        assert L.hmk_download(ctx._h, C.byref(out), err, 256) == _lib.STATUS_BAD_ARG
    finally:
        ctx.close()


def test_two_contexts_from_two_threads_on_one_device(blosum62):
    """'re-entrant per handle': two host threads drive their own contexts on the same device at the same time"""
    work = [_workload(6000, 12, 12, seed=11), _workload(5000, 7, 12, seed=12)]
    want = [_oracle(d, blosum62, T, X, 0, K) for d, (T, X, K) in work]
    got = [None, None]
    errs = []

    def run(i):
        try:
            d, (T, X, K) = work[i]
            ctx = hb.GreedyContext(0)
            try:
                for _ in range(3):
                    ctx.upload(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K)
                    ctx.run()
                    got[i] = ctx.download()
                    assert _same(want[i], got[i]), i
            finally:
                ctx.close()
        except Exception as e:      # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_a_small_context_does_not_lower_another_contexts_shared_memory_limit(blosum62):
    """cudaFuncSetAttribute REPLACES the opt-in shared-memory size of a kernel on the device: a context that needs little
    must not undercut a context that configured a full tile before (the limit is tracked per device, process-wide)"""
    big, (T, X, K) = _workload(30000, 12, 12, seed=61)
    small, (Ts, Xs, Ks) = _workload(40, 12, 12, seed=62)
    Rb, Rs = _oracle(big, blosum62, T, X, 0, K), _oracle(small, blosum62, Ts, Xs, 0, Ks)
    a = hb.GreedyContext(0)
    b = hb.GreedyContext(0)
    try:
        for _ in range(2):
            a.upload(big["residues"], big["offsets"], big["abundance"], blosum62, T, X, 0, K)
            a.run()
            assert _same(Rb, a.download())
            b.upload(small["residues"], small["offsets"], small["abundance"], blosum62, Ts, Xs, 0, Ks)
            rc, _ = b.run_status()
            assert rc == Rs.status
            if rc == 0:
                assert _same(Rs, b.download())
    finally:
        a.close()
        b.close()


def test_kept_hit_overflow_is_reported_and_repaired(blosum62):
    """an overflow of the buffer that keeps the phase-1 hits for phase 2 must show in hmk_stats.flags, give the same
    result through the separate founder pass, and the next run on the context must size the buffer right"""
    d, (T, X, K) = _workload(20000, 12, 12, seed=21)
    R = _oracle(d, blosum62, T, X, 0, K)
    ctx = hb.GreedyContext(0, xhit_cap=4096)
    try:
        ctx.upload(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K)
        ctx.run()
        st = ctx.stats()
        assert _same(R, ctx.download())
        assert st["flags"] & _lib.FLAG_XHIT_OVERFLOW and not st["flags"] & _lib.FLAG_P2_REUSED, st
        assert st["xhits_kept"] > st["xhits_capacity"] >= 4096
        ctx.set_option("xhit_cap", 0)          # automatic again: sized from what the run above needed
        ctx.run()
        st = ctx.stats()
        assert _same(R, ctx.download())
        assert st["flags"] & _lib.FLAG_P2_REUSED and not st["flags"] & _lib.FLAG_XHIT_OVERFLOW, st
        assert st["xhits_kept"] <= st["xhits_capacity"]
    finally:
        ctx.close()


def test_two_devices_in_one_process(blosum62):
    """device 0, then device 1, then device 0 again from one process (the opt-in shared-memory size is a per-device
    kernel attribute: it must be configured on each device)"""
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    d, (T, X, K) = _workload(8000, 12, 12, seed=31)
    R = _oracle(d, blosum62, T, X, 0, K)
    for dev in (0, 1, 0, 1):
        rc, G, _, err = hb.greedy_cluster_arrays(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K, device=dev)
        assert rc == 0, (dev, err)
        assert _same(R, G), dev
    for dev in (1, 0):
        ctx = hb.GreedyContext(dev)
        try:
            ctx.upload(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K)
            ctx.run()
            assert _same(R, ctx.download()), dev
        finally:
            ctx.close()
    _lib.load().hmk_release_cached()


@pytest.mark.parametrize("shape", [(30000, 12, 12), (12000, 7, 12), (900, 12, 12)])
def test_one_process_multi_gpu_entry(blosum62, shape):
    """hmk_greedy_cluster_multi: one call, worker threads and the NCCL communicator inside the library; result ==
    one-GPU result == oracle.  Called twice (the second call reuses the cached group), then on one device again."""
    ng = _ndev()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    d, (T, X, K) = _workload(*shape, seed=41)
    R = _oracle(d, blosum62, T, X, 0, K)
    for devs in ([0, 1], [0, 1], list(range(ng))[::-1]):
        rc, G, _, err = hb.greedy_cluster_arrays(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K, devices=devs)
        assert rc == 0, (devs, err)
        assert _same(R, G), devs
    rc, G, _, err = hb.greedy_cluster_arrays(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K, device=0)
    assert rc == 0 and _same(R, G)
    # the reference's failure modes come back through the multi entry as well
    res, offs = O.pack(["WWWWWWWWWW", "AAAAAAAAAA", "CCCCCCCCCC"])
    rc, _, step, _ = hb.greedy_cluster_arrays(res, offs, np.array([3, 2, 1], np.int32), blosum62, 24, 2, 0, 2, devices=[0, 1])
    assert rc == _lib.STATUS_NULL_CLUSTER and step == 0
    _lib.load().hmk_release_cached()


@pytest.mark.parametrize("lookahead", [0, 1, 2])
def test_ranks_as_threads_match_one_gpu(blosum62, lookahead):
    """the multi-rank path through the handle API (what one process per GPU does), with the look-ahead off and on:
    every rank returns the one-GPU result"""
    ng = min(_ndev(), 4)
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    d, (T, X, K) = _workload(40000, 12, 12, seed=51)
    R = _oracle(d, blosum62, T, X, 0, K)
    uid = hb.host.nccl_unique_id()
    L = _lib.load()
    got, errs = [None] * ng, []

    def rank(r):
        try:
            ctx = hb.GreedyContext(r, lookahead=lookahead, batch=96)
            try:
                err = C.create_string_buffer(256)
                buf = (C.c_char * 128).from_buffer_copy(uid)
                assert L.hmk_init_distributed(ctx._h, r, ng, buf, err, 256) == 0, err.value
                ctx.upload(d["residues"], d["offsets"], d["abundance"], blosum62, T, X, 0, K)
                ctx.run()
                got[r] = ctx.download()
            finally:
                ctx.close()
        except Exception as e:      # noqa: BLE001
            errs.append((r, repr(e)))

    th = [threading.Thread(target=rank, args=(r,)) for r in range(ng)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for r in range(ng):
        assert _same(R, got[r]), r
