"""N > 1 host logic on CPU: world_size-2 gloo rendezvous of the NCCL id, and the sharding / best-hit
merge rules the library applies on the GPUs (the collectives themselves run inside the CUDA library
over NCCL and are covered by scripts/dist_check.py on multi-GPU boxes)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hammock_b200 import distributed as hd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = hd.exchange_unique_id(dist, rank, lambda: bytes((7 * i + 3) % 256 for i in range(128)))
    # every rank scans its stripe of a common score vector and keeps its top-k; the gathered lists
    # must merge to the global top-k (what phase 1 does with ncclAllGather + hmk_topk_merge)
    rng = np.random.default_rng(5)
    n, kb = 1000, 8
    scores = rng.integers(-30, 60, size=n)
    lo, hi = hd.shard_range(n, world, rank)
    keys = np.array([hd.make_key(int(scores[i]), i) for i in range(lo, hi) if scores[i] >= 20], dtype=np.uint64)
    keys = np.sort(keys)[::-1]
    mine = np.zeros(kb, dtype=np.int64)
    mine[:min(kb, len(keys))] = keys[:kb].view(np.int64)
    cnt = torch.tensor([min(kb, len(keys)), int(len(keys) > kb)])
    g_keys = [torch.zeros(kb, dtype=torch.int64) for _ in range(world)]
    g_cnt = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(g_keys, torch.from_numpy(mine))
    dist.all_gather(g_cnt, cnt)
    lists = [g_keys[r].numpy().view(np.uint64)[:int(g_cnt[r][0])] for r in range(world)]
    merged, ovf = hd.merge_best_hits(lists, [int(g_cnt[r][1]) for r in range(world)], kb)
    q.put((rank, uid, merged.tolist(), ovf))
    dist.destroy_process_group()


def test_world2_gloo_rendezvous_and_best_hit_merge():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_uid = bytes((7 * i + 3) % 256 for i in range(128))
    assert all(g[1] == expect_uid for g in got)
    # reference: global top-k over the unsharded vector
    rng = np.random.default_rng(5)
    scores = rng.integers(-30, 60, size=1000)
    allk = np.sort(np.array([hd.make_key(int(s), i) for i, s in enumerate(scores) if s >= 20], dtype=np.uint64))[::-1]
    assert got[0][2] == got[1][2] == allk[:8].tolist()
    assert got[0][3] is True and got[1][3] is True


def test_key_order_is_the_reference_tie_break():
    """score desc, then (abundance desc, id asc) via the tie rank (ClinkageSequenceClusterer.java:258-293)"""
    k = hd.make_key
    assert k(30, 5) > k(29, 0) and k(30, 4) > k(30, 5) and k(-5, 0) > k(-6, 0) and k(0, 0) > k(-1, 0)


def test_shard_ranges():
    for n in (0, 5, 1000, 999983):
        for w in (1, 2, 4, 8):
            parts = [hd.shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n and all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
