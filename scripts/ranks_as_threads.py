"""Debugging aid: run a workload on N GPUs with one host THREAD per rank (handle API + hmk_init_distributed) and compare
every rank's result and counters with the one-GPU run.  usage: ranks_as_threads.py antibodies|n:lo:hi N [N ...] [opt=v ...]"""
import ctypes as C
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import _lib, synth

spec = sys.argv[1]
worlds = [int(a) for a in sys.argv[2:] if "=" not in a]
opts = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[2:] if "=" in a}
M = synth.blosum62()
if spec == "antibodies":
    z = np.load(os.path.join(ROOT, "tests", "golden", "antibodies.npz"))
    d = {k: z[k] for k in ("residues", "offsets", "abundance")}
    T, X, P, K = (int(v) for v in z["params"])
else:
    n, lo, hi = (int(v) for v in spec.split(":"))
    d = synth.generate(n, lo, hi)
    T, X, K = synth.default_params(d["lengths"]); P = 0
L = _lib.load()
c1 = hb.GreedyContext(0, **opts)
c1.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
c1.run(); g1 = c1.download(); s1 = c1.stats(); c1.close()
print("1 GPU:", hb.result_digest(g1.cluster_id, g1.member_rank, g1.result_order)[:12], {k: s1[k] for k in ("p1_steps", "p1_joins", "p1_batches", "p2_queries", "p2_candidates", "p2_assigned", "p2_rounds", "flags")}, flush=True)
for world in worlds:
    for rep in range(2):
        uid = hb.host.nccl_unique_id()
        out = [None] * world
        def rank(r):
            ctx = hb.GreedyContext(r, **opts)
            err = C.create_string_buffer(256)
            buf = (C.c_char * 128).from_buffer_copy(uid)
            assert L.hmk_init_distributed(ctx._h, r, world, buf, err, 256) == 0, err.value
            ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
            ctx.run()
            g = ctx.download(); st = ctx.stats()
            out[r] = (hb.result_digest(g.cluster_id, g.member_rank, g.result_order)[:12], {k: st[k] for k in ("p1_steps", "p1_joins", "p1_batches", "p2_queries", "p2_candidates", "p2_assigned", "p2_rounds", "flags")},
                      int((g.cluster_id != g1.cluster_id).sum()))
            ctx.close()
        th = [threading.Thread(target=rank, args=(r,)) for r in range(world)]
        [t.start() for t in th]; [t.join() for t in th]
        print(f"world {world} rep {rep}:", flush=True)
        for r in range(world):
            print("   rank", r, out[r], flush=True)
