"""Quick GPU bring-up check: prints one line per case instead of stopping at the first failure."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O

M = synth.blosum62()
G = os.path.join(ROOT, "tests", "golden")
NT = os.cpu_count() or 1
print("cpus", NT, flush=True)


def run(d, T, X, P, K, **opts):
    ctx = hb.GreedyContext(0, **opts)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
    t = time.time(); rc, msg = ctx.run_status(); wall = time.time() - t
    st = ctx.stats()
    g = ctx.download() if rc == 0 else None
    ctx.close()
    return rc, g, st, wall


def cmp(name, d, T, X, P, K, ref=None, **opts):
    try:
        rc, g, st, wall = run(d, T, X, P, K, **opts)
        if ref is None:
            R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K, nthreads=NT)
            ref = (R.status, R.cluster_id, R.member_rank, R.result_order)
        ok = rc == ref[0] and (rc != 0 or ((g.cluster_id == ref[1]).all() and (g.member_rank == ref[2]).all()
                                            and len(g.result_order) == len(ref[3]) and (g.result_order == ref[3]).all()))
        extra = ""
        if rc == 0 and not ok:
            bad = np.nonzero(g.cluster_id != ref[1])[0]
            extra = f" first_bad={bad[:8].tolist()} gpu={g.cluster_id[bad[:8]].tolist()} ref={ref[1][bad[:8]].tolist()} nbad={len(bad)}"
        keys = ("p1_steps", "p1_joins", "p1_new_clusters", "p1_orphans", "p1_batches", "p1_restarts", "p2_queries",
                "p2_assigned", "p2_rounds", "p2_hits", "p2_candidates", "bulk_pairs", "scalar_pairs", "fast_path", "lane_bits")
        print(f"[{'OK' if ok else 'FAIL'}] {name} opts={opts} rc={rc} wall={wall*1e3:.1f}ms dev={st['total_ms']:.1f}ms "
              f"p1={st['phase1_ms']:.1f} p2={st['phase2_ms']:.1f} " + " ".join(f"{k}={st[k]}" for k in keys) + extra, flush=True)
    except Exception:
        print(f"[EXC] {name} opts={opts}"); traceback.print_exc(); sys.stdout.flush()


z = np.load(os.path.join(G, "musi.npz"))
# 1. pair scores
for generic in (0, 1):
    try:
        ctx = hb.GreedyContext(0, force_generic=generic)
        ctx.upload(z["residues"], z["offsets"], z["abundance"], M, 20, 3, 0, 61)
        ids = np.arange(0, 2457, 9, dtype=np.int32)
        got = ctx.score_block(ids, ids)
        res, offs = z["residues"], z["offsets"]
        nbad = 0
        for i, a in enumerate(ids):
            for j, b in enumerate(ids):
                exp = O.score_with_shift(res[offs[a]:offs[a+1]], res[offs[b]:offs[b+1]], M, 3, 0)[0]
                if got[i, j] != exp:
                    nbad += 1
                    if nbad < 5: print("  score mismatch", a, b, got[i, j], exp)
        print(f"[{'OK' if nbad == 0 else 'FAIL'}] score_block generic={generic} n={len(ids)**2} bad={nbad}", flush=True)
        ctx.close()
    except Exception:
        traceback.print_exc()

T, X, P, K = (int(v) for v in z["params"])
ref = (0, z["cluster_id"], z["member_rank"], z["result_order"])
for opts in ({}, {"force_generic": 1}, {"batch": 7, "kb": 1}, {"batch": 64, "kb": 2, "qt": 16}, {"p2_chunk": 1024, "hit_cap": 2048, "p2_window": 100, "capq": 1}):
    cmp("musi", z, T, X, P, K, ref, **opts)
d = synth.generate(3000, 7, 12, seed=5)
T2, X2, K2 = synth.default_params(d["lengths"])
cmp("synth3000_mixed", d, T2, X2, 0, K2)
d = synth.generate(20000)
T2, X2, K2 = synth.default_params(d["lengths"])
cmp("synth20k", d, T2, X2, 0, K2)
cmp("synth20k", d, T2, X2, -1, K2, batch=64)
z = np.load(os.path.join(G, "antibodies.npz"))
T, X, P, K = (int(v) for v in z["params"])
cmp("antibodies", z, T, X, P, K, (0, z["cluster_id"], z["member_rank"], z["result_order"]))
if "--big" in sys.argv:
    d = synth.generate(1000000)
    T2, X2, K2 = synth.default_params(d["lengths"])
    for opts in ({"profile": 1}, {"profile": 1, "batch": 96}, {"profile": 1, "batch": 384}):
        rc, g, st, wall = run(d, T2, X2, 0, K2, **opts)
        print("S1M", opts, rc, f"wall={wall:.3f}s", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
