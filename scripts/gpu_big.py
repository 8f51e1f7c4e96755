import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
M = synth.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
variants = []
for a in sys.argv[2:]:
    variants.append({kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.split(",") if kv})
if not variants:
    variants = [{"profile": 1}]
t = time.time(); d = synth.generate(n); print("gen", time.time() - t, flush=True)
T, X, K = synth.default_params(d["lengths"])
for opts in variants:
    ctx = hb.GreedyContext(0, **opts)
    t = time.time(); ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, 0, K); up = time.time() - t
    for rep in range(3):
        t = time.time(); rc, msg = ctx.run_status(); wall = time.time() - t
        st = ctx.stats()
        sec = ctx.section_ms()
        gc = st["bulk_cells"] / (st["bulk_kernel_ms"] * 1e-3) / 1e9 if st["bulk_kernel_ms"] else 0
        print(opts, "rep", rep, "rc", rc, f"upload={up*1e3:.1f}ms run={wall*1e3:.1f}ms bulkGCUPS={gc:.0f} bulk_ms={st['bulk_kernel_ms']:.1f} "
              f"p1={st['phase1_ms']:.1f} p2={st['phase2_ms']:.1f} batches={st['p1_batches']} iters={st['p2_rounds']} launches={st['total_launches']}", flush=True)
        print("   sections", {k: round(v, 1) for k, v in sec.items()}, flush=True)
    t = time.time(); g = ctx.download(); dn = time.time() - t
    print("  download", dn, "checksum", int(g.cluster_id.astype(np.int64).sum()), int(g.member_rank.astype(np.int64).sum()), g.n_multi, flush=True)
    ctx.close()
