"""usage: python scripts/gpu_case2.py name:n:lo:hi:matrix '{"opt": v}' ...  -- one workload, several option sets"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
mats = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
name, n, lo, hi, m = sys.argv[1].split(":")
n, lo, hi = int(n), int(lo), int(hi)
d = synth.generate(n, lo, hi)
T, X, K = synth.default_params(d["lengths"])
first = None
for o in sys.argv[2:] or ["{}"]:
    opts = json.loads(o)
    ctx = hb.GreedyContext(0, profile=1, **opts)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, 0, K)
    best = None
    for rep in range(3):
        ctx.run()
        st = ctx.stats()
        if best is None or st["total_ms"] < best[0]["total_ms"]:
            best = (st, ctx.section_ms())
    g = ctx.download()
    dg = hb.result_digest(g.cluster_id, g.member_rank, g.result_order)
    first = first or dg
    st, sec = best
    print(name, opts, "ms", round(st["total_ms"], 2), "bulk_ms", round(st["bulk_kernel_ms"], 2), "launches", st["total_launches"], "same", dg == first,
          {k: round(v, 1) for k, v in sec.items() if v >= 0.3}, flush=True)
    ctx.close()
