import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
mats = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
for spec in sys.argv[1:]:
    name, n, lo, hi, m = spec.split(":")
    n, lo, hi = int(n), int(lo), int(hi)
    d = synth.generate(n, lo, hi, seed=4242 + lo * 31 + hi)
    T, X, K = synth.default_params(d["lengths"])
    ctx = hb.GreedyContext(0, profile=1)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, 0, K)
    for rep in range(2):
        rc, _ = ctx.run_status()
    st = ctx.stats(); sec = ctx.section_ms()
    g = ctx.download()
    print(name, "rc", rc, "ms", round(st["total_ms"], 1), {k: st[k] for k in ("p1_batches", "p1_restarts", "p2_rounds", "p2_hits", "p2_candidates", "p2_assigned", "scalar_pairs", "bulk_pairs", "fast_path")},
          {k: round(v, 1) for k, v in sec.items() if v >= 0.5}, "chk", int(g.cluster_id.astype(np.int64).sum()), flush=True)
    ctx.close()
