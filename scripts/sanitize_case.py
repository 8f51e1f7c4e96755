"""small cases for compute-sanitizer (memcheck): every kernel family once"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
mats = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
for name, n, lo, hi, m, opts in [("uniform12", 3000, 12, 12, "blosum62", {}), ("mixed7-12", 2500, 7, 12, "blosum62", {"batch": 48}),
                                 ("long20", 1500, 20, 20, "blosum62", {}), ("generic7-30", 1200, 7, 30, "blosum62", {}),
                                 ("s16lanes", 2000, 12, 12, "blosum30", {"batch": 64, "kb": 2})]:
    d = synth.generate(n, lo, hi, seed=77 + n)
    T, X, K = synth.default_params(d["lengths"])
    ctx = hb.GreedyContext(0, **opts)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, 0, K)
    rc, msg = ctx.run_status()
    g = ctx.download()
    sc = ctx.score_block(np.arange(0, 50, dtype=np.int32), np.arange(10, 40, dtype=np.int32))
    print(name, rc, g.n_multi, int(sc.sum()), ctx.stats()["fast_path"], flush=True)
    ctx.close()
