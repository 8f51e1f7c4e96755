"""ncu target: one clustering of a synthetic workload.  usage: gpu_prof_case.py n lo hi [opt=v ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hammock_b200 as hb
from hammock_b200 import synth
n, lo, hi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
opts = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[4:]}
d = synth.generate(n, lo, hi)
T, X, K = synth.default_params(d["lengths"])
ctx = hb.GreedyContext(0, **opts)
ctx.upload(d["residues"], d["offsets"], d["abundance"], synth.blosum62(), T, X, 0, K)
rc, _ = ctx.run_status()
print(n, lo, hi, opts, rc, ctx.stats()["total_ms"])
ctx.close()
