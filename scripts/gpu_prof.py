import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
M = synth.blosum62()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
opts = {}
for a in sys.argv[2:]:
    k, v = a.split("="); opts[k] = int(v)
d = synth.generate(n)
T, X, K = synth.default_params(d["lengths"])
ctx = hb.GreedyContext(0, **opts)
ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, 0, K)
t = time.time(); rc, msg = ctx.run_status(); wall = time.time() - t
st = ctx.stats()
print(n, opts, rc, f"run={wall*1e3:.1f}ms", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
ctx.close()
