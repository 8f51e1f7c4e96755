"""Debug helper: one synthetic case vs the oracle, printing counters and the first mismatches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O
mats = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
n, lo, hi, m, P, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[7:]}
d = synth.generate(n, lo, hi, seed=1000 + n + hi)
T, X, K0 = synth.default_params(d["lengths"])
K = K0 if K < 0 else K
R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K, nthreads=os.cpu_count())
ctx = hb.GreedyContext(0, **opts)
ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K)
rc, msg = ctx.run_status()
st = ctx.stats()
print("rc", rc, msg, R.status, {k: st[k] for k in ("p1_steps", "p1_joins", "p1_new_clusters", "p1_orphans", "p1_batches", "p1_restarts", "p2_assigned")})
print("oracle", R.counters)
if rc == 0:
    g = ctx.download()
    bad = np.nonzero(g.cluster_id != R.cluster_id)[0]
    print("nbad", len(bad), bad[:20].tolist(), g.cluster_id[bad[:20]].tolist(), R.cluster_id[bad[:20]].tolist())
    bad = np.nonzero(g.member_rank != R.member_rank)[0]
    print("rank bad", len(bad), bad[:20].tolist())
