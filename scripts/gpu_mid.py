"""One case of tests/test_gpu_fuzz.py::test_fuzz_midsize_default_options in its own process (debugging aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O
z = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz")); mats = {k: z[k] for k in z.files}
want = int(sys.argv[1])
extra = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[2:]}
rng = np.random.default_rng(77)
for it in range(10):
    n = int(rng.integers(3000, 20000))
    kind = int(rng.integers(0, 4))
    lo, hi = [(12, 12), (9, 9), (7, 12), (16, 16)][kind]
    m = str(rng.choice(["blosum62", "blosum62", "blosum45", "pam250"]))
    d = synth.generate(n, lo, hi, seed=int(rng.integers(1, 1 << 30)), top_abundance=int(rng.choice([50, 100000])))
    T0, X0, K0 = synth.default_params(d["lengths"])
    T = int(T0 + rng.choice([0, 0, -4, 5]))
    K = int(rng.choice([K0, K0, n // 10, n // 3, n]))
    P = int(rng.choice([0, 0, -1]))
    opts = {} if it % 3 else {"kb": int(rng.choice([2, 4]))}
    if it != want:
        continue
    opts.update(extra)
    R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K, nthreads=os.cpu_count())
    ctx = hb.GreedyContext(0, **opts)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K)
    for rep in range(2):
        t = time.time(); rc, msg = ctx.run_status(); dt = time.time() - t
        st = ctx.stats()
        ok = rc == R.status
        if rc == 0:
            G = ctx.download()
            ok = ok and (G.cluster_id == R.cluster_id).all() and (G.member_rank == R.member_rank).all()
        print(f"case {it} rep {rep} n={n} len={lo}-{hi} {m} T={T} K={K} opts={opts}: rc={rc} {msg} ok={ok} {dt*1e3:.0f} ms "
              f"batches={st['p1_batches']} restarts={st['p1_restarts']} path={st['fast_path']} lane={st['lane_bits']}", flush=True)
    ctx.close()
