"""Extra randomised mid-size differential runs (same generator family as tests/test_gpu_fuzz.py's mid-size test,
other seeds):  python scripts/gpu_fuzz_mid.py SEED NCASES   -- set HMK_DEBUG_HANG=1 to exit instead of hanging."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O
z = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz")); mats = {k: z[k] for k in z.files}
seed, ncases = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
bad = 0
for it in range(ncases):
    n = int(rng.integers(2000, 16000))
    lo, hi = [(12, 12), (12, 12), (9, 9), (7, 12), (16, 16), (10, 10)][int(rng.integers(0, 6))]
    m = str(rng.choice(["blosum62", "blosum62", "blosum45", "pam250", "blosum80"]))
    d = synth.generate(n, lo, hi, seed=int(rng.integers(1, 1 << 30)), top_abundance=int(rng.choice([50, 100000])))
    if rng.random() < 0.2:
        d["abundance"] = np.ascontiguousarray(rng.permutation(d["abundance"]))
    T0, X0, K0 = synth.default_params(d["lengths"])
    T = int(T0 + rng.choice([0, 0, -4, -6, 5]))
    K = int(rng.choice([K0, K0, n // 10, n // 3, n]))
    P = int(rng.choice([0, 0, -1]))
    opts = {}
    if rng.random() < 0.4:
        opts = {"kb": int(rng.choice([2, 4, 8])), "batch": int(rng.choice([0, 96, 200, 448]))}
    R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K, nthreads=os.cpu_count())
    ctx = hb.GreedyContext(0, **opts)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X0, P, K)
    for rep in range(2):
        t = time.time(); rc, msg = ctx.run_status(); dt = time.time() - t
        st = ctx.stats()
        ok = rc == R.status
        if rc == 0:
            G = ctx.download()
            ok = ok and (G.cluster_id == R.cluster_id).all() and (G.member_rank == R.member_rank).all() and (G.result_order == R.result_order).all()
        if not ok:
            bad += 1
        print(f"[{'OK' if ok else 'FAIL'}] seed {seed} case {it} rep {rep} n={n} len={lo}-{hi} {m} T={T} P={P} K={K} opts={opts}: rc={rc} {msg} "
              f"{dt*1e3:.0f} ms batches={st['p1_batches']} restarts={st['p1_restarts']} path={st['fast_path']} lane={st['lane_bits']}", flush=True)
    ctx.close()
print("done, failures:", bad)
