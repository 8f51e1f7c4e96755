"""BASELINE configs[1] and configs[3]: S100k (lengths 7-12) and the matrix / length sweep.
Times the GPU path (warm, second run) and checks parity against the oracle where the oracle
finishes in seconds.  Output: one JSON line per case (kept under profiles/)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hammock_b200 as hb
from hammock_b200 import synth
from oracle import oracle as O

mats = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
NT = os.cpu_count() or 1
cases = [("S100k_len7-12", 100000, 7, 12, "blosum62", 0, True)]
for L in (7, 9, 12, 16, 20, 25, 30):
    cases.append((f"sweep_len{L}", 50000, L, L, "blosum62", 0, True))
cases.append(("sweep_len7-30", 50000, 7, 30, "blosum62", 0, True))
cases.append(("sweep_len13-18", 50000, 13, 18, "blosum62", 0, True))
for m in ("blosum30", "blosum45", "blosum80", "blosum100", "pam250", "mcla71"):
    cases.append((f"sweep_{m}", 50000, 12, 12, m, 0, True))
cases.append(("sweep_blosum62_P-1", 50000, 12, 12, "blosum62", -1, True))
quick = "--quick" in sys.argv
for name, n, lo, hi, m, P, check in cases:
    d = synth.generate(n, lo, hi, seed=4242 + lo * 31 + hi)
    T, X, K = synth.default_params(d["lengths"])
    ctx = hb.GreedyContext(0, profile=1)
    ctx.upload(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K)
    rc, _ = ctx.run_status()
    rc, _ = ctx.run_status()
    st = ctx.stats()
    g = ctx.download() if rc == 0 else None
    ctx.close()
    line = {"case": name, "n": n, "len": [lo, hi], "matrix": m, "P": P, "T": T, "X": X, "K": K, "rc": rc,
            "gpu_ms": round(st["total_ms"], 2), "seq_per_s": round(n / (st["total_ms"] * 1e-3)), "fast_path": st["fast_path"],
            "lane_bits": st["lane_bits"], "bulk_pairs": st["bulk_pairs"], "bulk_kernel_ms": round(st["bulk_kernel_ms"], 2)}
    if check and not quick and rc == 0:
        t = time.time()
        R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K, nthreads=NT)
        line["oracle_s"] = round(time.time() - t, 2)
        line["parity"] = bool(R.status == 0 and (R.cluster_id == g.cluster_id).all() and (R.member_rank == g.member_rank).all()
                              and (R.result_order == g.result_order).all())
    print(json.dumps(line), flush=True)
