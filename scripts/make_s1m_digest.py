"""Runs the FULL CPU oracle on a synthetic workload (default: the S1M headline config) and writes the
golden digest tests/golden/<name>_digest.json.  CPU only (OpenMP, all cores); ~10-20 min for S1M on
8 cores.  The digest is sha256 over cluster_id || member_rank || result_order (little-endian int32) --
the same bytes hammock_b200.host.result_digest() hashes for a GPU result.

usage: python scripts/make_s1m_digest.py [n=1000000] [min_len=12] [max_len=12] [name=s1m]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hammock_b200 import synth          # noqa: E402  (the workload generator only)
from oracle import oracle as O          # noqa: E402


def digest(cluster_id, member_rank, result_order):
    h = hashlib.sha256()
    for a in (cluster_id, member_rank, result_order):
        h.update(np.ascontiguousarray(a, dtype="<i4").tobytes())
    return h.hexdigest()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    name = sys.argv[4] if len(sys.argv) > 4 else "s1m"
    d = synth.generate(n, lo, hi)
    T, X, K = synth.default_params(d["lengths"])
    M = synth.blosum62()
    t0 = time.time()
    R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], M, T, X, 0, K, nthreads=os.cpu_count() or 1)
    dt = time.time() - t0
    assert R.status == 0, R.status
    out = {
        "workload": f"synth.generate({n},{lo},{hi}) seed 20260101, BLOSUM62, T={T} X={X} P=0 K={K}",
        "producer": "oracle/hammock_oracle.c full run (parity unpinned: C restatement of the Java)",
        "n": n, "threshold": T, "max_shift": X, "shift_penalty": 0, "max_clusters": K,
        "sha256": digest(R.cluster_id, R.member_rank, R.result_order),
        "n_result": int(len(R.result_order)), "n_multi": int(R.n_multi),
        "counters": R.counters, "oracle_seconds": round(dt, 1), "oracle_threads": os.cpu_count(),
    }
    path = os.path.join(ROOT, "tests", "golden", f"{name}_digest.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(ROOT, "scratch", f"{name}_oracle_result.npz"), cluster_id=R.cluster_id,
                        member_rank=R.member_rank, result_order=R.result_order)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
