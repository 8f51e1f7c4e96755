"""Re-runs one case of tests/test_gpu_fuzz.py::test_fuzz_against_oracle several times and reports where the GPU result
differs from the oracle.  usage: python scripts/gpu_fuzz_one.py ITER [REPEATS] ['{"opt": v}']"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hammock_b200 as hb                      # noqa: E402
from oracle import oracle as O                 # noqa: E402
from tests.test_gpu_fuzz import _case          # noqa: E402


def main():
    target = int(sys.argv[1])
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    override = json.loads(sys.argv[3]) if len(sys.argv) > 3 else None
    z = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
    mats = {k: z[k] for k in z.files}
    names = sorted(mats)
    asym = mats["blosum62"].copy()
    asym[np.triu_indices(24, 1)] -= 1
    mats["asym"] = asym
    names.append("asym")
    mats["big"] = mats["blosum62"] * 700
    names.append("big")
    rng = np.random.default_rng(20260101)
    for it in range(target + 1):
        d, M, T, X, P, K, opts, tag = _case(rng, mats, names)
        if tag[3] == "big":
            T *= 700
    if override is not None:
        opts = override
    R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K, nthreads=1)
    print("case", tag, "T", T, "X", X, "P", P, "K", K, "opts", opts, "oracle status", R.status, R.counters)
    for rep in range(reps):
        ctx = hb.GreedyContext(0, **opts)
        ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, P, K)
        rc, _ = ctx.run_status()
        st = ctx.stats()
        if rc == 0:
            G = ctx.download()
            bad = np.nonzero(G.cluster_id != R.cluster_id)[0]
            badr = np.nonzero(G.member_rank != R.member_rank)[0]
            print(f"rep {rep}: rc {rc} p1 steps {st['p1_steps']} (oracle {R.counters['p1_steps']}) joins {st['p1_joins']} p2_assigned {st['p2_assigned']} "
                  f"(oracle {R.counters['p2_assigned']}) rounds {st['p2_rounds']} cluster_id diffs {len(bad)} rank diffs {len(badr)}")
            for i in bad[:8]:
                print("   id", i, "gpu", G.cluster_id[i], G.member_rank[i], "oracle", R.cluster_id[i], R.member_rank[i])
        else:
            print(f"rep {rep}: rc {rc} oracle {R.status}")
        ctx.close()


if __name__ == "__main__":
    main()
