"""Times the 1 M workload under several tuning variants (one context, profile=1) and checks every result against the
golden digest.  usage: python scripts/gpu_variants.py [n] ['{"opt": v, ...}' ...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hammock_b200 as hb                      # noqa: E402
from hammock_b200 import synth                 # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    variants = [json.loads(a) for a in sys.argv[2:]] or [{}]
    d = synth.generate(n, 12, 12)
    T, X, K = synth.default_params(d["lengths"])
    M = synth.blosum62()
    gold = None
    try:
        gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "s1m_digest.json")))["sha256"] if n == 1000000 else None
    except Exception:
        pass
    first = None
    for v in variants:
        ctx = hb.GreedyContext(0, profile=1, **v)
        ctx.upload(d["residues"], d["offsets"], d["abundance"], M, T, X, 0, K)
        best = None
        for rep in range(4):
            ctx.run()
            st = ctx.stats()
            if rep and (best is None or st["total_ms"] < best[0]["total_ms"]):
                best = (st, ctx.section_ms())
        G = ctx.download()
        dg = hb.result_digest(G.cluster_id, G.member_rank, G.result_order)
        first = first or dg
        st, sec = best
        print(json.dumps({"opts": v, "total_ms": round(st["total_ms"], 2), "p1_ms": round(st["phase1_ms"], 2), "p2_ms": round(st["phase2_ms"], 2),
                          "bulk_ms": round(st["bulk_kernel_ms"], 2), "batches": st["p1_batches"], "restarts": st["p1_restarts"],
                          "p2_rounds": st["p2_rounds"], "scalar_pairs": st["scalar_pairs"], "launches": st["total_launches"],
                          "flags": st["flags"], "digest_ok": (dg == gold) if gold else None, "same_as_first": dg == first,
                          "sections": {k: round(x, 2) for k, x in sec.items() if x}}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
