"""Generates the committed fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; the GPU box never reads /root/reference).

What the fixtures are -- and are not.  The reference (Java) cannot be executed here (no JVM)
and ships no tests or expected outputs, so NOTHING below is reference OUTPUT.  The files hold
 * reference INPUT data re-encoded (examples/MUSI, examples/antibodies, matrices/*.txt), and
 * results of the CPU oracle (oracle/hammock_oracle.c), cross-checked against the independent
   numpy restatement (oracle/pyref.py) and against the hand-derived pins of SURVEY.md 8c /
   Appendix A.  They pin the ORACLE (regressions), not the reference: "parity unpinned".
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle.pyref import PyRef  # noqa: E402

REF = "/root/reference"


def dataset(fasta, name, check_pyref):
    strs, res, offs, ab = O.load_fasta(fasta)
    perm = O.sort_order_size(res, offs, ab)
    ordered = [strs[i] for i in perm]
    r2, o2 = O.pack(ordered)
    a2 = ab[perm]
    M = O.load_matrix(f"{REF}/matrices/blosum62.txt")
    T, X, K = O.default_params(o2)
    R = O.greedy_cluster(r2, o2, a2, M, T, X, 0, K, nthreads=8)
    assert R.status == 0
    if check_pyref:
        P = PyRef([r2[o2[i]:o2[i + 1]] for i in range(len(ordered))], a2, M, T, X, 0, K).run()
        assert (P["cluster_id"] == R.cluster_id).all() and (P["member_rank"] == R.member_rank).all()
        assert (P["result_order"] == R.result_order).all()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), residues=r2, offsets=o2, abundance=a2,
                        input_perm=perm.astype(np.int32), cluster_id=R.cluster_id, member_rank=R.member_rank,
                        result_order=R.result_order, params=np.array([T, X, 0, K], dtype=np.int32),
                        n_multi=np.int32(R.n_multi))
    return {"n": len(ordered), "T": T, "X": X, "K": K, "counters": R.counters, "first": ordered[:4]}


def main():
    meta = {}
    mats = {}
    for fn in sorted(os.listdir(f"{REF}/matrices")):
        if not fn.endswith(".txt"):
            continue
        try:
            mats[fn[:-4]] = O.load_matrix(f"{REF}/matrices/{fn}")
        except O.OracleError as e:
            meta.setdefault("matrix_rejected", {})[fn] = e.status
    np.savez_compressed(os.path.join(HERE, "matrices.npz"), **mats)
    meta["matrices"] = {k: {"min": int(v.min()), "max": int(v.max()), "symmetric": bool((v == v.T).all())}
                        for k, v in mats.items()}
    meta["musi"] = dataset(f"{REF}/examples/MUSI/musi.fa", "musi", True)
    meta["antibodies"] = dataset(f"{REF}/examples/antibodies/antibodies.fa", "antibodies", False)
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1)[:3000])


if __name__ == "__main__":
    main()
