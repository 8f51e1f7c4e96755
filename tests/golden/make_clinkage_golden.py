"""Golden result of the EXACT complete-linkage clusterer on examples/MUSI (the default initial stage for that fixture:
2 457 <= 10 000 sequences, Hammock.java:371-373), produced by the C oracle (oracle/clinkage_oracle.c) and cross-checked
here against the second restatement (oracle/pyref_clinkage.py).  PARITY UNPINNED: no JVM, no reference outputs.
Input order = first-occurrence order of examples/MUSI/musi.fa, as Hammock.runClinkageClustering sees it (no sorting).

    python tests/golden/make_clinkage_golden.py      (needs /root/reference for the fasta; the .npz travels)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O, pyref_clinkage as PC      # noqa: E402


def main():
    strs, res, offs, ab = O.load_fasta("/root/reference/examples/MUSI/musi.fa")
    M = O.load_matrix("/root/reference/matrices/blosum62.txt")
    T, X, _ = O.default_params(offs)          # setClinkageThreshold == setGreedyThreshold: round(1.7 * mean length)
    R = O.clinkage_cluster(res, offs, ab, M, T, X, 0)
    assert R.status == 0
    c, r, o = PC.clinkage_cluster([res[offs[i]:offs[i + 1]] for i in range(len(ab))], ab, M, T, X, 0)
    assert (c == R.cluster_id).all() and (r == R.member_rank).all() and (o == R.result_order).all(), "the two restatements differ"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "musi_clinkage.npz"), residues=res, offsets=offs, abundance=ab,
                        params=np.array([T, X, 0], np.int32), cluster_id=R.cluster_id, member_rank=R.member_rank,
                        result_order=R.result_order, nearest_searches=np.int64(R.nearest_searches))
    sizes = np.bincount(R.cluster_id)
    print("n", len(ab), "T", T, "X", X, "clusters", len(R.result_order), "multi-member", int((sizes > 1).sum()), "largest", int(sizes.max()),
          "searches", R.nearest_searches, "first ids", R.result_order[:8].tolist())


if __name__ == "__main__":
    main()
