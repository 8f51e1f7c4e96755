"""Offline differential campaigns between the two independent restatements of the reference (test infrastructure; the
CPU suite runs a small fixed subset of this in tests/test_oracle.py and tests/test_clinkage.py):

  greedy    C oracle (oracle/hammock_oracle.c) vs oracle/pyref.py -- random sizes 1..220, lengths 4..30 (mixed), thresholds
            around the default, max shift 0..shortest-1 (sometimes too big), penalties 0 / -1 / -3 / +2, every bundled matrix
            plus an asymmetric one, K from 0 to n + 5; compares statuses (incl. the null-object step), assignment, ranks,
            result order and the phase counters
  clinkage  C oracle (oracle/clinkage_oracle.c) vs oracle/pyref_clinkage.py -- shuffled input order, sizes 1..160

usage: python scripts/oracle_campaign.py greedy|clinkage SEED CASES
round 2: greedy seeds 1, 2 (150 + 600 cases) and clinkage seeds 1, 2 (120 + 600 cases): 0 mismatches
(profiles/r02_oracle_campaign.log)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hammock_b200 import synth                                  # noqa: E402
from oracle import oracle as O, pyref_clinkage as PC            # noqa: E402
from oracle.pyref import PyRef                                  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "matrices.npz"))
mats = {k: z[k] for k in z.files}
asym = mats["blosum62"].copy()
asym[np.triu_indices(24, 1)] -= 2


def greedy(seed, cases):
    rng = np.random.default_rng(seed)
    bad = 0; t0 = time.time(); stat = {}
    for it in range(cases):
        n = int(rng.integers(1, 220))
        lo = int(rng.integers(4, 20)); hi = int(min(30, lo + rng.integers(0, 12)))
        d = synth.generate(n, lo, hi, seed=int(rng.integers(0, 1 << 30)))
        T0, X0, K0 = synth.default_params(d["lengths"])
        T = int(T0 + rng.integers(-8, 9)); X = int(rng.integers(0, lo + (1 if rng.random() < 0.05 else 0))); P = int(rng.choice([0, 0, -1, -3, 2]))
        K = int(rng.choice([0, 1, 2, K0, max(1, n // 3), n, n + 5]))
        m = names[int(rng.integers(0, len(names)))]
        R = O.greedy_cluster(d["residues"], d["offsets"], d["abundance"], mats[m], T, X, P, K)
        seqs = [d["residues"][d["offsets"][i]:d["offsets"][i + 1]] for i in range(n)]
        try:
            Pr = PyRef(seqs, d["abundance"], mats[m], T, X, P, K).run()
            pst = Pr["status"]
        except Exception as e:
            Pr = None; pst = "exc:" + type(e).__name__ + ":" + str(e)[:60]
        stat[(R.status, str(pst))] = stat.get((R.status, str(pst)), 0) + 1
        ok = True
        if Pr is None: ok = R.status != 0
        elif R.status != Pr["status"]: ok = False
        elif R.status == 0:
            ok = (R.cluster_id == Pr["cluster_id"]).all() and (R.member_rank == Pr["member_rank"]).all() and (R.result_order == Pr["result_order"]).all() \
                 and R.counters["p1_steps"] == Pr["p1_steps"] and R.counters["p2_assigned"] == Pr["p2_assigned"]
        elif R.status == 2: ok = R.counters["npe_step"] == Pr["npe_step"]
        if not ok:
            bad += 1
            print("MISMATCH", it, n, lo, hi, T, X, P, K, m, R.status, pst, flush=True)
    print("cases done, mismatches:", bad, "statuses:", stat, "sec", round(time.time() - t0, 1))


def clinkage(seed, cases):
    rng = np.random.default_rng(seed)
    bad = 0; t0 = time.time(); stat = {}
    for it in range(cases):
        n = int(rng.integers(1, 160))
        lo = int(rng.integers(4, 16)); hi = int(min(30, lo + rng.integers(0, 8)))
        d = synth.generate(n, lo, hi, seed=int(rng.integers(0, 1 << 30)))
        perm = rng.permutation(n)
        res, offs = d["residues"], d["offsets"]
        seqs = [res[offs[i]:offs[i + 1]] for i in perm]
        o = np.zeros(n + 1, np.int32); o[1:] = np.cumsum([len(s) for s in seqs])
        ab = np.ascontiguousarray(d["abundance"][perm])
        T0, X0, _ = synth.default_params(d["lengths"])
        T = int(T0 + rng.integers(-10, 9)); X = int(rng.integers(0, lo)); P = int(rng.choice([0, 0, -1, -3]))
        m = names[int(rng.integers(0, len(names)))]
        R = O.clinkage_cluster(np.concatenate(seqs), o, ab, mats[m], T, X, P)
        try:
            c, r, od = PC.clinkage_cluster(seqs, ab, mats[m], T, X, P)
            ok = R.status == 0 and (c == R.cluster_id).all() and (r == R.member_rank).all() and (od == R.result_order).all()
            pst = 0
        except AssertionError as e:
            pst = "treeify"; ok = R.status == O.ERR_TREEIFIED if hasattr(O, "ERR_TREEIFIED") else R.status != 0
        stat[(R.status, pst)] = stat.get((R.status, pst), 0) + 1
        if not ok:
            bad += 1; print("MISMATCH", it, n, lo, hi, T, X, P, m, R.status, pst, flush=True)
    print("clinkage cases done, mismatches:", bad, stat, "sec", round(time.time() - t0, 1))


if __name__ == "__main__":
    mode, seed, cases = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    names = sorted(mats)
    if mode == "greedy":
        mats["asym"] = asym
        names = sorted(mats)
        greedy(seed, cases)
    else:
        clinkage(seed, cases)
