"""CPU study of the design space around the upper-bound filter of hmk_bulk_filter ("run the cheap
experiment the numpy filter model supports ... and build it only if the shared-memory traffic per pair drops").

The kernel is bound by shared-memory WAVEFRONTS (ncu: 73 % of the wavefront peak), not by bytes: a conflict-free
warp-wide LDS costs one wavefront whatever its width up to 32 bits, so a cheaper filter needs FEWER or conflict-free
look-ups per pair -- and must not pass many more pairs, because a passed pair costs V wavefronts in the verify stage
(per-lane random profile rows: V is taken from the measured 15.7 wavefronts per pair = 12 + pass rate x V).

Designs (all are upper bounds of every diagonal's score, so none has false negatives; numpy checks that):
  current      12 look-ups, full alphabet (24-word rows: conflict free), a byte bounds 2 neighbouring diagonals
  groups of 4  12 look-ups, a byte bounds 4 neighbouring diagonals (2 groups)            -- looser, same look-ups
  one group    12 look-ups, one bound for all 7 diagonals                                  -- loosest
  pairs/full   6 look-ups indexed by TWO adjacent residues (576-word rows: bank conflicts), tighter bound
  pairs/C      6 look-ups indexed by two adjacent residue CLASSES (C x C-word rows), class maxima: looser
  hybrid       positions 0-3 as two PAIRS (the database is walked in clustering order, whose neighbours share their first
               residues: those two look-ups are nearly conflict free), positions 4-11 as now: 10 look-ups
  exact        the 7 diagonals themselves = the true hit rate

usage: python scripts/filter_design_study.py [queries=300] [db=20000]   -> profiles/r02_filter_design_study.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hammock_b200 import synth          # noqa: E402

L, X, T = 12, 3, 20
MEASURED_WAVEFRONTS_PER_PAIR = 15.7    # profiles/r01_ncu_bulk_filter_full.json (12 ideal)


def residue_classes(M, C):
    """agglomerative grouping of the 24 residues by the similarity of their matrix rows (for the class-maximum bounds)"""
    groups = [[r] for r in range(24)]
    while len(groups) > C:
        best = None
        for a in range(len(groups)):
            for b in range(a + 1, len(groups)):
                rows = M[groups[a] + groups[b]]
                cost = float((rows.max(axis=0) - rows.min(axis=0)).sum())      # slack the class maximum introduces
                if best is None or cost < best[0]:
                    best = (cost, a, b)
        _, a, b = best
        groups[a] += groups[b]
        del groups[b]
    cls = np.zeros(24, np.int64)
    for c, g in enumerate(groups):
        cls[g] = c
    return cls, groups


def diag_cells(q, M):
    """cells[k, j, r] = M[r][q[j + k]] where diagonal k has a cell at database position j, else 0 (as the kernel counts
    it); valid[k, j]"""
    cells = np.zeros((2 * X + 1, L, 24), np.int64)
    valid = np.zeros((2 * X + 1, L), bool)
    for lam in range(2 * X + 1):
        k = lam - X
        for j in range(L):
            if 0 <= j + k < L:
                cells[lam, j] = M[:, q[j + k]]
                valid[lam, j] = True
    return cells, valid


def warp_wavefronts(idx, words_per_row):
    """average wavefronts of one warp-wide LDS.32 into a row of `words_per_row` words, lanes = 32 consecutive items:
    per bank the number of DISTINCT words requested, maximum over the banks"""
    n = len(idx) // 32 * 32
    w = idx[:n].reshape(-1, 32)
    out = np.zeros(len(w))
    for i, lanes in enumerate(w):
        words = np.unique(lanes)
        out[i] = np.bincount(words % 32, minlength=32).max()
    return float(out.mean())


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    ndb = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    M = synth.blosum62().astype(np.int64)
    d = synth.generate(200000, L, L)
    R = d["residues"].reshape(-1, L).astype(np.int64)
    rng = np.random.default_rng(5)
    qs = R[rng.choice(len(R), nq, replace=False)]
    D = R[rng.choice(len(R), ndb, replace=False)]
    jj = np.arange(L)
    designs = ["exact", "current (2 diagonals per byte)", "groups of 4 diagonals", "one group of 7", "pairs / full alphabet",
               "hybrid (pairs for positions 0-3)"]
    class_sets = {C: residue_classes(M, C) for C in (6, 8, 12)}
    designs += [f"pairs / {C} classes" for C in class_sets] + [f"single / {C} classes" for C in class_sets]
    passed = {k: 0 for k in designs}
    false_neg = {k: 0 for k in designs}
    groups2 = [(0, 1), (2, 3), (4, 5), (6,)]
    groups4 = [(0, 1, 2, 3), (4, 5, 6)]
    for q in qs:
        cells, valid = diag_cells(q, M)
        per = cells[:, jj[None, :], D]                           # [diag, item, pos]
        exact = per.sum(axis=2)                                  # [diag, item]
        hit = (exact >= T).any(axis=0)
        passed["exact"] += int(hit.sum())

        def single(groups, table):
            ok = np.zeros(ndb, bool)
            for g in groups:
                ok |= table[list(g)].max(axis=0).sum(axis=1) >= T
            return ok

        def paired(groups, table):
            ok = np.zeros(ndb, bool)
            pairsum = table[:, :, 0::2] + table[:, :, 1::2]      # adjacent positions (0,1), (2,3), ...
            for g in groups:
                ok |= pairsum[list(g)].max(axis=0).sum(axis=1) >= T
            return ok
        def hybrid(groups, table):
            ok = np.zeros(ndb, bool)
            head = table[:, :, 0:4:2] + table[:, :, 1:4:2]       # (0,1), (2,3)
            for g in groups:
                ok |= head[list(g)].max(axis=0).sum(axis=1) + table[list(g)][:, :, 4:].max(axis=0).sum(axis=1) >= T
            return ok
        res = {"hybrid (pairs for positions 0-3)": hybrid(groups2, per), "current (2 diagonals per byte)": single(groups2, per), "groups of 4 diagonals": single(groups4, per),
               "one group of 7": single([tuple(range(7))], per), "pairs / full alphabet": paired(groups2, per)}
        for C, (cls, groups) in class_sets.items():
            cmax = np.zeros((2 * X + 1, L, 24), np.int64)        # class maximum of every cell
            for g in groups:
                cmax[:, :, g] = cells[:, :, g].max(axis=2, keepdims=True)
            pc = cmax[:, jj[None, :], D]
            res[f"pairs / {C} classes"] = paired(groups2, pc)
            res[f"single / {C} classes"] = single(groups2, pc)
        for k, ok in res.items():
            passed[k] += int(ok.sum())
            false_neg[k] += int((hit & ~ok).sum())
    total = nq * ndb
    rate = {k: passed[k] / total for k in designs}
    # look-ups and bank conflicts: lanes = 32 consecutive database items of the synthetic set in clustering order (what
    # a warp of the kernel holds), averaged over all positions / position pairs -- the sorted order makes the first
    # positions of neighbouring items agree, the later ones are as good as random
    warps = R[len(R) // 2:len(R) // 2 + 32 * 2000]
    mean = lambda f, n: float(np.mean([f(p) for p in range(n)]))
    conflict = {"24-word rows (one residue)": mean(lambda p: warp_wavefronts(warps[:, p], 24), L),
                "576-word rows (two residues)": mean(lambda p: warp_wavefronts(warps[:, 2 * p] * 24 + warps[:, 2 * p + 1], 576), L // 2)}
    for C, (cls, _) in class_sets.items():
        conflict[f"{C * C}-word rows (two classes of {C})"] = mean(
            lambda p: warp_wavefronts(cls[warps[:, 2 * p]] * C + cls[warps[:, 2 * p + 1]], C * C), L // 2)
        conflict[f"{C}-word rows (one class of {C})"] = mean(lambda p: warp_wavefronts(cls[warps[:, p]], C), L)
    head_pairs = float(np.mean([warp_wavefronts(warps[:, 2 * p] * 24 + warps[:, 2 * p + 1], 576) for p in range(2)]))
    conflict["576-word rows, positions 0-3 only"] = head_pairs
    p0 = rate["current (2 diagonals per byte)"]
    V = (MEASURED_WAVEFRONTS_PER_PAIR - 12.0) / p0               # verify-stage wavefronts per passed pair, from the measurement
    lookups = {"current (2 diagonals per byte)": (12, conflict["24-word rows (one residue)"]),
               "groups of 4 diagonals": (12, conflict["24-word rows (one residue)"]),
               "one group of 7": (12, conflict["24-word rows (one residue)"]),
               "pairs / full alphabet": (6, conflict["576-word rows (two residues)"])}
    lookups["hybrid (pairs for positions 0-3)"] = (10, (2 * head_pairs + 8 * conflict["24-word rows (one residue)"]) / 10)
    for C in class_sets:
        lookups[f"pairs / {C} classes"] = (6, conflict[f"{C * C}-word rows (two classes of {C})"])
        lookups[f"single / {C} classes"] = (12, conflict[f"{C}-word rows (one class of {C})"])
    table = []
    for k in designs:
        if k == "exact":
            continue
        n, cf = lookups[k]
        table.append({"design": k, "pass_rate": round(rate[k], 5), "false_negatives": false_neg[k], "lookups_per_pair": n,
                      "wavefronts_per_lookup": round(cf, 3), "filter_wavefronts": round(n * cf, 2),
                      "modelled_wavefronts_per_pair": round(n * cf + rate[k] * V, 2)})
    out = {"workload": f"{nq} x {ndb} pairs of the synthetic 12-mer set, BLOSUM62, T={T}, X={X}",
           "true_hit_rate": round(rate["exact"], 6), "verify_wavefronts_per_passed_pair_from_ncu": round(V, 1),
           "measured_wavefronts_per_pair_current": MEASURED_WAVEFRONTS_PER_PAIR,
           "bank_conflict_model": {k: round(v, 3) for k, v in conflict.items()}, "designs": table,
           "if_the_verify_stage_cost_half": round(12.0 + p0 * V / 2, 2),
           "classes": {str(C): [[synth.ALPHABET[r] for r in g] for g in groups] for C, (_, groups) in class_sets.items()}}
    path = os.path.join(ROOT, "profiles", "r02_filter_design_study.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
