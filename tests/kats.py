"""Known-answer vectors for ShiftedScorer.scoreWithShift (reference ShiftedScorer.java:48-95).

SURVEY-DERIVED, NOT JVM-VERIFIED: the reference ships no tests; these were derived from the
Java source by hand / by two independent restatements.  (seq1, seq2, maxShift,
shiftPenalty, matrix, score, shift-or-None)
"""
SCORER_KATS = [
    ("WVTAPRSLPVLP", "RSPIVRQLPSLP", 3, 0, "blosum62", 16, 0),
    ("YYYKTRGLPAVP", "YSYKTRGLPAVP", 3, 0, "blosum62", 59, 0),   # hand check: 7-2+7+5+5+5+6+4+7+4+4+7
    ("GSWVVDISNVED", "GSWAVDISNVED", 3, 0, "blosum62", 60, 0),
    ("AKSRPLPMVGLV", "RALPVMPTNGPM", 3, 0, "blosum62", 11, 3),
    ("RSLPVLP", "WVTAPRSLPVLP", 2, 0, "blosum62", 35, -5),
    ("RSLPVLP", "WVTAPRSLPVLP", 2, -1, "blosum62", 30, -5),
    ("WVTAPRSLPVLP", "RSLPVLP", 2, -1, "blosum62", 30, 5),
    ("ACDEFGH", "CDEFGHI", 2, -2, "blosum62", 36, 1),
    ("ACDEFGH", "CDEFGHI", 0, 0, "blosum62", -12, 0),
    ("W" * 7, "W" * 30, 5, 0, "blosum62", 77, 0),
    ("AB*XZ", "ZX*BA", 4, 0, "blosum62", 4, -4),
    ("WVTAPRSLPVLP", "RSPIVRQLPSLP", 3, 0, "pam250", 27, None),
]

# (cells, shifts) pins from SURVEY.md 8c
CELL_KATS = [((12, 12, 3), 72, 7), ((7, 12, 2), 64, 10), ((7, 7, 2), 29, 5), ((7, 7, 0), 7, 1),
             ((7, 30, 5), 208, 34), ((5, 5, 4), 25, 9)]

# Appendix A micro-fixture: (sequence, abundance) in arbitrary input order; BLOSUM62, T=24, X=2, P=0
MICRO = [("WVTAPRSLPVLP", 9), ("WVTAPRSLPVLA", 9), ("WVTAPRSLPVLG", 3), ("AVTAPRSLPVLP", 3), ("GSWVVDISNVED", 5),
         ("GSWAVDISNVED", 2), ("GSWVVDISNVEE", 2), ("RSLPVLP", 4), ("TAPRSLPVL", 1), ("MMMMMMMCCCCC", 1),
         ("HHHHHHHWWWWW", 7), ("WVTAPRSLPKKK", 1), ("KKKAPRSLPVLP", 1), ("GSWVVDIKKKKK", 1)]
MICRO_ORDER = ["WVTAPRSLPVLP", "WVTAPRSLPVLA", "HHHHHHHWWWWW", "GSWVVDISNVED", "RSLPVLP", "WVTAPRSLPVLG",
               "AVTAPRSLPVLP", "GSWVVDISNVEE", "GSWAVDISNVED", "WVTAPRSLPKKK", "TAPRSLPVL", "MMMMMMMCCCCC",
               "KKKAPRSLPVLP", "GSWVVDIKKKKK"]
# K -> (clusters as member-id lists in creation order, remaining singletons in list order)
MICRO_EXPECT = {
    2: ([[0, 1, 4, 5, 6, 10, 12], [3, 7, 8, 13]], [2, 9, 11]),
    5: ([[0, 1, 5, 9], [3, 7, 8, 13], [4, 6], [10, 12]], [2, 11]),
}


def result_to_lists(cluster_id, member_rank, result_order, n_multi):
    import numpy as np
    clusters = []
    for c in result_order[:n_multi]:
        mem = np.nonzero(cluster_id == c)[0]
        clusters.append([int(m) for m in mem[np.argsort(member_rank[mem])]])
    return clusters, [int(x) for x in result_order[n_multi:]]
