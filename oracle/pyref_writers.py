"""CPU ORACLE (test infrastructure only) for the greedy stage's result files -- SURVEY.md 8(f) N2.

A statement-by-statement Python restatement of the reference's writers (paths relative to
/root/reference/src/cz/krejciadam/hammock/), used by tests/ to check the host-side writers of the product
(hammock_b200/host.py and host_cpp/hammock_host.hpp) byte for byte.  PARITY UNPINNED like the rest of oracle/: the
reference ships no expected output files and cannot be run here (no JVM).

Data model of this module (deliberately NOT the product's classes):
    sequence = (string, {label: count})          UniqueSequence; size() = sum of the counts, Java int
    cluster  = (id, [sequence, ...])             Cluster; members in getSequences() order
    alignments = {cluster id: [msa line, ...]}   what FileIOManager.getAlignmentsMap returns: the lines of the
                 cluster's Clustal-Omega .aln for multi-member clusters (FileIOManager.java:770-775) -- absent when that
                 step has not run -- and the bare sequence for one-member clusters (Cluster.getFastaString,
                 Cluster.java:167-176)
"""
from __future__ import annotations

import functools

SEP = "\t"          # Hammock.CSV_SEPARATOR


def _i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def seq_size(seq) -> int:
    """UniqueSequence.size() (UniqueSequence.java:81-88): sum over the label map, Java int arithmetic"""
    t = 0
    for v in seq[1].values():
        t = _i32(t + v)
    return t


def cluster_size(cl) -> int:
    """Cluster.size() (Cluster.java:156-158; accumulated in the constructor / insert, :31-41, 50-74)"""
    t = 0
    for s in cl[1]:
        t = _i32(t + seq_size(s))
    return t


def _java_string_compare(a: str, b: str) -> int:
    """String.compareTo: first differing UTF-16 unit, else the length difference"""
    for x, y in zip(a, b):
        if x != y:
            return ord(x) - ord(y)
    return len(a) - len(b)


def _cluster_compare(c1, c2) -> int:
    """Cluster.compareTo (Cluster.java:197-204)"""
    s1, s2 = cluster_size(c1), cluster_size(c2)
    if s1 != s2:
        return _i32(s1 - s2)
    return _i32(c1[0] - c2[0])


def _size_alphabetic_compare(o1, o2) -> int:
    """UniqueSequenceSizeAlphabeticComparator (UniqueSequence.java:238-249): SizeComparator, then the strings"""
    r = _i32(seq_size(o1) - seq_size(o2))
    if r == 0:
        r = _java_string_compare(o1[0], o2[0])
    return r


def _unique_sequence_compare(a, b) -> int:
    """UniqueSequence.compareTo (UniqueSequence.java:160-171): size, then the REVERSED string order"""
    if seq_size(a) != seq_size(b):
        return _i32(seq_size(a) - seq_size(b))
    return -_java_string_compare(a[0], b[0])


def _sorted_reverse(items, cmp):
    """Collections.sort(list, Collections.reverseOrder(cmp)) -- a stable merge sort"""
    return sorted(items, key=functools.cmp_to_key(lambda x, y: cmp(y, x)))


def default_alignments(clusters):
    """FileIOManager.getAlignmentsMap when no Clustal-Omega run exists: one-member clusters contribute their sequence
    (getAllAlignmentLines -> getFastaString, lines 1, 3, ... = the sequences), multi-member clusters nothing"""
    return {cl[0]: [cl[1][0][0]] for cl in clusters if len(cl[1]) == 1}


def _write_cluster_sequences(sequences, clusters, labels, alignments) -> str:
    """FileIOManager.writeClusterSequencesToCsv (FileIOManager.java:594-638)"""
    msa_map = {}
    sequence_cluster_map = {}
    for cl in clusters:
        for seq in cl[1]:
            sequence_cluster_map[seq[0]] = cl
        if cl[0] in alignments:
            for line in alignments[cl[0]]:
                msa_map[line.replace("-", "")] = line
    out = ["cluster_id" + SEP + "sequence" + SEP + "alignment" + SEP + "sum" + "".join(SEP + lab for lab in labels) + "\n"]
    for seq in sequences:
        cluster = sequence_cluster_map.get(seq[0])
        if cluster is not None:
            row = str(cluster[0]) + SEP + seq[0] + SEP
            row += (msa_map[seq[0]] if seq[0] in msa_map else "NA") + SEP
        else:
            row = "NA" + SEP + seq[0] + SEP + "NA" + SEP
        row += str(seq_size(seq))
        for lab in labels:
            row += SEP + str(seq[1].get(lab, 0))
        out.append(row + "\n")
    return "".join(out)


def cluster_sequences_tsv(clusters, labels, alignments=None) -> str:
    """FileIOManager.saveClusterSequencesToCsv (FileIOManager.java:398-404): clusters by Cluster.compareTo descending,
    members by UniqueSequenceSizeAlphabeticComparator descending (getSortedSequences, :530-538)"""
    alignments = default_alignments(clusters) if alignments is None else alignments
    ordered = []
    for cl in _sorted_reverse(clusters, _cluster_compare):
        ordered.extend(_sorted_reverse(cl[1], _size_alphabetic_compare))
    return _write_cluster_sequences(ordered, clusters, labels, alignments)


def cluster_sequences_tsv_ordered(clusters, labels, ordered_sequences, alignments=None) -> str:
    """FileIOManager.saveClusterSequencesToCsvOrdered (FileIOManager.java:371-374)"""
    alignments = default_alignments(clusters) if alignments is None else alignments
    return _write_cluster_sequences(ordered_sequences, clusters, labels, alignments)


def clusters_tsv(clusters, labels) -> str:
    """FileIOManager.SaveClustersToCsv (FileIOManager.java:649-676) + getClusterLabelsMap (:685-699)"""
    out = ["cluster_id" + SEP + "main_sequence" + SEP + "sum" + "".join(SEP + lab for lab in labels) + "\n"]
    for cl in _sorted_reverse(clusters, _cluster_compare):
        seqs = _sorted_reverse(cl[1], _unique_sequence_compare)
        row = str(cl[0]) + SEP + seqs[0][0] + SEP + str(cluster_size(cl))
        counts = {}
        for seq in cl[1]:
            for lab, v in seq[1].items():
                counts[lab] = _i32(counts.get(lab, 0) + v)
        for lab in labels:
            row += SEP + str(counts.get(lab, 0))
        out.append(row + "\n")
    return "".join(out)


def input_statistics(sequences, labels) -> str:
    """FileIOManager.saveInputStatistics (FileIOManager.java:709-729) with getTotalLabelCounts / getUniqueLabelCounts:
    no newline after the last row"""
    total = {lab: 0 for lab in labels}
    unique = {lab: 0 for lab in labels}
    for seq in sequences:
        for lab, v in seq[1].items():
            if lab in total:
                total[lab] = _i32(total[lab] + v)
                unique[lab] += 1
    out = "".join(SEP + lab for lab in labels) + "\n"
    out += "total_count" + "".join(SEP + str(total[lab]) for lab in labels) + "\n"
    out += "unique_count" + "".join(SEP + str(unique[lab]) for lab in labels)
    return out
