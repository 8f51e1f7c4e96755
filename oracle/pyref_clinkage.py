"""Second, independently structured restatement of ClinkageSequenceClusterer.cluster -- TEST
INFRASTRUCTURE ONLY, PARITY UNPINNED (see oracle/hammock_oracle.h).  Cross-checks oracle/clinkage_oracle.c.

Structured differently from the C oracle on purpose:
  * no cluster-score matrix: the complete-linkage score of two clusters is recomputed from the sequence-pair scores of
    their member lists each time (what ClinkageClusterScorer does; the reference's cache only memoises it),
  * the java.util.HashSet is modelled as Java's node table itself: a list of bins, each a Python list in chain order,
    rebuilt on resize by HashMap's lo/hi split,
  * clusters are Python objects with member lists.

Citations: /root/reference/src/cz/krejciadam/hammock/<file>:<line>.
"""
from __future__ import annotations

import numpy as np

from .pyref import PyRef, JMIN


class JavaHashSet:
    """java.util.HashSet<Cluster> as implemented by OpenJDK 8+ (HashMap.putVal / resize / removeNode / HashIterator),
    for keys whose hashCode is Cluster.hashCode() = 79 * 7 + id (Cluster.java:179-183)."""

    def __init__(self):
        self.table = None          # allocated by the first add (HashMap.resize)
        self.threshold = 0
        self.size = 0

    @staticmethod
    def _hash(cid: int) -> int:
        h = (79 * 7 + cid) & 0xFFFFFFFF
        return h ^ (h >> 16)

    def _resize(self):
        if self.table is None:
            self.table = [[] for _ in range(16)]
            self.threshold = 12
            return
        old = self.table
        oc = len(old)
        new = [[] for _ in range(2 * oc)]
        for j, chain in enumerate(old):
            for cid in chain:                       # lo list stays at j, hi list goes to j + oldCap, order preserved
                new[j + oc if self._hash(cid) & oc else j].append(cid)
        self.table = new
        self.threshold = int(2 * oc * 0.75)

    def add(self, cid: int):
        if self.table is None:
            self._resize()
        chain = self.table[self._hash(cid) & (len(self.table) - 1)]
        if cid in chain:
            return
        assert len(chain) < 7, "bin would be treeified: iteration order no longer modelled"
        chain.append(cid)
        self.size += 1
        if self.size > self.threshold:
            self._resize()

    def remove(self, cid: int):
        chain = self.table[self._hash(cid) & (len(self.table) - 1)]
        if cid in chain:
            chain.remove(cid)
            self.size -= 1

    def __iter__(self):
        if self.table is not None:
            for chain in self.table:
                yield from chain

    def first(self) -> int:
        return next(iter(self))


class _Cluster:
    def __init__(self, cid, members, size):
        self.id, self.members, self.size = cid, members, size


def clinkage_cluster(seqs, abundance, matrix, threshold, max_shift, shift_penalty):
    """-> (cluster_id[n], member_rank[n], result_order) with the conventions of hmko_clinkage_cluster"""
    n = len(seqs)
    ref = PyRef(seqs, abundance, matrix, threshold, max_shift, shift_penalty, 0)
    T = int(threshold)
    # sequence-pair scores, row by row: P[a, b] = sequenceScore(seq a, seq b)
    pair = np.empty((n, n), dtype=np.int64)
    everyone = np.arange(n)
    for q in range(n):
        pair[:, q] = ref.scores(everyone, q)

    def cluster_score(c1, c2):                       # ClinkageClusterScorer.java:30-49
        sub = pair[np.ix_(c1.members, c2.members)]
        m = int(sub.min())
        return JMIN + 1 if m < T else m

    by_id = {}
    active, ready = JavaHashSet(), JavaHashSet()
    current_id = 1                                                        # ClinkageSequenceClusterer.java:48
    for i in range(n):                                                    # :49-54
        by_id[current_id] = _Cluster(current_id, [i], int(abundance[i]))
        active.add(current_id)
        current_id += 1
    stack = []
    while active.size > 1:                                                # :62
        stack.append(active.first())                                      # :69-70
        while stack:
            top = by_id[stack[-1]]
            # findNearestClusterParallel / NearestClusterRunner (:137-177, 258-293) as one sort key
            cand = [(-cluster_score(by_id[c], top), -by_id[c].size, c) for c in active if c != top.id]
            if cand:
                neg, _, near = min(cand)
                max_score = -neg
            else:
                near, max_score = None, JMIN
            if max_score < T:                                             # :85-91
                stack.pop()
                ready.add(top.id)
                active.remove(top.id)
                continue
            if len(stack) > 1 and stack[-2] == near:                      # :95-109
                current_id += 1
                stack.pop(); stack.pop()
                active.remove(top.id); active.remove(near)
                other = by_id[near]
                size = (top.size + other.size + 2 ** 31) % 2 ** 32 - 2 ** 31
                by_id[current_id] = _Cluster(current_id, top.members + other.members, size)
                active.add(current_id)
            else:
                stack.append(near)                                        # :111
    ready.add(active.first())                                             # :116
    cluster_id = np.zeros(n, dtype=np.int32)
    member_rank = np.zeros(n, dtype=np.int32)
    order = []
    for cid in ready:                                                     # :119-123
        order.append(cid)
        for r, m in enumerate(by_id[cid].members):
            cluster_id[m] = cid
            member_rank[m] = r
    return cluster_id, member_rank, np.array(order, dtype=np.int32)
