/*
 * hammock_oracle.h -- CPU ORACLE for Hammock's greedy initial clustering stage.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (hammock_b200/, libhammock_b200.so) never links, imports or calls anything here.
 *
 * It is a plain-C restatement of the reference's Java (Hammock v1.2.0, /root/reference),
 * written from the source semantics; every function cites the file:line it follows
 * (paths relative to src/cz/krejciadam/hammock/).
 *
 * PARITY UNPINNED: the reference ships no tests, no golden outputs and cannot be run
 * here (no JVM).  The oracle is pinned only against hand-checkable known-answer
 * vectors derived from the Java source and against a
 * second, independently written numpy restatement (oracle/pyref.py).
 */
#ifndef HAMMOCK_ORACLE_H
#define HAMMOCK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* UniqueSequence.java:23-26 -- residue code = index in this string */
#define HMKO_ALPHABET "ARNDCQEGHILKMFPSTWYVBZX*"
#define HMKO_NRES 24

/* status codes (shared numbering with include/hammock_b200.h) */
enum {
    HMKO_OK = 0,
    HMKO_ERR_SHIFT_TOO_BIG = 1,   /* DataException, ShiftedScorer.java:59-62            */
    HMKO_ERR_NULL_CLUSTER = 2,    /* NullPointerException, LimitedGreedy...java:104,108 */
    HMKO_ERR_BAD_RESIDUE = 3,     /* FileFormatException, UniqueSequence.java:51-54     */
    HMKO_ERR_FILE_FORMAT = 6,     /* FileFormatException in the loaders                 */
    HMKO_ERR_IO = 7,
    HMKO_ERR_EMPTY = 8,           /* clinkage: NoSuchElementException on an empty input (ClinkageSequenceClusterer.java:116) */
    HMKO_ERR_ASYMMETRIC = 9,      /* clinkage: asymmetric substitution matrix (the reference's cached scores become schedule dependent) */
    HMKO_ERR_TREEIFIED = 10       /* clinkage: a HashMap bin reached the tree threshold; the bucket-order emulation is not exact */
};

/* work counters */
typedef struct {
    int64_t p1_steps, p1_new_clusters, p1_joins, p1_orphans;
    int64_t p1_pairs;            /* pair scores executed in phase 1 (with reference early exit) */
    int64_t p2_queries, p2_assigned;
    int64_t p2_pairs_early;      /* pair scores executed in phase 2 (reference early exit)      */
    int64_t p2_pairs_dense;      /* pair scores a dense (no early exit) evaluation would do     */
    int64_t cells;               /* matrix cells summed over all executed pair scores           */
    int32_t npe_step;            /* phase-1 step index at which status 2 fired, else -1         */
} hmko_counters;

/* UniqueSequence.java:46-57 : letter -> code, case-insensitive; -1 = not in alphabet */
int hmko_encode_char(char c);
/* returns 0 or HMKO_ERR_BAD_RESIDUE; out must hold strlen(s) bytes */
int hmko_encode(const char* s, uint8_t* out);

/* FileIOManager.java:46-81 */
int hmko_load_matrix(const char* path, int32_t* matrix576, char* err, size_t errlen);

/* ShiftedScorer.java:48-95 ; seq1 = first argument, seq2 = second argument.
 * *shift (may be NULL) receives what scoreWithShift reports.  *status set on error. */
int32_t hmko_score_with_shift(const uint8_t* seq1, int len1, const uint8_t* seq2, int len2,
                              const int32_t* matrix576, int max_shift, int shift_penalty,
                              int* shift, int* status);

/* number of matrix cells / shifts one pair score sums */
int64_t hmko_pair_cells(int len1, int len2, int max_shift);
int hmko_pair_shifts(int len1, int len2, int max_shift);

/* UniqueSequence.java:176-203 + :238-261, order "size": permutation perm[0..n) such that
 * sequence perm[i] (input index) is i-th in clustering order.  Stable. */
void hmko_sort_order_size(int32_t n, const uint8_t* residues, const int32_t* offsets,
                          const int32_t* abundance, int32_t* perm);

/* Hammock.java:394-401, 1409-1434, 1554-1563 -- automatic parameters */
void hmko_default_params(int32_t n, const int32_t* offsets, int32_t* threshold,
                         int32_t* max_shift, int32_t* max_clusters);
/* Hammock.java:1421-1427 */
int32_t hmko_check_max_shift(int32_t n, const int32_t* offsets, int32_t max_shift);

/* LimitedGreedySequenceClusterer.java:39-120 (+ ClinkageClusterScorer.java:30-49,
 * ClinkageSequenceClusterer.java:137-177, 258-293).  Sequences are ALREADY in clustering
 * order.  nthreads > 1 parallelises the nearest-cluster search over candidates (OpenMP),
 * like findNearestClusterParallel does over its parts; the result does not depend on it.
 * Outputs (caller-allocated, n entries each): cluster_id = founder index, member_rank =
 * position in Cluster.getSequences(), result_order = ids of the returned List<Cluster>. */
int hmko_greedy_cluster(int32_t n, const uint8_t* residues, const int32_t* offsets,
                        const int32_t* abundance, const int32_t* matrix576,
                        int32_t threshold, int32_t max_shift, int32_t shift_penalty,
                        int32_t max_clusters, int32_t nthreads,
                        int32_t* cluster_id, int32_t* member_rank, int32_t* result_order,
                        int32_t* n_result, int32_t* n_multi, hmko_counters* counters);

/* Bounded-sample CPU baseline: runs the same algorithm but stops after max_p1_steps phase-1
 * steps and max_p2_queries phase-2 queries (<=0: no bound).  Outputs may be NULL. */
int hmko_greedy_cluster_bounded(int32_t n, const uint8_t* residues, const int32_t* offsets,
                                const int32_t* abundance, const int32_t* matrix576,
                                int32_t threshold, int32_t max_shift, int32_t shift_penalty,
                                int32_t max_clusters, int32_t nthreads,
                                int64_t max_p1_steps, int64_t max_p2_queries,
                                int32_t* cluster_id, int32_t* member_rank,
                                int32_t* result_order, int32_t* n_result, int32_t* n_multi,
                                hmko_counters* counters);

/* SURVEY.md 8(f) N1 -- ClinkageSequenceClusterer.cluster (ClinkageSequenceClusterer.java:43-124 with
 * CachedClusterScorer.java:38-125 and the java.util.HashSet iteration order, see clinkage_oracle.c).  Sequences in the
 * order the caller hands them over (Hammock.runClinkageClustering does NOT sort, Hammock.java:449-462).
 * cluster_id[i] = Cluster.getId() of sequence i's cluster (singletons keep i + 1, merged clusters get n + 2, n + 3, ...),
 * member_rank[i] = position in Cluster.getSequences(), result_order[0..*n_result) = ids of the returned list. */
int hmko_clinkage_cluster(int32_t n, const uint8_t* residues, const int32_t* offsets, const int32_t* abundance,
                          const int32_t* matrix576, int32_t threshold, int32_t max_shift, int32_t shift_penalty,
                          int32_t* cluster_id, int32_t* member_rank, int32_t* result_order, int32_t* n_result,
                          int64_t* nearest_searches);

/* FileIOManager.java:159-216 -- fasta reader.  Returns a malloc'ed table; free with
 * hmko_fasta_free.  Sequences in first-occurrence order; abundance = sum over labels. */
typedef struct {
    int32_t n;
    uint8_t* residues;
    int32_t* offsets;    /* n+1 */
    int32_t* abundance;  /* n   */
    char* text;          /* concatenated upper-case strings, same offsets */
} hmko_fasta;
int hmko_load_fasta(const char* path, hmko_fasta* out, char* err, size_t errlen);
void hmko_fasta_free(hmko_fasta* f);

#ifdef __cplusplus
}
#endif
#endif
