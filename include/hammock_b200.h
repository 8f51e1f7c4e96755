/*
 * hammock_b200.h -- C ABI of the B200-native greedy-clustering stage of Hammock.
 *
 * Drop-in boundary.  The reference has no FFI layer; its narrowest seam for this path is
 * the Java interface
 *     SequenceClusterer { List<Cluster> cluster(List<UniqueSequence>) }
 *         (reference src/cz/krejciadam/hammock/SequenceClusterer.java:15-26)
 * instantiated for greedy in exactly one place
 *         new ShiftedScorer(scoringMatrix, shiftPenalty, maxShift)           Hammock.java:402
 *         new LimitedGreedySequenceClusterer(scorer, threshold, limit)       Hammock.java:403
 *         clusterer.cluster(sequences)                                       Hammock.java:409
 * hmk_greedy_cluster() replaces that call: the host hands over the (already ordered,
 * UniqueSequence.sortSequences, UniqueSequence.java:176-203) sequences as residue codes plus
 * the scorer/clusterer constructor arguments and rebuilds List<Cluster> from cluster_id /
 * member_rank / result_order.  INTEGRATION.md shows the JNI / FFM stub.
 *
 * Everything runs on the GPU (CUDA, sm_100a).  There is NO CPU fallback: without a usable
 * device every entry point returns HMK_ERR_CUDA.
 *
 * All pointers are host memory owned by the caller and only read/written during the call.
 */
#ifndef HAMMOCK_B200_H
#define HAMMOCK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMK_ABI_VERSION 2

/* status codes: the reference's failure modes on this path */
#define HMK_STATUS_OK 0
#define HMK_STATUS_SHIFT_TOO_BIG 1 /* DataException "Shift too big"      ShiftedScorer.java:59-62 */
#define HMK_STATUS_NULL_CLUSTER 2  /* NullPointerException               LimitedGreedySequenceClusterer.java:104,108 */
#define HMK_STATUS_BAD_RESIDUE 3   /* residue code >= 24                  UniqueSequence.java:51-54 */
#define HMK_STATUS_CUDA 4          /* CUDA / NCCL failure (no CPU fallback) */
#define HMK_STATUS_BAD_ARG 5
#define HMK_STATUS_UNSUPPORTED 6   /* hmk_clinkage_cluster only, see there */

typedef struct {
    int32_t n;                /* sequences, ALREADY in clustering order                              */
    const uint8_t* residues;  /* concatenated codes 0..23, alphabet "ARNDCQEGHILKMFPSTWYVBZX*"        */
                              /*   (UniqueSequence.java:23-26)                                        */
    const int32_t* offsets;   /* n+1 prefix offsets into residues                                    */
    const int32_t* abundance; /* UniqueSequence.size() (UniqueSequence.java:81-88), Java int          */
    const int32_t* matrix;    /* 24*24 row-major, as FileIOManager.loadScoringMatrix returns it       */
                              /*   (FileIOManager.java:46-81)                                         */
    int32_t threshold;        /* LimitedGreedySequenceClusterer ctor  (...Clusterer.java:22-26)       */
    int32_t max_shift;        /* ShiftedScorer ctor                   (ShiftedScorer.java:28-32)      */
    int32_t shift_penalty;    /* ShiftedScorer ctor                                                   */
    int32_t max_clusters;     /* LimitedGreedySequenceClusterer ctor                                  */
} hmk_greedy_in;

typedef struct {
    int32_t* cluster_id;      /* [n] Cluster.getId() of the sequence's cluster = founder's index      */
    int32_t* member_rank;     /* [n] position in Cluster.getSequences() (0 founder, 1 partner, ...)   */
    int32_t* result_order;    /* [n] ids of the returned List<Cluster>, in list order                 */
    int32_t n_result;         /* length of that list                                                  */
    int32_t n_multi;          /* multi-member clusters (they form the list's prefix)                  */
    int32_t error_step;       /* phase-1 step index for HMK_STATUS_NULL_CLUSTER, else -1              */
} hmk_greedy_out;

/* executed-work and timing report of the last run */
typedef struct {
    int64_t bulk_pairs;       /* pair scores executed by the bulk kernels                             */
    int64_t scalar_pairs;     /* pair scores executed one at a time (member checks, resolvers)        */
    int64_t bulk_cells;       /* matrix cells summed by the bulk kernels                              */
    int64_t bulk_ops;         /* algorithmic int ops = cells + shifts                  */
    double bulk_kernel_ms;    /* sum of CUDA-event durations of the bulk scoring launches             */
    double total_ms;          /* whole hmk_run, device timed                                          */
    double phase1_ms, phase2_ms;
    int32_t bulk_launches;    /* bulk scoring kernel launches                                         */
    int32_t total_launches;   /* all kernel launches                                                  */
    int32_t p1_steps, p1_joins, p1_new_clusters, p1_orphans, p1_batches, p1_restarts;
    int32_t p2_queries, p2_assigned, p2_rounds;
    int64_t p2_hits, p2_candidates;
    int32_t fast_path;        /* 0: generic scalar kernel, 1: packed SWAR kernel (one length), 2: packed per length bucket */
    int32_t lane_bits;        /* 8 or 16 on the fast path                                             */
    int32_t error_step;       /* phase-1 step of HMK_STATUS_NULL_CLUSTER in the last run, else -1     */
    int32_t flags;            /* HMK_FLAG_* of the last run                                           */
    int64_t xhits_kept;       /* phase-1 partner-search hits the run wanted to keep for phase 2       */
    int64_t xhits_capacity;   /* entries the kept-hit buffer had; kept > capacity == overflow         */
} hmk_stats;
/* hmk_stats.flags */
#define HMK_FLAG_P2_REUSED 1      /* phase 2 took its founder hits from the phase-1 partner searches   */
#define HMK_FLAG_XHIT_OVERFLOW 2  /* the kept-hit buffer overflowed: this run paid the separate founder  */
                                  /* pass (about 1.4x slower); the next run on this context sizes it right */
#define HMK_FLAG_ASYMMETRIC 4     /* substitution matrix not symmetric: hit reuse not applicable        */

typedef struct hmk_ctx hmk_ctx;

int hmk_abi_version(void);

/* One blocking call == SequenceClusterer.cluster(sequences).  Uses CUDA device `device`.  The library
 * keeps one context per device between calls (device buffers stay allocated, so repeated calls do not
 * pay allocation again); hmk_release_cached frees them. */
int hmk_greedy_cluster(const hmk_greedy_in* in, hmk_greedy_out* out, int device, char* errbuf, size_t errlen);
/* The same blocking call on n_gpus devices of THIS process (SURVEY.md 8b's `n_gpus`): the reference's host is
 * one JVM that calls cluster() once, synchronously (Hammock.java:409), so the library itself runs one worker
 * thread per device and builds the NCCL communicator (one unique id, ncclCommInitRank per thread).
 * devices == NULL means devices 0 .. n_gpus-1.  Contexts and the communicator are kept for the next call
 * with the same device list.  The result is identical to the one-GPU result. */
int hmk_greedy_cluster_multi(const hmk_greedy_in* in, hmk_greedy_out* out, const int32_t* devices, int32_t n_gpus,
                             char* errbuf, size_t errlen);
/* SURVEY.md 8(f) N1 -- the EXACT complete-linkage initial clustering, Hammock's default initial stage for up to 10 000
 * unique sequences (Hammock.java:371-373):  new ClinkageSequenceClusterer(scorer, threshold).cluster(sequences)
 * (Hammock.java:457-462, ClinkageSequenceClusterer.java:43-124; nearest-neighbour chain over complete-linkage scores with
 * CachedClusterScorer.java:38-125).  Same input struct (max_clusters is ignored; the sequences come in the caller's order --
 * runClinkageClustering does not sort).  Outputs: cluster_id[i] = Cluster.getId() of sequence i's cluster (i + 1 for
 * singletons, n + 2, n + 3, ... for merged clusters, in merge order), member_rank[i] = position in getSequences(),
 * result_order[0 .. n_result) = the ids of the returned list IN ITS ORDER, n_multi = clusters with more than one member.
 * The chain start and the order of the returned list come from java.util.HashSet iteration; they are reproduced for the
 * OpenJDK 8+ HashMap.  HMK_STATUS_UNSUPPORTED: asymmetric matrix (the reference's cached scores then depend on the thread
 * schedule), more than 32768 sequences, or a hash bin that the JDK would turn into a tree.  An empty input (the reference
 * throws NoSuchElementException) returns HMK_STATUS_BAD_ARG. */
int hmk_clinkage_cluster(const hmk_greedy_in* in, hmk_greedy_out* out, int device, char* errbuf, size_t errlen);
/* Frees the contexts the two calls above keep.  Call it before the process unloads CUDA if the memory
 * matters; the library never tears them down from a static destructor. */
void hmk_release_cached(void);

/* Handle API: the same work split into upload / device-resident run / download, so that a
 * host can keep the sequences resident and time the stages separately. */
int hmk_create(hmk_ctx** ctx, int device, char* errbuf, size_t errlen);
void hmk_destroy(hmk_ctx* ctx);
int hmk_upload(hmk_ctx* ctx, const hmk_greedy_in* in, char* errbuf, size_t errlen);
int hmk_run(hmk_ctx* ctx, char* errbuf, size_t errlen);
int hmk_download(hmk_ctx* ctx, hmk_greedy_out* out, char* errbuf, size_t errlen);
int hmk_get_stats(hmk_ctx* ctx, hmk_stats* stats);
/* Multi-GPU: one process per GPU.  Rank 0 obtains a 128-byte NCCL unique id, the host distributes
 * it (any channel), every rank calls hmk_init_distributed on its own context and then makes the
 * SAME upload / run / download calls with the SAME input.  Phase 1 stripes the later singletons
 * across ranks and all-gathers the per-rank best-hit lists; phase 2 shards the queries and
 * all-gathers the candidate lists; decisions are replicated, so every rank returns the full,
 * identical result. */
int hmk_nccl_unique_id(void* id128, char* errbuf, size_t errlen);
int hmk_init_distributed(hmk_ctx* ctx, int rank, int world, const void* id128, char* errbuf, size_t errlen);

/* CUDA-event stopwatch on the library's stream (brackets upload + run + download for end-to-end
 * timing); hmk_timer_end synchronises and returns milliseconds. */
int hmk_timer_begin(hmk_ctx* ctx);
int hmk_timer_end(hmk_ctx* ctx, double* ms);
/* roofline denominators measured on this device: out4[0] IADD3 lane-instructions/s (the INT32 ALU
 * peak), out4[1] lane-instructions/s of an IADD3 + IMAD mix (both integer pipes), out4[2]
 * conflict-free shared-memory load bytes/s, out4[3] SM count */
int hmk_measure_peaks(hmk_ctx* ctx, double* out4, char* errbuf, size_t errlen);
/* device time per section of the last run (only filled when option "profile" is 1):
 * p1_select, p1_partner, p1_cluster, p1_intra, p1_resolve, p2_setup, p2_filter, p2_check, p2_sort,
 * p2_base, p2_iterate, p2_commit, final */
#define HMK_NSECTIONS 13
int hmk_get_section_ms(hmk_ctx* ctx, double* out, int n);
/* tuning knobs; unknown names and out-of-range values return HMK_STATUS_BAD_ARG.  None of them changes the result.
 *   batch      phase-1 queries per batch (0 = automatic: 8 profile tiles, at most 512 and what the resolver's
 *              shared memory holds)
 *   kb         partner candidates kept per query (1..32, default 8)
 *   capq       initial capacity of the per-query cluster-candidate arrays (grown on demand)
 *   lookahead  batches prepared ahead on a side stream while batch i is resolved: 0, 1 or 2 (default 2)
 *   filter     1 = upper-bound filter + exact verify kernel where it applies (default), 0 = exact kernel only
 *   reuse      1 = phase 2 takes its founder hits from the phase-1 partner searches when the matrix is
 *              symmetric (default), 0 = always run the separate founder pass
 *   qt, waves  profiles per shared-memory tile (0 = as many as fit), grid waves per launch
 *   reserve    SMs the look-ahead bulk launches leave free for the main stream (resolver + small kernels; default 4)
 *   min_iters, bucket_aux   static grids: block iterations per stripe at least; mixed lengths: per-length launches on separate streams
 *   persistent  1 = the filter kernel runs as one CTA per SM taking database chunks dynamically (default), 0 = static grid
 *   p2_chunk, p2_window, hit_cap   phase-2 chunking / window size / initial hit-buffer size
 *   p2_first   queries in the first phase-2 window (0 = automatic; later windows grow towards p2_window)
 *   xhit_cap   entries of the buffer that keeps the phase-1 hits for phase 2 (0 = automatic: what the last run on this
 *              context needed, else 96 per sequence); an overflow is reported in hmk_stats.flags
 *   force_generic  1 = scalar kernel for everything (correctness path)
 *   profile    1 = time every bulk launch with CUDA events (hmk_stats.bulk_kernel_ms) */
int hmk_set_option(hmk_ctx* ctx, const char* name, int64_t value);

/* SequenceScorer.sequenceScore over a block of pairs (SequenceScorer.java:12-15,
 * ShiftedScorer.java:97-100): scores[a * n_second + b] = sequenceScore(seq1 = first_ids[a],
 * seq2 = second_ids[b]) for the uploaded sequences.  Used by the parity tests of the kernel. */
int hmk_score_block(hmk_ctx* ctx, const int32_t* first_ids, int32_t n_first, const int32_t* second_ids,
                    int32_t n_second, int32_t* scores, char* errbuf, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
