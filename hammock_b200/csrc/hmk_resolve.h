// hmk_resolve.h -- in-order resolution passes of the greedy engine (host/device source).
//
// The bulk kernels score a whole batch of queries against the PRE-batch state; these
// routines then replay the reference's sequential decisions in reference order, repairing
// every intra-batch dependency (partners consumed by earlier queries, clusters created or
// grown inside the batch, the K limit, the null-object quirks), so that the result is
// bit-identical to LimitedGreedySequenceClusterer (reference
// LimitedGreedySequenceClusterer.java:39-120).
//
// This header holds the shared state description; the resolver kernels themselves are in
// hmk_kernels.cuh (hmk_p1_resolve_kernel, hmk_p2_*).
#pragma once
#include "hmk_common.h"

// ---------------------------------------------------------------- shared state
struct HmkState {
    // immutable problem
    int32_t n;
    const uint8_t* res;
    const int32_t* off;
    const int32_t* ab;         // abundance == UniqueSequence.size()
    const int32_t* M;          // 24x24
    int32_t T, X, P, K;
    const int32_t* id_of_rank; // tierank -> id, NULL = identity
    // clustering state
    int32_t* slot;      // [n]  -1 = still a singleton, else cluster slot (creation order)
    int32_t* rank;      // [n]  position in Cluster.getSequences()
    int32_t* next;      // [n]  member chain in insertion order (-1 terminates)
    int32_t* c_founder; // [K]  Cluster.getId()
    int32_t* c_size;    // [K]  abundance-weighted Cluster.size()
    int32_t* c_count;   // [K]  getUniqueSize()
    int32_t* c_tail;    // [K]
    int32_t* qbatch;    // [n]  id of the phase-1 batch in which the sequence was resolved as a query (-1: never); may be NULL
    HmkCtl* ctl;
};

HMK_HD int32_t hmk_state_score(const HmkState& S, int32_t member, int32_t query) {
    return hmk_pair_score(S.res + S.off[member], S.off[member + 1] - S.off[member],
                          S.res + S.off[query], S.off[query + 1] - S.off[query], S.M, S.X, S.P);
}

// NearestClusterRunner order (ClinkageSequenceClusterer.java:258-293): score desc, size desc, id asc
struct HmkBestCluster {
    int32_t score, size, fid, slot;
};
HMK_HD void hmk_consider(HmkBestCluster& b, int32_t score, int32_t size, int32_t fid, int32_t slot) {
    bool take;
    if (b.slot < 0) take = true;
    else if (score != b.score) take = score > b.score;
    else if (size != b.size) take = size > b.size;
    else take = fid < b.fid;
    if (take) { b.score = score; b.size = size; b.fid = fid; b.slot = slot; }
}

