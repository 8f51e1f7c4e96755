// hmk_resolve.h -- in-order resolution passes of the greedy engine (host/device source).
//
// The bulk kernels score a whole batch of queries against the PRE-batch state; these
// routines then replay the reference's sequential decisions in reference order, repairing
// every intra-batch dependency (partners consumed by earlier queries, clusters created or
// grown inside the batch, the K limit, the null-object quirks), so that the result is
// bit-identical to LimitedGreedySequenceClusterer (reference
// LimitedGreedySequenceClusterer.java:39-120).
//
// Each routine is executed by ONE warp with warp-uniform control flow; data-parallel inner
// loops are strided over lanes and combined with the reductions of the executor `W`.
// On the GPU W = HmkWarp (shuffles); the CPU test harness (tests/emu) instantiates the very
// same source with W = HmkSerial (one lane) to check the logic against the oracle.
#pragma once
#include "hmk_common.h"

// ---------------------------------------------------------------- executors
struct HmkSerial {
    HMK_HD int lane() const { return 0; }
    HMK_HD int nl() const { return 1; }
    HMK_HD int rmin(int v) const { return v; }
    HMK_HD bool all(bool p) const { return p; }
    HMK_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
    HMK_HD void sync() const {}
};

#if defined(__CUDACC__)
struct HmkWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int nl() const { return 32; }
    __device__ __forceinline__ int rmin(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ void sync() const { __threadfence_block(); __syncwarp(); }
};
#endif

HMK_HD int hmk_ffs(uint32_t m) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

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

// ---------------------------------------------------------------- phase 1
struct HmkP1Batch {
    int32_t nq;
    const int32_t* qid;       // [nq] ascending ids, all singletons at batch start
    // partner (B) search: per query the best `kb` alive singletons j > q at batch start
    int32_t kb;
    const uint64_t* bk_key;   // [nq][kb] sorted descending
    const int32_t* bk_cnt;    // [nq]
    const int32_t* bk_ovf;    // [nq] 1 = more hits existed than the list holds
    // cluster (A) search: per query the pre-batch clusters whose every pre-batch member
    // scores >= T, as linked lists (ac_head[b] -> entries)
    const int32_t* ac_head;   // [nq]
    const int32_t* ac_next;
    const int32_t* ac_slot;
    const int32_t* ac_score;  // min over the pre-batch members
    // intra-batch scores ib[b*ib_stride + b2] = S(member = qid[b2], query = qid[b])
    const int32_t* ib;
    int32_t ib_stride;
    // scratch (device global, single warp): members added during this batch
    int32_t* bm_slot;         // [2*nq]
    int32_t* bm_ref;          // >= 0: batch index of the member; < 0: ~sequence id (a partner)
    int32_t* nf_b;            // [nq] founders of clusters created in this batch (batch index)
    int32_t* nf_slot;         // [nq]
};

// complete-linkage over the members a cluster received during this batch
template <class W>
HMK_HD bool hmk_p1_eval_added(const HmkState& S, const HmkP1Batch& B, const W& w, int32_t b, int32_t q,
                              int32_t c, int32_t bm_n, int32_t& cl, int64_t& npairs) {
    bool ok = true;
    int32_t mn = cl;
    for (int32_t e = w.lane(); e < bm_n; e += w.nl()) {
        if (B.bm_slot[e] != c) continue;
        int32_t ref = B.bm_ref[e];
        int32_t s;
        if (ref >= 0) s = B.ib[(int64_t)b * B.ib_stride + ref];
        else { s = hmk_state_score(S, ~ref, q); npairs++; }
        if (s < S.T) ok = false;
        if (s < mn) mn = s;
    }
    ok = w.all(ok);
    cl = w.rmin(mn);
    return ok;
}

// firstPhase loop body for one batch (LimitedGreedySequenceClusterer.java:90-116)
template <class W>
HMK_HD void hmk_p1_resolve(const HmkState& S, const HmkP1Batch& B, const W& w) {
    HmkCtl* ctl = S.ctl;
    int32_t ncl = ctl->ncl, unproc = ctl->unproc_alive;
    int32_t steps = ctl->steps, joins = ctl->joins, creates = ctl->creates, orphans = ctl->orphans;
    int32_t status = HMK_P1_CONTINUE, npe_step = -1, cur = ctl->cur;
    int32_t bm_n = 0, nf_n = 0;
    int64_t npairs = 0;
    w.sync();

    int32_t b = 0;
    for (; b < B.nq; b++) {
        const int32_t q = B.qid[b];
        if (ncl >= S.K) { status = HMK_P1_DONE; cur = q; break; }                  // :90
        if (S.slot[q] >= 0) continue;  // consumed as a partner earlier in this batch (:101,110)

        // ---- B: nearest among initialList[index+1 ..]                              (:93)
        int bkind = 0;  // 0 = Java null, 1 = found, 2 = (null cluster, MIN_VALUE) object
        int32_t bscore = HMK_JMIN, bid = -1;
        if (unproc - 1 == 0) bkind = 2;           // empty sub-list (ClinkageSequenceClusterer.java:138-140)
        else {
            const int32_t cnt = B.bk_cnt[b];
            int first = 0x7fffffff;
            for (int32_t e = w.lane(); e < cnt; e += w.nl()) {
                uint32_t r = hmk_key_rank(B.bk_key[(int64_t)b * B.kb + e]);
                int32_t id = S.id_of_rank ? S.id_of_rank[r] : (int32_t)r;
                if (S.slot[id] < 0) { first = e; break; }
            }
            first = w.rmin(first);
            if (first != 0x7fffffff) {
                uint64_t key = B.bk_key[(int64_t)b * B.kb + first];
                uint32_t r = hmk_key_rank(key);
                bid = S.id_of_rank ? S.id_of_rank[r] : (int32_t)r;
                bscore = hmk_key_score(key);
                bkind = 1;
            } else if (B.bk_ovf[b]) {
                status = HMK_P1_RESTART; cur = q; break;   // list truncated: rescore from q
            }
        }

        // ---- A: nearest among actualClusters (complete linkage)                     (:92)
        int akind = 0;
        HmkBestCluster best;
        best.score = HMK_JMIN; best.size = 0; best.fid = 0; best.slot = -1;
        if (ncl == 0) akind = 2;
        else {
            for (int32_t e = B.ac_head[b]; e >= 0; e = B.ac_next[e]) {   // pre-batch clusters
                int32_t c = B.ac_slot[e], cl = B.ac_score[e];
                if (hmk_p1_eval_added(S, B, w, b, q, c, bm_n, cl, npairs))
                    hmk_consider(best, cl, S.c_size[c], S.c_founder[c], c);
            }
            for (int32_t base = 0; base < nf_n; base += w.nl()) {        // clusters born in this batch
                int32_t i = base + w.lane();
                bool hit = false;
                if (i < nf_n) hit = B.ib[(int64_t)b * B.ib_stride + B.nf_b[i]] >= S.T;
                uint32_t m = w.ballot(hit);
                while (m) {
                    int l = hmk_ffs(m);
                    m &= m - 1;
                    int32_t c = B.nf_slot[base + l], cl = HMK_JMAX;
                    if (hmk_p1_eval_added(S, B, w, b, q, c, bm_n, cl, npairs))
                        hmk_consider(best, cl, S.c_size[c], S.c_founder[c], c);
                }
            }
            if (best.slot >= 0) akind = 1;
        }

        // ---- decision                                                              (:94-114)
        const int32_t ascore = akind == 1 ? best.score : HMK_JMIN;
        bool join = false, create = false;
        if (akind != 0) {
            if (bkind != 0) { if (ascore >= bscore) join = true; else create = true; }
            else join = true;
        } else if (bkind != 0) create = true;
        if ((join && akind == 2) || (create && bkind == 2)) {   // null.insertAll / null.getSequences
            status = HMK_P1_NPE; npe_step = steps; cur = q; break;
        }
        if (w.lane() == 0) {
            if (join) {
                const int32_t c = best.slot;
                S.next[S.c_tail[c]] = q; S.next[q] = -1; S.c_tail[c] = q;
                S.rank[q] = S.c_count[c]; S.c_count[c] += 1;
                S.c_size[c] = hmk_wadd(S.c_size[c], S.ab[q]);
                S.slot[q] = c;
                B.bm_slot[bm_n] = c; B.bm_ref[bm_n] = b;
            } else if (create) {
                const int32_t c = ncl;
                S.c_founder[c] = q; S.c_tail[c] = bid; S.c_count[c] = 2;
                S.c_size[c] = hmk_wadd(S.ab[q], S.ab[bid]);
                S.next[q] = bid; S.next[bid] = -1;
                S.slot[q] = c; S.rank[q] = 0; S.slot[bid] = c; S.rank[bid] = 1;
                B.bm_slot[bm_n] = c; B.bm_ref[bm_n] = b;
                B.bm_slot[bm_n + 1] = c; B.bm_ref[bm_n + 1] = ~bid;
                B.nf_b[nf_n] = b; B.nf_slot[nf_n] = c;
            }
        }
        if (join) { bm_n += 1; joins++; }
        else if (create) { bm_n += 2; nf_n += 1; ncl++; unproc--; creates++; }
        else orphans++;
        steps++;
        unproc--;
        cur = q + 1;
        w.sync();
    }
    if (status == HMK_P1_CONTINUE && (ncl >= S.K || unproc <= 0)) status = HMK_P1_DONE;
    if (w.lane() == 0) {
        ctl->cur = cur; ctl->ncl = ncl; ctl->unproc_alive = unproc; ctl->status = status;
        if (npe_step >= 0) ctl->npe_step = npe_step;
        ctl->steps = steps; ctl->joins = joins; ctl->creates = creates; ctl->orphans = orphans;
        if (status == HMK_P1_RESTART) ctl->restarts += 1;
    }
#if defined(__CUDA_ARCH__)
    if (npairs) atomicAdd((unsigned long long*)&ctl->scalar_pairs, (unsigned long long)npairs);
#else
    ctl->scalar_pairs += npairs;
#endif
    w.sync();
}

// ---------------------------------------------------------------- phase 2
// Candidate pairs (single q, cluster c) = founder AND every phase-1 member score >= T.  They
// are held twice: grouped by query (cq_*) and grouped by cluster in query order (cc_*).
// A query is resolved once it heads the list of every cluster it is a candidate of; that
// reproduces the reference's sequential order (LimitedGreedySequenceClusterer.java:59-66)
// because a query's decision only depends on earlier queries that share a candidate cluster.
struct HmkP2 {
    int32_t ncl;
    const int32_t* singles;   // [ns] ascending ids of the phase-2 queries
    const int32_t* qstart;    // [ns+1] into cq_*
    const int32_t* cq_c;      // candidate cluster slot (ascending inside a query)
    const int32_t* cq_s;      // complete-linkage min over the phase-1 members
    const int32_t* cstart;    // [ncl+1] into cc_q / dyn
    const int32_t* cc_q;      // query index (into singles), ascending inside a cluster
    const int32_t* head_cur;  // [ncl] snapshot of this round
    int32_t* head_nxt;        // [ncl]
    int32_t* dyn;             // members joined in phase 2: dyn[cstart[c] + t]
    int32_t* dyn_n;           // [ncl]
    int32_t* remaining;       // set to 1 while any list is non-empty
    int32_t* progress;        // set to 1 when any head advanced this round
};

template <class W>
HMK_HD void hmk_p2_resolve_query(const HmkState& S, const HmkP2& P, const W& w, int32_t qi, int64_t& npairs) {
    const int32_t q = P.singles[qi];
    HmkBestCluster best;
    best.score = HMK_JMIN; best.size = 0; best.fid = 0; best.slot = -1;
    for (int32_t i = P.qstart[qi]; i < P.qstart[qi + 1]; i++) {
        const int32_t c = P.cq_c[i];
        int32_t mn = P.cq_s[i];
        bool ok = true;
        const int32_t nd = P.dyn_n[c];
        const int32_t* dm = P.dyn + P.cstart[c];
        for (int32_t t = w.lane(); t < nd; t += w.nl()) {
            int32_t s = hmk_state_score(S, dm[t], q);   // ClinkageClusterScorer.java:36-44
            npairs++;
            if (s < S.T) ok = false;
            if (s < mn) mn = s;
        }
        ok = w.all(ok);
        mn = w.rmin(mn);
        if (ok) hmk_consider(best, mn, S.c_size[c], S.c_founder[c], c);
    }
    if (best.slot >= 0 && w.lane() == 0) {   // foundCluster.getScore() >= threshold holds by construction (:61)
        const int32_t c = best.slot;
        P.dyn[P.cstart[c] + P.dyn_n[c]] = q;
        P.dyn_n[c] += 1;
        S.rank[q] = S.c_count[c]; S.c_count[c] += 1;
        S.c_size[c] = hmk_wadd(S.c_size[c], S.ab[q]);
        S.slot[q] = c;
    }
    w.sync();
}

// one round for cluster c (one warp)
template <class W>
HMK_HD void hmk_p2_round(const HmkState& S, const HmkP2& P, const W& w, int32_t c) {
    const int32_t h0 = P.head_cur[c], end = P.cstart[c + 1];
    int32_t h = h0;
    int64_t npairs = 0;
    while (h < end) {
        const int32_t qi = P.cc_q[h];
        const int32_t qs = P.qstart[qi], qe = P.qstart[qi + 1];
        if (qe - qs == 1) {   // c is the only candidate: no other cluster is involved
            hmk_p2_resolve_query(S, P, w, qi, npairs);
            h++;
            continue;
        }
        if (h != h0) break;   // snapshot semantics: shared queries only at the round's head
        bool ready = true;
        for (int32_t i = qs + w.lane(); i < qe; i += w.nl()) {
            int32_t c2 = P.cq_c[i];
            if (P.cc_q[P.head_cur[c2]] != qi) ready = false;
        }
        ready = w.all(ready);
        if (ready) {
            if (P.cq_c[qs] == c) hmk_p2_resolve_query(S, P, w, qi, npairs);   // owner = lowest slot
            h++;
        }
        break;
    }
    if (w.lane() == 0) {
        P.head_nxt[c] = h;
        if (h < end) *P.remaining = 1;
        if (h != h0) *P.progress = 1;
    }
    // every lane adds the pair scores it computed itself
#if defined(__CUDA_ARCH__)
    if (npairs) atomicAdd((unsigned long long*)&S.ctl->scalar_pairs, (unsigned long long)npairs);
#else
    S.ctl->scalar_pairs += npairs;
#endif
}
