// hmk_kernels.cuh -- sm_100a kernels of the greedy-clustering engine.
//
// Hot kernel: hmk_bulk_fast<NW, MODE>.  One thread owns one database peptide (5-bit packed,
// one 64-bit word, <= 12 residues) and scores it against a tile of "profile" sequences held
// in shared memory.  A profile is the substitution matrix pre-arranged per diagonal:
// entry[h][j][r] (32 bit) holds, for database position j carrying residue r, the
// contributions to 4 (u8 lanes) or 2 (s16 lanes) of the 2X+1 shift diagonals.  One pair score
// = L x NW conflict-free LDS.32 + 3-input adds on packed lanes; lanes are biased so that
// "score >= threshold" is the lane's top bit, i.e. the common path ends in a single
// LOP3/branch and only hits (rare) are decoded.  This replaces the reference's scalar
// triple loop (ShiftedScorer.java:67-90) and the per-candidate ClinkageClusterScorer call
// (ClinkageClusterScorer.java:30-49) for all candidates of a batch at once.
//
// Profile tiles are staged global->shared with the TMA bulk-copy engine
// (cp.async.bulk + mbarrier; SASS UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hmk_common.h"
#include "hmk_resolve.h"

#define HMK_MAXL1 12           // residues per 64-bit packed word
#define HMK_BULK_THREADS 512

enum { HMK_MODE_TOPK = 0, HMK_MODE_EMIT = 1, HMK_MODE_DENSE = 2 };
enum { HMK_PROF_QUERY = 0, HMK_PROF_MEMBER = 1 };

// ---------------------------------------------------------------- scoring scheme
struct HmkScheme {
    int32_t L;        // uniform sequence length of the fast path
    int32_t X, P, T;
    int32_t nw;       // 32-bit words per profile entry
    int32_t lane16;   // 0: four u8 lanes per word, 1: two s16 lanes per word
    int32_t bias;     // added to every valid cell so entries are >= 0
    int32_t half;     // 128 or 32768: lane value >= half  <=>  score >= T
    int32_t prof_words;  // nw * HMK_MAXL1 * 24 (fixed row stride so LDS offsets are immediates)
};

// ---------------------------------------------------------------- profile builder
// prof[t][h][j][r] with j < HMK_MAXL1 (rows j >= L are zero); mode QUERY: profile sequence is the reference's seq2 (compared/query,
// the "shorter" one for equal lengths), threads will carry seq1 (member/candidate):
//   lane k gets M[p[j-k]][r];   mode MEMBER: profile sequence is seq1 (member), threads carry
// seq2 (query): lane k gets M[r][p[j+k]]   (M[shorter][longer], ShiftedScorer.java:71,75,110)
__global__ void hmk_build_profiles(HmkScheme sc, int mode, const int32_t* __restrict__ ids, int nq,
                                   const uint8_t* __restrict__ res, const int32_t* __restrict__ off,
                                   const int32_t* __restrict__ M, uint32_t* __restrict__ prof) {
    __shared__ int32_t sM[HMK_NRES * HMK_NRES];
    __shared__ uint8_t sp[HMK_MAXL1];
    const int t = blockIdx.x;
    if (t >= nq) return;
    for (int i = threadIdx.x; i < HMK_NRES * HMK_NRES; i += blockDim.x) sM[i] = M[i];
    const int32_t id = ids[t];
    if (threadIdx.x < sc.L) sp[threadIdx.x] = res[off[id] + threadIdx.x];
    __syncthreads();
    const int L = sc.L, lanes_per_word = sc.lane16 ? 2 : 4, lane_bits = sc.lane16 ? 16 : 8;
    uint32_t* out = prof + (size_t)t * sc.prof_words;
    for (int e = threadIdx.x; e < sc.prof_words; e += blockDim.x) {
        const int r = e % HMK_NRES, j = (e / HMK_NRES) % HMK_MAXL1, h = e / (HMK_NRES * HMK_MAXL1);
        uint32_t word = 0;
        if (j >= L) { out[e] = 0; continue; }
        for (int b = 0; b < lanes_per_word; b++) {
            const int lam = h * lanes_per_word + b;
            if (lam > 2 * sc.X) continue;
            const int k = lam - sc.X, ak = k < 0 ? -k : k;
            int32_t val = 0;
            const int pi = mode == HMK_PROF_QUERY ? j - k : j + k;
            if (pi >= 0 && pi < L)
                val = (mode == HMK_PROF_QUERY ? sM[sp[pi] * HMK_NRES + r] : sM[r * HMK_NRES + sp[pi]]) + sc.bias;
            if (j == 0) val += 2 * sc.P * ak + sc.half - sc.T - (L - ak) * sc.bias;   // lane constant
            word |= ((uint32_t)val & ((1u << lane_bits) - 1u)) << (b * lane_bits);
        }
        out[e] = word;
    }
}

// ---------------------------------------------------------------- TMA bulk copy helpers
__device__ __forceinline__ uint32_t hmk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void hmk_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hmk_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hmk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hmk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hmk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     hmk_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(hmk_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void hmk_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(hmk_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ---------------------------------------------------------------- bulk arguments
struct HmkBulkArgs {
    HmkScheme sc;
    // profile side
    const uint32_t* prof;      // [nq][prof_words]
    int32_t nq, qt, nqt;       // profiles, profiles per tile, tiles
    // thread side ("database")
    const uint64_t* packed;    // by sequence id
    const int32_t* db_ids;     // NULL: id = db_begin + i
    int32_t db_begin, ndb;
    int32_t nstripes, chunk;   // items per stripe
    const int32_t* slot;       // non-NULL: skip items whose slot >= 0 (no longer singletons)
    const int32_t* q_minid;    // non-NULL: only ids > q_minid[t] qualify (initialList[index+1..])
    const int32_t* tierank;    // NULL: tierank == id
    // MODE_TOPK
    int32_t kb;
    uint64_t* tk_key;          // [nstripes][nq][kb] descending
    int32_t* tk_cnt;           // [nstripes][nq]
    int32_t* tk_ovf;           // [nstripes][nq]
    // MODE_EMIT
    int4* hits;                // (t, i, score, 0)
    unsigned int* hit_count;
    unsigned int hit_cap;
    // MODE_DENSE
    int32_t* dense;            // dense[t * dense_stride + i]
    int32_t dense_stride;
    unsigned long long* pair_counter;
};

// ---------------------------------------------------------------- hit handling
struct HmkTopkSmem {
    uint64_t* key;    // [qt][kb]
    uint64_t* minkey; // [qt]
    int* cnt;         // [qt]
    int* lock;        // [qt]
    int* ovf;         // [qt]
};

__device__ __forceinline__ void hmk_topk_insert(const HmkTopkSmem& s, int t, int kb, uint64_t key) {
    volatile uint64_t* keys = s.key + (size_t)t * kb;
    volatile uint64_t* mink = s.minkey + t;
    volatile int* cnt = s.cnt + t;
    if (*cnt >= kb && key <= *mink) { s.ovf[t] = 1; return; }   // cheap reject (minkey only grows)
    bool done = false;
    while (!done) {
        if (atomicCAS(s.lock + t, 0, 1) == 0) {
            __threadfence_block();
            int c = *cnt;
            if (c < kb) {
                keys[c] = key;
                c++;
                if (c == kb) {
                    uint64_t m = keys[0];
                    for (int i = 1; i < kb; i++) { uint64_t v = keys[i]; m = v < m ? v : m; }
                    *mink = m;
                }
                *cnt = c;
            } else {
                s.ovf[t] = 1;
                if (key > *mink) {
                    int mi = 0;
                    uint64_t m = keys[0];
                    for (int i = 1; i < kb; i++) { uint64_t v = keys[i]; if (v < m) { m = v; mi = i; } }
                    keys[mi] = key;
                    m = keys[0];
                    for (int i = 1; i < kb; i++) { uint64_t v = keys[i]; m = v < m ? v : m; }
                    *mink = m;
                }
            }
            __threadfence_block();
            atomicExch(s.lock + t, 0);
            done = true;
        }
    }
}

template <int MODE>
__device__ __noinline__ void hmk_handle_hit(const HmkBulkArgs& a, const HmkTopkSmem& tk, int tl, int tg,
                                               int32_t i, int32_t id, int32_t score) {
    if (MODE == HMK_MODE_TOPK) {
        if (a.q_minid && id <= a.q_minid[tg]) return;
        uint32_t rk = a.tierank ? (uint32_t)a.tierank[id] : (uint32_t)id;
        hmk_topk_insert(tk, tl, a.kb, hmk_key_make(score, rk));
    } else if (MODE == HMK_MODE_EMIT) {
        unsigned int pos = atomicAdd(a.hit_count, 1u);
        if (pos < a.hit_cap) a.hits[pos] = make_int4(tg, i, score, 0);
    }
}

__device__ __forceinline__ int32_t hmk_lane_max(const uint32_t* acc, int nw, int lane16) {
    if (lane16) {
        uint32_t m = acc[0];
        for (int w = 1; w < nw; w++) m = __vmaxu2(m, acc[w]);
        uint32_t lo = m & 0xffffu, hi = m >> 16;
        return (int32_t)(lo > hi ? lo : hi);
    }
    uint32_t m = acc[0];
    for (int w = 1; w < nw; w++) m = __vmaxu4(m, acc[w]);
    uint32_t a = m & 0xffu, b = (m >> 8) & 0xffu, c = (m >> 16) & 0xffu, d = m >> 24;
    a = a > b ? a : b;
    c = c > d ? c : d;
    return (int32_t)(a > c ? a : c);
}

// common CTA prologue/epilogue for the top-k buffers
__device__ __forceinline__ void hmk_topk_init(const HmkTopkSmem& tk, int qn, int kb) {
    for (int i = threadIdx.x; i < qn; i += blockDim.x) { tk.cnt[i] = 0; tk.lock[i] = 0; tk.ovf[i] = 0; tk.minkey[i] = 0; }
}
__device__ __forceinline__ void hmk_topk_flush(const HmkBulkArgs& a, const HmkTopkSmem& tk, int q0, int qn, int stripe) {
    for (int t = threadIdx.x; t < qn; t += blockDim.x) {
        uint64_t* k = tk.key + (size_t)t * a.kb;
        int c = tk.cnt[t];
        for (int i = 1; i < c; i++) {   // insertion sort, descending
            uint64_t v = k[i];
            int j = i - 1;
            while (j >= 0 && k[j] < v) { k[j + 1] = k[j]; j--; }
            k[j + 1] = v;
        }
        size_t o = (size_t)stripe * a.nq + q0 + t;
        for (int i = 0; i < c; i++) a.tk_key[o * a.kb + i] = k[i];
        a.tk_cnt[o] = c;
        a.tk_ovf[o] = tk.ovf[t];
    }
}

// ---------------------------------------------------------------- fast bulk kernel
#define HMK_ROWB (HMK_NRES * 4)                    // bytes of one (h, j) row: 24 residues x u32

template <int NW, int MODE>
__global__ void __launch_bounds__(HMK_BULK_THREADS) hmk_bulk_fast(const __grid_constant__ HmkBulkArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t PWB = NW * HMK_MAXL1 * HMK_ROWB;   // bytes per profile (compile time)
    const int qtile = blockIdx.x % a.nqt, stripe = blockIdx.x / a.nqt;
    const int q0 = qtile * a.qt;
    const int qn = min(a.qt, a.nq - q0);
    if (qn <= 0) return;
    const int L = a.sc.L;
    size_t o = ((size_t)a.qt * PWB + 15) & ~(size_t)15;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + o);
    o += 16;
    HmkTopkSmem tk;
    tk.key = reinterpret_cast<uint64_t*>(smem_raw + o);  o += (size_t)a.qt * a.kb * 8;
    tk.minkey = reinterpret_cast<uint64_t*>(smem_raw + o); o += (size_t)a.qt * 8;
    tk.cnt = reinterpret_cast<int*>(smem_raw + o);  o += (size_t)a.qt * 4;
    tk.lock = reinterpret_cast<int*>(smem_raw + o); o += (size_t)a.qt * 4;
    tk.ovf = reinterpret_cast<int*>(smem_raw + o);

    // ---- stage the profile tile with the TMA bulk-copy engine
    if (threadIdx.x == 0) hmk_mbar_init(bar, 1);
    if (MODE == HMK_MODE_TOPK) hmk_topk_init(tk, qn, a.kb);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t total = (uint32_t)qn * PWB;
        hmk_mbar_expect_tx(bar, total);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.prof) + (size_t)q0 * PWB;
        uint32_t done = 0;
        while (done < total) {
            uint32_t n = min(total - done, 32768u);
            hmk_bulk_g2s(smem_raw + done, src + done, n, bar);
            done += n;
        }
    }
    hmk_mbar_wait(bar, 0);

    const int i_begin = stripe * a.chunk;
    const int i_end = min(a.ndb, i_begin + a.chunk);
    unsigned long long scored = 0;
    const uint32_t topmask = a.sc.lane16 ? 0x80008000u : 0x80808080u;
    const int32_t dec = a.sc.T - a.sc.half;
    const unsigned char* sbase = smem_raw;

    for (int i = i_begin + threadIdx.x; i < i_end; i += blockDim.x) {
        const int32_t id = a.db_ids ? a.db_ids[i] : a.db_begin + i;
        if (a.slot && a.slot[id] >= 0) continue;
        const uint64_t w = a.packed[id];
        // rowp[j] = &profile[0][h=0][j][residue_j]; later profiles/halves are immediates away
        const unsigned char* rowp[HMK_MAXL1];
#pragma unroll
        for (int j = 0; j < HMK_MAXL1; j++)
            rowp[j] = sbase + j * HMK_ROWB + (uint32_t)((w >> (5 * j)) & 31u) * 4u;
        scored += qn;

        auto score = [&](const int tu, uint32_t (&acc)[NW]) {
#pragma unroll
            for (int h = 0; h < NW; h++) acc[h] = 0;
#pragma unroll
            for (int j = 0; j < HMK_MAXL1; j++) {
                if (j < L) {
#pragma unroll
                    for (int h = 0; h < NW; h++)
                        acc[h] += *reinterpret_cast<const uint32_t*>(rowp[j] + tu * PWB + h * (HMK_MAXL1 * HMK_ROWB));
                }
            }
        };
        auto finish = [&](const int t, const uint32_t (&acc)[NW]) {
            uint32_t any = acc[0];
#pragma unroll
            for (int h = 1; h < NW; h++) any |= acc[h];
            if (MODE == HMK_MODE_DENSE) {
                a.dense[(size_t)(q0 + t) * a.dense_stride + i] = hmk_lane_max(acc, NW, a.sc.lane16) + dec;
            } else if (any & topmask) {
                hmk_handle_hit<MODE>(a, tk, t, q0 + t, i, id, hmk_lane_max(acc, NW, a.sc.lane16) + dec);
            }
        };

        int t = 0;
        for (; t + 4 <= qn; t += 4) {
            uint32_t a0[NW], a1[NW], a2[NW], a3[NW];
            score(0, a0); score(1, a1); score(2, a2); score(3, a3);
            finish(t, a0); finish(t + 1, a1); finish(t + 2, a2); finish(t + 3, a3);
#pragma unroll
            for (int j = 0; j < HMK_MAXL1; j++) rowp[j] += 4 * PWB;
        }
        for (; t < qn; t++) {
            uint32_t a0[NW];
            score(0, a0);
            finish(t, a0);
#pragma unroll
            for (int j = 0; j < HMK_MAXL1; j++) rowp[j] += PWB;
        }
    }
    if (a.pair_counter) {
        for (int s = 16; s > 0; s >>= 1) scored += __shfl_xor_sync(0xffffffffu, scored, s);
        if ((threadIdx.x & 31) == 0 && scored) atomicAdd(a.pair_counter, scored);
    }
    if (MODE == HMK_MODE_TOPK) {
        __syncthreads();
        hmk_topk_flush(a, tk, q0, qn, stripe);
    }
}

// ---------------------------------------------------------------- generic bulk kernel
// Any lengths, any int32 matrix, any penalty: scalar Java-int arithmetic per pair.  Same
// outputs as the fast kernel.  prof_ids[t] are the profile-side sequence ids; `prof_is_query`
// says whether they are the reference's seq2 (query) or seq1 (member).
struct HmkGenericArgs {
    HmkBulkArgs b;
    const int32_t* prof_ids;
    int32_t prof_is_query;
    const uint8_t* res;
    const int32_t* off;
    const int32_t* M;
    int32_t maxlen;
};

template <int MODE>
__global__ void __launch_bounds__(256) hmk_bulk_generic(const __grid_constant__ HmkGenericArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const HmkBulkArgs& a = g.b;
    const int qtile = blockIdx.x % a.nqt, stripe = blockIdx.x / a.nqt;
    const int q0 = qtile * a.qt;
    const int qn = min(a.qt, a.nq - q0);
    if (qn <= 0) return;
    int32_t* sM = reinterpret_cast<int32_t*>(smem_raw);
    size_t o = HMK_NRES * HMK_NRES * 4;
    int32_t* slen = reinterpret_cast<int32_t*>(smem_raw + o); o += (size_t)a.qt * 4;
    HmkTopkSmem tk;
    tk.key = reinterpret_cast<uint64_t*>(smem_raw + ((o + 7) & ~(size_t)7)); o = ((o + 7) & ~(size_t)7) + (size_t)a.qt * a.kb * 8;
    tk.minkey = reinterpret_cast<uint64_t*>(smem_raw + o); o += (size_t)a.qt * 8;
    tk.cnt = reinterpret_cast<int*>(smem_raw + o);  o += (size_t)a.qt * 4;
    tk.lock = reinterpret_cast<int*>(smem_raw + o); o += (size_t)a.qt * 4;
    tk.ovf = reinterpret_cast<int*>(smem_raw + o);  o += (size_t)a.qt * 4;
    uint8_t* sres = smem_raw + o;   // [qt][maxlen]
    for (int i = threadIdx.x; i < HMK_NRES * HMK_NRES; i += blockDim.x) sM[i] = g.M[i];
    for (int t = threadIdx.x; t < qn; t += blockDim.x) {
        int32_t id = g.prof_ids[q0 + t];
        int len = g.off[id + 1] - g.off[id];
        slen[t] = len;
        for (int j = 0; j < len; j++) sres[(size_t)t * g.maxlen + j] = g.res[g.off[id] + j];
    }
    if (MODE == HMK_MODE_TOPK) hmk_topk_init(tk, qn, a.kb);
    __syncthreads();
    const int i_begin = stripe * a.chunk;
    const int i_end = min(a.ndb, i_begin + a.chunk);
    unsigned long long scored = 0;
    for (int i = i_begin + threadIdx.x; i < i_end; i += blockDim.x) {
        const int32_t id = a.db_ids ? a.db_ids[i] : a.db_begin + i;
        if (a.slot && a.slot[id] >= 0) continue;
        const uint8_t* dres = g.res + g.off[id];
        const int dlen = g.off[id + 1] - g.off[id];
        scored += qn;
        for (int t = 0; t < qn; t++) {
            const uint8_t* pres = sres + (size_t)t * g.maxlen;
            int32_t s = g.prof_is_query ? hmk_pair_score(dres, dlen, pres, slen[t], sM, a.sc.X, a.sc.P)
                                        : hmk_pair_score(pres, slen[t], dres, dlen, sM, a.sc.X, a.sc.P);
            if (MODE == HMK_MODE_DENSE) a.dense[(size_t)(q0 + t) * a.dense_stride + i] = s;
            else if (s >= a.sc.T) hmk_handle_hit<MODE>(a, tk, t, q0 + t, i, id, s);
        }
    }
    if (a.pair_counter) {
        for (int s = 16; s > 0; s >>= 1) scored += __shfl_xor_sync(0xffffffffu, scored, s);
        if ((threadIdx.x & 31) == 0 && scored) atomicAdd(a.pair_counter, scored);
    }
    if (MODE == HMK_MODE_TOPK) {
        __syncthreads();
        hmk_topk_flush(a, tk, q0, qn, stripe);
    }
}

// ---------------------------------------------------------------- top-k merge across stripes
// one warp per query: kb rounds of "largest key below the previous pick" (keys are distinct)
__global__ void hmk_topk_merge(int nq, int nstripes, int kb, const uint64_t* __restrict__ tk_key,
                               const int32_t* __restrict__ tk_cnt, const int32_t* __restrict__ tk_ovf,
                               uint64_t* __restrict__ out_key, int32_t* __restrict__ out_cnt,
                               int32_t* __restrict__ out_ovf) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= nq) return;
    int total = 0, ovf = 0;
    for (int s = lane; s < nstripes; s += 32) {
        total += tk_cnt[(size_t)s * nq + t];
        ovf |= tk_ovf[(size_t)s * nq + t];
    }
    for (int s = 16; s > 0; s >>= 1) { total += __shfl_xor_sync(0xffffffffu, total, s); ovf |= __shfl_xor_sync(0xffffffffu, ovf, s); }
    uint64_t prev = ~0ull;
    const int take = total < kb ? total : kb;
    for (int r = 0; r < take; r++) {
        uint64_t best = 0;
        for (int e = lane; e < nstripes * kb; e += 32) {
            const int s = e / kb, i = e % kb;
            if (i < tk_cnt[(size_t)s * nq + t]) {
                uint64_t v = tk_key[((size_t)s * nq + t) * kb + i];
                if (v < prev && v > best) best = v;
            }
        }
        for (int s = 16; s > 0; s >>= 1) { uint64_t v = __shfl_xor_sync(0xffffffffu, best, s); best = v > best ? v : best; }
        if (lane == 0) out_key[(size_t)t * kb + r] = best;
        prev = best;
    }
    if (lane == 0) { out_cnt[t] = take; out_ovf[t] = (ovf || total > kb) ? 1 : 0; }
}

// ---------------------------------------------------------------- batch selection (phase 1)
// single CTA: the next `want` ids >= cur that are still singletons, ascending
__global__ void hmk_select_queries(const int32_t* __restrict__ slot, int n, const HmkCtl* ctl, int want,
                                   int32_t* __restrict__ qid, int32_t* __restrict__ nq_out) {
    __shared__ int warp_cnt[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = ctl->cur; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool f = i < n && slot[i] < 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w2 = 0; w2 < nwarps; w2++) { int c = warp_cnt[w2]; if (w2 < wid) pre += c; tot += c; }
        const int pos = base + pre + __popc(m & ((1u << lane) - 1u));
        if (f && pos < want) qid[pos] = i;
        __syncthreads();
        if (threadIdx.x == 0) base += tot;
        __syncthreads();
        if (base >= want) break;
    }
    if (threadIdx.x == 0) *nq_out = base < want ? base : want;
}

// ---------------------------------------------------------------- member check (complete linkage)
// One thread per founder hit: does the query also score >= T against every other current
// member of that cluster?  (ClinkageClusterScorer.java:30-49; early exit keeps it cheap.)
struct HmkCheckArgs {
    HmkState S;
    const int4* hits;
    const unsigned int* hit_count;
    unsigned int hit_cap;
    int32_t hit_t_is_query;      // 1: hit.x = query index, hit.y = cluster slot; 0: the reverse
    const int32_t* qids;         // query index -> sequence id
    // phase 1 output: linked lists per query
    int32_t* ac_head; int32_t* ac_next; int32_t* ac_slot; int32_t* ac_score;
    // phase 2 output: flat candidate arrays
    unsigned long long* cand_key_q;   // (query index << 32) | slot
    unsigned long long* cand_key_c;   // (slot << 32) | query index
    int32_t* cand_score;
    unsigned int* cand_count;
    unsigned int cand_cap;
    int32_t linked;
    int32_t q_index_offset;      // added to the query index in the flat candidate keys
};

__global__ void hmk_member_check(const HmkCheckArgs a) {
    const unsigned int nh = min(*a.hit_count, a.hit_cap);
    long long npairs = 0;
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < nh; e += gridDim.x * blockDim.x) {
        const int4 h = a.hits[e];
        const int qi = a.hit_t_is_query ? h.x : h.y, c = a.hit_t_is_query ? h.y : h.x;
        const int32_t q = a.qids[qi];
        int32_t cl = h.z;
        bool ok = true;
        for (int32_t m = a.S.next[a.S.c_founder[c]]; m >= 0; m = a.S.next[m]) {
            int32_t s = hmk_state_score(a.S, m, q);
            npairs++;
            if (s < cl) cl = s;
            if (s < a.S.T) { ok = false; break; }
        }
        if (!ok) continue;
        unsigned int pos = atomicAdd(a.cand_count, 1u);
        if (pos >= a.cand_cap) continue;
        if (a.linked) {
            a.ac_slot[pos] = c;
            a.ac_score[pos] = cl;
            a.ac_next[pos] = atomicExch(a.ac_head + qi, (int32_t)pos);
        } else {
            const uint32_t gq = (uint32_t)(qi + a.q_index_offset);
            a.cand_key_q[pos] = ((unsigned long long)gq << 32) | (uint32_t)c;
            a.cand_key_c[pos] = ((unsigned long long)(uint32_t)c << 32) | gq;
            a.cand_score[pos] = cl;
        }
    }
    for (int s = 16; s > 0; s >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, s);
    if ((threadIdx.x & 31) == 0 && npairs) atomicAdd((unsigned long long*)&a.S.ctl->scalar_pairs, (unsigned long long)npairs);
}

// ---------------------------------------------------------------- resolvers
__global__ void hmk_p1_resolve_kernel(const HmkState S, const HmkP1Batch B) {
    HmkWarp w;
    hmk_p1_resolve(S, B, w);
}

__global__ void hmk_p2_round_kernel(const HmkState S, const HmkP2 P) {
    HmkWarp w;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c < P.ncl) hmk_p2_round(S, P, w, c);
}

// ---------------------------------------------------------------- small utilities
__global__ void hmk_fill_i32(int32_t* p, int32_t v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void hmk_pack_sequences(int n, const uint8_t* __restrict__ res, const int32_t* __restrict__ off,
                                   uint64_t* __restrict__ packed, int32_t* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int len = off[i + 1] - off[i];
    uint64_t w = 0;
    for (int j = 0; j < len; j++) {
        uint32_t r = res[off[i] + j];
        if (r >= HMK_NRES) { atomicExch(bad, 1); r = 0; }
        if (j < HMK_MAXL1) w |= (uint64_t)r << (5 * j);
    }
    packed[i] = w;
}

// flags -> exclusive positions, three-step scan (per-block counts, single-block scan, scatter)
__global__ void hmk_count_unassigned(const int32_t* __restrict__ slot, int n, int32_t* __restrict__ block_cnt) {
    __shared__ int wsum[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = i < n && slot[i] < 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += wsum[w];
        block_cnt[blockIdx.x] = t;
    }
}
__global__ void hmk_scan_blocks(int32_t* block_cnt, int nblocks, int32_t* total) {
    // single thread block, sequential chunks (nblocks is n/1024: tiny)
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nblocks; b++) { int c = block_cnt[b]; block_cnt[b] = run; run += c; }
        *total = run;
    }
}
__global__ void hmk_scatter_unassigned(const int32_t* __restrict__ slot, int n, const int32_t* __restrict__ block_off,
                                       int32_t* __restrict__ out) {
    __shared__ int wsum[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = i < n && slot[i] < 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wsum[wid] = __popc(m);
    __syncthreads();
    int pre = 0;
    for (int w = 0; w < wid; w++) pre += wsum[w];
    if (f) out[block_off[blockIdx.x] + pre + __popc(m & ((1u << lane) - 1u))] = i;
}

// first index whose key's high half is >= s, for s = 0..nseg (start[nseg] = n)
__global__ void hmk_segment_starts(const unsigned long long* __restrict__ keys, int n, int nseg, int32_t* __restrict__ start) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > nseg) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> 32) < (int64_t)s) lo = mid + 1; else hi = mid;
    }
    start[s] = lo;
}
__global__ void hmk_split_keys_lo(const unsigned long long* __restrict__ keys, int n, int32_t* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lo[i] = (int32_t)(uint32_t)keys[i];
}

__global__ void hmk_finalize(int n, const int32_t* __restrict__ slot, const int32_t* __restrict__ rank,
                             const int32_t* __restrict__ c_founder, int32_t* __restrict__ cluster_id,
                             int32_t* __restrict__ member_rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slot[i];
    cluster_id[i] = s >= 0 ? c_founder[s] : i;
    member_rank[i] = s >= 0 ? rank[i] : 0;
}
