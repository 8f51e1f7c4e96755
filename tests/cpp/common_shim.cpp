// Test shim (CPU tests only): the host/device primitives of hammock_b200/csrc/hmk_common.h and hmk_resolve.h -- the scalar
// scorer the generic kernels run on the device, the partner key, the cluster arg-max order -- compiled with plain g++
// and exported with C linkage so that tests/test_host.py can compare them with the oracle.  Never linked into the product.
#include "../../hammock_b200/csrc/hmk_common.h"
#include "../../hammock_b200/csrc/hmk_resolve.h"

extern "C" {
int32_t shim_pair_score(const uint8_t* s1, int l1, const uint8_t* s2, int l2, const int32_t* M, int X, int P) {
    return hmk_pair_score(s1, l1, s2, l2, M, X, P);
}
int32_t shim_pair_score_strided(const uint8_t* s1, int st1, int l1, const uint8_t* s2, int st2, int l2, const int32_t* M, int X, int P) {
    return hmk_pair_score_strided(s1, st1, l1, s2, st2, l2, M, X, P);
}
int64_t shim_pair_cells(int l1, int l2, int X) { return hmk_pair_cells(l1, l2, X); }
uint64_t shim_key_make(int32_t score, uint32_t rank) { return hmk_key_make(score, rank); }
int32_t shim_key_score(uint64_t key) { return hmk_key_score(key); }
uint32_t shim_key_rank(uint64_t key) { return hmk_key_rank(key); }
int32_t shim_wadd(int32_t a, int32_t b) { return hmk_wadd(a, b); }
int32_t shim_wmul(int32_t a, int32_t b) { return hmk_wmul(a, b); }
// the one-at-a-time scorers on packed words (member checks, phase 2, mixed-length dense tables)
int32_t shim_packed_pair_score(uint64_t w1, uint64_t w2, const int32_t* M, int X, int P) { return hmk_packed_pair_score(w1, w2, M, X, P); }
int32_t shim_score12x3(uint64_t wquery, uint64_t wmember, const int32_t* M, int P) {
    int32_t qrow[HMK_MAXL1];
    hmk_qrow12(wquery, qrow);
    return hmk_score12x3(qrow, wmember, M, P);
}
// arg-max over n candidates under hmk_consider, fed in the given order; returns the winning slot (-1: none)
int32_t shim_best_cluster(int n, const int32_t* score, const int32_t* size, const int32_t* fid) {
    HmkBestCluster b;
    b.score = HMK_JMIN; b.size = 0; b.fid = 0; b.slot = -1;
    for (int i = 0; i < n; i++) hmk_consider(b, score[i], size[i], fid[i], i);
    return b.slot;
}
}
