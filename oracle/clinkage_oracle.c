/*
 * clinkage_oracle.c -- CPU ORACLE (test infrastructure, see hammock_oracle.h) for Hammock's EXACT complete-linkage
 * initial clustering, SURVEY.md 8(f) N1: ClinkageSequenceClusterer.cluster (nearest-neighbour chain), the default
 * initial stage for <= 10 000 unique sequences (Hammock.java:371-373).  Paths are relative to
 * src/cz/krejciadam/hammock/ of /root/reference.
 *
 * PARITY UNPINNED: written from the Java source; the reference cannot be run here (no JVM).
 *
 * Two things make the reference's result depend on more than the scores:
 *   * the chain is started from  activeClusters.iterator().next()  (ClinkageSequenceClusterer.java:70) and the returned
 *     list is  new ArrayList(readyClusters)  (:119-123); both are java.util.HashSet<Cluster>, so the result depends on
 *     HashMap's bucket order.  This file emulates the OpenJDK 8+ HashMap (hash spreading h ^ h>>>16, power-of-two table,
 *     append at the bin's tail, order-preserving split on resize, default capacity 16 / load factor 0.75) with
 *     Cluster.hashCode() = 79 * 7 + id (Cluster.java:179-183).  Bins never reach the tree threshold (8) here: the hash
 *     codes are consecutive integers.  (A Java 7 runtime iterates differently; the jar's manifest does not pin one.)
 *   * CachedClusterScorer (CachedClusterScorer.java:38-125) keeps ONE value per unordered pair of clusters, whichever
 *     argument order computed it first; the values it merges are minima, so for a symmetric substitution matrix -- every
 *     matrix shipped in matrices/ -- it returns exactly ClinkageClusterScorer.clusterScore.  Asymmetric matrices are
 *     rejected here (status HMKO_ERR_ASYMMETRIC): the reference's result would depend on the thread schedule.
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "hammock_oracle.h"

#define JMIN INT32_MIN
#define BELOW (INT32_MIN + 1) /* ClinkageClusterScorer.java:42 */

static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* ---------------------------------------------------------------- java.util.HashSet<Cluster>, OpenJDK 8+ */
typedef struct {
    int32_t cap, size, thr; /* cap == 0: table not allocated yet (HashMap.resize on first put) */
    int32_t* head;          /* [cap] first key of the bin, -1 = empty */
    int32_t* tail;          /* [cap] */
    int32_t* next;          /* [maxkey + 1] chain links, by key */
    int32_t* prev;
    uint8_t* in;            /* [maxkey + 1] membership */
    int32_t maxkey;
    int treeified;          /* a bin reached 8 entries: emulation no longer exact */
} jset;

static uint32_t cluster_hash(int32_t id) { /* Cluster.hashCode (Cluster.java:179-183) + HashMap.hash */
    uint32_t h = (uint32_t)(79 * 7 + id);
    return h ^ (h >> 16);
}

static void jset_init(jset* s, int32_t maxkey) {
    memset(s, 0, sizeof *s);
    s->maxkey = maxkey;
    s->next = (int32_t*)malloc(sizeof(int32_t) * (size_t)(maxkey + 1));
    s->prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)(maxkey + 1));
    s->in = (uint8_t*)calloc((size_t)(maxkey + 1), 1);
}
static void jset_free(jset* s) { free(s->head); free(s->tail); free(s->next); free(s->prev); free(s->in); }

static void jset_link_tail(jset* s, int32_t b, int32_t key) {
    s->next[key] = -1;
    s->prev[key] = s->tail[b];
    if (s->tail[b] >= 0) s->next[s->tail[b]] = key; else s->head[b] = key;
    s->tail[b] = key;
}

static void jset_resize(jset* s) { /* HashMap.resize: lo / hi split keeps the relative order of a bin */
    const int32_t oldcap = s->cap, newcap = oldcap ? oldcap * 2 : 16;
    int32_t* oh = s->head;
    int32_t* ot = s->tail;
    s->head = (int32_t*)malloc(sizeof(int32_t) * (size_t)newcap);
    s->tail = (int32_t*)malloc(sizeof(int32_t) * (size_t)newcap);
    for (int32_t i = 0; i < newcap; i++) s->head[i] = s->tail[i] = -1;
    s->cap = newcap;
    s->thr = (int32_t)(newcap * 0.75f);
    for (int32_t b = 0; b < oldcap; b++)
        for (int32_t k = oh[b], nx; k >= 0; k = nx) {
            nx = s->next[k];
            jset_link_tail(s, (int32_t)(cluster_hash(k) & (uint32_t)(newcap - 1)), k);
        }
    free(oh);
    free(ot);
}

static void jset_add(jset* s, int32_t key) { /* HashMap.putVal */
    if (s->cap == 0) jset_resize(s);
    if (s->in[key]) return;
    const int32_t b = (int32_t)(cluster_hash(key) & (uint32_t)(s->cap - 1));
    int chain = 0;
    for (int32_t k = s->head[b]; k >= 0; k = s->next[k]) chain++;
    if (chain >= 7) s->treeified = 1; /* TREEIFY_THRESHOLD - 1 */
    jset_link_tail(s, b, key);
    s->in[key] = 1;
    if (++s->size > s->thr) jset_resize(s);
}

static void jset_remove(jset* s, int32_t key) {
    if (!s->in[key]) return;
    const int32_t b = (int32_t)(cluster_hash(key) & (uint32_t)(s->cap - 1));
    if (s->prev[key] >= 0) s->next[s->prev[key]] = s->next[key]; else s->head[b] = s->next[key];
    if (s->next[key] >= 0) s->prev[s->next[key]] = s->prev[key]; else s->tail[b] = s->prev[key];
    s->in[key] = 0;
    s->size--;
}

static int32_t jset_first(const jset* s) { /* iterator().next() */
    for (int32_t b = 0; b < s->cap; b++)
        if (s->head[b] >= 0) return s->head[b];
    return -1;
}

/* ---------------------------------------------------------------- the clusterer */
int hmko_clinkage_cluster(int32_t n, const uint8_t* residues, const int32_t* offsets, const int32_t* abundance,
                          const int32_t* M, int32_t T, int32_t X, int32_t P, int32_t* cluster_id, int32_t* member_rank,
                          int32_t* result_order, int32_t* n_result, int64_t* nearest_searches) {
    *n_result = 0;
    if (nearest_searches) *nearest_searches = 0;
    if (n <= 0) return HMKO_ERR_EMPTY; /* activeClusters.iterator().next() on an empty set (:116): NoSuchElementException */
    for (int32_t i = 0; i < n; i++)
        for (int32_t p = offsets[i]; p < offsets[i + 1]; p++)
            if (residues[p] >= HMKO_NRES) return HMKO_ERR_BAD_RESIDUE;
    for (int a = 0; a < HMKO_NRES; a++)
        for (int b = 0; b < a; b++)
            if (M[a * HMKO_NRES + b] != M[b * HMKO_NRES + a]) return HMKO_ERR_ASYMMETRIC;
    if (n >= 2) /* the first nearest-neighbour search scores the chain start against everything (ShiftedScorer.java:59-62) */
        for (int32_t i = 0; i < n; i++)
            if (offsets[i + 1] - offsets[i] <= X) return HMKO_ERR_SHIFT_TOO_BIG;

    /* D[a][b]: ClinkageClusterScorer.clusterScore of the clusters in slots a and b (ClinkageClusterScorer.java:30-49):
     * min over the member pairs, BELOW as soon as one pair scores < T.  Slot i starts as the cluster {sequence i}. */
    int32_t* D = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * (size_t)n);
#pragma omp parallel for schedule(dynamic, 16)
    for (int32_t a = 0; a < n; a++)
        for (int32_t b = 0; b <= a; b++) {
            int st = 0;
            int32_t s = hmko_score_with_shift(residues + offsets[a], offsets[a + 1] - offsets[a], residues + offsets[b],
                                              offsets[b + 1] - offsets[b], M, X, P, NULL, &st);
            if (s < T) s = BELOW;
            D[(size_t)a * n + b] = D[(size_t)b * n + a] = s;
        }
    const int32_t maxid = 2 * n + 2;
    int32_t* slot_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)(maxid + 1)); /* cluster id -> slot */
    int32_t* id_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* size_of = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);           /* Cluster.size(): abundance weighted */
    int32_t* mhead = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);             /* member list of the slot's cluster */
    int32_t* mtail = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* mnext = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));       /* cluster ids */
    int32_t* ready_order = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    jset active, ready;
    jset_init(&active, maxid);
    jset_init(&ready, maxid);
    int32_t current_id = 1;                                                               /* :48 */
    for (int32_t i = 0; i < n; i++) {                                                     /* :49-54 */
        id_of[i] = current_id; slot_of[current_id] = i; size_of[i] = abundance[i];
        mhead[i] = mtail[i] = i; mnext[i] = -1;
        jset_add(&active, current_id);
        current_id++;
    }
    int32_t sp = 0;
    while (active.size > 1) {                                                             /* :62 */
        stack[sp++] = jset_first(&active);                                                /* :69-70 */
        while (sp > 0) {                                                                  /* :71 */
            const int32_t top = stack[sp - 1], ts = slot_of[top];
            /* findNearestClusterParallel + NearestClusterRunner (:137-177, 258-293): arg-max of the score over the
             * other active clusters, ties by Cluster.size() (larger), then id (smaller) */
            int32_t best = -1, bscore = JMIN;
            if (nearest_searches) (*nearest_searches)++;
            for (int32_t b = 0; b < active.cap; b++)
                for (int32_t k = active.head[b]; k >= 0; k = active.next[k]) {
                    if (k == top) continue;
                    const int32_t ks = slot_of[k], s = D[(size_t)ts * n + ks];
                    int take;
                    if (best < 0) take = 1;
                    else if (s != bscore) take = s > bscore;
                    else if (size_of[ks] != size_of[slot_of[best]]) take = size_of[ks] > size_of[slot_of[best]];
                    else take = k < best;
                    if (take) { best = k; bscore = s; }
                }
            if (best < 0) bscore = JMIN;                                                  /* :77-82 (null result) */
            if (bscore < T) {                                                             /* :85-91 */
                sp--;
                jset_add(&ready, top);
                jset_remove(&active, top);
                continue;
            }
            if (sp > 1 && stack[sp - 2] == best) {                                        /* :95-109 */
                current_id++;
                sp -= 2;
                jset_remove(&active, top);
                jset_remove(&active, best);
                const int32_t bs = slot_of[best];
                /* CachedClusterScorer.join (:96-125): the merged cluster's row is the element-wise minimum */
                for (int32_t k = 0; k < n; k++) {
                    const int32_t a = D[(size_t)ts * n + k], b2 = D[(size_t)bs * n + k], m = a < b2 ? a : b2;
                    D[(size_t)ts * n + k] = D[(size_t)k * n + ts] = m;
                }
                /* new Cluster(top.getSequences() ++ nearest.getSequences(), currentId) (:104-106) */
                mnext[mtail[ts]] = mhead[bs];
                mtail[ts] = mtail[bs];
                size_of[ts] = wadd(size_of[ts], size_of[bs]);
                id_of[ts] = current_id;
                slot_of[current_id] = ts;
                jset_add(&active, current_id);
            } else {
                stack[sp++] = best;                                                       /* :111 */
            }
        }
    }
    jset_add(&ready, jset_first(&active));                                                /* :116 */
    /* new ArrayList(readyClusters) (:119-123): HashSet iteration order */
    int32_t nr = 0;
    for (int32_t b = 0; b < ready.cap; b++)
        for (int32_t k = ready.head[b]; k >= 0; k = ready.next[k]) ready_order[nr++] = k;
    for (int32_t r = 0; r < nr; r++) {
        const int32_t s = slot_of[ready_order[r]];
        result_order[r] = ready_order[r];
        int32_t rank = 0;
        for (int32_t m = mhead[s]; m >= 0; m = mnext[m]) { cluster_id[m] = ready_order[r]; member_rank[m] = rank++; }
    }
    *n_result = nr;
    const int tre = active.treeified || ready.treeified;
    jset_free(&active); jset_free(&ready);
    free(D); free(slot_of); free(id_of); free(size_of); free(mhead); free(mtail); free(mnext); free(stack); free(ready_order);
    return tre ? HMKO_ERR_TREEIFIED : HMKO_OK;
}
