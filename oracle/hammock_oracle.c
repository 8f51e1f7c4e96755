/*
 * hammock_oracle.c -- CPU ORACLE (test infrastructure, see hammock_oracle.h).
 *
 * Plain-C restatement of Hammock v1.2.0's greedy initial clustering path.  Written from
 * the semantics of the Java source; no code is copied.  Citations are file:line under
 * /root/reference/src/cz/krejciadam/hammock/.  PARITY UNPINNED (no JVM, no reference
 * goldens) -- pinned only by source-derived known answers and a second restatement.
 */
#include "hammock_oracle.h"

#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Java int arithmetic wraps; do every add in uint32 and reinterpret. */
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

static void set_err(char* err, size_t errlen, const char* msg) {
    if (err && errlen) {
        strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

/* ---------------------------------------------------------------- alphabet */

/* UniqueSequence.java:23-35,49-55: toUpperCase() then map lookup */
int hmko_encode_char(char c) {
    static const char* alpha = HMKO_ALPHABET;
    char u = (char)toupper((unsigned char)c);
    for (int i = 0; i < HMKO_NRES; i++)
        if (alpha[i] == u) return i;
    return -1;
}

int hmko_encode(const char* s, uint8_t* out) {
    for (size_t i = 0; s[i]; i++) {
        int c = hmko_encode_char(s[i]);
        if (c < 0) return HMKO_ERR_BAD_RESIDUE;
        out[i] = (uint8_t)c;
    }
    return HMKO_OK;
}

/* ---------------------------------------------------------------- matrix loader */

static int java_ws(int c) { return c == ' ' || c == '\t' || c == '\n' || c == 0x0B || c == '\f' || c == '\r'; }

/* Integer.parseInt: optional sign, decimal digits, must fit int32 */
static int parse_int_java(const char* tok, size_t len, int32_t* out) {
    size_t i = 0;
    int neg = 0;
    if (len == 0) return -1;
    if (tok[0] == '-') { neg = 1; i = 1; } else if (tok[0] == '+') { i = 1; }
    if (i >= len) return -1;
    int64_t v = 0;
    for (; i < len; i++) {
        if (tok[i] < '0' || tok[i] > '9') return -1;
        v = v * 10 + (tok[i] - '0');
        if (v > (int64_t)INT32_MAX + 1) return -1;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return -1;
    *out = (int32_t)v;
    return 0;
}

/* FileIOManager.java:46-81.  Quirks kept: rows are taken in FILE ORDER with no label check
 * (the header test at :53-58 compares against the literal two characters "\s" and can never
 * fire); lines starting with '#', ' ' or TAB are skipped (:60); every other line -- also an
 * empty one -- must split into exactly 25 whitespace-separated tokens (:61-64); a 25th data
 * row overflows the 24x24 array -> FileFormatException (:71-79); fewer rows leave zeros. */
int hmko_load_matrix(const char* path, int32_t* M, char* err, size_t errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { set_err(err, errlen, "cannot open matrix file"); return HMKO_ERR_IO; }
    memset(M, 0, sizeof(int32_t) * HMKO_NRES * HMKO_NRES);
    char* line = NULL;
    size_t cap = 0;
    ssize_t got;
    int row = 0, rc = HMKO_OK;
    while ((got = getline(&line, &cap, f)) >= 0) {
        size_t len = (size_t)got;
        /* BufferedReader.readLine strips \n, \r or \r\n */
        if (len && line[len - 1] == '\n') len--;
        if (len && line[len - 1] == '\r') len--;
        line[len] = 0;
        if (len && (line[0] == '#' || line[0] == ' ' || line[0] == '\t')) continue;
        /* String.split("\\s+"): leading empty token kept when the line starts with
         * whitespace (cannot happen past the skip above except for VT/FF), trailing
         * empty tokens dropped. */
        const char* toks[64];
        size_t tlen[64];
        int nt = 0;
        size_t i = 0;
        if (len == 0) { toks[0] = line; tlen[0] = 0; nt = 1; }
        else {
            if (java_ws((unsigned char)line[0])) { toks[0] = line; tlen[0] = 0; nt = 1; }
            while (i < len) {
                while (i < len && java_ws((unsigned char)line[i])) i++;
                if (i >= len) break;
                size_t s = i;
                while (i < len && !java_ws((unsigned char)line[i])) i++;
                if (nt < 64) { toks[nt] = line + s; tlen[nt] = i - s; }
                nt++;
            }
        }
        if (nt != 25) {
            set_err(err, errlen, "Scoring matrix should always have 24 columns (plus 1 column describing AAs).");
            rc = HMKO_ERR_FILE_FORMAT;
            break;
        }
        if (row >= HMKO_NRES) { /* ArrayIndexOutOfBounds -> FileFormatException (:75-78) */
            set_err(err, errlen, "Scoring matrix should always have 24 rows (plus 1 column describing AAs).");
            rc = HMKO_ERR_FILE_FORMAT;
            break;
        }
        for (int c = 1; c < 25; c++) {
            int32_t v;
            if (parse_int_java(toks[c], tlen[c], &v)) {
                set_err(err, errlen, "NumberFormatException in scoring matrix");
                rc = HMKO_ERR_FILE_FORMAT;
                break;
            }
            M[row * HMKO_NRES + (c - 1)] = v;
        }
        if (rc) break;
        row++;
    }
    free(line);
    fclose(f);
    return rc;
}

/* ---------------------------------------------------------------- pair score */

/* ShiftedScorer.java:48-95 -- literal loop structure */
int32_t hmko_score_with_shift(const uint8_t* seq1, int len1, const uint8_t* seq2, int len2,
                              const int32_t* M, int max_shift, int P, int* shift, int* status) {
    const uint8_t *s, *l;
    int ls, ll, shorter_is_seq2;
    if (len1 >= len2) { s = seq2; ls = len2; l = seq1; ll = len1; shorter_is_seq2 = 1; }   /* :51-53 */
    else              { s = seq1; ls = len1; l = seq2; ll = len2; shorter_is_seq2 = 0; }   /* :54-57 */
    if (max_shift >= ls) {                                                                /* :59-62 */
        if (status) *status = HMKO_ERR_SHIFT_TOO_BIG;
        return 0;
    }
    int32_t best = INT32_MIN;
    int best_shift = 0;
    int d = ll - ls;                                                                      /* :66 */
    for (int k = -max_shift; k <= max_shift + d; k++) {                                   /* :67 */
        int32_t sc = 0;
        if (k <= 0) {                                                                     /* :69-72 */
            for (int i = 0; i < ls + k; i++) sc = wrap_add(sc, M[s[i - k] * HMKO_NRES + l[i]]);
        } else {                                                                          /* :73-77 */
            int lim = ls < ll - k ? ls : ll - k;
            for (int i = 0; i < lim; i++) sc = wrap_add(sc, M[s[i] * HMKO_NRES + l[i + k]]);
        }
        sc = wrap_add(sc, wrap_mul(d, P));                                                /* :79 */
        if (k < 0) sc = wrap_add(sc, wrap_mul(wrap_mul(-k, 2), P));                       /* :80-82 */
        if (k > d) sc = wrap_add(sc, wrap_mul(wrap_mul(k - d, 2), P));                    /* :83-85 */
        if (sc > best) { best = sc; best_shift = k; }                                     /* :86-89 */
    }
    if (!shorter_is_seq2) best_shift = -best_shift;                                       /* :91-93 */
    if (shift) *shift = best_shift;
    return best;
}

int hmko_pair_shifts(int len1, int len2, int X) {
    int d = len1 > len2 ? len1 - len2 : len2 - len1;
    int n = 2 * X + d + 1;
    return n > 0 ? n : 0;
}

int64_t hmko_pair_cells(int len1, int len2, int X) {
    int ls = len1 < len2 ? len1 : len2, ll = len1 < len2 ? len2 : len1;
    int d = ll - ls;
    int64_t cells = 0;
    for (int k = -X; k <= X + d; k++) {
        int c;
        if (k <= 0) c = ls + k;
        else c = ls < ll - k ? ls : ll - k;
        if (c > 0) cells += c;
    }
    return cells;
}

/* ---------------------------------------------------------------- ordering + defaults */

typedef struct {
    const uint8_t* residues;
    const int32_t* offsets;
    const int32_t* abundance;
} sort_ctx;

/* UniqueSequenceSizeAlphabeticComparator (UniqueSequence.java:238-248): size, then
 * String.compareTo of the upper-case strings.  getSequenceString() rebuilds the string
 * from codes (:103-109), so comparing alphabet letters of the codes is the same thing. */
static int cmp_size_alpha(const sort_ctx* c, int32_t a, int32_t b) {
    static const char* alpha = HMKO_ALPHABET;
    int32_t sa = c->abundance[a], sb = c->abundance[b];
    if (sa != sb) return sa < sb ? -1 : 1; /* SizeComparator.java:15-20 (no overflow for sizes >= 1) */
    int la = c->offsets[a + 1] - c->offsets[a], lb = c->offsets[b + 1] - c->offsets[b];
    const uint8_t* pa = c->residues + c->offsets[a];
    const uint8_t* pb = c->residues + c->offsets[b];
    int lim = la < lb ? la : lb;
    for (int i = 0; i < lim; i++) {
        int ca = alpha[pa[i]], cb = alpha[pb[i]];
        if (ca != cb) return ca - cb;
    }
    return la - lb;
}

/* Collections.sort = stable merge sort; reverseOrder(cmp) flips the sign (:180). */
static void merge_sort(const sort_ctx* c, int32_t* a, int32_t* tmp, int32_t lo, int32_t hi) {
    if (hi - lo < 2) return;
    int32_t mid = lo + (hi - lo) / 2;
    merge_sort(c, a, tmp, lo, mid);
    merge_sort(c, a, tmp, mid, hi);
    int32_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) {
        /* take from the right run only if it is strictly "smaller" under the reversed order */
        if (-cmp_size_alpha(c, a[j], a[i]) < 0) tmp[k++] = a[j++];
        else tmp[k++] = a[i++];
    }
    while (i < mid) tmp[k++] = a[i++];
    while (j < hi) tmp[k++] = a[j++];
    memcpy(a + lo, tmp + lo, sizeof(int32_t) * (size_t)(hi - lo));
}

void hmko_sort_order_size(int32_t n, const uint8_t* residues, const int32_t* offsets,
                          const int32_t* abundance, int32_t* perm) {
    sort_ctx c = {residues, offsets, abundance};
    for (int32_t i = 0; i < n; i++) perm[i] = i;
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    merge_sort(&c, perm, tmp, 0, n);
    free(tmp);
}

/* Math.round(double) = floor(x + 0.5) as long */
static int32_t java_round(double x) {
    double f = x + 0.5;
    int64_t r = (int64_t)f;
    if ((double)r > f) r--; /* floor for negatives */
    return (int32_t)r;
}

int32_t hmko_check_max_shift(int32_t n, const int32_t* offsets, int32_t max_shift) {
    int32_t min_len = INT32_MAX;                                   /* Hammock.java:1421-1427 */
    for (int32_t i = 0; i < n; i++) {
        int32_t l = offsets[i + 1] - offsets[i];
        if (l < min_len) min_len = l;
    }
    return max_shift < min_len - 1 ? max_shift : min_len - 1;
}

void hmko_default_params(int32_t n, const int32_t* offsets, int32_t* threshold,
                         int32_t* max_shift, int32_t* max_clusters) {
    int32_t count = 0, length_sum = 0;                             /* Hammock.java:1554-1563 */
    for (int32_t i = 0; i < n; i++) { count++; length_sum += offsets[i + 1] - offsets[i]; }
    double mean = ((double)length_sum) / ((double)count);
    if (threshold) *threshold = java_round(mean * 1.7);           /* :1409-1413 */
    if (max_shift) *max_shift = hmko_check_max_shift(n, offsets, java_round(mean / 4)); /* :1429-1434 */
    if (max_clusters) *max_clusters = java_round(n * 0.025);      /* :398-401 */
}

/* ---------------------------------------------------------------- clustering */

#define JMIN INT32_MIN

typedef struct {
    int32_t id;       /* Cluster.getId()  (Cluster.java:121-123)                    */
    int32_t size;     /* abundance-weighted Cluster.size() (Cluster.java:156-158)   */
    int32_t n, cap;   /* getUniqueSize()                                            */
    int32_t* members; /* getSequences(), insertion order                            */
} ocluster;

typedef struct {
    int32_t n;
    const uint8_t* residues;
    const int32_t* offsets;
    const int32_t* abundance;
    const int32_t* M;
    int32_t T, X, P;
    int nthreads;
    int status;
} octx;

static void cl_insert(ocluster* c, int32_t seq, const octx* x) {   /* Cluster.java:50-63 */
    if (c->n == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 4;
        c->members = (int32_t*)realloc(c->members, sizeof(int32_t) * (size_t)c->cap);
    }
    c->members[c->n++] = seq;
    c->size = wrap_add(c->size, x->abundance[seq]);
}

static inline int32_t seq_score(const octx* x, int32_t member, int32_t query, int* status) {
    /* ClinkageClusterScorer.java:38: scorer.sequenceScore(seq1 = member of the database
     * cluster, seq2 = member of the compared cluster) */
    return hmko_score_with_shift(x->residues + x->offsets[member], x->offsets[member + 1] - x->offsets[member],
                                 x->residues + x->offsets[query], x->offsets[query + 1] - x->offsets[query],
                                 x->M, x->X, x->P, NULL, status);
}

/* ClinkageClusterScorer.java:30-49 for cl2 = {query}.  Counts executed pair scores. */
static int32_t cluster_score(const octx* x, const int32_t* members, int32_t nm, int32_t query,
                             int64_t* pairs, int64_t* cells, int* status) {
    int32_t result = INT32_MAX;
    int lq = x->offsets[query + 1] - x->offsets[query];
    for (int32_t i = 0; i < nm; i++) {
        int32_t m = members[i];
        int32_t r = seq_score(x, m, query, status);
        (*pairs)++;
        *cells += hmko_pair_cells(x->offsets[m + 1] - x->offsets[m], lq, x->X);
        if (r < result) {
            result = r;
            if (result < x->T) return JMIN + 1;     /* :41-43 */
        }
    }
    return result;
}

typedef struct {
    int32_t score;
    int32_t size;
    int32_t id;
    int32_t idx;   /* index into the candidate array, -1 = Java null cluster */
} obest;

/* NearestClusterRunner.call (ClinkageSequenceClusterer.java:258-293): is candidate b
 * preferred over the current a?  score higher, then size bigger, then id smaller. */
static inline int better(const obest* b, const obest* a) {
    if (a->idx < 0) return b->score > a->score; /* nearestCluster == null: only `score > maxScore` takes it */
    if (b->score != a->score) return b->score > a->score;
    if (b->size != a->size) return b->size > a->size;
    return b->id < a->id;
}

/* findNearestClusterParallel (ClinkageSequenceClusterer.java:137-177) over multi-member
 * clusters.  Return: found=1 with *out = winner; found=0 -> Java null; found=2 -> the
 * non-null (null cluster, MIN_VALUE) object returned for an EMPTY candidate collection
 * (:138-140). */
static int nearest_among_clusters(octx* x, ocluster* cl, int32_t ncl, int32_t query, obest* out,
                                  int64_t* pairs, int64_t* cells) {
    if (ncl == 0) { out->score = JMIN; out->idx = -1; return 2; }
    obest best = {JMIN, 0, 0, -1};
    int64_t pr = 0, ce = 0;
    int st = 0;
#pragma omp parallel num_threads(x->nthreads) if (x->nthreads > 1 && ncl > 64)
    {
        obest lb = {JMIN, 0, 0, -1};
        int64_t lp = 0, lc = 0;
        int lst = 0;
#pragma omp for schedule(static) nowait
        for (int32_t i = 0; i < ncl; i++) {
            obest c;
            c.score = cluster_score(x, cl[i].members, cl[i].n, query, &lp, &lc, &lst);
            c.size = cl[i].size; c.id = cl[i].id; c.idx = i;
            if (better(&c, &lb)) lb = c;
        }
#pragma omp critical
        {
            /* merge loop :151-176 -- same strict total order, so part order is irrelevant */
            if (lb.idx >= 0 && better(&lb, &best)) best = lb;
            pr += lp; ce += lc;
            if (lst) st = lst;
        }
    }
    *pairs += pr; *cells += ce;
    if (st) x->status = st;
    /* sentinel MIN_VALUE+42 (:151,159-161): sub-threshold winners (MIN_VALUE+1) are dropped */
    if (best.idx < 0 || best.score < JMIN + 42) return 0;
    *out = best;
    return 1;
}

/* Same search over the singleton clusters initialList[index+1 ..] (each has one member, so
 * clusterScore is the pair score or MIN_VALUE+1).  alive[j] = still in initialList. */
static int nearest_among_singles(octx* x, const uint8_t* alive, int32_t from, int32_t n_after,
                                 int32_t query, obest* out, int64_t* pairs, int64_t* cells) {
    if (n_after == 0) { out->score = JMIN; out->idx = -1; return 2; }
    obest best = {JMIN, 0, 0, -1};
    int64_t pr = 0, ce = 0;
    int st = 0;
    int lq = x->offsets[query + 1] - x->offsets[query];
#pragma omp parallel num_threads(x->nthreads) if (x->nthreads > 1)
    {
        obest lb = {JMIN, 0, 0, -1};
        int64_t lp = 0, lc = 0;
        int lst = 0;
#pragma omp for schedule(static) nowait
        for (int32_t j = from; j < x->n; j++) {
            if (!alive[j]) continue;
            int32_t r = seq_score(x, j, query, &lst);
            lp++;
            lc += hmko_pair_cells(x->offsets[j + 1] - x->offsets[j], lq, x->X);
            obest c;
            c.score = r < x->T ? JMIN + 1 : r;   /* ClinkageClusterScorer.java:39-44 */
            c.size = x->abundance[j];            /* new Cluster({seq}, i).size()      */
            c.id = j; c.idx = j;
            if (better(&c, &lb)) lb = c;
        }
#pragma omp critical
        {
            if (lb.idx >= 0 && better(&lb, &best)) best = lb;
            pr += lp; ce += lc;
            if (lst) st = lst;
        }
    }
    *pairs += pr; *cells += ce;
    if (st) x->status = st;
    if (best.idx < 0 || best.score < JMIN + 42) return 0;
    *out = best;
    return 1;
}

int hmko_greedy_cluster_bounded(int32_t n, const uint8_t* residues, const int32_t* offsets,
                                const int32_t* abundance, const int32_t* M, int32_t T, int32_t X,
                                int32_t P, int32_t K, int32_t nthreads, int64_t max_p1_steps,
                                int64_t max_p2_queries, int32_t* cluster_id, int32_t* member_rank,
                                int32_t* result_order, int32_t* n_result, int32_t* n_multi,
                                hmko_counters* ctr) {
    hmko_counters c0;
    memset(&c0, 0, sizeof c0);
    c0.npe_step = -1;
    octx x = {n, residues, offsets, abundance, M, T, X, P, nthreads > 0 ? nthreads : 1, 0};
    for (int32_t i = 0; i < n; i++)
        for (int32_t p = offsets[i]; p < offsets[i + 1]; p++)
            if (residues[p] >= HMKO_NRES) { if (ctr) *ctr = c0; return HMKO_ERR_BAD_RESIDUE; }

    /* firstPhase (LimitedGreedySequenceClusterer.java:77-120).  initialList is the id-ordered
     * list of singleton clusters; `alive` marks the elements still in it, `cur` walks it
     * (initialList.get(index)), `list_size` is initialList.size(). */
    uint8_t* alive = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    memset(alive, 1, (size_t)(n > 0 ? n : 1));
    ocluster* actual = (ocluster*)calloc((size_t)(K > 0 ? K : 1), sizeof(ocluster));
    int32_t nactual = 0;
    int32_t* orphans = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t norph = 0;
    int32_t list_size = n, index = 0, cur = 0;
    int rc = HMKO_OK;

    while (index < list_size && nactual < K) {                                          /* :90 */
        if (max_p1_steps > 0 && c0.p1_steps >= max_p1_steps) break;   /* bounded baseline only */
        while (!alive[cur]) cur++;
        int32_t q = cur;                                                                /* :91 */
        obest A, B;
        int fa = nearest_among_clusters(&x, actual, nactual, q, &A, &c0.p1_pairs, &c0.cells);   /* :92 */
        int fb = nearest_among_singles(&x, alive, q + 1, list_size - index - 1, q, &B,
                                       &c0.p1_pairs, &c0.cells);                        /* :93 */
        if (x.status) { rc = x.status; break; }
        int join = 0, create = 0;
        if (fa != 0) {                                                                  /* :94 */
            if (fb != 0) {                                                              /* :95 */
                if (A.score >= B.score) join = 1; else create = 1;                      /* :96-102 */
            } else join = 1;                                                            /* :103-105 */
        } else {
            if (fb != 0) create = 1;                                                    /* :107-110 */
        }
        /* Java null dereference: getCluster() of the (null, MIN_VALUE) object */
        if ((join && fa == 2) || (create && fb == 2)) {
            rc = HMKO_ERR_NULL_CLUSTER;
            c0.npe_step = (int32_t)c0.p1_steps;
            break;
        }
        if (join) {
            cl_insert(&actual[A.idx], q, &x);                                           /* :97,104 */
            c0.p1_joins++;
        } else if (create) {
            ocluster* nc = &actual[nactual++];                                          /* :100,109 */
            nc->id = q; nc->size = 0; nc->n = 0;
            cl_insert(nc, q, &x);
            cl_insert(nc, B.idx, &x);                                                   /* :99,108 */
            alive[B.idx] = 0;                                                           /* :101,110 */
            list_size--;
            c0.p1_new_clusters++;
        } else {
            orphans[norph++] = q;                                                       /* :112 */
            c0.p1_orphans++;
        }
        c0.p1_steps++;
        index++;                                                                        /* :115 */
        cur++;
    }

    /* outputs */
    if (cluster_id) for (int32_t i = 0; i < n; i++) { cluster_id[i] = i; member_rank[i] = 0; }
    int32_t nres = 0;
    if (rc == HMKO_OK) {
        /* cluster(): split at the first size-1 cluster (:43-51); singles = orphans ++
         * initialList[index..] (:117-119), visited in that order (:59-66). */
        int32_t nsingles = norph;
        int32_t* singles = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        memcpy(singles, orphans, sizeof(int32_t) * (size_t)norph);
        {
            int32_t left = list_size - index, j = cur;
            while (left > 0) {
                while (!alive[j]) j++;
                singles[nsingles++] = j++;
                left--;
            }
        }
        int32_t* remaining = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        int32_t nrem = 0;
        for (int32_t s = 0; s < nsingles; s++) {
            if (max_p2_queries > 0 && c0.p2_queries >= max_p2_queries) {
                remaining[nrem++] = singles[s];
                continue;
            }
            int32_t q = singles[s];
            for (int32_t i = 0; i < nactual; i++) c0.p2_pairs_dense += actual[i].n;
            obest A;
            int fa = nearest_among_clusters(&x, actual, nactual, q, &A, &c0.p2_pairs_early, &c0.cells);
            if (x.status) { rc = x.status; break; }
            c0.p2_queries++;
            if (fa != 0 && A.score >= T) {                                              /* :61 */
                cl_insert(&actual[A.idx], q, &x);                                       /* :62 */
                c0.p2_assigned++;
            } else remaining[nrem++] = q;                                               /* :64 */
        }
        if (cluster_id) {
            for (int32_t i = 0; i < nactual; i++) {
                for (int32_t r = 0; r < actual[i].n; r++) {
                    cluster_id[actual[i].members[r]] = actual[i].id;
                    member_rank[actual[i].members[r]] = r;
                }
                result_order[nres++] = actual[i].id;                                    /* :67 */
            }
            for (int32_t i = 0; i < nrem; i++) result_order[nres++] = remaining[i];     /* :67-68 */
        }
        free(singles);
        free(remaining);
    }
    if (n_result) *n_result = nres;
    if (n_multi) *n_multi = (rc == HMKO_OK) ? nactual : 0;
    if (ctr) *ctr = c0;
    for (int32_t i = 0; i < nactual; i++) free(actual[i].members);
    free(actual);
    free(orphans);
    free(alive);
    return rc;
}

int hmko_greedy_cluster(int32_t n, const uint8_t* residues, const int32_t* offsets,
                        const int32_t* abundance, const int32_t* M, int32_t T, int32_t X, int32_t P,
                        int32_t K, int32_t nthreads, int32_t* cluster_id, int32_t* member_rank,
                        int32_t* result_order, int32_t* n_result, int32_t* n_multi,
                        hmko_counters* ctr) {
    return hmko_greedy_cluster_bounded(n, residues, offsets, abundance, M, T, X, P, K, nthreads, 0, 0,
                                       cluster_id, member_rank, result_order, n_result, n_multi, ctr);
}

/* ---------------------------------------------------------------- fasta loader */

/* Integer.decode: sign, then 0x/0X/# hex, leading-0 octal, else decimal */
static int decode_int_java(const char* s, size_t len, int32_t* out) {
    size_t i = 0;
    int neg = 0, radix = 10;
    if (len == 0) return -1;
    if (s[0] == '-') { neg = 1; i++; } else if (s[0] == '+') i++;
    if (i + 1 < len && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X')) { radix = 16; i += 2; }
    else if (i < len && s[i] == '#') { radix = 16; i++; }
    else if (i + 1 < len && s[i] == '0') { radix = 8; i++; }
    if (i >= len) return -1;
    int64_t v = 0;
    for (; i < len; i++) {
        int d;
        char ch = s[i];
        if (ch >= '0' && ch <= '9') d = ch - '0';
        else if (ch >= 'a' && ch <= 'f') d = ch - 'a' + 10;
        else if (ch >= 'A' && ch <= 'F') d = ch - 'A' + 10;
        else return -1;
        if (d >= radix) return -1;
        v = v * radix + d;
        if (v > (int64_t)INT32_MAX + 1) return -1;
    }
    if (neg) v = -v;
    if (v > INT32_MAX || v < INT32_MIN) return -1;
    *out = (int32_t)v;
    return 0;
}

static uint64_t fnv1a(const char* s, size_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; }
    return h;
}

static void trim_java(const char** s, size_t* n) { /* String.trim(): chars <= ' ' */
    while (*n && (unsigned char)(*s)[0] <= ' ') { (*s)++; (*n)--; }
    while (*n && (unsigned char)(*s)[*n - 1] <= ' ') (*n)--;
}

/* FileIOManager.java:159-216.  The map key is the RAW (case-sensitive) sequence string
 * (:168,192); repeated keys accumulate counts; iteration order = first occurrence
 * (LinkedHashMap).  Only the total over labels is kept here (UniqueSequence.size()). */
int hmko_load_fasta(const char* path, hmko_fasta* out, char* err, size_t errlen) {
    memset(out, 0, sizeof *out);
    FILE* f = fopen(path, "rb");
    if (!f) { set_err(err, errlen, "cannot open fasta file"); return HMKO_ERR_IO; }
    size_t ncap = 1024, tcap = 1 << 16, tlen = 0;
    int32_t n = 0;
    char* text = (char*)malloc(tcap);
    int32_t* offs = (int32_t*)malloc(sizeof(int32_t) * (ncap + 1));
    int32_t* abund = (int32_t*)malloc(sizeof(int32_t) * ncap);
    offs[0] = 0;
    size_t hcap = 4096;
    int32_t* table = (int32_t*)malloc(sizeof(int32_t) * hcap);
    for (size_t i = 0; i < hcap; i++) table[i] = -1;

    char* line = NULL;
    size_t lcap = 0;
    ssize_t got;
    char* seq = (char*)malloc(256);
    size_t scap = 256, slen = 0;
    int have_hdr = 0, rc = HMKO_OK, eof_flush = 0;
    int32_t count = 0;

    for (;;) {
        got = getline(&line, &lcap, f);
        int flush = 0;
        const char* p = NULL;
        size_t len = 0;
        if (got < 0) { flush = 1; eof_flush = 1; }
        else {
            len = (size_t)got;
            if (len && line[len - 1] == '\n') len--;
            if (len && line[len - 1] == '\r') len--;
            p = line;
            if (len && p[0] == '>') flush = (slen > 0);                              /* :170 */
        }
        if (flush) {
            if (eof_flush && !have_hdr) { /* updateLabelsMap(.., null count) -> NPE in the reference */
                set_err(err, errlen, "empty fasta input");
                rc = HMKO_ERR_FILE_FORMAT;
                break;
            }
            /* find / insert key */
            uint64_t h = fnv1a(seq, slen);
            size_t pos = (size_t)(h & (hcap - 1));
            int32_t found = -1;
            while (table[pos] >= 0) {
                int32_t e = table[pos];
                size_t el = (size_t)(offs[e + 1] - offs[e]);
                if (el == slen && memcmp(text + offs[e], seq, slen) == 0) { found = e; break; }
                pos = (pos + 1) & (hcap - 1);
            }
            if (found >= 0) abund[found] = wrap_add(abund[found], count);             /* :204-216 */
            else {
                if ((size_t)n == ncap) {
                    ncap *= 2;
                    offs = (int32_t*)realloc(offs, sizeof(int32_t) * (ncap + 1));
                    abund = (int32_t*)realloc(abund, sizeof(int32_t) * ncap);
                }
                while (tlen + slen + 1 > tcap) { tcap *= 2; text = (char*)realloc(text, tcap); }
                memcpy(text + tlen, seq, slen);
                tlen += slen;
                offs[n + 1] = (int32_t)tlen;
                abund[n] = count;
                table[pos] = n++;
                if ((size_t)n * 2 > hcap) { /* rehash */
                    hcap *= 2;
                    table = (int32_t*)realloc(table, sizeof(int32_t) * hcap);
                    for (size_t i = 0; i < hcap; i++) table[i] = -1;
                    for (int32_t e = 0; e < n; e++) {
                        size_t q = (size_t)(fnv1a(text + offs[e], (size_t)(offs[e + 1] - offs[e])) & (hcap - 1));
                        while (table[q] >= 0) q = (q + 1) & (hcap - 1);
                        table[q] = e;
                    }
                }
            }
            slen = 0;
        }
        if (got < 0) break;
        if (len && p[0] == '>') {
            /* header: line.trim().substring(1).split("\\|")  (:175-187) */
            const char* hp = p;
            size_t hl = len;
            trim_java(&hp, &hl);
            hp++; hl--;
            const char* fields[3] = {0, 0, 0};
            size_t flen[3] = {0, 0, 0};
            int nf = 0;
            size_t s = 0;
            int total_fields = 0;
            for (size_t i = 0; i <= hl; i++) {
                if (i == hl || hp[i] == '|') {
                    if (nf < 3) { fields[nf] = hp + s; flen[nf] = i - s; nf++; }
                    total_fields++;
                    s = i + 1;
                }
            }
            /* Java split drops TRAILING empty strings: recompute the effective length */
            int eff = 0;
            {
                /* walk again to find the last non-empty field index */
                size_t st = 0;
                int idx = 0;
                for (size_t i = 0; i <= hl; i++) {
                    if (i == hl || hp[i] == '|') {
                        if (i - st > 0) eff = idx + 1;
                        idx++;
                        st = i + 1;
                    }
                }
                if (eff == 0) eff = 1; /* "".split -> [""] */
            }
            (void)total_fields;
            if (eff >= 2) {
                const char* cp = fields[1];
                size_t cl = flen[1];
                trim_java(&cp, &cl);
                if (decode_int_java(cp, cl, &count)) {
                    set_err(err, errlen, "NumberFormatException in fasta header count");
                    rc = HMKO_ERR_FILE_FORMAT;
                    break;
                }
                if (count < 1) {
                    set_err(err, errlen, "Fasta header defines sequence count lower than 1.");
                    rc = HMKO_ERR_FILE_FORMAT;
                    break;
                }
            } else count = 1;
            have_hdr = 1;
        } else {
            if (!have_hdr) {                                                          /* :189-191 */
                set_err(err, errlen, "Incorrect fasta format. Maybe header or sequence line missing?");
                rc = HMKO_ERR_FILE_FORMAT;
                break;
            }
            const char* sp = p;
            size_t sl = len;
            trim_java(&sp, &sl);
            while (slen + sl + 1 > scap) { scap *= 2; seq = (char*)realloc(seq, scap); }
            memcpy(seq + slen, sp, sl);                                               /* :192 */
            slen += sl;
        }
    }
    free(line);
    free(seq);
    free(table);
    fclose(f);
    if (rc == HMKO_OK) {
        uint8_t* res = (uint8_t*)malloc(tlen ? tlen : 1);
        for (size_t i = 0; i < tlen; i++) {
            int c = hmko_encode_char(text[i]);                                        /* UniqueSequence.java:49-55 */
            if (c < 0) { rc = HMKO_ERR_BAD_RESIDUE; set_err(err, errlen, "not a valid letter from the amino acid alphabet code"); break; }
            res[i] = (uint8_t)c;
            text[i] = (char)toupper((unsigned char)text[i]);
        }
        if (rc == HMKO_OK) {
            out->n = n; out->residues = res; out->offsets = offs; out->abundance = abund; out->text = text;
            return HMKO_OK;
        }
        free(res);
    }
    free(text); free(offs); free(abund);
    return rc;
}

void hmko_fasta_free(hmko_fasta* f) {
    free(f->residues); free(f->offsets); free(f->abundance); free(f->text);
    memset(f, 0, sizeof *f);
}
