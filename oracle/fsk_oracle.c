/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the gapped k-mer kernel build.
 *
 * A plain-C restatement of the reference algorithm (QData/FastSK, commit d275a23), written
 * from the behaviour of the files cited below, with 64-bit indexing and 64-bit integer
 * accumulators so that it also covers N > 46341 (where the reference's `int` tri_access
 * overflows, shared.cpp:97-117) and never wraps (the reference's per-thread accumulator is
 * `unsigned int`, fastsk_kernel.cpp:175-176).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (fastsk_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (i)  oracle/_ref (the unmodified reference engine compiled by oracle/Makefile), when present,
 *   (ii) the committed fixtures in tests/golden/ that were generated from oracle/_ref,
 *   (iii) the known-answer vectors of SURVEY.md section 8c.
 *
 * Reference map
 *   window enumeration (seq-major, position-minor)     shared.cpp:55-91   (extractFeatures)
 *   C(n,k)                                             shared.cpp:335-345 (nchoosek)
 *   lexicographic k-subsets of kept positions          shared.cpp:347-360 (getCombinations)
 *   kept-position gather                               fastsk_kernel.cpp:224-228
 *   stable LSD counting sort, last column first        shared.cpp:156-191 (cntsrtna)
 *   run scan, per-sequence counts, K += c_i*c_j        shared.cpp:268-333 (countAndUpdateTri)
 *   per-stream loop, stop rules, merge                 fastsk_kernel.cpp:145-322
 *   Welford mean / variance statistic                  fastsk_kernel.cpp:108-143
 *   sd, convergence test, stdevs                       fastsk_kernel.cpp:243-262
 *   normalisation K_ij / sqrt(K_ii*K_jj)               fastsk_kernel.cpp:96-103
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FSKO_MAX_G 64

static inline int64_t tri(int64_t i, int64_t j) { /* i >= j; packed lower triangle, shared.cpp:97-117 */
    return i * (i + 1) / 2 + j;
}

/* shared.cpp:335-345 -- exact for the g <= 20 range the reference supports; done in 64 bit here. */
int64_t fsko_nchoosek(int n, int k) {
    if (k > n || k < 0) return 0;
    if (k * 2 > n) k = n - k;
    int64_t r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

/* Kept positions of combination `idx` in the order getCombinations emits them
 * (shared.cpp:347-360: recursive, ascending positions, depth-first => lexicographic). */
int fsko_combination(int g, int k, int64_t idx, int* pos) {
    if (k < 1 || k > g || idx < 0 || idx >= fsko_nchoosek(g, k)) return -1;
    int start = 0;
    for (int d = 0; d < k; ++d) {
        for (int p = start;; ++p) {
            int64_t below = fsko_nchoosek(g - 1 - p, k - 1 - d); /* subsets that fix position p here */
            if (idx < below) { pos[d] = p; start = p + 1; break; }
            idx -= below;
        }
    }
    return 0;
}

typedef struct {
    int64_t nfeat;
    int32_t* wseq;   /* sequence id of each window */
    int64_t* wstart; /* offset of the window's first character in codes */
} windows_t;

/* shared.cpp:55-91: windows in (sequence, position) order; sequences shorter than g yield none. */
static int build_windows(const int64_t* offsets, int64_t n, int g, windows_t* w) {
    int64_t nfeat = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t len = offsets[i + 1] - offsets[i];
        if (len >= g) nfeat += len - g + 1;
    }
    w->nfeat = nfeat;
    w->wseq = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nfeat ? nfeat : 1));
    w->wstart = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nfeat ? nfeat : 1));
    if (!w->wseq || !w->wstart) return -1;
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t len = offsets[i + 1] - offsets[i];
        for (int64_t p = 0; p + g <= len; ++p) {
            w->wseq[c] = (int32_t)i;
            w->wstart[c] = offsets[i] + p;
            ++c;
        }
    }
    return 0;
}

typedef struct {
    int64_t* order;  /* permutation being sorted */
    int64_t* tmp;
    int64_t* count;  /* radix histogram */
    int32_t* col;    /* digit of each element of `order` for the current column */
    int radix;
} sorter_t;

/* One combination: gather kept columns, stable LSD counting sort (shared.cpp:156-191),
 * run scan + outer-product update on the packed triangle (shared.cpp:268-333). */
static void partial_kernel(const int32_t* codes, const windows_t* w, int k, const int* pos,
                           sorter_t* s, uint64_t* K) {
    const int64_t r = w->nfeat;
    for (int64_t i = 0; i < r; ++i) s->order[i] = i;
    for (int c = k - 1; c >= 0; --c) {
        memset(s->count, 0, sizeof(int64_t) * (size_t)(s->radix + 1));
        for (int64_t i = 0; i < r; ++i) {
            int32_t d = codes[w->wstart[s->order[i]] + pos[c]];
            s->col[i] = d;
            s->count[d + 1]++;
        }
        for (int d = 0; d < s->radix; ++d) s->count[d + 1] += s->count[d];
        for (int64_t i = 0; i < r; ++i) s->tmp[s->count[s->col[i]]++] = s->order[i];
        int64_t* t = s->order; s->order = s->tmp; s->tmp = t;
    }
    /* runs of equal k-tuples; stability keeps sequence ids ascending inside a run */
    int64_t i = 0;
    while (i < r) {
        int64_t j = i + 1;
        const int64_t a = w->wstart[s->order[i]];
        for (; j < r; ++j) {
            const int64_t b = w->wstart[s->order[j]];
            int same = 1;
            for (int c = 0; c < k; ++c)
                if (codes[a + pos[c]] != codes[b + pos[c]]) { same = 0; break; }
            if (!same) break;
        }
        /* [i, j) is one run.  Compact it to (seq, count) in place in tmp (ascending seq). */
        int64_t nd = 0;
        for (int64_t p = i; p < j;) {
            int32_t sq = w->wseq[s->order[p]];
            int64_t q = p;
            while (q < j && w->wseq[s->order[q]] == sq) ++q;
            s->tmp[2 * nd] = sq;
            s->tmp[2 * nd + 1] = q - p;
            ++nd;
            p = q;
        }
        /* shared.cpp:316-327: every pair of distinct sequences in the run, diagonal included
         * (a length-1 run degenerates to K[s][s] += 1). */
        for (int64_t x = 0; x < nd; ++x)
            for (int64_t y = x; y < nd; ++y)
                K[tri(s->tmp[2 * y], s->tmp[2 * x])] += (uint64_t)(s->tmp[2 * x + 1] * s->tmp[2 * y + 1]);
        i = j;
    }
}

/* fastsk_kernel.cpp:108-143, literally (including the use of this iteration's sum only and
 * the iter == 1 constant).  `variances[]` / max_variance never influence any output and are omitted. */
static double welford_step(const uint64_t* Ks, double* K_hat, int64_t n_pairs, int64_t n_train_pairs, int iter) {
    double acc = 0;
    int64_t count = 0;
    for (int64_t p = 0; p < n_pairs; ++p) {
        double delta = (double)Ks[p] - K_hat[p];
        K_hat[p] += delta / iter;
        if (p < n_train_pairs) {
            double delta2 = (double)Ks[p] - K_hat[p];
            acc += delta * delta2;
            ++count;
        }
    }
    acc /= count;
    if (iter == 1) acc = 9999999;
    else acc /= iter - 1;
    return acc;
}

/* fastsk_kernel.cpp:96-103: off-diagonals first using the raw diagonals, then the diagonals. */
void fsko_normalise(double* K, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < i; ++j)
            K[tri(i, j)] = K[tri(i, j)] / sqrt(K[tri(i, i)] * K[tri(j, j)]);
    for (int64_t i = 0; i < n; ++i)
        K[tri(i, i)] = K[tri(i, i)] / sqrt(K[tri(i, i)] * K[tri(i, i)]);
}

/* Integer partial kernel of ONE combination given by its kept positions (per-combination counts). */
int fsko_partial(const int32_t* codes, const int64_t* offsets, int64_t n, int g, int k, const int* pos,
                 uint64_t* K_out /* zeroed by caller or accumulated into */) {
    windows_t w;
    if (build_windows(offsets, n, g, &w)) return -4;
    int radix = 1;
    for (int64_t i = 0; i < offsets[n]; ++i) if (codes[i] + 1 > radix) radix = codes[i] + 1;
    sorter_t s;
    size_t cap = (size_t)(w.nfeat > 2 * n ? w.nfeat : 2 * n) + 2;
    s.order = (int64_t*)malloc(sizeof(int64_t) * cap);
    s.tmp = (int64_t*)malloc(sizeof(int64_t) * cap);
    s.col = (int32_t*)malloc(sizeof(int32_t) * cap);
    s.count = (int64_t*)malloc(sizeof(int64_t) * (size_t)(radix + 2));
    s.radix = radix;
    partial_kernel(codes, &w, k, pos, &s, K_out);
    free(s.order); free(s.tmp); free(s.col); free(s.count); free(w.wseq); free(w.wstart);
    return 0;
}

/*
 * The engine (fastsk_kernel.cpp:145-322), with the T reference threads replaced by T virtual
 * streams evaluated one after the other: stream `tid` takes queue[tid], queue[tid+T], ...
 *   exact, or approx && skip_variance : integer sum over the stream's combinations
 *   approx && !skip_variance         : Ks zeroed every iteration, Welford mean K_hat,
 *                                      stop when delta/sd > 1.96; the merged result is the
 *                                      sum over streams of K_hat (fp64).
 * Merge `Ksfinal[p] += val` only when val != 0 (fastsk_kernel.cpp:306-307).
 * K_out: fp64 packed lower triangle (as the reference's K); K_int_out (optional): the integer
 * sum, valid in the two integer modes.  Returns 0, or <0 on bad arguments.
 */
int fsko_build(const int32_t* codes, const int64_t* offsets, int64_t n_train, int64_t n_test,
               int g, int m, int T, int approx, double delta, int max_iters, int skip_variance,
               const int32_t* queue, int nq, int normalise,
               double* K_out, uint64_t* K_int_out, double* stdevs_out, int64_t stdevs_cap, int64_t* n_stdevs) {
    const int64_t n = n_train + n_test;
    const int k = g - m;
    if (g < 1 || g > FSKO_MAX_G || k < 1 || nq < 1) return -1;
    for (int64_t i = 0; i < n; ++i) if (offsets[i + 1] - offsets[i] < g) return -3;
    const int64_t n_pairs = n * (n + 1) / 2;
    const int64_t n_train_pairs = n_train * (n_train + 1) / 2;
    if (T < 1) T = 20;   /* fastsk_kernel.cpp:54-60 */
    if (T > nq) T = nq;  /* fastsk_kernel.cpp:61 */

    windows_t w;
    if (build_windows(offsets, n, g, &w)) return -4;
    int radix = 1;
    for (int64_t i = 0; i < offsets[n]; ++i) {
        if (codes[i] < 0) return -1;
        if (codes[i] + 1 > radix) radix = codes[i] + 1;
    }
    sorter_t s;
    size_t cap = (size_t)(w.nfeat > 2 * n ? w.nfeat : 2 * n) + 2;
    s.order = (int64_t*)malloc(sizeof(int64_t) * cap);
    s.tmp = (int64_t*)malloc(sizeof(int64_t) * cap);
    s.col = (int32_t*)malloc(sizeof(int32_t) * cap);
    s.count = (int64_t*)malloc(sizeof(int64_t) * (size_t)(radix + 2));
    s.radix = radix;
    uint64_t* Ks = (uint64_t*)calloc((size_t)n_pairs, sizeof(uint64_t));
    const int welford = approx && !skip_variance;
    double* K_hat = welford ? (double*)malloc(sizeof(double) * (size_t)n_pairs) : NULL;
    if (!s.order || !s.tmp || !s.col || !s.count || !Ks || (welford && !K_hat)) return -4;

    memset(K_out, 0, sizeof(double) * (size_t)n_pairs);
    if (K_int_out) memset(K_int_out, 0, sizeof(uint64_t) * (size_t)n_pairs);
    int64_t ns = 0;
    int pos[FSKO_MAX_G];

    for (int tid = 0; tid < T; ++tid) {
        memset(Ks, 0, sizeof(uint64_t) * (size_t)n_pairs);
        if (welford) memset(K_hat, 0, sizeof(double) * (size_t)n_pairs);
        int item = tid, iter = 1, working = 1;
        while (working) {
            if (welford) memset(Ks, 0, sizeof(uint64_t) * (size_t)n_pairs);
            if (fsko_combination(g, k, queue[item], pos)) return -1;
            partial_kernel(codes, &w, k, pos, &s, Ks);
            if (welford) {
                double sd = welford_step(Ks, K_hat, n_pairs, n_train_pairs, iter);
                sd = sqrt(sd / iter);
                if (tid == 0) {
                    if (stdevs_out && ns < stdevs_cap) stdevs_out[ns] = sd;
                    ++ns;
                }
                if (delta / sd > 1.96) working = 0;
            }
            if (approx && max_iters != -1 && iter >= max_iters) working = 0;
            item += T;
            if (item >= nq) working = 0;
            ++iter;
        }
        for (int64_t p = 0; p < n_pairs; ++p) {
            double val = welford ? K_hat[p] : (double)Ks[p];
            if (val != 0) K_out[p] += val;
            if (K_int_out && !welford) K_int_out[p] += Ks[p];
        }
    }
    if (n_stdevs) *n_stdevs = ns;
    if (normalise) fsko_normalise(K_out, n);
    free(s.order); free(s.tmp); free(s.col); free(s.count); free(w.wseq); free(w.wstart);
    free(Ks); free(K_hat);
    return 0;
}
