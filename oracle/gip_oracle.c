/* oracle/gip_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the GIP scorer + top-k of castorini/dhr
 * retrieval/gip_retrieval.py (reference @ e236f3d):
 *   exact branch   :110-126  (idx equality mask, masked row dot, top-k)
 *   dense-only     :60-85    (row dot, descending sort, keep k)
 *   --IP branch    :139      (masked == 0)
 * Scores are accumulated in double so they are "exact" to ~1e-13 for fp16 corpus
 * values and fp32 query values.  Top-k tie rule: (score desc, row asc).
 *
 * Parity: pinned against outputs of the real reference, see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { IDX_NONE = 0, IDX_U8 = 1, IDX_I8 = 2, IDX_I16 = 3, IDX_U16 = 4, IDX_I32 = 5, IDX_I64 = 6 };

static inline int64_t load_idx(const void *base, int dtype, int64_t off) {
    switch (dtype) {
    case IDX_U8:  return ((const uint8_t *)base)[off];
    case IDX_I8:  return ((const int8_t *)base)[off];
    case IDX_I16: return ((const int16_t *)base)[off];
    case IDX_U16: return ((const uint16_t *)base)[off];
    case IDX_I32: return ((const int32_t *)base)[off];
    default:      return ((const int64_t *)base)[off];
    }
}

static inline double h2d(uint16_t bits) {
    _Float16 h;
    memcpy(&h, &bits, 2);
    return (double)h;
}

/* score of one (query, passage) pair */
static inline double pair_score(const uint16_t *cv, const void *ci, int ci_dt, int64_t ci_off,
                                const float *qv, const int64_t *qi, int S, int G, int C, int masked) {
    double acc = 0.0;
    int D = S * G;
    if (masked) {
        for (int s = 0; s < S; ++s) {
            if (load_idx(ci, ci_dt, ci_off + s) == qi[s]) {            /* :119 equality on integer value */
                for (int g = 0; g < G; ++g) acc += (double)qv[s * G + g] * h2d(cv[s * G + g]);
            }
        }
    } else {
        for (int j = 0; j < D; ++j) acc += (double)qv[j] * h2d(cv[j]);  /* :139 */
    }
    for (int c = 0; c < C; ++c) acc += (double)qv[D + c] * h2d(cv[D + c]);  /* dense tail, mask always true (:110-113) */
    return acc;
}

typedef struct { double s; int64_t r; } cand_t;

static int cand_cmp(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->r > y->r) - (x->r < y->r);
}

/* Exact scores of query `q` for all rows -> out[N] (double). */
void gip_oracle_scores(int64_t N, int S, int G, int C, const uint16_t *c_vals, const void *c_idx,
                       int c_idx_dtype, const float *q_vals_row, const void *q_idx_row_base, int q_idx_dtype,
                       int64_t q_idx_off, int masked, double *out) {
    int W = S * G + C;
    int use_mask = masked && S > 0 && c_idx && q_idx_row_base;
    int64_t *qi = (int64_t *)malloc(sizeof(int64_t) * (S > 0 ? S : 1));
    if (use_mask)
        for (int s = 0; s < S; ++s) qi[s] = load_idx(q_idx_row_base, q_idx_dtype, q_idx_off + s);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < N; ++p)
        out[p] = pair_score(c_vals + p * W, c_idx, c_idx_dtype, p * S, q_vals_row, qi, S, G, C, use_mask);
    free(qi);
}

/* Full search: for each of Q queries the top-k rows by (score desc, row asc).
 * out_scores [Q,k] double, out_rows [Q,k] int64; entries beyond min(k,N) are (-inf,-1).
 * Returns 0 on success. */
int gip_oracle_search(int64_t N, int S, int G, int C, const uint16_t *c_vals, const void *c_idx, int c_idx_dtype,
                      int Q, const float *q_vals, const void *q_idx, int q_idx_dtype, int masked, int k,
                      double *out_scores, int64_t *out_rows) {
    int W = S * G + C;
    double *sc = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(N > 0 ? N : 1));
    if (!sc || !cand) { free(sc); free(cand); return 1; }
    for (int q = 0; q < Q; ++q) {
        gip_oracle_scores(N, S, G, C, c_vals, c_idx, c_idx_dtype, q_vals + (int64_t)q * W, q_idx, q_idx_dtype,
                          (int64_t)q * S, masked, sc);
        /* threshold prefilter keeps the sort small: k-th largest via nth-element-like pass */
        int64_t kk = k < N ? k : N;
        int64_t m = 0;
        if (N > 8 * (int64_t)k) {
            /* sample-free exact prefilter: find the k-th best score with a partial selection on a copy */
            double *tmp = (double *)malloc(sizeof(double) * (size_t)N);
            memcpy(tmp, sc, sizeof(double) * (size_t)N);
            int64_t lo = 0, hi = N - 1, target = kk - 1;           /* quickselect, descending */
            while (lo < hi) {
                double pivot = tmp[(lo + hi) / 2];
                int64_t i = lo, j = hi;
                while (i <= j) {
                    while (tmp[i] > pivot) ++i;
                    while (tmp[j] < pivot) --j;
                    if (i <= j) { double t = tmp[i]; tmp[i] = tmp[j]; tmp[j] = t; ++i; --j; }
                }
                if (target <= j) hi = j; else if (target >= i) lo = i; else break;
            }
            double kth = tmp[target];
            free(tmp);
            for (int64_t p = 0; p < N; ++p)
                if (sc[p] >= kth) { cand[m].s = sc[p]; cand[m].r = p; ++m; }
        } else {
            for (int64_t p = 0; p < N; ++p) { cand[m].s = sc[p]; cand[m].r = p; ++m; }
        }
        qsort(cand, (size_t)m, sizeof(cand_t), cand_cmp);
        for (int64_t j = 0; j < k; ++j) {
            if (j < kk) { out_scores[(int64_t)q * k + j] = cand[j].s; out_rows[(int64_t)q * k + j] = cand[j].r; }
            else        { out_scores[(int64_t)q * k + j] = -INFINITY; out_rows[(int64_t)q * k + j] = -1; }
        }
    }
    free(sc); free(cand);
    return 0;
}

int gip_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
