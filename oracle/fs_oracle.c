/*
 * fs_oracle.c -- CPU restatement of FastSelect's Relief-family scoring path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fastselect_b200/ may import, link or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against outputs of the reference itself (tests/golden .npz files, generated
 * by tests/golden/make_golden.py importing /root/reference/src/fast_select),
 * including the known-answer vectors of SURVEY.md section 8(c).
 *
 * All file:line citations are into /root/reference/src/fast_select/.
 *
 * Third-party arithmetic on this path: ReliefF orders neighbours with
 * numba's np.argsort (numba >=0.56 per pyproject.toml:45-50; 0.65.0 installed
 * when the golden vectors were made).  Its quicksort (numba/misc/quicksort.py:
 * median-of-three partition, explicit stack, insertion sort below 15 items)
 * is restated in fso_argsort_numba() so that tie order matches the reference.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off; no
 * -ffast-math, so every rounding below is the one written).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FSO_API __attribute__((visibility("default")))

/* mask codes written to mask_out (one int8 per (target, j)) */
enum { FSO_NONE = 0, FSO_NEAR_HIT = 1, FSO_NEAR_MISS = 2, FSO_FAR_MISS = 3, FSO_FAR_HIT = 4 };

FSO_API void fso_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

FSO_API int fso_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* a1: per-feature diff.  MultiSURF.py:184-188, ReliefF.py:151-154 (x is    */
/* float32: f32 subtract, abs, f32 multiply; discrete = exact 0/1).         */
static inline double diff_f32(const float *xi, const float *xj, int64_t f,
                              const float *recip, const uint8_t *isd) {
    float a = xi[f], b = xj[f];
    float t = fabsf(a - b) * recip[f];
    double ne = (a != b) ? 1.0 : 0.0;
    return isd[f] ? ne : (double)t;
}

/* SURF.py:153-156 (x is float64, recip float32 promoted: f64 arithmetic). */
static inline double diff_f64(const double *xi, const double *xj, int64_t f,
                              const float *recip, const uint8_t *isd) {
    double a = xi[f], b = xj[f];
    double t = fabs(a - b) * (double)recip[f];
    double ne = (a != b) ? 1.0 : 0.0;
    return isd[f] ? ne : t;
}

/* a2: d_ij = sum_f diff_f in float64 (MultiSURF.py:181-191, ReliefF.py:149-155).
 * numba compiles the reference with fastmath=True, i.e. the reduction order
 * is the compiler's; we fix it to 8 interleaved partial sums (vectorisable,
 * deterministic).  The difference to any other order is O(1e-16) relative. */
static double dist_f32(const float *xi, const float *xj, int64_t p,
                       const float *recip, const uint8_t *isd) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t f = 0;
    for (; f + 8 <= p; f += 8)
        for (int l = 0; l < 8; ++l) s[l] += diff_f32(xi, xj, f + l, recip, isd);
    for (; f < p; ++f) s[f & 7] += diff_f32(xi, xj, f, recip, isd);
    return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
}

static double dist_f64(const double *xi, const double *xj, int64_t p,
                       const float *recip, const uint8_t *isd) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t f = 0;
    for (; f + 8 <= p; f += 8)
        for (int l = 0; l < 8; ++l) s[l] += diff_f64(xi, xj, f + l, recip, isd);
    for (; f < p; ++f) s[f & 7] += diff_f64(xi, xj, f, recip, isd);
    return ((s[0] + s[4]) + (s[2] + s[6])) + ((s[1] + s[5]) + (s[3] + s[7]));
}

/* ------------------------------------------------------------------------ */
/* MultiSURF / MultiSURF*  (MultiSURF.py:165-253)                           */
/*
 * One target i.  out32 (reference arithmetic: float32 accumulators,
 * MultiSURF.py:198-251) or out64 (float64 accumulators, used to check the
 * GPU path tightly) receives W_i[f] = miss_diffs[f] - hit_diffs[f].
 */
static void multisurf_target(const float *x, int64_t n, int64_t p, const int64_t *y,
                             const float *recip, const uint8_t *isd, int use_star,
                             int64_t i, float *out32, double *out64,
                             double *thresh_out, int8_t *mask_row, double *dist_row,
                             int64_t *counts /* nH, nM, nF */) {
    const float *xi = x + i * p;
    /* pass 1: MultiSURF.py:175-196 */
    double sum_d = 0.0, sum_d2 = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        if (j == i) continue;
        double d = dist_f32(xi, x + j * p, p, recip, isd);
        sum_d += d;
        sum_d2 += d * d;
    }
    /* MultiSURF.py:193-196 as numba (fastmath=True) compiles it on x86-64 with
     * FMA (observed with inspect_asm(), numba 0.65): both divisions become a
     * multiply by inv = 1/(n-1), mu*mu is rounded, and the subtraction is fused
     * (vfmsub231sd).  This matters when T_i is mathematically an integer and the
     * distances are integers (fixture A, discrete_limit=10 -- pinned by the
     * golden vectors); the GPU path performs the same four operations. */
    double inv = 1.0 / (double)(n - 1);
    double mu = sum_d * inv;
    double var = fma(sum_d2, inv, -(mu * mu));
    if (!(var > 0.0)) var = 0.0;
    double thresh = mu - 0.5 * sqrt(var);
    if (thresh_out) *thresh_out = thresh;

    float *h32 = NULL, *m32 = NULL;
    double *h64 = NULL, *m64 = NULL;
    if (out32) { h32 = calloc(p, sizeof(float)); m32 = calloc(p, sizeof(float)); }
    if (out64) { h64 = calloc(p, sizeof(double)); m64 = calloc(p, sizeof(double)); }
    int64_t n_hits = 0, n_miss = 0, n_far = 0;

    /* pass 2: MultiSURF.py:203-243 (distance recomputed, as the reference does) */
    for (int64_t j = 0; j < n; ++j) {
        if (mask_row) mask_row[j] = FSO_NONE;
        if (dist_row) dist_row[j] = 0.0;
        if (j == i) continue;
        const float *xj = x + j * p;
        double d = dist_f32(xi, xj, p, recip, isd);
        if (dist_row) dist_row[j] = d;
        int is_hit = (y[i] == y[j]);
        int code = FSO_NONE;
        if (d < thresh) code = is_hit ? FSO_NEAR_HIT : FSO_NEAR_MISS;     /* :217 strict < */
        else if (use_star && !is_hit) code = FSO_FAR_MISS;                /* :236 */
        if (mask_row) mask_row[j] = (int8_t)code;
        if (code == FSO_NONE) continue;
        if (code == FSO_NEAR_HIT) n_hits++;
        else if (code == FSO_NEAR_MISS) n_miss++;
        else n_far++;
        for (int64_t f = 0; f < p; ++f) {
            double t = diff_f32(xi, xj, f, recip, isd);
            if (out32) {   /* float32 array element += float64 scalar */
                if (code == FSO_NEAR_HIT) h32[f] = (float)((double)h32[f] + t);
                else if (code == FSO_NEAR_MISS) m32[f] = (float)((double)m32[f] + t);
                else m32[f] = (float)((double)m32[f] - t);
            }
            if (out64) {
                if (code == FSO_NEAR_HIT) h64[f] += t;
                else if (code == FSO_NEAR_MISS) m64[f] += t;
                else m64[f] -= t;
            }
        }
    }
    /* MultiSURF.py:245-251: zero-count guards; far-miss shares /n_miss */
    for (int64_t f = 0; f < p; ++f) {
        if (out32) {
            float h = h32[f], m = m32[f];
            if (n_hits > 0) h = (float)((double)h / (double)n_hits);
            if (n_miss > 0) m = (float)((double)m / (double)n_miss);
            out32[f] = m - h;
        }
        if (out64) {
            double h = h64[f], m = m64[f];
            if (n_hits > 0) h /= (double)n_hits;
            if (n_miss > 0) m /= (double)n_miss;
            out64[f] = m - h;
        }
    }
    if (counts) { counts[0] = n_hits; counts[1] = n_miss; counts[2] = n_far; }
    free(h32); free(m32); free(h64); free(m64);
}

/* Full reference pipeline: temp_scores[n,p] float32, float32 column sums in
 * row order (MultiSURF.py:250-253), then / n_samples in float32 (:270). */
FSO_API int fso_multisurf_scores(const float *x, int64_t n, int64_t p, const int64_t *y,
                                 const float *recip, const uint8_t *isd, int use_star,
                                 float *scores_out) {
    if (n < 2 || p < 1) return -1;
    float *temp = malloc((size_t)n * p * sizeof(float));
    if (!temp) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i)
        multisurf_target(x, n, p, y, recip, isd, use_star, i, temp + i * p, NULL,
                         NULL, NULL, NULL, NULL);
    for (int64_t f = 0; f < p; ++f) {
        float s = 0.0f;
        for (int64_t i = 0; i < n; ++i) s += temp[i * p + f];
        scores_out[f] = s / (float)n;
    }
    free(temp);
    return 0;
}

/* Per-target view for parity tests: the targets listed, float64 accumulators.
 * wsum_out[p] = sum over targets of W_i[f] (not divided by n).  Optional:
 * thresh_out[nt], mask_out[nt*n], dist_out[nt*n], counts_out[nt*3]. */
FSO_API int fso_multisurf_targets(const float *x, int64_t n, int64_t p, const int64_t *y,
                                  const float *recip, const uint8_t *isd, int use_star,
                                  const int64_t *targets, int64_t nt,
                                  double *wsum_out, double *thresh_out, int8_t *mask_out,
                                  double *dist_out, int64_t *counts_out) {
    if (n < 2 || p < 1) return -1;
    double *rows = malloc((size_t)nt * p * sizeof(double));
    if (!rows) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nt; ++t)
        multisurf_target(x, n, p, y, recip, isd, use_star, targets[t], NULL, rows + t * p,
                         thresh_out ? thresh_out + t : NULL,
                         mask_out ? mask_out + t * n : NULL,
                         dist_out ? dist_out + t * n : NULL,
                         counts_out ? counts_out + t * 3 : NULL);
    for (int64_t f = 0; f < p; ++f) {
        double s = 0.0;
        for (int64_t t = 0; t < nt; ++t) s += rows[t * p + f];
        wsum_out[f] = s;
    }
    free(rows);
    return 0;
}

/*
 * The same per-target computation for a matrix of one-byte genotype codes in which
 * EVERY column is discrete (MultiSURF.py:184-185: diff = [a != b]) -- the form the
 * reference's float32 kernel takes on such data, restated on the bytes so that the
 * full-width benchmark shapes (20 000 x 500 000 int8 = 10 GB; 40 GB as float32)
 * fit in host memory.  cols (nullable) restricts the computation to a column
 * subset, as TuRF's X[:, active] does (TuRF.py:110).  Distances and neighbour
 * counts are exact integers; the per-feature sums are integer counts divided once
 * (float64), i.e. the out64 arithmetic of multisurf_target.  Pinned against
 * multisurf_target on the genotype fixtures by tests/test_oracle_golden.py.
 */
static void multisurf_target_u8(const uint8_t *x, int64_t n, int64_t ld, const int64_t *cols, int64_t p,
                                const int64_t *y, int use_star, int64_t i, double *out64,
                                double *thresh_out, int8_t *mask_row, double *dist_row) {
    const uint8_t *xi = x + i * ld;
    uint8_t *ti = NULL;          /* target row restricted to cols */
    if (cols) {
        ti = malloc((size_t)p);
        for (int64_t f = 0; f < p; ++f) ti[f] = xi[cols[f]];
    }
    int32_t *drow = malloc((size_t)n * sizeof(int32_t));
    /* pass 1: MultiSURF.py:175-196 */
    double sum_d = 0.0, sum_d2 = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        const uint8_t *xj = x + j * ld;
        int32_t d = 0;
        if (cols) for (int64_t f = 0; f < p; ++f) d += ti[f] != xj[cols[f]];
        else for (int64_t f = 0; f < p; ++f) d += xi[f] != xj[f];
        drow[j] = d;
        if (j == i) continue;
        sum_d += (double)d;
        sum_d2 += (double)d * (double)d;
    }
    double inv = 1.0 / (double)(n - 1);
    double mu = sum_d * inv;
    double var = fma(sum_d2, inv, -(mu * mu));
    if (!(var > 0.0)) var = 0.0;
    double thresh = mu - 0.5 * sqrt(var);
    if (thresh_out) *thresh_out = thresh;
    int32_t *h = calloc((size_t)p, sizeof(int32_t)), *m = calloc((size_t)p, sizeof(int32_t));
    int64_t n_hits = 0, n_miss = 0;
    /* pass 2: MultiSURF.py:203-243 */
    for (int64_t j = 0; j < n; ++j) {
        if (mask_row) mask_row[j] = FSO_NONE;
        if (dist_row) dist_row[j] = 0.0;
        if (j == i) continue;
        const uint8_t *xj = x + j * ld;
        double d = (double)drow[j];
        if (dist_row) dist_row[j] = d;
        int is_hit = (y[i] == y[j]);
        int code = FSO_NONE;
        if (d < thresh) code = is_hit ? FSO_NEAR_HIT : FSO_NEAR_MISS;
        else if (use_star && !is_hit) code = FSO_FAR_MISS;
        if (mask_row) mask_row[j] = (int8_t)code;
        if (code == FSO_NONE) continue;
        if (code == FSO_NEAR_HIT) n_hits++;
        else if (code == FSO_NEAR_MISS) n_miss++;
        int32_t *acc = code == FSO_NEAR_HIT ? h : m;
        const int32_t sgn = code == FSO_FAR_MISS ? -1 : 1;
        if (cols) for (int64_t f = 0; f < p; ++f) acc[f] += sgn * (int32_t)(ti[f] != xj[cols[f]]);
        else for (int64_t f = 0; f < p; ++f) acc[f] += sgn * (int32_t)(xi[f] != xj[f]);
    }
    for (int64_t f = 0; f < p; ++f) {      /* MultiSURF.py:245-251 */
        double hh = (double)h[f], mm = (double)m[f];
        if (n_hits > 0) hh /= (double)n_hits;
        if (n_miss > 0) mm /= (double)n_miss;
        out64[f] = mm - hh;
    }
    free(h); free(m); free(drow); free(ti);
}

/* x: n rows of ld bytes (uint8 or int8 codes: only equality is used). */
FSO_API int fso_multisurf_targets_u8(const uint8_t *x, int64_t n, int64_t ld, const int64_t *cols, int64_t p,
                                     const int64_t *y, int use_star, const int64_t *targets, int64_t nt,
                                     double *wsum_out, double *thresh_out, int8_t *mask_out, double *dist_out) {
    if (n < 2 || p < 1) return -1;
    double *rows = malloc((size_t)nt * p * sizeof(double));
    if (!rows) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nt; ++t)
        multisurf_target_u8(x, n, ld, cols, p, y, use_star, targets[t], rows + t * p,
                            thresh_out ? thresh_out + t : NULL, mask_out ? mask_out + t * n : NULL,
                            dist_out ? dist_out + t * n : NULL);
    for (int64_t f = 0; f < p; ++f) {
        double s = 0.0;
        for (int64_t t = 0; t < nt; ++t) s += rows[t * p + f];
        wsum_out[f] = s;
    }
    free(rows);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* SURF / SURF*  (SURF.py:131-218)                                          */
/*
 * sum_mode 0: float32 running sum of the float32 distances in j order -- the
 *             literal reading of np.sum(dists_from_i) (SURF.py:162); the
 *             reference's actual order is whatever LLVM vectorises to.
 * sum_mode 1: the same sum taken in float64 and rounded once to float32
 *             (order-independent).
 * sum_mode 2: the order LLVM actually emits on an AVX2 host (see
 *             sum_f32_avx2_order); this is what the reference computes here.
 */
/* np.sum(float32[n]) as LLVM vectorises it for an AVX2 host under fastmath
 * (observed in the JIT'd code, numba 0.65): 32 interleaved float32 partial sums
 * over the leading multiple of 32, folded 32->8->4->2->1; then a 4-wide loop
 * whose lane 0 starts from that value; then a scalar tail. */
static float sum_f32_avx2_order(const float *a, int64_t n) {
    float s = 0.0f;
    int64_t done = 0;
    if (n >= 32) {
        float p[32];
        for (int l = 0; l < 32; ++l) p[l] = 0.0f;
        int64_t m = n & ~(int64_t)31;
        for (int64_t t = 0; t < m; t += 32)
            for (int l = 0; l < 32; ++l) p[l] += a[t + l];
        float q[8], r[4], u[2];
        for (int l = 0; l < 8; ++l) q[l] = (p[l] + p[l + 8]) + (p[l + 24] + p[l + 16]);
        for (int l = 0; l < 4; ++l) r[l] = q[l + 4] + q[l];
        for (int l = 0; l < 2; ++l) u[l] = r[l + 2] + r[l];
        s = u[1] + u[0];
        done = m;
    }
    if (n - done >= 4) {
        float v[4] = {s, 0.0f, 0.0f, 0.0f};
        int64_t m = done + ((n - done) & ~(int64_t)3);
        for (int64_t t = done; t < m; t += 4)
            for (int l = 0; l < 4; ++l) v[l] += a[t + l];
        float u0 = v[2] + v[0], u1 = v[3] + v[1];
        s = u1 + u0;
        done = m;
    }
    for (int64_t t = done; t < n; ++t) s += a[t];
    return s;
}

static void surf_target(const double *x, int64_t n, int64_t p, const int32_t *y,
                        const float *recip, const uint8_t *isd, int use_star, int sum_mode,
                        int64_t i, float *out32, double *out64,
                        double *thresh_out, int8_t *mask_row, double *dist_row) {
    const double *xi = x + i * p;
    float *dists = malloc((size_t)n * sizeof(float));
    /* SURF.py:146-160; d_ii = 0 is part of the sum (:148) */
    for (int64_t j = 0; j < n; ++j)
        dists[j] = (j == i) ? 0.0f : (float)dist_f64(xi, x + j * p, p, recip, isd);
    float sum_d;
    if (sum_mode == 0) {
        sum_d = 0.0f;
        for (int64_t j = 0; j < n; ++j) sum_d += dists[j];
    } else if (sum_mode == 2) {
        sum_d = sum_f32_avx2_order(dists, n);
    } else {
        double s = 0.0;
        for (int64_t j = 0; j < n; ++j) s += (double)dists[j];
        sum_d = (float)s;
    }
    /* :163 float32 / int64 -> float64; compiled as a multiply by 1/(n-1) */
    double avg = (double)sum_d * (1.0 / (double)(n - 1));
    if (thresh_out) *thresh_out = avg;

    float *acc32 = NULL; double *acc64 = NULL;             /* [4][p]: nh, nm, fh, fm */
    if (out32) acc32 = calloc((size_t)4 * p, sizeof(float));
    if (out64) acc64 = calloc((size_t)4 * p, sizeof(double));
    for (int64_t j = 0; j < n; ++j) {
        if (mask_row) mask_row[j] = FSO_NONE;
        if (dist_row) dist_row[j] = (double)dists[j];
        if (j == i) continue;
        int is_hit = (y[i] == y[j]);
        int is_near = ((double)dists[j] < avg);            /* :176 */
        int slot;
        if (is_near) slot = is_hit ? 0 : 1;
        else if (use_star) slot = is_hit ? 2 : 3;
        else continue;
        if (mask_row) {
            static const int8_t codes[4] = {FSO_NEAR_HIT, FSO_NEAR_MISS, FSO_FAR_HIT, FSO_FAR_MISS};
            mask_row[j] = codes[slot];
        }
        const double *xj = x + j * p;
        for (int64_t f = 0; f < p; ++f) {
            float t = (float)diff_f64(xi, xj, f, recip, isd);   /* diffs_from_i is float32 (:144,158) */
            if (out32) acc32[slot * p + f] += t;
            if (out64) acc64[slot * p + f] += (double)t;
        }
    }
    /* SURF.py:191-193 */
    for (int64_t f = 0; f < p; ++f) {
        if (out32) {
            float u = acc32[1 * p + f] - acc32[0 * p + f];
            if (use_star) u += (acc32[2 * p + f] - acc32[3 * p + f]);
            out32[f] = u;
        }
        if (out64) {
            double u = acc64[1 * p + f] - acc64[0 * p + f];
            if (use_star) u += (acc64[2 * p + f] - acc64[3 * p + f]);
            out64[f] = u;
        }
    }
    free(dists); free(acc32); free(acc64);
}

/* Full pipeline with one accumulation thread: private_scores[0] += update in
 * row order (SURF.py:195), / n_samples (:218; float32 array / int). */
FSO_API int fso_surf_scores(const double *x, int64_t n, int64_t p, const int32_t *y,
                            const float *recip, const uint8_t *isd, int use_star, int sum_mode,
                            float *scores_out) {
    if (n < 2 || p < 1) return -1;
    float *temp = malloc((size_t)n * p * sizeof(float));
    if (!temp) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i)
        surf_target(x, n, p, y, recip, isd, use_star, sum_mode, i, temp + i * p, NULL,
                    NULL, NULL, NULL);
    for (int64_t f = 0; f < p; ++f) {
        float s = 0.0f;
        for (int64_t i = 0; i < n; ++i) s += temp[i * p + f];
        scores_out[f] = s / (float)n;
    }
    free(temp);
    return 0;
}

FSO_API int fso_surf_targets(const double *x, int64_t n, int64_t p, const int32_t *y,
                             const float *recip, const uint8_t *isd, int use_star, int sum_mode,
                             const int64_t *targets, int64_t nt,
                             double *wsum_out, double *thresh_out, int8_t *mask_out,
                             double *dist_out) {
    if (n < 2 || p < 1) return -1;
    double *rows = malloc((size_t)nt * p * sizeof(double));
    if (!rows) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nt; ++t)
        surf_target(x, n, p, y, recip, isd, use_star, sum_mode, targets[t], NULL, rows + t * p,
                    thresh_out ? thresh_out + t : NULL,
                    mask_out ? mask_out + t * n : NULL,
                    dist_out ? dist_out + t * n : NULL);
    for (int64_t f = 0; f < p; ++f) {
        double s = 0.0;
        for (int64_t t = 0; t < nt; ++t) s += rows[t * p + f];
        wsum_out[f] = s;
    }
    free(rows);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* numba np.argsort restated (numba/misc/quicksort.py, is_argsort=True,      */
/* lt = a < b; no NaNs reach it).  R is the index permutation.               */
#define SMALL_QUICKSORT 15
#define MAX_STACK 100

static void nb_insertion_sort(const float *A, int64_t *R, int64_t low, int64_t high) {
    if (high <= low) return;
    for (int64_t i = low + 1; i <= high; ++i) {
        int64_t k = R[i];
        float v = A[k];
        int64_t j = i;
        while (j > low && v < A[R[j - 1]]) { R[j] = R[j - 1]; --j; }
        R[j] = k;
    }
}

static int64_t nb_partition(const float *A, int64_t *R, int64_t low, int64_t high) {
    int64_t mid = (low + high) >> 1, t;
#define SWAP(a, b) do { t = R[a]; R[a] = R[b]; R[b] = t; } while (0)
    if (A[R[mid]] < A[R[low]]) SWAP(low, mid);
    if (A[R[high]] < A[R[mid]]) SWAP(high, mid);
    if (A[R[mid]] < A[R[low]]) SWAP(low, mid);
    float pivot = A[R[mid]];
    SWAP(high, mid);
    int64_t i = low, j = high - 1;
    for (;;) {
        while (i < high && A[R[i]] < pivot) ++i;
        while (j >= low && pivot < A[R[j]]) --j;
        if (i >= j) break;
        SWAP(i, j);
        ++i; --j;
    }
    SWAP(i, high);
#undef SWAP
    return i;
}

FSO_API void fso_argsort_numba(const float *A, int64_t n, int64_t *R) {
    for (int64_t i = 0; i < n; ++i) R[i] = i;
    if (n < 2) return;
    int64_t stack_lo[MAX_STACK], stack_hi[MAX_STACK];
    int sp = 1;
    stack_lo[0] = 0; stack_hi[0] = n - 1;
    while (sp > 0) {
        --sp;
        int64_t low = stack_lo[sp], high = stack_hi[sp];
        while (high - low >= SMALL_QUICKSORT) {
            int64_t i = nb_partition(A, R, low, high);
            if (high - i > i - low) {
                if (high > i) { stack_lo[sp] = i + 1; stack_hi[sp] = high; ++sp; }
                high = i - 1;
            } else {
                if (i > low) { stack_lo[sp] = low; stack_hi[sp] = i - 1; ++sp; }
                low = i + 1;
            }
        }
        nb_insertion_sort(A, R, low, high);
    }
}

/* stable order by (distance, index): the GPU path's documented tie rule */
typedef struct { float d; int64_t j; } fso_dj;
static int cmp_dj(const void *a, const void *b) {
    const fso_dj *u = a, *v = b;
    if (u->d < v->d) return -1;
    if (u->d > v->d) return 1;
    return (u->j > v->j) - (u->j < v->j);
}

/* ------------------------------------------------------------------------ */
/* ReliefF  (ReliefF.py:137-220)                                             */
/* tie_mode 0: numba quicksort order (the reference); 1: (distance, index). */
static void relieff_target(const float *x, int64_t n, int64_t p, const int32_t *y_enc,
                           const float *recip, const uint8_t *isd, int32_t k,
                           const float *class_probs, int32_t n_classes, int tie_mode,
                           int64_t i, float *out32, double *out64,
                           int8_t *mask_row, double *dist_row) {
    const float *xi = x + i * p;
    float *dists = malloc((size_t)n * sizeof(float));
    int64_t *order = malloc((size_t)n * sizeof(int64_t));
    for (int64_t j = 0; j < n; ++j)                                   /* :144-155 */
        dists[j] = (j == i) ? INFINITY : (float)dist_f32(xi, x + j * p, p, recip, isd);
    if (tie_mode == 0) {
        fso_argsort_numba(dists, n, order);                           /* :157 */
    } else {
        fso_dj *dj = malloc((size_t)n * sizeof(fso_dj));
        for (int64_t j = 0; j < n; ++j) { dj[j].d = dists[j]; dj[j].j = j; }
        qsort(dj, (size_t)n, sizeof(fso_dj), cmp_dj);
        for (int64_t j = 0; j < n; ++j) order[j] = dj[j].j;
        free(dj);
    }
    int32_t lbl_i = y_enc[i];
    int32_t *hits = malloc((size_t)k * sizeof(int32_t));
    int32_t *misses = malloc((size_t)n_classes * k * sizeof(int32_t));
    int32_t *m_found = calloc((size_t)n_classes, sizeof(int32_t));
    int32_t h_found = 0;
    /* :164-175.  The scan stops when hits and EVERY class (the target's own
     * included, which never fills) have k entries, i.e. it runs to the end;
     * order[] includes i itself (distance inf, a "hit" of its own class). */
    for (int64_t q = 0; q < n; ++q) {
        int64_t idx = order[q];
        int32_t lbl = y_enc[idx];
        if (lbl == lbl_i) {
            if (h_found < k) hits[h_found++] = (int32_t)idx;
        } else if (m_found[lbl] < k) {
            misses[lbl * k + m_found[lbl]++] = (int32_t)idx;
        }
        int all = (h_found == k);
        for (int32_t c = 0; c < n_classes && all; ++c) all = (m_found[c] >= k);
        if (all) break;
    }
    if (mask_row) {
        memset(mask_row, 0, (size_t)n);
        for (int32_t q = 0; q < h_found; ++q) mask_row[hits[q]] = FSO_NEAR_HIT;
        for (int32_t c = 0; c < n_classes; ++c)
            for (int32_t q = 0; q < m_found[c]; ++q) mask_row[misses[c * k + q]] = FSO_NEAR_MISS;
    }
    if (dist_row) for (int64_t j = 0; j < n; ++j) dist_row[j] = (j == i) ? 0.0 : (double)dists[j];

    double denom = 1.0 - (double)class_probs[lbl_i];                  /* :177-179 */
    if (denom == 0) denom = 1.0;
    for (int64_t f = 0; f < p; ++f) {                                 /* :181-216 */
        double hit_sum = 0.0;
        for (int32_t q = 0; q < h_found; ++q)
            hit_sum += diff_f32(xi, x + (int64_t)hits[q] * p, f, recip, isd);
        double miss_sum = 0.0;
        for (int32_t c = 0; c < n_classes; ++c) {
            if (c == lbl_i) continue;
            double weight = (double)class_probs[c] / denom;
            double cur = 0.0;
            for (int32_t q = 0; q < m_found[c]; ++q)
                cur += diff_f32(xi, x + (int64_t)misses[c * k + q] * p, f, recip, isd);
            miss_sum += weight * cur;
        }
        double update = 0.0;
        if (h_found > 0) update -= hit_sum / (double)h_found;
        if (k > 0) update += miss_sum / (double)k;                    /* :213-214 divides by k */
        if (out32) out32[f] = (float)update;
        if (out64) out64[f] = update;
    }
    free(dists); free(order); free(hits); free(misses); free(m_found);
}

FSO_API int fso_relieff_scores(const float *x, int64_t n, int64_t p, const int32_t *y_enc,
                               const float *recip, const uint8_t *isd, int32_t k,
                               const float *class_probs, int32_t n_classes, int tie_mode,
                               float *scores_out) {
    if (n < 2 || p < 1 || k < 1) return -1;
    float *temp = malloc((size_t)n * p * sizeof(float));
    if (!temp) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i)
        relieff_target(x, n, p, y_enc, recip, isd, k, class_probs, n_classes, tie_mode, i,
                       temp + i * p, NULL, NULL, NULL);
    for (int64_t f = 0; f < p; ++f) {                                 /* :219-220, :236 */
        float s = 0.0f;
        for (int64_t i = 0; i < n; ++i) s += temp[i * p + f];
        scores_out[f] = s / (float)n;
    }
    free(temp);
    return 0;
}

FSO_API int fso_relieff_targets(const float *x, int64_t n, int64_t p, const int32_t *y_enc,
                                const float *recip, const uint8_t *isd, int32_t k,
                                const float *class_probs, int32_t n_classes, int tie_mode,
                                const int64_t *targets, int64_t nt,
                                double *wsum_out, int8_t *mask_out, double *dist_out) {
    if (n < 2 || p < 1 || k < 1) return -1;
    double *rows = malloc((size_t)nt * p * sizeof(double));
    if (!rows) return -2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < nt; ++t)
        relieff_target(x, n, p, y_enc, recip, isd, k, class_probs, n_classes, tie_mode,
                       targets[t], NULL, rows + t * p,
                       mask_out ? mask_out + t * n : NULL,
                       dist_out ? dist_out + t * n : NULL);
    for (int64_t f = 0; f < p; ++f) {
        double s = 0.0;
        for (int64_t t = 0; t < nt; ++t) s += rows[t * p + f];
        wsum_out[f] = s;
    }
    free(rows);
    return 0;
}

/* ======================================================================== */
/* feature x feature joint-count tables and the     */
/* two statistics the reference derives from them.                           */
/*   mutual information  -- mutual_information.py:26-46 (_mi_pair_cpu) and   */
/*                          :49-63 (_batch_mi_cpu: relevance + redundancy)   */
/*   symmetrical uncertainty -- CFS.py:26-77 (_entropy, _mutual_information, */
/*                          _symmetrical_uncertainty) and :81-104            */
/* The count tables are integers and restated exactly.  The statistics are   */
/* restated in float64.  The reference computes MI in float64 too (numba     */
/* fastmath only reassociates: golden vectors agree to 1e-15); its SU path   */
/* keeps the probabilities and log2 terms in float32 (numba types p, p_xy,   */
/* p_x, p_y as float32 arrays), which puts +-1e-7 of rounding noise on values */
/* in [0, 1] -- the golden tests allow 1e-6 absolute for SU.                 */
/* x is row-major [n, p] non-negative codes; states that never occur add 0.  */
/* ------------------------------------------------------------------------ */
#define FSO_MAX_STATES 64

/* table[a * kb + b] = #{i : xa[i] = a, xb[i] = b}  (mutual_information.py:31-33, CFS.py:51-53) */
FSO_API int fso_joint_counts(const int32_t *xa, const int32_t *xb, int64_t n, int32_t ka, int32_t kb,
                             int64_t *table) {
    if (n < 1 || ka < 1 || kb < 1 || ka > FSO_MAX_STATES || kb > FSO_MAX_STATES) return -1;
    memset(table, 0, (size_t)ka * (size_t)kb * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        if (xa[i] < 0 || xa[i] >= ka || xb[i] < 0 || xb[i] >= kb) return -1;
        table[(int64_t)xa[i] * kb + xb[i]] += 1;
    }
    return 0;
}

/* mutual_information.py:35-46 on an integer count table */
static double mi_from_table(const int64_t *t, int ka, int kb, int64_t n, double log_base) {
    double p1[FSO_MAX_STATES], p2[FSO_MAX_STATES];
    for (int a = 0; a < ka; ++a) p1[a] = 0.0;
    for (int b = 0; b < kb; ++b) p2[b] = 0.0;
    for (int a = 0; a < ka; ++a)
        for (int b = 0; b < kb; ++b) {
            const double pxy = (double)t[a * kb + b] / (double)n;      /* :35 table /= n */
            p1[a] += pxy;                                              /* :36 */
            p2[b] += pxy;                                              /* :37 */
        }
    double mi = 0.0;
    const double eps = 1e-12;
    for (int a = 0; a < ka; ++a)
        for (int b = 0; b < kb; ++b) {
            const double pxy = (double)t[a * kb + b] / (double)n;
            if (pxy > eps) mi += pxy * log(pxy / (p1[a] * p2[b] + eps));   /* :44-45 */
        }
    return mi / log_base;                                                  /* :46 */
}

static double entropy_counts(const int64_t *c, int k, int64_t n) {          /* CFS.py:26-41 */
    double e = 0.0;
    for (int s = 0; s < k; ++s) {
        const double pr = (double)c[s] / (double)n;
        if (pr > 1e-12) e -= pr * log2(pr);
    }
    return e;
}

/* CFS.py:44-77 on an integer count table */
static double su_from_table(const int64_t *t, int ka, int kb, int64_t n) {
    int64_t ca[FSO_MAX_STATES], cb[FSO_MAX_STATES];
    double pa[FSO_MAX_STATES], pb[FSO_MAX_STATES];
    for (int a = 0; a < ka; ++a) { ca[a] = 0; pa[a] = 0.0; }
    for (int b = 0; b < kb; ++b) { cb[b] = 0; pb[b] = 0.0; }
    for (int a = 0; a < ka; ++a)
        for (int b = 0; b < kb; ++b) {
            ca[a] += t[a * kb + b];
            cb[b] += t[a * kb + b];
            const double pxy = (double)t[a * kb + b] / (double)n;
            pa[a] += pxy;                                                   /* :56 */
            pb[b] += pxy;                                                   /* :57 */
        }
    const double hx = entropy_counts(ca, ka, n), hy = entropy_counts(cb, kb, n);
    if (hx + hy < 1e-12) return 0.0;                                        /* :73-74 */
    double mi = 0.0;
    for (int a = 0; a < ka; ++a)
        for (int b = 0; b < kb; ++b) {
            const double pxy = (double)t[a * kb + b] / (double)n;
            if (pxy > 1e-12 && pa[a] > 1e-12 && pb[b] > 1e-12) mi += pxy * log2(pxy / (pa[a] * pb[b]));   /* :62-63 */
        }
    return 2.0 * mi / (hx + hy);                                            /* :77 */
}

/* columns of x (and y as column p) as contiguous int32 vectors + their state counts */
static int32_t *columns_of(const int32_t *x, int64_t n, int64_t p, const int32_t *y, int32_t *k_out) {
    int32_t *cols = malloc((size_t)(p + 1) * n * sizeof(int32_t));
    if (!cols) return NULL;
    for (int64_t f = 0; f <= p; ++f) {
        int32_t mx = 0;
        for (int64_t i = 0; i < n; ++i) {
            const int32_t v = f < p ? x[i * p + f] : y[i];
            if (v < 0 || v >= FSO_MAX_STATES) { free(cols); return NULL; }
            cols[f * n + i] = v;
            mx = v > mx ? v : mx;
        }
        k_out[f] = mx + 1;                     /* mutual_information.py:28-29: k = max + 1 */
    }
    return cols;
}

/* kind 0: MI (relevance / redundancy in units of log_base), kind 1: SU (r_cf / r_ff).
 * vec_out[p] = statistic(feature f, y); mat_out[p*p] symmetric with a zero diagonal
 * (mutual_information.py:52-63, CFS.py:87-104).  mat_out may be NULL. */
FSO_API int fso_joint_matrices(const int32_t *x, int64_t n, int64_t p, const int32_t *y, int kind,
                               double log_base, double *vec_out, double *mat_out) {
    if (n < 1 || p < 1 || (kind != 0 && kind != 1)) return -1;
    int32_t *k = malloc((size_t)(p + 1) * sizeof(int32_t));
    if (!k) return -2;
    int32_t *cols = columns_of(x, n, p, y, k);
    if (!cols) { free(k); return -1; }
    if (mat_out) memset(mat_out, 0, (size_t)p * p * sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t f = 0; f < p; ++f) {
        int64_t t[FSO_MAX_STATES * FSO_MAX_STATES];
        fso_joint_counts(cols + f * n, cols + p * n, n, k[f], k[p], t);
        vec_out[f] = kind == 0 ? mi_from_table(t, k[f], k[p], n, log_base) : su_from_table(t, k[f], k[p], n);
        if (!mat_out) continue;
        for (int64_t g = f + 1; g < p; ++g) {
            fso_joint_counts(cols + f * n, cols + g * n, n, k[f], k[g], t);
            const double v = kind == 0 ? mi_from_table(t, k[f], k[g], n, log_base) : su_from_table(t, k[f], k[g], n);
            mat_out[f * p + g] = v;
            mat_out[g * p + f] = v;
        }
    }
    free(cols);
    free(k);
    return 0;
}
