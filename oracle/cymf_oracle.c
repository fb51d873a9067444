/* cymf_oracle.c -- CPU restatement of minatosato/cymf's factor-update hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cymf_b200/ may link, import or call this file; it is the
 * checker that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg compare the CUDA path
 * against.  Every function names the reference file:line (paths relative to /root/reference) whose
 * arithmetic it restates.  All arithmetic is IEEE double, evaluated in the reference's own order
 * (single thread, num_threads=1 semantics), compiled with -ffp-contract=off so that no FMA is formed,
 * matching the reference's plain x86-64 build.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks every entry point against vectors produced by
 * the compiled reference itself (oracle/_ref, built by oracle/build_ref.py; generator script
 * tests/golden/make_golden.py), plus the libstdc++-13 RNG known-answer vector of SURVEY.md section 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * RNG: std::mt19937(seed) + std::uniform_int_distribution<long>(0, n-1)   (cymf/math.pyx:12-18,
 * cymf/math.pxd:31-39).  libstdc++ >= 11 maps a 32-bit engine onto [0,n) with Lemire's
 * multiply-shift and rejection of the biased low words.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mt[624];
    int idx;
} oracle_rng;

void oracle_rng_seed(oracle_rng *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int t = 1; t < 624; ++t)
        g->mt[t] = 1812433253u * (g->mt[t - 1] ^ (g->mt[t - 1] >> 30)) + (uint32_t)t;
    g->idx = 624;
}

static void rng_refill(oracle_rng *g) {
    uint32_t *mt = g->mt;
    for (int t = 0; t < 624; ++t) {
        uint32_t y = (mt[t] & 0x80000000u) | (mt[(t + 1) % 624] & 0x7fffffffu);
        uint32_t v = mt[(t + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        mt[t] = v;
    }
    g->idx = 0;
}

uint32_t oracle_rng_u32(oracle_rng *g) {
    if (g->idx >= 624) rng_refill(g);
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* one draw of uniform_int_distribution<long>(0, n-1); n < 2^32 */
int64_t oracle_rng_below(oracle_rng *g, uint32_t n) {
    uint64_t prod = (uint64_t)oracle_rng_u32(g) * (uint64_t)n;
    uint32_t low = (uint32_t)prod;
    if (low < n) {
        uint32_t threshold = (uint32_t)(0u - n) % n;
        while (low < threshold) {
            prod = (uint64_t)oracle_rng_u32(g) * (uint64_t)n;
            low = (uint32_t)prod;
        }
    }
    return (int64_t)(prod >> 32);
}

/* one draw of uniform_int_distribution<long>(0, n-1) for any n >= 1 (libstdc++ >= 11, bits/uniform_int_dist.h):
 * ranges narrower than the 32-bit engine use the multiply-shift map above; a range of exactly 2^32 takes the raw
 * word; wider ranges draw the high part recursively over [0, (n-1) / 2^32], add a raw low word and redraw
 * while the sum exceeds n-1 (RelMF draws cells from [0, U*I), cymf/relmf.pyx:127). */
int64_t oracle_rng_below64(oracle_rng *g, uint64_t n) {
    const uint64_t urange = n - 1;
    if (urange < 0xffffffffull) return oracle_rng_below(g, (uint32_t)n);
    if (urange == 0xffffffffull) return (int64_t)oracle_rng_u32(g);
    uint64_t ret, tmp;
    do {
        tmp = 0x100000000ull * (uint64_t)oracle_rng_below64(g, urange / 0x100000000ull + 1);
        ret = tmp + (uint64_t)oracle_rng_u32(g);
    } while (ret > urange || ret < tmp);
    return (int64_t)ret;
}
void oracle_rng_fill_below64(oracle_rng *g, uint64_t n, int64_t *out, int64_t count) {
    for (int64_t t = 0; t < count; ++t) out[t] = oracle_rng_below64(g, n);
}

oracle_rng *oracle_rng_new(uint32_t seed) {
    oracle_rng *g = (oracle_rng *)malloc(sizeof(oracle_rng));
    if (g) oracle_rng_seed(g, seed);
    return g;
}
void oracle_rng_free(oracle_rng *g) { free(g); }
void oracle_rng_fill_u32(oracle_rng *g, uint32_t *out, int64_t count) {
    for (int64_t t = 0; t < count; ++t) out[t] = oracle_rng_u32(g);
}
void oracle_rng_fill_below(oracle_rng *g, uint32_t n, int32_t *out, int64_t count) {
    for (int64_t t = 0; t < count; ++t) out[t] = (int32_t)oracle_rng_below(g, n);
}

/* std::set<int>::find on a user's positives (cymf/bpr.pyx:140,146-147,166) == membership in the
 * sorted CSR row. */
static int row_contains(const int32_t *indices, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        int32_t v = indices[mid];
        if (v == key) return 1;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * BPR   (cymf/bpr.pyx:140-171, cymf/model.pyx:47-87, cymf/optimizer.pyx:52-58,66-82,127-160)
 *   optimizer: 0 = sgd, 1 = adagrad (accumulators start at ONE), 2 = adam (no timestep, constant
 *   1/(1-beta) correction).  Optimizer state is rebuilt on every call, as every fit() does.
 *   negatives_in (num_epochs*N, may be NULL) replaces the mt19937 stream with a caller-supplied one
 *   (used to check the Hogwild kernel, whose negatives come from Philox, one triplet at a time).
 *   negatives_out / applied_out (each num_epochs*N, may be NULL) record the triplet stream so the
 *   CUDA replay kernel can be driven with exactly the reference's (u, i, j, skip) sequence.
 *   loss_out (num_epochs, may be NULL): accum_loss / N of bpr.pyx:168-171.
 * ---------------------------------------------------------------------------------------------- */
int oracle_bpr_fit(double *W, double *H, int32_t U, int32_t I, int32_t K,
                   const int32_t *users, const int32_t *positives, int64_t N,
                   const int32_t *indptr, const int32_t *indices,
                   int32_t num_epochs, double lr, double wd, int32_t optimizer, uint32_t seed,
                   const int32_t *negatives_in, int32_t *negatives_out, uint8_t *applied_out, double *loss_out) {
    const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    double *sW1 = NULL, *sH1 = NULL, *sW2 = NULL, *sH2 = NULL;
    size_t nW = (size_t)U * K, nH = (size_t)I * K;
    if (optimizer == 1) {
        sW1 = (double *)malloc(nW * sizeof(double));
        sH1 = (double *)malloc(nH * sizeof(double));
        if (!sW1 || !sH1) return -1;
        for (size_t t = 0; t < nW; ++t) sW1[t] = 1.0;
        for (size_t t = 0; t < nH; ++t) sH1[t] = 1.0;
    } else if (optimizer == 2) {
        sW1 = (double *)calloc(nW, sizeof(double));
        sH1 = (double *)calloc(nH, sizeof(double));
        sW2 = (double *)calloc(nW, sizeof(double));
        sH2 = (double *)calloc(nH, sizeof(double));
        if (!sW1 || !sH1 || !sW2 || !sH2) return -1;
    }
    oracle_rng gen;
    oracle_rng_seed(&gen, seed);                                   /* bpr.pyx:141 */
    for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
        double accum = 0.0;
        for (int64_t l = 0; l < N; ++l) {
            int32_t u = users[l], i = positives[l];
            int32_t j = negatives_in ? negatives_in[(int64_t)epoch * N + l]
                                     : (int32_t)oracle_rng_below(&gen, (uint32_t)I);   /* bpr.pyx:165 */
            int hit = row_contains(indices, indptr[u], indptr[u + 1], j);   /* bpr.pyx:166 */
            if (negatives_out) negatives_out[(int64_t)epoch * N + l] = j;
            if (applied_out) applied_out[(int64_t)epoch * N + l] = (uint8_t)!hit;
            if (hit) continue;
            double *wu = W + (size_t)u * K, *hi = H + (size_t)i * K, *hj = H + (size_t)j * K;
            double x = 0.0, l2 = 0.0;
            for (int32_t k = 0; k < K; ++k) {                                /* model.pyx:55-57 */
                x += wu[k] * (hi[k] - hj[k]);
                l2 += wu[k] * wu[k] + hi[k] * hi[k] + hj[k] * hj[k];
            }
            accum += -log(1.0 / (1.0 + exp(-x))) + wd * l2;                 /* model.pyx:59 */
            double s = 1.0 / (1.0 + exp(x));                                /* model.pyx:78 */
            for (int32_t k = 0; k < K; ++k) {
                double gw = -(s * (hi[k] - hj[k]) - wd * wu[k]);            /* model.pyx:81-83 */
                double gi = -(s * wu[k] - wd * hi[k]);
                double gj = -(s * (-wu[k]) - wd * hj[k]);
                if (optimizer == 0) {                                       /* optimizer.pyx:52-58 */
                    wu[k] -= lr * gw;
                    hi[k] -= lr * gi;
                    hj[k] -= lr * gj;
                } else if (optimizer == 1) {                                /* optimizer.pyx:74-82 */
                    double *aw = sW1 + (size_t)u * K + k, *ai = sH1 + (size_t)i * K + k,
                           *aj = sH1 + (size_t)j * K + k;
                    *aw += gw * gw; wu[k] -= lr * gw / sqrt(*aw);
                    *ai += gi * gi; hi[k] -= lr * gi / sqrt(*ai);
                    *aj += gj * gj; hj[k] -= lr * gj / sqrt(*aj);
                } else {                                                    /* optimizer.pyx:150-160 */
                    double *mw = sW1 + (size_t)u * K + k, *vw = sW2 + (size_t)u * K + k;
                    double *mi = sH1 + (size_t)i * K + k, *vi = sH2 + (size_t)i * K + k;
                    double *mj = sH1 + (size_t)j * K + k, *vj = sH2 + (size_t)j * K + k;
                    *mw = beta1 * *mw + (1 - beta1) * gw;
                    *vw = beta2 * *vw + (1 - beta2) * (gw * gw);
                    wu[k] -= lr * (*mw / (1 - beta1)) / (sqrt(*vw / (1 - beta2)) + eps);
                    *mi = beta1 * *mi + (1 - beta1) * gi;
                    *vi = beta2 * *vi + (1 - beta2) * (gi * gi);
                    hi[k] -= lr * (*mi / (1 - beta1)) / (sqrt(*vi / (1 - beta2)) + eps);
                    *mj = beta1 * *mj + (1 - beta1) * gj;
                    *vj = beta2 * *vj + (1 - beta2) * (gj * gj);
                    hj[k] -= lr * (*mj / (1 - beta1)) / (sqrt(*vj / (1 - beta2)) + eps);
                }
            }
        }
        if (loss_out) loss_out[epoch] = accum / (double)N;                  /* bpr.pyx:171 */
    }
    free(sW1); free(sH1); free(sW2); free(sH2);
    return 0;
}

/* position of `key` in the sorted run indices[lo, hi), or -1 */
static int64_t row_find(const int32_t *indices, int64_t lo, int64_t hi, int32_t key) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        int32_t v = indices[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return -1;
}

/* one element of the reference's optimizers (cymf/optimizer.pyx:52-58, 74-82, 150-160) */
static void opt_update(double *theta, double g, double lr, int32_t optimizer, double *s1, double *s2) {
    if (optimizer == 0) {
        *theta -= lr * g;
    } else if (optimizer == 1) {
        *s1 += g * g;
        *theta -= lr * g / sqrt(*s1);
    } else {
        const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
        *s1 = beta1 * *s1 + (1 - beta1) * g;
        *s2 = beta2 * *s2 + (1 - beta2) * (g * g);
        *theta -= lr * (*s1 / (1 - beta1)) / (sqrt(*s2 / (1 - beta2)) + eps);
    }
}

/* ------------------------------------------------------------------------------------------------
 * RelMF   (cymf/relmf.pyx:107-148, cymf/model.pyx:89-142; optimizers as for BPR)
 *   Every epoch draws U*I cells r ~ U[0, U*I) from one mt19937(seed) that lives for the whole fit,
 *   u = r / I, i = r % I (relmf.pyx:144-146); the label X[u,i] is read from the dense matrix in the
 *   reference and from the CSR here (absent cell = 0.0, `values` NULL = all stored cells are 1.0).
 *   t = W_u . H_i (k ascending); c = X[u,i] / max(p_i, clip);
 *   g_w = -(c (1 - t) H_ik + (1 - c)(0 - t) H_ik) + wd W_uk, g_h symmetric, both from pre-update
 *   values (model.pyx:129-141).
 *   cells_in (num_epochs*n, may be NULL) replaces the mt19937 stream; cells_out records it;
 *   loss_out (num_epochs, may be NULL) = accum_loss of relmf.pyx:150-152 (sum, not mean).
 * ---------------------------------------------------------------------------------------------- */
int oracle_relmf_fit(double *W, double *H, int32_t U, int32_t I, int32_t K,
                     const int32_t *indptr, const int32_t *indices, const double *values,
                     const double *propensities, int64_t n_samples,
                     int32_t num_epochs, double lr, double wd, double clip, int32_t optimizer, uint32_t seed,
                     const int64_t *cells_in, int64_t *cells_out, double *loss_out) {
    double *sW1 = NULL, *sH1 = NULL, *sW2 = NULL, *sH2 = NULL;
    size_t nW = (size_t)U * K, nH = (size_t)I * K;
    if (optimizer == 1) {
        sW1 = (double *)malloc(nW * sizeof(double));
        sH1 = (double *)malloc(nH * sizeof(double));
        if (!sW1 || !sH1) return -1;
        for (size_t t = 0; t < nW; ++t) sW1[t] = 1.0;
        for (size_t t = 0; t < nH; ++t) sH1[t] = 1.0;
    } else if (optimizer == 2) {
        sW1 = (double *)calloc(nW, sizeof(double));
        sH1 = (double *)calloc(nH, sizeof(double));
        sW2 = (double *)calloc(nW, sizeof(double));
        sH2 = (double *)calloc(nH, sizeof(double));
        if (!sW1 || !sH1 || !sW2 || !sH2) return -1;
    }
    oracle_rng gen;
    oracle_rng_seed(&gen, seed);                                              /* relmf.pyx:127 */
    const uint64_t cells = (uint64_t)U * (uint64_t)I;
    for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
        double accum = 0.0;
        for (int64_t l = 0; l < n_samples; ++l) {
            int64_t r = cells_in ? cells_in[(int64_t)epoch * n_samples + l]
                                 : oracle_rng_below64(&gen, cells);         /* relmf.pyx:144 */
            if (cells_out) cells_out[(int64_t)epoch * n_samples + l] = r;
            int32_t u = (int32_t)(r / I), i = (int32_t)(r % I);               /* relmf.pyx:145-146 */
            int64_t pos = row_find(indices, indptr[u], indptr[u + 1], i);
            double x = pos < 0 ? 0.0 : (values ? values[pos] : 1.0);          /* X[u, i] */
            double p = propensities[i];
            double pm = p >= clip ? p : clip;                                 /* dmax, math.pxd:47-51 */
            double *wu = W + (size_t)u * K, *hi = H + (size_t)i * K;
            double t = 0.0, l2 = 0.0;
            for (int32_t k = 0; k < K; ++k) {                                 /* model.pyx:114-116 */
                t += wu[k] * hi[k];
                l2 += wu[k] * wu[k] + hi[k] * hi[k];
            }
            accum += (x / pm) * ((1. - t) * (1. - t)) + (1 - x / pm) * (t * t) + wd * l2;   /* model.pyx:118 */
            for (int32_t k = 0; k < K; ++k) {                                 /* model.pyx:129-141 */
                double gw = -((x / pm) * (1. - t) * hi[k] + (1 - x / pm) * (0. - t) * hi[k]) + wd * wu[k];
                double gh = -((x / pm) * (1. - t) * wu[k] + (1 - x / pm) * (0. - t) * wu[k]) + wd * hi[k];
                size_t ow = (size_t)u * K + k, oh = (size_t)i * K + k;
                opt_update(wu + k, gw, lr, optimizer, sW1 ? sW1 + ow : NULL, sW2 ? sW2 + ow : NULL);
                opt_update(hi + k, gh, lr, optimizer, sH1 ? sH1 + oh : NULL, sH2 ? sH2 + oh : NULL);
            }
        }
        if (loss_out) loss_out[epoch] = accum;
    }
    free(sW1); free(sH1); free(sW2); free(sH2);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * WMF ALS half sweep   (cymf/wmf.pyx:136-174; cymf/linalg.pyx:144-163 -> LAPACK dgesv)
 *   X[r,:] = solve(YtY + wd*I + (weight-1) * sum_{c in row r} y_c y_c^T ,  weight * sum y_c)
 *   empty row -> zeros (wmf.pyx:154-156).  dgesv = LU with partial (row) pivoting; A is symmetric so
 *   the reference's row-major fill of a column-major LAPACK matrix is immaterial.
 * ---------------------------------------------------------------------------------------------- */
static int lu_solve(double *A, double *b, int32_t K) {
    /* right-looking LU, partial pivoting, unit lower; then forward/back substitution */
    for (int32_t c = 0; c < K; ++c) {
        int32_t p = c;
        double best = fabs(A[(size_t)c * K + c]);
        for (int32_t r = c + 1; r < K; ++r) {
            double v = fabs(A[(size_t)r * K + c]);
            if (v > best) { best = v; p = r; }
        }
        if (best == 0.0) return c + 1;
        if (p != c) {
            for (int32_t q = 0; q < K; ++q) {
                double t = A[(size_t)c * K + q]; A[(size_t)c * K + q] = A[(size_t)p * K + q]; A[(size_t)p * K + q] = t;
            }
            double t = b[c]; b[c] = b[p]; b[p] = t;
        }
        double inv = 1.0 / A[(size_t)c * K + c];
        for (int32_t r = c + 1; r < K; ++r) {
            double f = A[(size_t)r * K + c] * inv;
            A[(size_t)r * K + c] = f;
            if (f != 0.0)
                for (int32_t q = c + 1; q < K; ++q) A[(size_t)r * K + q] -= f * A[(size_t)c * K + q];
        }
    }
    for (int32_t r = 1; r < K; ++r) {
        double acc = b[r];
        for (int32_t q = 0; q < r; ++q) acc -= A[(size_t)r * K + q] * b[q];
        b[r] = acc;
    }
    for (int32_t r = K - 1; r >= 0; --r) {
        double acc = b[r];
        for (int32_t q = r + 1; q < K; ++q) acc -= A[(size_t)r * K + q] * b[q];
        b[r] = acc / A[(size_t)r * K + r];
    }
    return 0;
}

int oracle_als_half(const int64_t *indptr, const int32_t *indices, double *X, const double *Y,
                    int64_t rows, int64_t n, int32_t K, double wd, double weight) {
    size_t KK = (size_t)K * K;
    double *G = (double *)calloc(KK, sizeof(double));
    double *A = (double *)malloc(KK * sizeof(double));
    double *b = (double *)malloc((size_t)K * sizeof(double));
    if (!G || !A || !b) return -1;
    for (int64_t r = 0; r < n; ++r) {                                       /* wmf.pyx:142 */
        const double *y = Y + (size_t)r * K;
        for (int32_t a = 0; a < K; ++a)
            for (int32_t c = 0; c < K; ++c) G[(size_t)a * K + c] += y[a] * y[c];
    }
    for (int32_t a = 0; a < K; ++a) G[(size_t)a * K + a] += wd;            /* wmf.pyx:143 */
    for (int64_t r = 0; r < rows; ++r) {
        double *x = X + (size_t)r * K;
        if (indptr[r] == indptr[r + 1]) { memset(x, 0, sizeof(double) * (size_t)K); continue; }
        memcpy(A, G, KK * sizeof(double));
        memset(b, 0, sizeof(double) * (size_t)K);
        for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p) {               /* wmf.pyx:161-166 */
            const double *y = Y + (size_t)indices[p] * K;
            for (int32_t a = 0; a < K; ++a) {
                b[a] += y[a] * weight;
                for (int32_t c = 0; c < K; ++c) A[(size_t)a * K + c] += y[a] * y[c] * (weight - 1.0);
            }
        }
        lu_solve(A, b, K);                                                  /* wmf.pyx:168 (info dropped) */
        memcpy(x, b, sizeof(double) * (size_t)K);
    }
    free(G); free(A); free(b);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * GloVe   (cymf/glove.pyx:117-156, cymf/model.pyx:34-35,166-204, cymf/optimizer.pyx:91-123)
 *   AdaGrad accumulators all start at ONE; both biases are updated INSIDE the k loop, i.e. K times
 *   per sample (model.pyx:195-204).  loss_out (num_epochs, may be NULL) = sum_l loss[l] / N.
 * ---------------------------------------------------------------------------------------------- */
int oracle_glove_fit(const int32_t *central, const int32_t *context, const double *counts, int64_t N,
                     double *W, double *bw, double *H, double *bh, int64_t Vw, int64_t Vh, int32_t K,
                     int32_t num_epochs, double lr, double x_max, double alpha, double *loss_out) {
    size_t nW = (size_t)Vw * K, nH = (size_t)Vh * K;
    double *aW = (double *)malloc(nW * sizeof(double)), *aH = (double *)malloc(nH * sizeof(double));
    double *abw = (double *)malloc((size_t)Vw * sizeof(double)), *abh = (double *)malloc((size_t)Vh * sizeof(double));
    if (!aW || !aH || !abw || !abh) return -1;
    for (size_t t = 0; t < nW; ++t) aW[t] = 1.0;
    for (size_t t = 0; t < nH; ++t) aH[t] = 1.0;
    for (int64_t t = 0; t < Vw; ++t) abw[t] = 1.0;
    for (int64_t t = 0; t < Vh; ++t) abh[t] = 1.0;
    for (int32_t it = 0; it < num_epochs; ++it) {
        double accum = 0.0;
        for (int64_t l = 0; l < N; ++l) {
            int32_t c = central[l], x = context[l];
            double *wc = W + (size_t)c * K, *hx = H + (size_t)x * K;
            double d = 0.0;
            for (int32_t k = 0; k < K; ++k) d += wc[k] * hx[k];              /* model.pyx:174-175 */
            d += bw[c] + bh[x];
            d -= log(counts[l]);
            double raw = d;
            d *= fmin(pow(counts[l] / x_max, alpha), 1.0);                   /* model.pyx:34-35,179 */
            accum += 0.5 * d * raw;
            for (int32_t k = 0; k < K; ++k) {                                /* model.pyx:195-204 */
                double gw = d * hx[k], gh = d * wc[k];
                double *pw = aW + (size_t)c * K + k, *ph = aH + (size_t)x * K + k;
                *pw += gw * gw; wc[k] -= lr * gw / sqrt(*pw);
                *ph += gh * gh; hx[k] -= lr * gh / sqrt(*ph);
                abw[c] += d * d; bw[c] -= lr * d / sqrt(abw[c]);
                abh[x] += d * d; bh[x] -= lr * d / sqrt(abh[x]);
            }
        }
        if (loss_out) loss_out[it] = accum / (double)N;
    }
    free(aW); free(aH); free(abw); free(abh);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Co-occurrence counting of read_text   (cymf/glove.pyx:218-221)
 *   for every kept token j of a line, for k = max(0, j - window) .. j-1 (same line):
 *       M[x_j + x_k * V] += 1.0 / (j - k)          (std::unordered_map<long, double>)
 *   tokens: kept-word ids, lines concatenated; line_ptr[n_lines + 1]: offsets of the lines.
 *   Cells are returned sorted by (row = x_j, col = x_k); every cell is summed in corpus order, as the
 *   map does.  Returns the number of cells, or -1 (memory) / -2 (capacity).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t key; double val; } cooc_cell;
static int cmp_cooc(const void *a, const void *b) {
    int64_t x = ((const cooc_cell *)a)->key, y = ((const cooc_cell *)b)->key;
    return x < y ? -1 : (x > y ? 1 : 0);
}
int64_t oracle_cooc_count(const int32_t *tokens, const int64_t *line_ptr, int64_t n_lines, int32_t V, int32_t window,
                          int32_t *rows, int32_t *cols, double *vals, int64_t capacity) {
    int64_t cap = 1024, used = 0;
    cooc_cell *tab = (cooc_cell *)malloc((size_t)cap * sizeof(cooc_cell));
    if (!tab) return -1;
    for (int64_t t = 0; t < cap; ++t) tab[t].key = -1;
    for (int64_t l = 0; l < n_lines; ++l) {
        const int32_t *x = tokens + line_ptr[l];
        int64_t n = line_ptr[l + 1] - line_ptr[l];
        for (int64_t j = 0; j < n; ++j)
            for (int64_t k = (j - window > 0 ? j - window : 0); k < j; ++k) {
                if (2 * (used + 1) > cap) {                                   /* grow and rehash */
                    int64_t ncap = cap * 2;
                    cooc_cell *nt = (cooc_cell *)malloc((size_t)ncap * sizeof(cooc_cell));
                    if (!nt) { free(tab); return -1; }
                    for (int64_t t = 0; t < ncap; ++t) nt[t].key = -1;
                    for (int64_t t = 0; t < cap; ++t)
                        if (tab[t].key >= 0) {
                            uint64_t h = ((uint64_t)tab[t].key * 0x9E3779B97F4A7C15ull) & (uint64_t)(ncap - 1);
                            while (nt[h].key >= 0) h = (h + 1) & (uint64_t)(ncap - 1);
                            nt[h] = tab[t];
                        }
                    free(tab); tab = nt; cap = ncap;
                }
                /* key as the reference forms it, but row-major so that sorting by key is (row, col) order */
                int64_t key = (int64_t)x[j] * V + x[k];
                uint64_t h = ((uint64_t)key * 0x9E3779B97F4A7C15ull) & (uint64_t)(cap - 1);
                while (tab[h].key >= 0 && tab[h].key != key) h = (h + 1) & (uint64_t)(cap - 1);
                if (tab[h].key < 0) { tab[h].key = key; tab[h].val = 0.0; ++used; }
                tab[h].val += 1.0 / (double)(j - k);                          /* glove.pyx:221 */
            }
    }
    int64_t m = 0;
    for (int64_t t = 0; t < cap; ++t) if (tab[t].key >= 0) tab[m++] = tab[t];
    qsort(tab, (size_t)m, sizeof(cooc_cell), cmp_cooc);
    if (m > capacity) { free(tab); return -2; }
    for (int64_t t = 0; t < m; ++t) {
        rows[t] = (int32_t)(tab[t].key / V);
        cols[t] = (int32_t)(tab[t].key % V);
        vals[t] = tab[t].val;
    }
    free(tab);
    return m;
}

/* ------------------------------------------------------------------------------------------------
 * Evaluator   (cymf/evaluator.pyx:57-139, cymf/metrics.pyx:24-43,71-85,109-125), unbiased=False.
 *   Candidates of user u = its test positives (CSR order) followed by num_negatives items drawn
 *   with replacement from one sequential generator, rejecting test+train positives.
 *   Ranking = descending score; equal scores are ordered by DESCENDING candidate position (the
 *   reverse of a stable ascending argsort -- NumPy's own unstable sort leaves tie order undefined).
 * ---------------------------------------------------------------------------------------------- */
int64_t oracle_eval_candidates(int32_t U, int32_t I, const int32_t *test_indptr, const int32_t *test_indices,
                               const int32_t *all_indptr, const int32_t *all_indices,
                               int32_t num_negatives, uint32_t seed,
                               int64_t *cand_ptr /* U+1 */, int32_t *cand_items /* nnz_test + U*num_neg */) {
    oracle_rng gen;
    oracle_rng_seed(&gen, seed);                                             /* evaluator.pyx:82 */
    int64_t w = 0;
    cand_ptr[0] = 0;
    for (int32_t u = 0; u < U; ++u) {
        if (test_indptr[u] != test_indptr[u + 1]) {                         /* evaluator.pyx:92-93 */
            for (int32_t p = test_indptr[u]; p < test_indptr[u + 1]; ++p) cand_items[w++] = test_indices[p];
            for (int32_t t = 0; t < num_negatives; ++t) {                   /* evaluator.pyx:106-111 */
                int32_t item = (int32_t)oracle_rng_below(&gen, (uint32_t)I);
                while (row_contains(all_indices, all_indptr[u], all_indptr[u + 1], item))
                    item = (int32_t)oracle_rng_below(&gen, (uint32_t)I);
                cand_items[w++] = item;
            }
        }
        cand_ptr[u + 1] = w;
    }
    return w;
}

double oracle_dcg_at_k(const int32_t *y, int64_t n, int32_t k) {            /* metrics.pyx:24-43 */
    double score = (double)y[0], counter = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        if (t >= 1 && t < k) score += (double)y[t] / log2((double)t + 1.0);
        counter += (double)y[t];
    }
    return counter == 0.0 ? 0.0 : score / counter;
}
double oracle_recall_at_k(const int32_t *y, int64_t n, int32_t k) {         /* metrics.pyx:71-85 */
    double score = 0.0, counter = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        if (t < k) score += (double)y[t];
        counter += (double)y[t];
    }
    return counter == 0.0 ? 0.0 : score / counter;
}
double oracle_ap_at_k(const int32_t *y, int64_t n, int32_t k) {             /* metrics.pyx:109-125 */
    double score = 0.0, counter = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        counter += (double)y[t];
        if (t < k && y[t] == 1) score += counter / ((double)t + 1.0);
    }
    return counter == 0.0 ? 0.0 : score / counter;
}

typedef struct { double score; int64_t pos; } scored;
static int cmp_scored(const void *a, const void *b) {
    const scored *p = (const scored *)a, *q = (const scored *)b;
    if (p->score > q->score) return -1;
    if (p->score < q->score) return 1;
    return (p->pos > q->pos) ? -1 : (p->pos < q->pos);
}

/* out_metrics: [nk][3] = DCG, Recall, MAP, each the mean over ALL U users (evaluator.pyx:135-137).
 * order_out (same length as cand_items, may be NULL): per user, candidate positions by rank. */
int oracle_eval_rank(const double *W, const double *H, int32_t U, int32_t K,
                     const int32_t *test_indptr, const int64_t *cand_ptr, const int32_t *cand_items,
                     const int32_t *ks, int32_t nk, double *out_metrics, int32_t *order_out) {
    int64_t maxc = 0;
    for (int32_t u = 0; u < U; ++u) if (cand_ptr[u + 1] - cand_ptr[u] > maxc) maxc = cand_ptr[u + 1] - cand_ptr[u];
    scored *buf = (scored *)malloc((size_t)(maxc ? maxc : 1) * sizeof(scored));
    int32_t *y = (int32_t *)malloc((size_t)(maxc ? maxc : 1) * sizeof(int32_t));
    if (!buf || !y) return -1;
    for (int32_t t = 0; t < nk * 3; ++t) out_metrics[t] = 0.0;
    for (int32_t u = 0; u < U; ++u) {
        int64_t c0 = cand_ptr[u], n = cand_ptr[u + 1] - c0;
        if (n == 0) continue;
        int64_t npos = test_indptr[u + 1] - test_indptr[u];
        const double *wu = W + (size_t)u * K;
        for (int64_t t = 0; t < n; ++t) {                                    /* evaluator.pyx:113 */
            const double *h = H + (size_t)cand_items[c0 + t] * K;
            double acc = 0.0;
            for (int32_t k = 0; k < K; ++k) acc += h[k] * wu[k];
            buf[t].score = acc; buf[t].pos = t;
        }
        qsort(buf, (size_t)n, sizeof(scored), cmp_scored);
        for (int64_t t = 0; t < n; ++t) {
            y[t] = buf[t].pos < npos ? 1 : 0;                                /* evaluator.pyx:114 */
            if (order_out) order_out[c0 + t] = (int32_t)buf[t].pos;
        }
        for (int32_t q = 0; q < nk; ++q) {
            out_metrics[q * 3 + 0] += oracle_dcg_at_k(y, n, ks[q]);
            out_metrics[q * 3 + 1] += oracle_recall_at_k(y, n, ks[q]);
            out_metrics[q * 3 + 2] += oracle_ap_at_k(y, n, ks[q]);
        }
    }
    for (int32_t t = 0; t < nk * 3; ++t) out_metrics[t] /= (double)U;
    free(buf); free(y);
    return 0;
}
