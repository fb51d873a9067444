/* Stand-in for the system CBLAS header that cymf/linalg.pxd:22-34 includes.
 * The two entry points declared there are never reached from the BPR / WMF /
 * GloVe / evaluator paths (only `solvep` -> LAPACK dgesv is), so this header
 * exists only to let linalg.pyx compile and import when building oracle/_ref.
 * Test infrastructure only. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

static inline double cblas_ddot(int n, double *x, int incx, double *y, int incy) {
    double acc = 0.0;
    for (int t = 0; t < n; ++t) acc += x[t * incx] * y[t * incy];
    return acc;
}

static inline double cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb,
                                 int m, int n, int k, double alpha, double *a, int lda,
                                 double *b, int ldb, double beta, double *c, int ldc) {
    (void)order; /* row-major only: all callers in linalg.pyx pass CblasRowMajor */
    for (int r = 0; r < m; ++r)
        for (int q = 0; q < n; ++q) {
            double acc = 0.0;
            for (int t = 0; t < k; ++t) {
                double av = (ta == CblasNoTrans) ? a[r * lda + t] : a[t * lda + r];
                double bv = (tb == CblasNoTrans) ? b[t * ldb + q] : b[q * ldb + t];
                acc += av * bv;
            }
            c[r * ldc + q] = alpha * acc + beta * c[r * ldc + q];
        }
    return 0.0;
}
#endif
