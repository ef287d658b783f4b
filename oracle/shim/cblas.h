/* oracle/shim/cblas.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Header-only, plain-loop implementation of the five CBLAS entry points the
 * reference's math wrappers call (ffm.cpp:21-60: daxpy, dscal, dgemm, dgemv,
 * ddot).  The reference expects an OpenBLAS/MKL <cblas.h> that is not in this
 * image; this shim lets the UNMODIFIED reference sources compile where they
 * lie under /root/reference (recipe: oracle/Makefile, outputs in oracle/_ref).
 * Only standard BLAS semantics are implemented; summation order is the naive
 * left-to-right one, which is deterministic and thread-count independent.
 *
 * The shim also counts calls, so the harness can read the number of CG
 * iterations the reference ran (cblas_dscal is called exactly once per CG
 * iteration, ffm.cpp:810, and nowhere else on the path).
 */
#ifndef OCFFM_ORACLE_SHIM_CBLAS_H
#define OCFFM_ORACLE_SHIM_CBLAS_H
#include <stddef.h>

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

#ifdef __cplusplus
extern "C" {
#endif
extern long ocffm_shim_dscal_calls;   /* defined in ref_harness.cpp / shim_counters.c */
#ifdef __cplusplus
}
#endif

static inline void cblas_daxpy(const long n, const double alpha, const double *x,
                               const long incx, double *y, const long incy) {
    for (long i = 0; i < n; i++) y[i * incy] += alpha * x[i * incx];
}

static inline void cblas_dscal(const long n, const double alpha, double *x, const long incx) {
    ocffm_shim_dscal_calls++;
    for (long i = 0; i < n; i++) x[i * incx] *= alpha;
}

static inline double cblas_ddot(const long n, const double *x, const long incx,
                                const double *y, const long incy) {
    double s = 0;
    for (long i = 0; i < n; i++) s += x[i * incx] * y[i * incy];
    return s;
}

/* y = alpha*op(A)*x + beta*y, A is M x N (row-major only; that is all the reference uses) */
static inline void cblas_dgemv(const CBLAS_ORDER order, const CBLAS_TRANSPOSE trans,
                               const long M, const long N, const double alpha,
                               const double *A, const long lda, const double *x, const long incx,
                               const double beta, double *y, const long incy) {
    (void)order;
    if (trans == CblasNoTrans) {
        for (long i = 0; i < M; i++) {
            double s = 0;
            for (long j = 0; j < N; j++) s += A[i * lda + j] * x[j * incx];
            y[i * incy] = alpha * s + (beta == 0 ? 0 : beta * y[i * incy]);
        }
    } else {
        for (long j = 0; j < N; j++) y[j * incy] = (beta == 0 ? 0 : beta * y[j * incy]);
        for (long i = 0; i < M; i++) {
            const double xi = alpha * x[i * incx];
            for (long j = 0; j < N; j++) y[j * incy] += A[i * lda + j] * xi;
        }
    }
}

/* C = alpha*op(A)*op(B) + beta*C, row-major, C is M x N, inner dimension K */
static inline void cblas_dgemm(const CBLAS_ORDER order, const CBLAS_TRANSPOSE ta,
                               const CBLAS_TRANSPOSE tb, const long M, const long N, const long K,
                               const double alpha, const double *A, const long lda,
                               const double *B, const long ldb, const double beta,
                               double *C, const long ldc) {
    (void)order;
    for (long i = 0; i < M; i++)
        for (long j = 0; j < N; j++)
            C[i * ldc + j] = (beta == 0 ? 0 : beta * C[i * ldc + j]);
    if (ta == CblasNoTrans && tb == CblasNoTrans) {
        for (long i = 0; i < M; i++)
            for (long l = 0; l < K; l++) {
                const double a = alpha * A[i * lda + l];
                for (long j = 0; j < N; j++) C[i * ldc + j] += a * B[l * ldb + j];
            }
    } else if (ta == CblasTrans && tb == CblasNoTrans) {
        /* A stored K x M */
        for (long l = 0; l < K; l++)
            for (long i = 0; i < M; i++) {
                const double a = alpha * A[l * lda + i];
                for (long j = 0; j < N; j++) C[i * ldc + j] += a * B[l * ldb + j];
            }
    } else {
        for (long i = 0; i < M; i++)
            for (long j = 0; j < N; j++) {
                double s = 0;
                for (long l = 0; l < K; l++) {
                    const double a = (ta == CblasNoTrans) ? A[i * lda + l] : A[l * lda + i];
                    const double b = (tb == CblasNoTrans) ? B[l * ldb + j] : B[j * ldb + l];
                    s += a * b;
                }
                C[i * ldc + j] += alpha * s;
            }
    }
}
#endif
