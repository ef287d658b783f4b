/* oracle/shim_openblas/cblas.h -- TEST INFRASTRUCTURE (CPU-baseline timing only).
 *
 * Routes the five CBLAS calls of the reference (ffm.cpp:21-60) to the optimised
 * OpenBLAS that ships inside scipy's wheel (symbols are prefixed scipy_, LP64),
 * because the image has no system OpenBLAS/MKL (the reference's Makefile:18-19
 * expects /opt/OpenBLAS).  Used for the TIMED reference build so the CPU baseline
 * is not handicapped by the plain-loop shim; goldens use oracle/shim instead.
 */
#ifndef OCFFM_ORACLE_SHIM_OPENBLAS_CBLAS_H
#define OCFFM_ORACLE_SHIM_OPENBLAS_CBLAS_H
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
#ifdef __cplusplus
extern "C" {
#endif
extern long ocffm_shim_dscal_calls;
void scipy_cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
void scipy_cblas_dscal(int n, double alpha, double *x, int incx);
double scipy_cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
void scipy_cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE trans, int m, int n, double alpha,
                       const double *a, int lda, const double *x, int incx, double beta,
                       double *y, int incy);
void scipy_cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n,
                       int k, double alpha, const double *a, int lda, const double *b, int ldb,
                       double beta, double *c, int ldc);
#ifdef __cplusplus
}
#endif
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_ddot scipy_cblas_ddot
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_dgemm scipy_cblas_dgemm
static inline void cblas_dscal(int n, double alpha, double *x, int incx) {
    ocffm_shim_dscal_calls++;
    scipy_cblas_dscal(n, alpha, x, incx);
}
#endif
