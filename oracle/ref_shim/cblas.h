/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product path.
 *
 * Minimal CBLAS/LAPACK declarations so that the UNMODIFIED reference sources under
 * /root/reference (which `#include "cblas.h"` when built with -DCALS_OPENBLAS=1,
 * reference include/cals_blas.h:21-23) can be compiled against the OpenBLAS that ships
 * inside the scipy wheel (libscipy_openblas, LP64, every symbol prefixed `scipy_`).
 * Only the routines the reference calls are declared; each public name is mapped to its
 * `scipy_`-prefixed symbol with a #define.
 */
#ifndef CALS_B200_ORACLE_REF_SHIM_CBLAS_H
#define CALS_B200_ORACLE_REF_SHIM_CBLAS_H

#include <stddef.h>

#define cblas_dasum scipy_cblas_dasum
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_idamax scipy_cblas_idamax
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_dcopy scipy_cblas_dcopy
#define cblas_dscal scipy_cblas_dscal
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dtrsm scipy_cblas_dtrsm
#define dpotrf_ scipy_dpotrf_
#define dposv_ scipy_dposv_
#define openblas_set_num_threads scipy_openblas_set_num_threads
#define openblas_get_num_threads scipy_openblas_get_num_threads

#ifdef __cplusplus
extern "C" {
#endif

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef enum CBLAS_ORDER CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE CBLAS_TRANSPOSE;
typedef enum CBLAS_UPLO CBLAS_UPLO;
typedef enum CBLAS_DIAG CBLAS_DIAG;
typedef enum CBLAS_SIDE CBLAS_SIDE;

void openblas_set_num_threads(int n);
int openblas_get_num_threads(void);

double cblas_dasum(int n, const double *x, int incx);
double cblas_dnrm2(int n, const double *x, int incx);
size_t cblas_idamax(int n, const double *x, int incx);
void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
void cblas_dcopy(int n, const double *x, int incx, double *y, int incy);
void cblas_dscal(int n, double alpha, double *x, int incx);
void cblas_dgemv(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, int m, int n, double alpha, const double *a, int lda,
                 const double *x, int incx, double beta, double *y, int incy);
void cblas_dgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int m, int n, int k, double alpha,
                 const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc);
void cblas_dtrsm(CBLAS_ORDER order, CBLAS_SIDE side, CBLAS_UPLO uplo, CBLAS_TRANSPOSE ta, CBLAS_DIAG diag, int m,
                 int n, double alpha, const double *a, int lda, double *b, int ldb);

#ifdef __cplusplus
}
#endif
#endif
