/* Minimal <cblas.h> for the reference build against SciPy's bundled OpenBLAS (symbols are
 * renamed to scipy_* by -D flags in the Makefile).  TEST INFRASTRUCTURE. */
#ifndef PL_SHIM_CBLAS_H
#define PL_SHIM_CBLAS_H
#include <stddef.h>
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef size_t CBLAS_INDEX;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, double, const double*, int,
                 const double*, int, double, double*, int);
void cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, float, const float*, int,
                 const float*, int, float, float*, int);
void cblas_cgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void*, const void*, int,
                 const void*, int, const void*, void*, int);
void cblas_zgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, const void*, const void*, int,
                 const void*, int, const void*, void*, int);
void cblas_dscal(int, double, double*, int);
void cblas_sscal(int, float, float*, int);
void cblas_cscal(int, const void*, void*, int);
void cblas_zscal(int, const void*, void*, int);
#endif
