/* Minimal <lapacke.h> for the reference build (LP64 scipy-openblas).  TEST INFRASTRUCTURE.
 * Only the real double/float entry points get prototypes; the complex ones used by code that
 * is never called here are left to implicit declaration. */
#ifndef PL_SHIM_LAPACKE_H
#define PL_SHIM_LAPACKE_H
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_dgeqrf(int, lapack_int, lapack_int, double*, lapack_int, double*);
lapack_int LAPACKE_dorgqr(int, lapack_int, lapack_int, lapack_int, double*, lapack_int, const double*);
lapack_int LAPACKE_dgesvd(int, char, char, lapack_int, lapack_int, double*, lapack_int, double*, double*,
                          lapack_int, double*, lapack_int, double*);
lapack_int LAPACKE_dgesdd(int, char, lapack_int, lapack_int, double*, lapack_int, double*, double*, lapack_int,
                          double*, lapack_int);
lapack_int LAPACKE_sgeqrf(int, lapack_int, lapack_int, float*, lapack_int, float*);
lapack_int LAPACKE_sorgqr(int, lapack_int, lapack_int, lapack_int, float*, lapack_int, const float*);
lapack_int LAPACKE_sgesvd(int, char, char, lapack_int, lapack_int, float*, lapack_int, float*, float*,
                          lapack_int, float*, lapack_int, float*);
lapack_int LAPACKE_sgesdd(int, char, lapack_int, lapack_int, float*, lapack_int, float*, float*, lapack_int,
                          float*, lapack_int);
#endif
