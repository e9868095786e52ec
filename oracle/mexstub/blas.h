/* TEST INFRASTRUCTURE ONLY.  Stand-in for Matlab's blas.h (see mex.h here): the three routines tracemult.c calls, with Matlab's
 * 64-bit integer arguments, mapped onto the netlib-order shim oracle/blas_shim.c built with SHIM_INT = long long. */
#ifndef TTIRT_MEXSTUB_BLAS_H
#define TTIRT_MEXSTUB_BLAS_H
#include "mex.h"
#define dgemm dgemm_
#define dcopy dcopy_
#define zgemm zgemm_
void dgemm_(char *ta, char *tb, mwIndex *m, mwIndex *n, mwIndex *k, double *alpha, double *A, mwIndex *lda, double *B, mwIndex *ldb,
            double *beta, double *C, mwIndex *ldc);
void dcopy_(mwIndex *n, double *x, mwIndex *incx, double *y, mwIndex *incy);
void zgemm_(char *ta, char *tb, mwIndex *m, mwIndex *n, mwIndex *k, double *alpha, double *A, mwIndex *lda, double *B, mwIndex *ldb,
            double *beta, double *C, mwIndex *ldc);
#endif
