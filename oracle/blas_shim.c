/*
 * TEST INFRASTRUCTURE ONLY.  Four-routine Fortran-BLAS stand-in (dgemm_ N/N, dscal_,
 * daxpy_, dcopy_) so that the UNMODIFIED reference tt_irt1 sources
 * (/root/reference/python/tt_irt_py/tt_irt1_int32.c:28-31 declare exactly these)
 * link into a self-contained, deterministic library under oracle/_ref/.
 * Accumulation order is the netlib reference order: every output element of
 * dgemm is one sequential chain over the contraction index.  Build with
 * -ffp-contract=off.  SHIM_INT is `int` for the int32 ABI, `long long` for int64.
 */
#ifndef SHIM_INT
#define SHIM_INT int
#endif
typedef SHIM_INT bint;

void dgemm_(char *ta, char *tb, bint *m, bint *n, bint *k, double *alpha, double *A, bint *lda,
            double *B, bint *ldb, double *beta, double *C, bint *ldc) {
  (void)ta; (void)tb; /* the reference only ever passes 'N','N' */
  const bint M = *m, N = *n, K = *k;
  for (bint j = 0; j < N; j++) {
    double *c = C + (long long)j * *ldc;
    if (*beta == 0.0) for (bint i = 0; i < M; i++) c[i] = 0.0;
    else if (*beta != 1.0) for (bint i = 0; i < M; i++) c[i] *= *beta;
    for (bint l = 0; l < K; l++) {
      const double t = *alpha * B[l + (long long)j * *ldb];
      const double *a = A + (long long)l * *lda;
      for (bint i = 0; i < M; i++) c[i] += t * a[i];
    }
  }
}

void dscal_(bint *n, double *a, double *x, bint *incx) {
  for (bint i = 0; i < *n; i++) x[(long long)i * *incx] = *a * x[(long long)i * *incx];
}

void daxpy_(bint *n, double *a, double *x, bint *incx, double *y, bint *incy) {
  for (bint i = 0; i < *n; i++) y[(long long)i * *incy] += *a * x[(long long)i * *incx];
}

void dcopy_(bint *n, double *x, bint *incx, double *y, bint *incy) {
  for (bint i = 0; i < *n; i++) y[(long long)i * *incy] = x[(long long)i * *incx];
}
