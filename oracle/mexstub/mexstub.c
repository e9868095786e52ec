/* TEST INFRASTRUCTURE ONLY.  The mxArray functions declared in mex.h here, plus the host-side helpers oracle/mex_host.py uses
 * to build arguments and read results.  Real double arrays only; a complex request aborts (the paths this repo pins are real). */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "mex.h"

double *mxGetPr(const mxArray *a) { return a->pr; }
double *mxGetPi(const mxArray *a) { return a->pi; }
mwSize mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
const mwSize *mxGetDimensions(const mxArray *a) { return a->dims; }
bool mxIsComplex(const mxArray *a) { return a->is_complex != 0; }

mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c) {
  (void)cls;
  if (c != mxREAL || ndim > 8) { fprintf(stderr, "mexstub: complex / >8-d arrays are not supported\n"); abort(); }
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  size_t total = 1;
  a->ndim = ndim < 2 ? 2 : ndim;
  for (mwSize i = 0; i < 8; i++) a->dims[i] = 1;
  for (mwSize i = 0; i < ndim; i++) { a->dims[i] = dims[i]; total *= dims[i]; }
  a->pr = (double *)calloc(total ? total : 1, sizeof(double));
  return a;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
  const mwSize dims[2] = {m, n};
  return mxCreateNumericArray(2, dims, mxDOUBLE_CLASS, c);
}
size_t mxGetNumberOfElements(const mxArray *a) {
  size_t t = 1;
  for (mwSize i = 0; i < a->ndim; i++) t *= a->dims[i];
  return t;
}
size_t mxGetM(const mxArray *a) { return a->dims[0]; }
size_t mxGetN(const mxArray *a) {
  size_t t = 1;
  for (mwSize i = 1; i < a->ndim; i++) t *= a->dims[i];
  return t;
}
void *mxMalloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void mxFree(void *p) { free(p); }
int mexPrintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  const int r = vfprintf(stderr, fmt, ap);
  va_end(ap);
  return r;
}
void zgemm_(char *ta, char *tb, mwIndex *m, mwIndex *n, mwIndex *k, double *alpha, double *A, mwIndex *lda, double *B, mwIndex *ldb,
            double *beta, double *C, mwIndex *ldc) {
  (void)ta; (void)tb; (void)m; (void)n; (void)k; (void)alpha; (void)A; (void)lda; (void)B; (void)ldb; (void)beta; (void)C; (void)ldc;
  fprintf(stderr, "mexstub: zgemm (complex tracemult) is not supported\n");
  abort();
}

/* ---- host side (ctypes): wrap caller memory, run the MEX entry point, hand the result back ---- */
mxArray *mexstub_wrap(double *data, mwSize ndim, const mwSize *dims) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  a->ndim = ndim < 2 ? 2 : ndim;
  for (mwSize i = 0; i < 8; i++) a->dims[i] = 1;
  for (mwSize i = 0; i < ndim && i < 8; i++) a->dims[i] = dims[i];
  a->pr = data;
  return a;
}
void mexstub_free_wrapper(mxArray *a) { free(a); }
void mexstub_free_array(mxArray *a) { if (a) { free(a->pr); free(a); } }
mwSize mexstub_ndim(const mxArray *a) { return a->ndim; }
mwSize mexstub_dim(const mxArray *a, mwSize i) { return a->dims[i]; }
double *mexstub_data(const mxArray *a) { return a->pr; }
/* one output; returns NULL when the MEX function assigned none (it prints and returns on bad arguments) */
mxArray *mexstub_call1(int nrhs, const mxArray **prhs) {
  mxArray *plhs[1] = {NULL};
  mexFunction(1, plhs, nrhs, prhs);
  return plhs[0];
}
