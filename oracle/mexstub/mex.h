/*
 * TEST INFRASTRUCTURE ONLY.  Stand-in for Matlab's mex.h: just enough of the mxArray API for the UNMODIFIED reference MEX source
 * /root/reference/matlab/utils/tracemult.c to compile and run outside Matlab (oracle/Makefile builds it into
 * oracle/_ref/libref_tracemult.so; oracle/mex_host.py calls its mexFunction through ctypes).  Real arrays only.
 */
#ifndef TTIRT_MEXSTUB_MEX_H
#define TTIRT_MEXSTUB_MEX_H
#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>

typedef size_t mwIndex;
typedef size_t mwSize;

typedef struct mxArray_tag {
  double *pr, *pi;
  mwSize ndim;
  mwSize dims[8];
  int is_complex;
} mxArray;

typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

double *mxGetPr(const mxArray *a);
double *mxGetPi(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
bool mxIsComplex(const mxArray *a);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
size_t mxGetNumberOfElements(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);   /* product of the dimensions from the second on, as in Matlab */
void *mxMalloc(size_t bytes);
void mxFree(void *p);
int mexPrintf(const char *fmt, ...);

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#endif
