/* The MEX gateway INTEGRATION.md proposes for the squared-density transform (same pattern as the reference's
 * matlab/utils/tt_irt_mex.c:5-39).  In Matlab:  mex -largeArrayDims -O tt_irt_sqr_mex.c -L/path/to/tt-irt_b200/lib -ltt_irt1_int64
 * tests/test_mex_gateway.py compiles this very file against the stand-in mex.h of oracle/mexstub/ and, on a B200, checks it
 * against the outputs of the reference's tt_irt_sqr.m (tests/golden/matlab_sqr_*.npz). */
/* tt_irt_sqr_mex.c:  [xq, lFapp] = tt_irt_sqr_mex(n, xs, ttrank, ttcore, q)
 * build:  mex -largeArrayDims -O tt_irt_sqr_mex.c -L/path/to/tt-irt_b200/lib -ltt_irt1_int64 */
#include "mex.h"
extern void tt_irt_sqr(mwIndex d, mwIndex *n, mwIndex nxs, double *xs, mwIndex *ttrank, double *ttcore,
                       mwIndex M, mwIndex D, double *q, double *z, double *lFapp);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  mwIndex d, i, M, D, nxs, *n, *rk;
  if (nrhs < 5) { mexPrintf("Specify n, xs, ttrank, ttcore, q\n"); return; }
  d = mxGetNumberOfElements(prhs[0]);
  n = (mwIndex *)mxMalloc(sizeof(mwIndex) * d);
  rk = (mwIndex *)mxMalloc(sizeof(mwIndex) * (d + 1));
  for (i = 0; i < d; i++) n[i] = (mwIndex)mxGetPr(prhs[0])[i];       /* f.n  (doubles in Matlab) */
  for (i = 0; i <= d; i++) rk[i] = (mwIndex)mxGetPr(prhs[2])[i];     /* f.r */
  nxs = mxGetNumberOfElements(prhs[1]);                             /* cell2mat(xsf): sum(n) or sum(n + 2) points */
  M = mxGetM(prhs[4]); D = mxGetN(prhs[4]);                         /* q is M x D, D <= d */
  plhs[0] = mxCreateDoubleMatrix(M, D, mxREAL);
  plhs[1] = mxCreateDoubleMatrix(M, 1, mxREAL);
  tt_irt_sqr(d, n, nxs, mxGetPr(prhs[1]), rk, mxGetPr(prhs[3]), M, D, mxGetPr(prhs[4]), mxGetPr(plhs[0]), mxGetPr(plhs[1]));
  mxFree(n); mxFree(rk);
}
