/*
 * Plain-C caller of the squared-density transforms (include/tt_irt_sqr.h; reference matlab/samplers/tt_irt_sqr.m and
 * tt_rt_sqr.m), the calls a MEX gateway makes:
 *
 *   gcc -O2 -DTTIRT_INT=int        examples/call_tt_irt_sqr.c -Iinclude tt-irt_b200/tt_irt_py/tt_irt1_int32.so -lm -o sqr32
 *   gcc -O2 "-DTTIRT_INT=long long" examples/call_tt_irt_sqr.c -Iinclude -Ltt-irt_b200/lib -ltt_irt1_int64 -lm -o sqr64
 *
 * Builds a small random TT of the square root of a density with a fixed xorshift generator, samples M points with
 * tt_irt_sqr, maps them back with tt_rt_sqr and prints checksums plus the largest round-trip error |q' - q|.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>

#include "tt_irt_sqr.h"

static uint64_t s64 = 88172645463325252ULL;
static double next_u(void) {
  s64 ^= s64 << 13; s64 ^= s64 >> 7; s64 ^= s64 << 17;
  return (double)(s64 >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
  const TTIRT_INT d = 5, nn = 17, r = 8;
  const TTIRT_INT M = argc > 1 ? (TTIRT_INT)atoll(argv[1]) : 4096;
  TTIRT_INT n[5], rk[6];
  size_t ncore = 0;
  TTIRT_INT k, j;
  long long m, outside = 0;
  double *xs, *core, *q, *z, *lf, *qb, *lb, sz = 0.0, sl = 0.0, rt = 0.0, dl = 0.0;
  for (k = 0; k < d; k++) n[k] = nn;
  for (k = 0; k <= d; k++) rk[k] = (k == 0 || k == d) ? 1 : r;
  for (k = 0; k < d; k++) ncore += (size_t)rk[k] * n[k] * rk[k + 1];
  xs = malloc(sizeof(double) * d * nn);
  core = malloc(sizeof(double) * ncore);
  q = malloc(sizeof(double) * (size_t)M * d);
  z = calloc((size_t)M * d, sizeof(double));
  qb = calloc((size_t)M * d, sizeof(double));
  lf = calloc((size_t)M, sizeof(double));
  lb = calloc((size_t)M, sizeof(double));
  if (!xs || !core || !q || !z || !qb || !lf || !lb) return 2;
  for (k = 0; k < d; k++)
    for (j = 0; j < nn; j++) xs[k * nn + j] = -1.0 + 2.0 * (double)j / (double)(nn - 1);
  for (m = 0; m < (long long)ncore; m++) core[m] = next_u();
  for (m = 0; m < (long long)M * d; m++) q[m] = next_u();

  tt_irt_sqr(d, n, d * nn, xs, rk, core, M, d, q, z, lf);     /* [z, lf]  = tt_irt_sqr(xsf, f, q) */
  tt_rt_sqr(d, n, d * nn, xs, rk, core, M, d, z, qb, lb);     /* [qb, lb] = tt_rt_sqr(xsf, f, z)  */

  for (m = 0; m < (long long)M; m++) {
    sl += lf[m];
    if (fabs(lb[m] - lf[m]) > dl || !(lb[m] == lb[m])) dl = fabs(lb[m] - lf[m]);
    for (k = 0; k < d; k++) {
      const double v = z[m + (long long)M * k];
      const double e = fabs(qb[m + (long long)M * k] - q[m + (long long)M * k]);
      sz += v;
      if (e > rt || !(e == e)) rt = e;
      if (!(v >= -1.0 && v <= 1.0)) outside++;
    }
  }
  printf("M=%lld sumZ=%.15e sumlF=%.15e outside=%lld roundtrip=%.3e dlF=%.3e launches=%lld\n", (long long)M, sz, sl, outside, rt, dl,
         (long long)ttirt_kernel_launches());
  free(xs); free(core); free(q); free(z); free(qb); free(lf); free(lb);
  return (outside == 0 && isfinite(sl) && rt < 1e-9) ? 0 : 1;
}
