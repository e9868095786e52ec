/*
 * Plain-C caller of the drop-in symbol, exactly as a C program would call the reference routine
 * (python/tt_irt_py/tt_irt1_int32.c:34 / matlab/utils/tt_irt1_int64.c:34):
 *
 *   gcc -O2 -DTTIRT_INT=int        examples/call_tt_irt1.c -Iinclude tt-irt_b200/tt_irt_py/tt_irt1_int32.so -lm -o call32
 *   gcc -O2 "-DTTIRT_INT=long long" examples/call_tt_irt1.c -Iinclude -Ltt-irt_b200/lib -ltt_irt1_int64 -lm -o call64
 *
 * Builds a small random TT density with a fixed xorshift generator, samples M points and prints checksums
 * (sum of Z, sum of lPz, and the number of samples outside their grid) that the test compares with the Python path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>

#include "tt_irt1.h"

static uint64_t s64 = 88172645463325252ULL;
static double next_u(void) {
  s64 ^= s64 << 13; s64 ^= s64 >> 7; s64 ^= s64 << 17;
  return (double)(s64 >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
  const TTIRT_INT d = 6, nn = 17, r = 8;
  const TTIRT_INT M = argc > 1 ? (TTIRT_INT)atoll(argv[1]) : 4096;
  TTIRT_INT n[6], rk[7];
  size_t ncore = 0;
  TTIRT_INT k, j;
  long long m, outside = 0;
  double *xs, *core, *q, *z, *lpz, sz = 0.0, sl = 0.0;
  for (k = 0; k < d; k++) n[k] = nn;
  for (k = 0; k <= d; k++) rk[k] = (k == 0 || k == d) ? 1 : r;
  for (k = 0; k < d; k++) ncore += (size_t)rk[k] * n[k] * rk[k + 1];
  xs = malloc(sizeof(double) * d * nn);
  core = malloc(sizeof(double) * ncore);
  q = malloc(sizeof(double) * (size_t)M * d);
  z = calloc((size_t)M * d, sizeof(double));
  lpz = calloc((size_t)M, sizeof(double));
  if (!xs || !core || !q || !z || !lpz) return 2;
  for (k = 0; k < d; k++)
    for (j = 0; j < nn; j++) xs[k * nn + j] = -1.0 + 2.0 * (double)j / (double)(nn - 1);
  for (m = 0; m < (long long)ncore; m++) core[m] = next_u();
  for (m = 0; m < (long long)M * d; m++) q[m] = next_u();

  tt_irt1(d, n, xs, rk, core, M, q, z, lpz);

  for (m = 0; m < (long long)M; m++) {
    sl += lpz[m];
    for (k = 0; k < d; k++) {
      const double v = z[m + (long long)M * k];
      sz += v;
      if (!(v >= -1.0 - 1e-9 && v <= 1.0 + 1e-9)) outside++;
    }
  }
  printf("M=%lld sumZ=%.15e sumlPz=%.15e outside=%lld launches=%lld\n", (long long)M, sz, sl, outside,
         (long long)ttirt_kernel_launches());
  free(xs); free(core); free(q); free(z); free(lpz);
  return (outside == 0 && isfinite(sl)) ? 0 : 1;
}
