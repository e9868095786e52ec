/*
 * The reference's entry point, unchanged:  void tt_irt1(lapackint d, lapackint *n, double *xs,
 * lapackint *ttrank, double *ttcore, lapackint M, double *q, double *z, double *lPz)
 * (python/tt_irt_py/tt_irt1_int32.c:34, matlab/utils/tt_irt1_int64.c:34).  Plain C host code: it
 * widens the integer arguments and hands the call to the CUDA engine through the extended C-ABI in
 * include/tt_irt1.h.  Built twice: -DTTIRT_INT=int and -DTTIRT_INT="long long".
 * No CPU fallback: on failure the outputs are NaN-filled and one line goes to stderr.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tt_irt1.h"
#include "../../include/tt_irt_sqr.h"

static void nan_fill(double *p, int64_t count) {
  int64_t i;
  if (!p) return;
  for (i = 0; i < count; i++) p[i] = NAN;
}

/*
 * How many devices a call uses.  TTIRT_DEVICES=<count> or "all" is taken as given; unset or "auto": the samples are
 * independent (reference tt_irt1_int32.c:88-181), so a batch large enough to keep a full pipeline busy on each device
 * (2^22 seed points: four chunks of 2^20) is sharded over all visible GPUs from `first` on, a smaller one over fewer,
 * down to one.  Never more devices than `min_rows`-row blocks.
 */
static int device_policy(const char *e, int first, int64_t M, int64_t min_rows) {
  int ndev = 1, visible;
  if (e != NULL && strcmp(e, "all") != 0 && strcmp(e, "auto") != 0) {
    ndev = atoi(e);
  } else {
    visible = ttirt_device_count() - first;
    if (e != NULL && strcmp(e, "all") == 0) {
      ndev = visible;
    } else {
      ndev = ttirt_auto_devices(M, visible);
    }
  }
  if (ndev < 1) ndev = 1;
  while (ndev > 1 && M / ndev < min_rows) ndev--;
  return ndev;
}

void tt_irt1(TTIRT_INT d, TTIRT_INT *n, double *xs, TTIRT_INT *ttrank, double *ttcore, TTIRT_INT M,
             double *q, double *z, double *lPz) {
  int64_t *n64, *r64, k;
  int mode = TTIRT_MODE_FAST, first = 0, ndev = 1, rc;
  const char *e;

  if (d < 1 || M < 0 || !n || !xs || !ttrank || !ttcore || (M > 0 && (!q || !z || !lPz))) {
    fprintf(stderr, "tt_irt1[b200]: invalid arguments\n");
    if (d >= 1 && M > 0) { nan_fill(z, (int64_t)M * d); nan_fill(lPz, M); }
    return;
  }
  if (M == 0) return;
  n64 = (int64_t *)malloc(sizeof(int64_t) * (size_t)d);
  r64 = (int64_t *)malloc(sizeof(int64_t) * ((size_t)d + 1));
  if (!n64 || !r64) {
    fprintf(stderr, "tt_irt1[b200]: out of host memory\n");
    free(n64); free(r64);
    nan_fill(z, (int64_t)M * d); nan_fill(lPz, M);
    return;
  }
  for (k = 0; k < d; k++) n64[k] = (int64_t)n[k];
  for (k = 0; k <= d; k++) r64[k] = (int64_t)ttrank[k];

  if ((e = getenv("TTIRT_MODE")) != NULL && strcmp(e, "strict") == 0) mode = TTIRT_MODE_STRICT;
  if ((e = getenv("TTIRT_DEVICE")) != NULL) first = atoi(e);
  ndev = device_policy(getenv("TTIRT_DEVICES"), first, (int64_t)M, 64);

  rc = ttirt_run_host(d, n64, xs, r64, ttcore, M, q, z, lPz, NULL, mode, first, ndev);
  if (rc != 0) {
    nan_fill(z, (int64_t)M * d);
    nan_fill(lPz, M);
  } else if (getenv("TTIRT_VERBOSE") != NULL) {
    fprintf(stderr, "tt_irt1[b200]: M=%lld d=%lld mode=%s devices=%d launches=%lld\n", (long long)M, (long long)d,
            mode == TTIRT_MODE_STRICT ? "strict" : "fast", ndev, (long long)ttirt_kernel_launches());
  }
  free(n64);
  free(r64);
}

/*
 * The squared-density transforms, reference matlab/samplers/tt_irt_sqr.m:1 ([xq, lFapp] = tt_irt_sqr(xsf, f, q)) and
 * matlab/samplers/tt_rt_sqr.m:1 ([q, lFapp] = tt_rt_sqr(xsf, f, x)).  Matlab-only in the reference; these are the C entry
 * points a MEX gateway binds.
 */
static void sqr_entry(const char *name, int forward, TTIRT_INT d, TTIRT_INT *n, TTIRT_INT nxs, double *xs, TTIRT_INT *ttrank,
                      double *ttcore, TTIRT_INT M, TTIRT_INT D, double *in, double *out, double *lFapp) {
  int64_t *n64, *r64, k;
  int first = 0, ndev = 1, rc;
  const char *e;

  if (d < 1 || M < 0 || D < 1 || D > d || !n || !xs || !ttrank || !ttcore || (M > 0 && (!in || !out || !lFapp))) {
    fprintf(stderr, "%s[b200]: invalid arguments\n", name);
    if (D >= 1 && M > 0) { nan_fill(out, (int64_t)M * D); nan_fill(lFapp, M); }
    return;
  }
  if (M == 0) return;
  n64 = (int64_t *)malloc(sizeof(int64_t) * (size_t)d);
  r64 = (int64_t *)malloc(sizeof(int64_t) * ((size_t)d + 1));
  if (!n64 || !r64) {
    fprintf(stderr, "%s[b200]: out of host memory\n", name);
    free(n64); free(r64);
    nan_fill(out, (int64_t)M * D); nan_fill(lFapp, M);
    return;
  }
  for (k = 0; k < d; k++) n64[k] = (int64_t)n[k];
  for (k = 0; k <= d; k++) r64[k] = (int64_t)ttrank[k];
  if ((e = getenv("TTIRT_DEVICE")) != NULL) first = atoi(e);
  ndev = device_policy(getenv("TTIRT_DEVICES"), first, (int64_t)M, 128);   /* never more devices than 128-sample tiles */
  rc = forward ? ttirt_sqr_run_forward_host(d, n64, (int64_t)nxs, xs, r64, ttcore, M, D, in, out, lFapp, first, ndev)
               : ttirt_sqr_run_host(d, n64, (int64_t)nxs, xs, r64, ttcore, M, D, in, out, lFapp, first, ndev);
  if (rc != 0) {
    nan_fill(out, (int64_t)M * D);
    nan_fill(lFapp, M);
  }
  free(n64);
  free(r64);
}

void tt_irt_sqr(TTIRT_INT d, TTIRT_INT *n, TTIRT_INT nxs, double *xs, TTIRT_INT *ttrank, double *ttcore, TTIRT_INT M,
                TTIRT_INT D, double *q, double *z, double *lFapp) {
  sqr_entry("tt_irt_sqr", 0, d, n, nxs, xs, ttrank, ttcore, M, D, q, z, lFapp);
}

void tt_rt_sqr(TTIRT_INT d, TTIRT_INT *n, TTIRT_INT nxs, double *xs, TTIRT_INT *ttrank, double *ttcore, TTIRT_INT M,
               TTIRT_INT D, double *x, double *q, double *lFapp) {
  sqr_entry("tt_rt_sqr", 1, d, n, nxs, xs, ttrank, ttcore, M, D, x, q, lFapp);
}
