/*
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement of the reference's linear-spline inverse Rosenblatt transform
 * `tt_irt1` (reference: python/tt_irt_py/tt_irt1_int32.c:34-193 and the
 * byte-identical matlab/utils/tt_irt1_int64.c:34-193).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's library; nothing under tt-irt_b200/ links or calls it.
 *
 * Parity pin: this restatement is checked BIT FOR BIT (Z and lPz) against the
 * UNMODIFIED reference sources compiled from /root/reference and linked to the
 * netlib-order BLAS shim in oracle/blas_shim.c (oracle/_ref/, built by
 * oracle/Makefile), and to the noise floor against the same sources linked to
 * OpenBLAS; see tests/test_oracle.py and tests/golden/.  The reference ships no
 * golden vectors of its own, so those builds are the pin.
 *
 * The reference blocks samples by 64 and calls BLAS; a sample's arithmetic does
 * not depend on its block mates, so this file is written sample-major instead:
 * one marginalisation sweep, then for every sample a walk over the dimensions.
 * Every floating-point operation is performed in the order the reference
 * performs it when its BLAS accumulates each output element as one sequential
 * chain over the contraction index (the netlib order, no FMA contraction):
 * compile with -O2 -ffp-contract=off.
 *
 * Beyond (z, lPz) it can export, for the parity protocol, the grid-interval
 * index chosen per (sample, dimension), the cancellation factor kappa of the
 * reference's quadratic formula (tt_irt1_int32.c:150-156), and the distance of
 * q to the nearest normalised-CDF node (how robust the index is), and a first-order
 * condition estimate of xk (the entry-wise Z tolerance of the parity tests).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long long oidx;

typedef struct {
  oidx d;
  const oidx *n, *r;
  const double *xs, *core;
  oidx *off_x;    /* start of grid k in xs            (ref :43-49) */
  oidx *off_c;    /* start of core k in ttcore        (ref :55-57) */
  double **marg;  /* marg[k]: r_{k+1} right marginal  (ref C[k], :59-82) */
  double **pk;    /* pk[k]: r_k x n_k, core_k contracted with marg[k] (ref fk, :72 and :98) */
  double **hh;    /* hh[k][j] = x_{j+1}-x_j           (ref :74, :101) */
  oidx rmax, nmax;
} model_t;

/* P = reshape(core_k, r_k n_k, r_{k+1}) * marg ; one sequential chain per output
   element over the right rank index (ref dgemm at :72 / :98). */
static void contract_right(const model_t *md, oidx k, const double *marg, double *out) {
  const oidx rows = md->r[k] * md->n[k], rr = md->r[k + 1];
  const double *ck = md->core + md->off_c[k];
  for (oidx i = 0; i < rows; i++) out[i] = 0.0;
  for (oidx l = 0; l < rr; l++) {
    const double t = marg[l];
    for (oidx i = 0; i < rows; i++) out[i] += t * ck[i + l * rows];
  }
}

static int model_build(model_t *md) {
  const oidx d = md->d;
  md->off_x = (oidx *)malloc(sizeof(oidx) * (d + 1));
  md->off_c = (oidx *)malloc(sizeof(oidx) * (d + 1));
  md->marg = (double **)calloc(d, sizeof(double *));
  md->pk = (double **)calloc(d, sizeof(double *));
  md->hh = (double **)calloc(d, sizeof(double *));
  if (!md->off_x || !md->off_c || !md->marg || !md->pk || !md->hh) return -1;
  md->off_x[0] = 0; md->off_c[0] = 0; md->rmax = md->r[0]; md->nmax = 0;
  for (oidx k = 0; k < d; k++) {
    md->off_x[k + 1] = md->off_x[k] + md->n[k];
    md->off_c[k + 1] = md->off_c[k] + md->r[k] * md->n[k] * md->r[k + 1];
    if (md->r[k + 1] > md->rmax) md->rmax = md->r[k + 1];
    if (md->n[k] > md->nmax) md->nmax = md->n[k];
  }
  for (oidx k = 0; k < d; k++) {
    md->hh[k] = (double *)malloc(sizeof(double) * md->n[k]);
    md->pk[k] = (double *)malloc(sizeof(double) * md->r[k] * md->n[k]);
    md->marg[k] = (double *)malloc(sizeof(double) * md->r[k + 1]);
    const double *x = md->xs + md->off_x[k];
    for (oidx j = 0; j + 1 < md->n[k]; j++) md->hh[k][j] = x[j + 1] - x[j];
  }
  /* right-to-left sweep: marg[d-1] = {1}; marg[k-1][a] = sum_j (P[a,j]+P[a,j+1])*h_j*0.5 (ref :59-82) */
  md->marg[d - 1][0] = 1.0;
  for (oidx k = d - 1; k >= 0; k--) {
    contract_right(md, k, md->marg[k], md->pk[k]);
    if (k == 0) break;
    const oidx rk = md->r[k];
    double *mo = md->marg[k - 1];
    for (oidx a = 0; a < rk; a++) mo[a] = 0.0;
    for (oidx j = 0; j + 1 < md->n[k]; j++)
      for (oidx a = 0; a < rk; a++)
        mo[a] += (md->pk[k][a + j * rk] + md->pk[k][a + (j + 1) * rk]) * md->hh[k][j] * 0.5;
  }
  return 0;
}

static void model_free(model_t *md) {
  for (oidx k = 0; k < md->d; k++) {
    if (md->hh) free(md->hh[k]);
    if (md->pk) free(md->pk[k]);
    if (md->marg) free(md->marg[k]);
  }
  free(md->hh); free(md->pk); free(md->marg); free(md->off_x); free(md->off_c);
}

/* One sample, all dimensions.  left: r_k left-interface vector (ref row of fkm1). */
static void walk_sample(const model_t *md, oidx M, oidx m, const double *q, double *z, double *lPz,
                        int *idx, double *kappa, double *gap, double *cond, double *lsens,
                        double *left, double *next, double *p, double *cdf, double *slab) {
  double lp = 0.0;
  left[0] = 1.0; /* ref :90, assumes r_0 = 1 */
  for (oidx k = 0; k < md->d; k++) {
    const oidx rk = md->r[k], nk = md->n[k], rn = md->r[k + 1];
    const double *x = md->xs + md->off_x[k], *h = md->hh[k], *P = md->pk[k];
    const double *ck = md->core + md->off_c[k];
    const double qk = q[m + M * k];
    /* conditional pdf on the grid: p_j = | sum_a left[a] P[a,j] |  (ref :103-105) */
    for (oidx j = 0; j < nk; j++) {
      double s = 0.0;
      for (oidx a = 0; a < rk; a++) s += P[a + j * rk] * left[a];
      p[j] = fabs(s);
    }
    /* trapezoid prefix: cdf_j = (cdf_{j-1} + h/2 p_{j-1}) + h/2 p_j  (ref :107-113) */
    cdf[0] = 0.0;
    for (oidx j = 1; j < nk; j++) {
      const double hq = h[j - 1] * 0.5;
      double c = cdf[j - 1];
      c += hq * p[j - 1];
      c += hq * p[j];
      cdf[j] = c;
    }
    /* zero-mass fallback in index space, then normalise by reciprocal (ref :116-130) */
    if (cdf[nk - 1] == 0.0) {
      const double u = 1.0 / (double)(nk - 1);
      for (oidx j = 0; j < nk; j++) { p[j] = 1.0 * u; cdf[j] = (double)j * u; }
    }
    {
      const double s = 1.0 / cdf[nk - 1];
      for (oidx j = 0; j < nk; j++) { cdf[j] *= s; p[j] *= s; }
    }
    /* bisection with strict '>' (ref :134-142) */
    oidx lo = 0, hi = nk - 1;
    while (hi - lo > 1) {
      const oidx mid = (oidx)(int)((double)(lo + hi) * 0.5);
      if (qk > cdf[mid]) lo = mid; else hi = mid;
    }
    /* closed-form root of the piecewise-quadratic CDF, reference formula verbatim in
       operation order (ref :146-159); deliberately NOT the stable variant */
    const double x1 = x[lo], x2 = x[lo + 1];
    const double c1 = p[lo], c2 = p[lo + 1];
    const double hq = x2 - x1;
    const double Aq = 0.5 * (c2 - c1) / hq;
    const double Bq = (c1 * x2 - c2 * x1) / hq;
    double Dq = 2.0 * Aq * x1 + Bq;
    Dq *= Dq;
    Dq += 4.0 * Aq * (qk - cdf[lo]);
    const double root = sqrt(fabs(Dq));
    double xk = 0.5 * (-Bq + root) / Aq;
    if (Aq == 0.0) xk = x1 + (qk - cdf[lo]) / Bq;
    z[m + M * k] = xk;
    if (idx) idx[m + M * k] = (int)lo;
    if (kappa) {
      const double den = fabs(-Bq + root);
      kappa[m + M * k] = (Aq == 0.0) ? 1.0 : (den > 0.0 ? fabs(Bq) / den : INFINITY);
    }
    if (gap) {
      const double g1 = qk - cdf[lo], g2 = cdf[lo + 1] - qk;
      gap[m + M * k] = g1 < g2 ? g1 : g2;
    }
    /* log-density of the interpolated conditional (ref :161-165) */
    const double w1 = (x2 - xk) / hq, w2 = (xk - x1) / hq;
    lp += log(fabs(p[lo] * w1 + p[lo + 1] * w2));
    if (cond) {
      /* first-order sensitivity of xk to O(eps) perturbations of (cdf, p): the true
         root sensitivity (1 + max(c1,c2) h)/p(xk) plus the formula's cancellation term */
      const double pint = fabs(p[lo] * w1 + p[lo + 1] * w2);
      const double den = fabs(-Bq + root);
      const double kap = (Aq == 0.0) ? 1.0 : (den > 0.0 ? fabs(Bq) / den : INFINITY);
      const double xa = fabs(x1) > fabs(x2) ? fabs(x1) : fabs(x2);
      cond[m + M * k] = (1.0 + (c1 > c2 ? c1 : c2) * hq) / pint + kap * xa;
    }
    if (lsens) {
      /* |d log p(x_k) / d x_k| of the interpolated conditional: how a perturbation of xk shows up in lPz */
      const double pint = fabs(p[lo] * w1 + p[lo + 1] * w2);
      lsens[m + M * k] = pint > 0.0 ? fabs(c2 - c1) / (hq * pint) : INFINITY;
    }
    /* interface update: slab = w1*core[:,lo,:] + w2*core[:,lo+1,:]; left <- left*slab (ref :167-177) */
    if (k < md->d - 1) {
      for (oidx b = 0; b < rn; b++) {
        const double *s1 = ck + lo * rk + b * rk * nk, *s2 = s1 + rk;
        for (oidx a = 0; a < rk; a++) {
          double t = w1 * s1[a];
          t += w2 * s2[a];
          slab[a + b * rk] = t;
        }
      }
      for (oidx b = 0; b < rn; b++) {
        double s = 0.0;
        for (oidx a = 0; a < rk; a++) s += slab[a + b * rk] * left[a];
        next[b] = s;
      }
      for (oidx b = 0; b < rn; b++) left[b] = next[b];
    }
  }
  lPz[m] = lp;
}

/* Full-featured entry: 64-bit sizes; idx/kappa/gap may be NULL.  Rows [m_begin, m_end)
   only (M is still the leading dimension), so callers can shard over host threads. */
int tt_irt1_oracle_rows(oidx d, const oidx *n, const double *xs, const oidx *ttrank, const double *ttcore,
                        oidx M, oidx m_begin, oidx m_end, const double *q, double *z, double *lPz,
                        int *idx, double *kappa, double *gap, double *cond, double *lsens) {
  model_t md; memset(&md, 0, sizeof(md));
  md.d = d; md.n = n; md.r = ttrank; md.xs = xs; md.core = ttcore;
  if (d < 1 || model_build(&md) != 0) { model_free(&md); return -1; }
  const oidx rm = md.rmax > 1 ? md.rmax : 1, nm = md.nmax;
  double *left = (double *)malloc(sizeof(double) * rm), *next = (double *)malloc(sizeof(double) * rm);
  double *p = (double *)malloc(sizeof(double) * nm), *cdf = (double *)malloc(sizeof(double) * nm);
  double *slab = (double *)malloc(sizeof(double) * rm * rm);
  for (oidx m = m_begin; m < m_end; m++)
    walk_sample(&md, M, m, q, z, lPz, idx, kappa, gap, cond, lsens, left, next, p, cdf, slab);
  free(left); free(next); free(p); free(cdf); free(slab);
  model_free(&md);
  return 0;
}

int tt_irt1_oracle(oidx d, const oidx *n, const double *xs, const oidx *ttrank, const double *ttcore,
                   oidx M, const double *q, double *z, double *lPz, int *idx, double *kappa, double *gap, double *cond, double *lsens) {
  return tt_irt1_oracle_rows(d, n, xs, ttrank, ttcore, M, 0, M, q, z, lPz, idx, kappa, gap, cond, lsens);
}

/* The right marginals and core*marginal products alone (checks the device sweep). */
int tt_irt1_oracle_sweep(oidx d, const oidx *n, const double *xs, const oidx *ttrank, const double *ttcore,
                         double *pk_out /* sum r_k n_k */, double *marg_out /* sum r_{k+1} */) {
  model_t md; memset(&md, 0, sizeof(md));
  md.d = d; md.n = n; md.r = ttrank; md.xs = xs; md.core = ttcore;
  if (d < 1 || model_build(&md) != 0) { model_free(&md); return -1; }
  oidx op = 0, om = 0;
  for (oidx k = 0; k < d; k++) {
    memcpy(pk_out + op, md.pk[k], sizeof(double) * ttrank[k] * n[k]); op += ttrank[k] * n[k];
    memcpy(marg_out + om, md.marg[k], sizeof(double) * ttrank[k + 1]); om += ttrank[k + 1];
  }
  model_free(&md);
  return 0;
}
