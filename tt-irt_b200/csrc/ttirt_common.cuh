// Shared declarations for the B200-native tt_irt1 engine (internal; the public C-ABI is include/tt_irt1.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttirt {

// Per-dimension metadata, resident on the device next to the cores.
struct DimInfo {
  int n;            // grid size n_k
  int r0, r1;       // r_k, r_{k+1}
  int pad;
  int64_t off_x;    // start of grid k in xs        (reference psxs, tt_irt1_int32.c:43-49)
  int64_t off_c;    // start of core k in ttcore    (reference pstt, :55-57)
  int64_t off_p;    // start of P_k (r_k x n_k)     in the stacked product buffer
  int64_t off_m;    // start of C_k (r_{k+1})       in the stacked marginal buffer
  int64_t off_pw;   // start of the node-weighted copy of P_k (fast path; 16-byte aligned: even offsets)
};

struct CellOut {
  double xk, w1, w2, logp;
};

// The reference's per-sample tail (tt_irt1_int32.c:146-165): closed-form root of the piecewise-quadratic
// CDF on [x1, x2], interpolation weights and log of the interpolated conditional density.  Written with
// explicit round-to-nearest intrinsics so that no FMA contraction can change the operation order: the
// reference's formula cancels and is mirrored verbatim, not "fixed".
__device__ __forceinline__ CellOut invert_cell(double qk, double cdf_lo, double c1, double c2, double x1, double x2) {
  CellOut o;
  const double hq = __dsub_rn(x2, x1);
  const double Aq = __ddiv_rn(__dmul_rn(0.5, __dsub_rn(c2, c1)), hq);
  const double Bq = __ddiv_rn(__dsub_rn(__dmul_rn(c1, x2), __dmul_rn(c2, x1)), hq);
  double Dq = __dadd_rn(__dmul_rn(__dmul_rn(2.0, Aq), x1), Bq);
  Dq = __dmul_rn(Dq, Dq);
  const double dq = __dsub_rn(qk, cdf_lo);
  Dq = __dadd_rn(Dq, __dmul_rn(__dmul_rn(4.0, Aq), dq));
  const double root = __dsqrt_rn(fabs(Dq));
  double xk = __ddiv_rn(__dmul_rn(0.5, __dadd_rn(-Bq, root)), Aq);
  if (Aq == 0.0) xk = __dadd_rn(x1, __ddiv_rn(dq, Bq));
  o.xk = xk;
  o.w1 = __ddiv_rn(__dsub_rn(x2, xk), hq);
  o.w2 = __ddiv_rn(__dsub_rn(xk, x1), hq);
  o.logp = log(fabs(__dadd_rn(__dmul_rn(c1, o.w1), __dmul_rn(c2, o.w2))));
  return o;
}

// Fast-path variant.  It works on the conditional scaled by an exact power of two instead of normalised by its
// mass (the root x_k and the weights are invariant under a common scaling of c1, c2 and q - cdf_lo), so no
// reciprocal of the mass is needed.  The four divisions by the cell width h = x2 - x1 become multiplications by a
// tabulated 1/h and the division by Aq a multiplication by its reciprocal, which is taken as soon as Aq exists and
// so leaves the dependent chain (one extra rounding each, the size of the roundings the reference's own
// divisions commit).  The logarithm is not taken here: the tail returns the interpolated (scaled) conditional
// density and the caller keeps the running product in the split form of lp_accumulate().
//   dq = q * mass - cdf_lo  (scaled),  c1, c2 = |pdf| at the cell's nodes (scaled)
#ifndef TTIRT_XK_DIV
#define TTIRT_XK_DIV 0
#endif
struct CellFast {
  double xk, w1, w2, dens;
};
__device__ __forceinline__ CellFast invert_cell_fast(double dq, double c1, double c2, double x1, double x2, double ih) {
  // reference :146-159 with A2 = 2 Aq:  xk = (-Bq + sqrt((A2 x1 + Bq)^2 + 2 A2 dq)) / A2.  Scalar FP64 instructions
  // are the scarce resource next to a DMMA stream (see ttirt_fast.cu), so products feed fused multiply-adds here.
  CellFast o;
  const double A2 = __dmul_rn(__dsub_rn(c2, c1), ih);
  const double rA = __drcp_rn(A2);
  const double Bq = __dmul_rn(fma(c1, x2, -__dmul_rn(c2, x1)), ih);
  const double t = fma(A2, x1, Bq);
  const double Dq = fma(__dadd_rn(A2, A2), dq, __dmul_rn(t, t));
  const double root = __dsqrt_rn(fabs(Dq));
  double xk = __dmul_rn(__dsub_rn(root, Bq), rA);
  if ((__double_as_longlong(A2) << 1) == 0) xk = __dadd_rn(x1, __ddiv_rn(dq, Bq));   // Aq == 0.0: linear CDF
  o.xk = xk;
  o.w1 = __dmul_rn(__dsub_rn(x2, xk), ih);
  o.w2 = __dmul_rn(__dsub_rn(xk, x1), ih);
  o.dens = fabs(fma(c1, o.w1, __dmul_rn(c2, o.w2)));
  return o;
}

// Running log-density of the fast path: lPz = sum_k log(dens_k / mass_k) (reference tt_irt1_int32.c:116-130,
// 161-165) is carried as two running products in split form, Pn * 2^E over the densities and Pd over the masses,
// both mantissas in [1, 2), and the one logarithm is taken after the last dimension.  Zero, infinite and NaN
// products stay as they are and give -inf / inf / NaN as a sum of logarithms would.
static __device__ __noinline__ int split_exponent_subnormal(double &P) {
  P = __dmul_rn(P, 18014398509481984.0);  // 2^54
  const int hi = __double2hiint(P);
  const int ex = (hi >> 20) & 0x7ff;
  P = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(P));
  return ex - 1023 - 54;
}
__device__ __forceinline__ int split_exponent(double &P) {
  const int hi = __double2hiint(P);
  const int ex = (hi >> 20) & 0x7ff;
  if (ex == 0) {  // zero stays zero; a subnormal product is rescaled before splitting (rare: out of line)
    if (((hi & 0x7fffffff) | __double2loint(P)) == 0) return 0;
    return split_exponent_subnormal(P);
  }
  if (ex == 0x7ff) return 0;
  P = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(P));
  return ex - 1023;
}
__device__ __forceinline__ void lp_accumulate(double &Pn, double &Pd, int &E, double dens, double mass) {
  Pn = __dmul_rn(Pn, dens);
  E += split_exponent(Pn);
  Pd = __dmul_rn(Pd, mass);
  E -= split_exponent(Pd);
}
__device__ __forceinline__ double lp_finish(double Pn, double Pd, int E) {
  return fma((double)E, 0.6931471805599453094, log(__ddiv_rn(Pn, Pd)));
}
// 2^-e(x) for a positive normal x (1.0 otherwise): scaling by it is exact
__device__ __forceinline__ double pow2_scale(double x) {
  const int ex = (__double2hiint(x) >> 20) & 0x7ff;
  return (ex != 0 && ex != 0x7ff) ? __hiloint2double((2046 - ex) << 20, 0) : 1.0;
}

// Trapezoid weight of grid node j (n nodes x[0..n-1]): the CDF of the piecewise-linear density with node values p is
//   cdf_j = sum_{i<j} w_i p_i + h_{j-1} p_j,   w_i = h_{i-1} + h_i,   h_i = (x_{i+1} - x_i) / 2   (h_{-1} = 0),
// and cdf_{n-1} (the mass) = sum_{i<n-1} w_i p_i + h_{n-2} p_{n-1}: the last node carries h_{n-2}, nodes beyond it zero.
__device__ __forceinline__ double node_weight(const double *__restrict__ x, int j, int n) {
  if (j >= n) return 0.0;
  const double hl = j >= 1 ? 0.5 * (x[j] - x[j - 1]) : 0.0;
  const double hr = j + 1 < n ? 0.5 * (x[j + 1] - x[j]) : 0.0;
  return hl + hr;
}

// Exclusive scans of an interval histogram (nb <= 96 bins) by ONE warp: bst[b] = first sorted row of bin b, bts[b] = first
// CTA tile of bin b (tiles never straddle bins), entry nb = totals.  Every CTA of the scatter and the transition kernel
// recomputes them from the global histogram; that is cheaper than a launch of its own.  bts may be null.
__device__ __forceinline__ void bin_offsets_warp(const int *__restrict__ hist, int nb, int rows_per_tile, int *bst, int *bts, int lane) {
  int cs = 0, ct = 0;
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int b = b0 + lane;
    const int c = b < nb ? hist[b] : 0;
    const int t = (c + rows_per_tile - 1) / rows_per_tile;
    int is = c, it = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int us = __shfl_up_sync(0xffffffffu, is, o), ut = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { is += us; it += ut; }
    }
    if (b < nb) { bst[b] = cs + is - c; if (bts) bts[b] = ct + it - t; }
    cs += __shfl_sync(0xffffffffu, is, 31); ct += __shfl_sync(0xffffffffu, it, 31);
  }
  if (lane == 0) { bst[nb] = cs; if (bts) bts[nb] = ct; }
}

// Arguments of one fused "transition" launch of the fast path: interface update through dimension k
// (binned by the interval chosen there) followed by the whole conditional step of dimension k+1.
struct TransArgs {
  const double *core;        // core_k, column-major r0 x n0 x r1
  const double *pnext;       // P_{k+1} with column j scaled by node_weight(j), column-major r1 x n1, 16-byte aligned
  const double *xnext;       // grid of dimension k+1 (n1)
  int r0, n0, r1, n1;
  int async_ok;              // core is 16-byte aligned: slabs may be staged by cp.async when r0 is even
  int last;                  // k+1 == d-1: no further interface update
  int rows;                  // samples in this chunk
  double *F;                 // left-interface rows, rows x ldf, updated in place
  int ldf;
  const int *perm;           // samples ordered by the interval chosen in dimension k
  const int *hist_cur;       // n0-1 counters: samples per interval chosen in dimension k (every CTA scans them itself)
  int *idx;                  // per sample: interval index (out: dimension k+1)
  double *w1, *w2;           // per sample: interpolation weights (in: dimension k, out: k+1)
  double *lp;                // per sample: running product of the scaled densities, mantissa (see lp_accumulate)
  double *lpd;               // per sample: running product of the scaled masses, mantissa
  int *lpe;                  // per sample: exponent of the ratio of the two products
  const double *q;           // column k+1 of q for this chunk
  double *z;                 // column k+1 of z
  int32_t *idx_out;          // column k+1 of the exported index array (may be NULL)
  double *lpz;               // final log-density (written when last)
  int *hist_next;            // n1-1 counters for binning dimension k+1 (unused when last)
};

// Arguments of the all-dimensions "walk" kernel (ttirt_walk.cu): small uniform-rank TTs, one launch per chunk.
struct WalkArgs {
  const double *pack;        // per-dimension operand blocks (walk_pack), 16-byte aligned
  int d, rows;
  const double *q;           // column-major rows x d, leading dimension ldq
  int64_t ldq;
  double *z;                 // column-major, leading dimension ldz
  int64_t ldz;
  double *lpz;
  int32_t *idx_out;          // may be NULL; leading dimension ldz
};
int walk_class_for(int r, int n);                    // -1: shape not served by the walk kernel
int64_t walk_pack_doubles(int cls, int d);
cudaError_t walk_init(int device);
cudaError_t walk_pack(int cls, const DimInfo *d_dims, int d, const double *xs, const double *core, const double *pk,
                      const double *p0, const double *cdf0, double *pack, cudaStream_t st);
cudaError_t launch_walk(int cls, const WalkArgs &a, int sm_count, cudaStream_t st);

// One dimension step k -> k+1 of the wide path (ttirt_wide.cu: ranks / grids beyond the fused transition kernel): grouped
// DMMA GEMM for the interface update, DMMA GEMM for the weighted conditional pdf, one-thread-per-sample tail.
struct WideArgs {
  const double *core;        // core_k, column-major r0 x n0 x r1
  const double *pnext;       // P_{k+1} with column j scaled by node_weight(j), column-major r1 x n1
  const double *xnext, *ihnext, *rwnext, *hrnext;   // grid of dimension k+1 and its tables (wide_tables)
  int r0, n0, r1, n1;
  int last, rows;
  const double *Fin;         // left-interface rows of dimension k (rows x ldf, zero-padded to r0 rounded to 8)
  double *Fout;              // ... of dimension k+1 (a different buffer)
  int ldf;
  double *pb;                // n1 x rows scratch: weighted signed pdf, node-major
  double *mass_part;         // wide_mass_slots(nmax) x rows scratch: shares of the rows' masses
  const int *perm, *hist_cur;
  int *idx;
  double *w1, *w2, *lp, *lpd;
  int *lpe;
  const double *q;           // column k+1 of q for this chunk
  double *z;
  int32_t *idx_out;
  double *lpz;
  int *hist_next;
};
cudaError_t launch_wide_step(const WideArgs &a, cudaStream_t st);   // three launches
int wide_mass_slots(int nmax);
cudaError_t wide_init(int device);                  // opt in to the GEMM's dynamic shared memory
cudaError_t wide_tables(const DimInfo *d_dims, int d, const double *xs, double *ih, double *rw, double *hr, cudaStream_t st);
constexpr int kWideClass = 3;

// fast-path shape classes: 0 - 2 the fused transition kernel (rank tiles of 8, grid tiles of 8), 3 the wide path
int fast_class_for(int rmax, int nmax);             // -1: shape outside the fast path
cudaError_t launch_transition(int cls, const TransArgs &a, int sm_count, cudaStream_t st);
cudaError_t fast_init(int device);                  // opt in to large dynamic shared memory

}  // namespace ttirt
