// B200-native squared-density inverse Rosenblatt transform behind include/tt_irt_sqr.h
// (reference matlab/samplers/tt_irt_sqr.m:1-208 and its MEX helper matlab/utils/tracemult.c).
//
//   model_create : extend the cores to the boundary (:53-60), right-to-left sweep (:41-82) on the device:
//                    core x R (:62-64) -> Householder QR of the weighted unfolding (:66-72) -> Cartesian square (:74-80),
//                  the square stored as the packed symmetric DMMA operand of the conditional-pdf contraction.
//   sample       : per chunk and dimension
//                    sqr_pdf_kernel     (f (x) f)' P_k  (:107-112) as an FP64 tensor-core GEMM whose A operand, the outer
//                                       product of the left interface with itself, is formed in registers and whose B
//                                       operand streams from L2 through a TMA-bulk / mbarrier ring
//                    sqr_tail_kernel    trapezoid CDF, zero-mass fallback, normalise, bisection, quadratic root, clamp,
//                                       log-density (:113-195), in the reference's operation order
//                    bin scan / scatter counting sort of the chunk by chosen interval
//                    sqr_update_kernel  interpolated interface update (:197-207): two FP64 tensor-core products per tile with
//                                       the two core slabs of a bin staged in shared memory
//   forward      : tt_rt_sqr.m (x -> q): the same kernels with the tail's other branch (tt_rt_sqr.m:129-166)
//   DIRT loops   : tt_dirt_sample.m:17-73 and tt_dirt_inverse.m:24-59 over resident level models
// No CPU fallback anywhere: without a device every entry point fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tt_irt_sqr.h"
#include "ttirt_common.cuh"

namespace ttirt {
int aux_fail(const char *fmt, ...);   // ttirt_engine.cu: records the thread's last error, one line on stderr
void aux_launched();                  // bumps the library's kernel-launch counter
}  // namespace ttirt
using ttirt::aux_fail;

#define CKS(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return aux_fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define LAUNCHED() ttirt::aux_launched()

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// device allocations: blocks are handed back to a per-device pool instead of cudaFree (which synchronises the device and,
// with cudaMalloc, costs more than the kernels of a small call): DIRT calls the transform once per layer with same-shaped
// TTs (tt_dirt_sample.m:46,71).  No results survive in it; TTIRT_CACHE=0 disables it, ttirt_cache_clear() empties it.
// ------------------------------------------------------------------------------------------------
struct DevPool {
  std::mutex mu;
  std::multimap<size_t, void *> idle;
  std::unordered_map<void *, size_t> live;
  size_t idle_bytes = 0;
};

// idle blocks a device may hold back (TTIRT_POOL_MB, default 8192): callers that keep changing shapes would otherwise
// accumulate one set of blocks per shape
size_t pool_limit() {
  static size_t v = 0;
  if (v == 0) { const char *e = getenv("TTIRT_POOL_MB"); v = ((e && atoll(e) > 0) ? (size_t)atoll(e) : (size_t)8192) << 20; }
  return v;
}
DevPool g_pool[64];

bool pool_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("TTIRT_CACHE"); v = (e && atoi(e) == 0 && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

template <typename T>
cudaError_t sqr_alloc(T **p, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes == 0) bytes = 8;
  if (pool_enabled() && dev >= 0 && dev < 64) {
    DevPool &pl = g_pool[dev];
    std::lock_guard<std::mutex> lock(pl.mu);
    auto it = pl.idle.find(bytes);
    if (it != pl.idle.end()) {
      *p = static_cast<T *>(it->second);
      pl.live[it->second] = bytes;
      pl.idle.erase(it);
      pl.idle_bytes -= bytes;
      return cudaSuccess;
    }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {   // out of memory: give the idle blocks back and retry once
      cudaGetLastError();
      for (auto &kv : pl.idle) cudaFree(kv.second);
      pl.idle.clear();
      pl.idle_bytes = 0;
      e = cudaMalloc(&q, bytes);
      if (e != cudaSuccess) return e;
    }
    pl.live[q] = bytes;
    *p = static_cast<T *>(q);
    return cudaSuccess;
  }
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes);
  *p = static_cast<T *>(q);
  return e;
}

void sqr_free(void *p) {
  if (!p) return;
  int dev = 0;
  cudaGetDevice(&dev);
  if (pool_enabled() && dev >= 0 && dev < 64) {
    DevPool &pl = g_pool[dev];
    std::lock_guard<std::mutex> lock(pl.mu);
    auto it = pl.live.find(p);
    if (it != pl.live.end()) {
      const size_t bytes = it->second;
      pl.live.erase(it);
      // over the limit: the largest idle blocks go back to the driver first (callers of sqr_free have synchronised)
      while (!pl.idle.empty() && pl.idle_bytes + bytes > pool_limit()) {
        auto big = std::prev(pl.idle.end());
        cudaFree(big->second);
        pl.idle_bytes -= big->first;
        pl.idle.erase(big);
      }
      if (bytes > pool_limit()) { cudaFree(p); return; }
      pl.idle.emplace(bytes, p);
      pl.idle_bytes += bytes;
      return;
    }
  }
  cudaFree(p);
}

// ------------------------------------------------------------------------------------------------
// geometry of the packed Gram operand
// ------------------------------------------------------------------------------------------------
// The conditional pdf at node j is  sum_{a,b} f_a f_b G[a,b,j]  (tt_irt_sqr.m:107-112, G = P{k} of :80, symmetric in a, b).
// The contraction index is walked in DMMA k-steps of four pairs (a, b), a <= b, each pair once:
//   off-diagonal part   for every column block c >= 1 (b = 4c .. 4c+3) and every a < 4c the k-step (c, a) holds the pairs
//                       (a, 4c + t), t = 0..3, with 2 G (they stand for (a, b) and (b, a));  2 nc (nc - 1) k-steps, nc = ceil(r / 4)
//   diagonal part       the ten pairs 4c <= a <= b <= 4c+3 of every block, packed four to a k-step in block order, with G on
//                       the diagonal and 2 G off it;  ceil(10 nc / 4) k-steps
// r (r + 1) / 2 k-rows for r a multiple of 4: 2080 instead of r^2 = 4096 at r = 64.  Pairs that reach past r carry zeros.
__host__ __device__ inline int sqr_offdiag_ksteps(int r0) {
  const int nc = (r0 + 3) >> 2;
  return 2 * nc * (nc - 1);
}
__host__ __device__ inline int sqr_ksteps(int r0) {
  const int nc = (r0 + 3) >> 2;
  return 2 * nc * (nc - 1) + ((10 * nc + 3) >> 2);
}
// pair p of the diagonal part -> (a, b): block p / 10, position in the block's upper triangle p % 10
__host__ __device__ inline void sqr_diag_pair(int p, int &a, int &b) {
  const int cc = p / 10, q = p - 10 * cc;
  a = 4 * cc + (int)((0x3221110000ULL >> (4 * q)) & 15);
  b = 4 * cc + (int)((0x3323213210ULL >> (4 * q)) & 15);
}
// k-row kr = 4 ks + t of the packed operand -> (a, b, weight); weight 0: padding
__host__ __device__ inline void sqr_krow_pair(int kr, int r0, int &a, int &b, double &w) {
  const int ks = kr >> 2, t = kr & 3, noff = sqr_offdiag_ksteps(r0);
  if (ks < noff) {
    int c = 1;
    while (ks >= 2 * (c + 1) * c) c++;        // block c owns the k-steps [2 c (c - 1), 2 (c + 1) c)
    a = ks - 2 * c * (c - 1); b = 4 * c + t; w = 2.0;
  } else {
    sqr_diag_pair(4 * (ks - noff) + t, a, b);
    w = a == b ? 1.0 : 2.0;
  }
  if (b >= r0 || a >= r0) w = 0.0;
}

struct SqrDim {
  int n;         // grid size of the dimension, boundary points included
  int n_in;      // mode size of the caller's core (n or n - 2)
  int r0, r1;    // r_k, r_{k+1}
  int s1;        // columns of the factor R' to the right of core k (1 for the last core)
  int s0;        // columns of the factor R' this dimension produces (min(n s1, r0)); unused for k = 0
  int ksteps;    // DMMA k-steps of the packed operand
  int pad;
  int64_t off_x, off_cin, off_c, off_g, off_r;
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: FP64 MMA, mbarrier, TMA bulk copy
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// sweep kernels (once per model)
// ------------------------------------------------------------------------------------------------
// tt_irt_sqr.m:53-60: cores without boundary nodes are extrapolated linearly to the two extra grid points.
// Operation order as written there (note the right-hand factor (h_n + h_{n-1}) / h_{n-1}: mirrored, not "fixed").
__global__ void sqr_extend_kernel(const double *__restrict__ cin, const double *__restrict__ x, double *cout, int r0, int n, int r1) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // over (a, l)
  if (e >= r0 * r1) return;
  const int a = e % r0, l = e / r0;
  const int nin = n - 2;
  const double *ci = cin + a + (int64_t)r0 * nin * l;
  double *co = cout + a + (int64_t)r0 * n * l;
  for (int j = 0; j < nin; j++) co[(int64_t)r0 * (j + 1)] = ci[(int64_t)r0 * j];
  const double h1 = __dsub_rn(x[1], x[0]), h2 = __dsub_rn(x[2], x[1]);
  const double hn = __dsub_rn(x[n - 1], x[n - 2]), hm = __dsub_rn(x[n - 2], x[n - 3]);
  const double f1 = ci[0], f2 = ci[r0];
  co[0] = __dsub_rn(f1, __ddiv_rn(__dmul_rn(__dsub_rn(f2, f1), h1), h2));
  const double g1 = ci[(int64_t)r0 * (nin - 1)], g2 = ci[(int64_t)r0 * (nin - 2)];
  co[(int64_t)r0 * (n - 1)] = __dadd_rn(g1, __ddiv_rn(__dmul_rn(__dsub_rn(g1, g2), __dadd_rn(hn, hm)), hm));
}

// h (h[0] = 0) and its running sum, the fallback CDF of :123-127
__global__ void sqr_grid_kernel(const double *__restrict__ x, double *h, double *hc, int n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  h[0] = 0.0; hc[0] = 0.0;
  for (int j = 1; j < n; j++) {
    const double hj = __dsub_rn(x[j], x[j - 1]);
    h[j] = hj;
    s = __dadd_rn(s, hj);
    hc[j] = s;
  }
}

// :62-64  Pm[(j + n s) + m a] = sum_l core[a, j, l] R'[l, s]   (m = n s1: the unfolding that the QR and the square read)
__global__ void sqr_contract_kernel(const double *__restrict__ core, const double *__restrict__ Rp, double *Pm, int r0, int n,
                                    int r1, int s1) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)r0 * n * s1) return;
  const int a = (int)(e % r0);
  const int j = (int)((e / r0) % n);
  const int s = (int)(e / ((int64_t)r0 * n));
  const double *c = core + a + (int64_t)r0 * j;
  const double *rr = Rp + (int64_t)r1 * s;
  double acc = 0.0;
  for (int l = 0; l < r1; l++) acc = fma(c[(int64_t)r0 * n * l], rr[l], acc);
  Pm[(int64_t)j + (int64_t)n * s + (int64_t)n * s1 * a] = acc;
}

// :49-51, :68  A = diag(sqrt(w / 2)) Pm, w the trapezoid node weights
__global__ void sqr_weight_kernel(const double *__restrict__ Pm, const double *__restrict__ h, double *A, int n, int m, int r0) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)m * r0) return;
  const int j = (int)((e % m) % n);
  const double hr = j + 1 < n ? h[j + 1] : 0.0;
  const double w = sqrt(__dmul_rn(__dadd_rn(h[j], hr), 0.5));
  A[e] = __dmul_rn(Pm[e], w);
}

// :70-71  R of the economy QR of A (m x r): communication-avoiding (TSQR) Householder.  Each CTA factors a block of rows
// held in shared memory (reflectors as LAPACK dgeqr2 / dlarfg form them, one warp per trailing column) and emits its R;
// the stacked R factors are factored again until one block is left.  Only R'R enters the result (the factor's row signs and
// the order of the rows are immaterial), and Householder QR of stacked R factors is as backward stable as the flat one.
//   in   : column-major, leading dimension lda, rows [b * rb, min(m, (b + 1) * rb)) belong to CTA b
//   out  : final == 0: stack (column-major (gridDim.x * r) x r), CTA b writes rows [b * r, (b + 1) * r) (zero below the factor)
//          final == 1: R' (r x rnew column-major, rnew = min(m, r)), single CTA
__global__ void __launch_bounds__(1024) sqr_qr_block_kernel(const double *__restrict__ in, int m, int lda, int r, int rb, double *out,
                                                            int ldo, int final, int rnew) {
  extern __shared__ __align__(16) double As[];     // mb x r, column-major, pitch mb
  __shared__ double red[32];
  __shared__ double s_tau, s_scale;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int row0 = blockIdx.x * rb;
  const int mb = min(rb, m - row0);
  for (int e = tid; e < mb * r; e += blockDim.x) {
    const int i = e % mb, c = e / mb;
    As[e] = in[(size_t)row0 + i + (size_t)lda * c];
  }
  __syncthreads();
  const int steps = mb < r ? mb : r;
  for (int c = 0; c < steps; c++) {
    double *col = As + (size_t)c * mb;
    double ss = 0.0;
    for (int i = c + 1 + tid; i < mb; i += blockDim.x) ss = fma(col[i], col[i], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    if (warp == 0) {
      double v = lane < nw ? red[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == 0) {
        const double alpha = col[c];
        if (v == 0.0) {
          s_tau = 0.0; s_scale = 0.0;
        } else {
          const double beta = -copysign(sqrt(fma(alpha, alpha, v)), alpha);
          s_tau = (beta - alpha) / beta;
          s_scale = 1.0 / (alpha - beta);
          col[c] = beta;
        }
      }
    }
    __syncthreads();
    const double tau = s_tau, scale = s_scale;
    if (tau != 0.0) {
      for (int i = c + 1 + tid; i < mb; i += blockDim.x) col[i] *= scale;
      __syncthreads();
      for (int cc = c + 1 + warp; cc < r; cc += nw) {
        double *cj = As + (size_t)cc * mb;
        double dot = 0.0;
        for (int i = c + 1 + lane; i < mb; i += 32) dot = fma(col[i], cj[i], dot);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
        dot += cj[c];
        const double t = tau * dot;
        if (lane == 0) cj[c] -= t;
        for (int i = c + 1 + lane; i < mb; i += 32) cj[i] = fma(-t, col[i], cj[i]);
      }
    }
    __syncthreads();
  }
  if (final) {
    for (int e = tid; e < r * rnew; e += blockDim.x) {
      const int a = e % r, t = e / r;
      out[e] = (t <= a && t < steps) ? As[(size_t)t + (size_t)mb * a] : 0.0;
    }
  } else {
    for (int e = tid; e < r * r; e += blockDim.x) {
      const int t = e % r, a = e / r;
      out[(size_t)blockIdx.x * r + t + (size_t)ldo * a] = (t <= a && t < steps) ? As[(size_t)t + (size_t)mb * a] : 0.0;
    }
  }
}

// :74-80 as the packed DMMA operand: k-row kr holds w G[a, b, :] for its pair (sqr_krow_pair), G[a, b, j] = sum_s Pm[a,j,s] Pm[b,j,s]
__global__ void sqr_gram_pack_kernel(const double *__restrict__ Pm, int n, int s1, int r0, double *gp, int pb, int ksteps) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)ksteps * 4 * n) return;
  const int j = (int)(e % n);
  const int kr = (int)(e / n);
  int a, b;
  double w;
  sqr_krow_pair(kr, r0, a, b, w);
  double v = 0.0;
  if (w != 0.0) {
    const int64_t m = (int64_t)n * s1;
    const double *pa = Pm + j + m * a, *pbp = Pm + j + m * b;
    for (int s = 0; s < s1; s++) v = fma(pa[(int64_t)n * s], pbp[(int64_t)n * s], v);
    v *= w;
  }
  gp[(int64_t)kr * pb + j] = v;
}

// ------------------------------------------------------------------------------------------------
// per-chunk kernels
// ------------------------------------------------------------------------------------------------
// fkm1 = ones(1, Mb)  (:103): r_0 = 1, the padding columns of the first 4-block are zero
__global__ void sqr_init_kernel(double *F, int ldf, int rows) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  double *f = F + (size_t)m * ldf;
  f[0] = 1.0; f[1] = 0.0; f[2] = 0.0; f[3] = 0.0;
}

constexpr int PDF_WARPS = 8;                  // MMA warps, two per SM sub-partition
// 8-row MMA tiles per warp (template parameter MT): 2 for the r <= 64 / n <= 72 class (16 column-tile accumulator chains per
// k-step), 4 for the two lighter classes, whose few column tiles would otherwise leave a k-step too short to hide the
// operand loads behind its DMMAs
constexpr int PDF_KS = 16;                    // k-steps per B slice
// ring depth (template parameter PDF_STAGES): three slices, two in the light two-CTAs-per-SM configuration

struct TailArgs {
  const double *pdf; double *cdf; int64_t ldp; int rows; int n;
  const double *x, *h, *hc;
  const double *q; double *z; int32_t *idx_out;
  int *idx; double *w1, *w2; double *lf;
  int first;        // dimension 0: lFapp starts here (:91)
  int forward;      // 0: inverse transform (tt_irt_sqr.m: q -> x), 1: forward transform (tt_rt_sqr.m: x -> q)
  int *hist;        // n - 1 counters for the sort by interval, or NULL when no interface update follows
};

// One increment of the trapezoid CDF, :114-117:  ((0.5 p_{j-1} + 0.5 p_j) h_j), no FMA contraction
__device__ __forceinline__ double cdf_step(double pprev, double pj, double hj) {
  return __dmul_rn(__dadd_rn(__dmul_rn(0.5, pprev), __dmul_rn(0.5, pj)), hj);
}

// Everything after the cell i0 is known, :129-130, :146-195 (inverse) / tt_rt_sqr.m:141-163 (forward), in the reference's
// operation order.  c1raw, p1raw, p2raw: unnormalised CDF and conditional at the cell's nodes; cmax the normalisation.
struct TailOut { int i0; double wa, wb; };
__device__ __forceinline__ TailOut tail_finish(const TailArgs &a, int m, int i0, double qk, double c1raw, double p1raw, double p2raw,
                                               double cmax, int *sh_hist) {
  const double C1 = __ddiv_rn(c1raw, cmax);                                                       // :146-149, :129-130
  const double f1 = __ddiv_rn(p1raw, cmax);
  const double f2 = __ddiv_rn(p2raw, cmax);
  const double x1 = a.x[i0], x2 = a.x[i0 + 1];                                                    // :157-159
  const double h3 = __dsub_rn(x2, x1);
  const double Aq = __ddiv_rn(__dmul_rn(0.5, __dsub_rn(f2, f1)), h3);                            // :161
  double xk, out;
  if (a.forward) {
    xk = qk;
    const double dx = __dsub_rn(xk, x1);
    out = __dadd_rn(__dadd_rn(__dmul_rn(Aq, __dmul_rn(dx, dx)), __dmul_rn(f1, dx)), C1);         // tt_rt_sqr.m:151
  } else {
    const double dq = __dsub_rn(qk, C1);
    const double Dq = __dadd_rn(__dmul_rn(f1, f1), __dmul_rn(__dmul_rn(4.0, Aq), dq));           // :162
    xk = __dadd_rn(x1, __ddiv_rn(__dadd_rn(-f1, __dsqrt_rn(fabs(Dq))), __dmul_rn(2.0, Aq)));     // :163
    if (Aq == 0.0) {                                                                              // :164-170
      xk = __dadd_rn(x1, __ddiv_rn(dq, f1));
      if (f1 == 0.0) xk = x1;
    }
    if (xk > x2) xk = x2;                                                                         // :173-182
    if (xk < x1) xk = x1;
    out = xk;
  }
  a.z[m] = out;                                                                                   // :184 / tt_rt_sqr.m:153
  const double wa = __ddiv_rn(__dsub_rn(x2, xk), h3), wb = __ddiv_rn(__dsub_rn(xk, x1), h3);      // :187-188
  const double dens = __dadd_rn(__dmul_rn(f1, wa), __dmul_rn(f2, wb));                            // :193
  const double lg = log(dens);                                                                    // :194
  a.lf[m] = a.first ? lg : __dadd_rn(a.lf[m], lg);
  a.idx[m] = i0; a.w1[m] = wa; a.w2[m] = wb;
  if (a.idx_out) a.idx_out[m] = i0;
  if (sh_hist) atomicAdd(&sh_hist[i0], 1);
  TailOut o; o.i0 = i0; o.wa = wa; o.wb = wb;
  return o;
}

// :113-195, one thread per sample, every operation in the reference's order (explicit round-to-nearest intrinsics: no
// FMA contraction).  The unnormalised CDF goes to a scratch column block that the bisection reads back.
// (The path of shapes whose pdf tile cannot be parked in shared memory; otherwise the tail runs inside sqr_pdf_kernel.)
__global__ void __launch_bounds__(256) sqr_tail_kernel(TailArgs a) {
  extern __shared__ int sh_hist[];
  const int n = a.n;
  if (a.hist) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
  }
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < a.rows) {
    const double *p = a.pdf + m;
    double *cd = a.cdf + m;
    double pprev = p[0];
    double c = __dmul_rn(__dmul_rn(0.5, pprev), a.h[0]);       // :114-116, h(1) = 0
    cd[0] = c;
    for (int j = 1; j < n; j++) {
      const double pj = p[(size_t)j * a.ldp];
      c = __dadd_rn(c, cdf_step(pprev, pj, a.h[j]));           // :117 cumsum
      cd[(size_t)j * a.ldp] = c;
      pprev = pj;
    }
    double cmax = c;                                           // :120
    const bool zero = cmax <= 0.0;                             // :121-127: conditional replaced by h, CDF by cumsum(h)
    if (zero) cmax = a.hc[n - 1];
    const double qk = a.q[m];                                  // the seed (inverse) or the point (forward, tt_rt_sqr.m:129)
    int i0 = 0, i2 = n - 1;                                    // :135-143
    if (a.forward) {
      while (i2 - i0 > 1) {                                    // tt_rt_sqr.m:132-138: the cell is found on the grid
        const int i1 = (i0 + i2) >> 1;
        if (qk > a.x[i1]) i0 = i1; else i2 = i1;
      }
    } else {
      while (i2 - i0 > 1) {
        const int i1 = (i0 + i2) >> 1;
        const double c1 = __ddiv_rn(zero ? a.hc[i1] : cd[(size_t)i1 * a.ldp], cmax);
        if (qk > c1) i0 = i1; else i2 = i1;
      }
    }
    tail_finish(a, m, i0, qk, zero ? a.hc[i0] : cd[(size_t)i0 * a.ldp], zero ? a.h[i0] : p[(size_t)i0 * a.ldp],
                zero ? a.h[i0 + 1] : p[(size_t)(i0 + 1) * a.ldp], cmax, a.hist ? sh_hist : nullptr);
  }
  if (a.hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < n - 1; i += blockDim.x)
      if (sh_hist[i]) atomicAdd(a.hist + i, sh_hist[i]);
  }
}

// The same tail for one row whose conditional sits in shared memory (pr[0..n)), run by one lane inside sqr_pdf_kernel:
// no scratch in global memory.  The CDF is never stored: pass 1 yields the mass; when every increment is non-negative the
// CDF is monotone and the reference's bisection (:135-143) returns the number of interior nodes below q, which pass 2
// counts against q * mass with a 4-ulp guard band on either side (no division); if the two counts differ, or an increment
// was negative, the bisection itself runs with the prefix sums recomputed in the same order.  Bit-identical to
// sqr_tail_kernel on the same conditional.
__device__ TailOut fused_tail_row(const TailArgs &a, int m, const double *pr, const double *sx, const double *sh, const double *shc,
                                  int *sh_hist) {
  const int n = a.n;
  double pprev = pr[0];
  const double c0 = __dmul_rn(__dmul_rn(0.5, pprev), sh[0]);
  double c = c0;
  bool mono = true;
  for (int j = 1; j < n; j++) {
    const double pj = pr[j];
    const double tj = cdf_step(pprev, pj, sh[j]);
    mono = mono && !(tj < 0.0);
    c = __dadd_rn(c, tj);
    pprev = pj;
  }
  double cmax = c;
  const bool zero = cmax <= 0.0;
  if (zero) cmax = shc[n - 1];
  auto prefix = [&](int i) -> double {       // C_i, the rounding sequence of pass 1
    double pp = pr[0], cc = c0;
    for (int j = 1; j <= i; j++) { const double pj = pr[j]; cc = __dadd_rn(cc, cdf_step(pp, pj, sh[j])); pp = pj; }
    return cc;
  };
  const double qk = a.q[m];
  int i0 = 0, i2 = n - 1;
  double c1raw;
  if (a.forward) {
    while (i2 - i0 > 1) { const int i1 = (i0 + i2) >> 1; if (qk > sx[i1]) i0 = i1; else i2 = i1; }
    c1raw = zero ? shc[i0] : prefix(i0);
  } else if (zero) {
    while (i2 - i0 > 1) { const int i1 = (i0 + i2) >> 1; if (qk > __ddiv_rn(shc[i1], cmax)) i0 = i1; else i2 = i1; }
    c1raw = shc[i0];
  } else {
    bool done = false;
    if (mono && c0 == 0.0) {
      const double qc = __dmul_rn(qk, cmax);
      const double lo = __dmul_rn(qc, 1.0 - 8.8817841970012523e-16), hi = __dmul_rn(qc, 1.0 + 8.8817841970012523e-16);
      int L = 0, U = 0;
      double cc = c0, clast = c0, pp = pr[0];
      for (int j = 1; j < n - 1; j++) {
        const double pj = pr[j];
        cc = __dadd_rn(cc, cdf_step(pp, pj, sh[j]));
        pp = pj;
        if (cc < lo) { L++; clast = cc; }
        if (cc < hi) U++;
      }
      if (L == U && qc > 0.0) { i0 = L; c1raw = clast; done = true; }
    }
    if (!done) {
      while (i2 - i0 > 1) { const int i1 = (i0 + i2) >> 1; if (qk > __ddiv_rn(prefix(i1), cmax)) i0 = i1; else i2 = i1; }
      c1raw = prefix(i0);
    }
  }
  return tail_finish(a, m, i0, qk, c1raw, zero ? sh[i0] : pr[i0], zero ? sh[i0 + 1] : pr[i0 + 1], cmax, sh_hist);
}

struct PdfArgs {
  const double *F; int ldf; int rows;
  const double *gp; int pb; int ksteps; int r0; int n;
  double *pdf; int64_t ldp;
  // fused tail (FUSE): the warp parks its pdf tile in shared memory (row pitch pp doubles, warp w at doubles
  // park_off + w * park_stride from the start of dynamic shared memory) and one lane per row finishes the dimension
  TailArgs tail; int pp; int park_off; int park_stride;
  // fused interface update (small cores only: the whole core k sits in shared memory at core_off, [node][a][l], l fastest):
  // after the tail every row multiplies its own two slabs (:197-207), no sort, no further kernel in this dimension
  int upd; const double *core; double *Fout; int r1; int core_off;
};

__host__ __device__ inline size_t pdf_smem_bytes(int rows_cta, int ldf, int pb, int stages) {
  return sizeof(double) * ((size_t)rows_cta * ldf + (size_t)stages * PDF_KS * 4 * pb) + 2 * stages * sizeof(uint64_t);
}
// with the fused tail: + grid tables and histogram, + the parked tiles unless they alias the interface rows
__host__ inline size_t pdf_fused_tables_bytes(int n) { return sizeof(double) * 3 * (size_t)n + sizeof(int) * (((size_t)n + 1) & ~(size_t)1); }

// Conditional pdf of one dimension for a chunk: pdf[j, m] = sum over pairs f_m[a] f_m[b] Gp[(a,b), j].
//   * persistent CTAs over tiles of 8 * 8 PDF_MT samples; eight MMA warps of 8 PDF_MT samples each and one producer warp;
//   * B = Gp (up to 1.3 MB, L2-resident) streams through a ring of PDF_STAGES 16-k-step slices, each one TMA bulk copy
//     signalled on an mbarrier, released by the eight warps on a second mbarrier;
//   * A is never stored: lane (g, t) multiplies f_g[a] (32-byte loads, four k-steps at a time) with f_g[4c + t] (one LDS per
//     column block);
//   * pitches: ldf = 4 (mod 8) and pb = 4 (mod 8) doubles make the fragment loads bank-conflict free without a swizzle.
//   * TAIL1 (n = 8 NT + 1, the usual 2^p + 1 grid): the lone last grid column is a DFMA dot product on the A values the lanes
//     already hold (one broadcast LDS and two DFMA per k-step) instead of a DMMA column tile that is 7/8 padding.
//   * FUSE: the tail of the dimension (:113-195) runs here too: the warp parks its finished pdf tile in shared memory and one
//     lane per row walks it (fused_tail_row), so the conditional never goes to global memory and sqr_tail_kernel is not launched.
//   * MINB = 2 (with two ring stages and 16 rows per warp): two CTAs per SM for the small-core class, whose warps walk
//     contraction, tail and update one after the other and need neighbours to overlap them with.
template <int NT, bool TAIL1, int PDF_MT, bool FUSE, int PDF_STAGES, int MINB>
__global__ void __launch_bounds__(32 * (PDF_WARPS + 1), MINB) sqr_pdf_kernel(PdfArgs a) {
  constexpr int PDF_WROWS = 8 * PDF_MT;               // samples per warp
  constexpr int PDF_ROWS = PDF_WARPS * PDF_WROWS;     // samples per CTA tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *Fs = reinterpret_cast<double *>(smem_raw);
  double *Bs = Fs + (size_t)PDF_ROWS * a.ldf;
  const int SB = PDF_KS * 4 * a.pb;           // doubles per stage
  uint64_t *full = reinterpret_cast<uint64_t *>(Bs + (size_t)PDF_STAGES * SB);
  uint64_t *empty = full + PDF_STAGES;
  // FUSE: grid, intervals and their running sum (3 n doubles), the interval histogram (n ints), then the parked tiles
  double *sx = reinterpret_cast<double *>(empty + PDF_STAGES), *sh = sx + a.n, *shc = sh + a.n;
  int *sh_hist = reinterpret_cast<int *>(shc + a.n);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (FUSE) {
    for (int i = tid; i < a.n; i += blockDim.x) { sx[i] = a.tail.x[i]; sh[i] = a.tail.h[i]; shc[i] = a.tail.hc[i]; sh_hist[i] = 0; }
    if (a.upd) {
      // [node][a][l] with a padded to a multiple of 4 (zero rows) and l to a multiple of 4 (zero columns)
      double *Cs = reinterpret_cast<double *>(smem_raw) + a.core_off;
      const int r1p = (a.r1 + 3) & ~3, r0p = (a.r0 + 3) & ~3;
      for (int e = tid; e < r0p * a.n * r1p; e += blockDim.x) {
        const int aa = e % r0p, j = (e / r0p) % a.n, l = e / (r0p * a.n);
        Cs[((size_t)j * r0p + aa) * r1p + l] = (l < a.r1 && aa < a.r0) ? a.core[aa + (size_t)a.r0 * (j + (size_t)a.n * l)] : 0.0;
      }
    }
  }
  const int ntiles = (a.rows + PDF_ROWS - 1) / PDF_ROWS;
  const int nslices = (a.ksteps + PDF_KS - 1) / PDF_KS;
  if (tid == 0) {
    for (int s = 0; s < PDF_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, PDF_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == PDF_WARPS) {                    // producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int s = 0; s < nslices; s++, it++) {
          const uint32_t stage = it % PDF_STAGES;
          if (it >= PDF_STAGES) mbar_wait(empty + stage, ((it / PDF_STAGES) - 1) & 1);
          const int steps = (a.ksteps - s * PDF_KS) < PDF_KS ? (a.ksteps - s * PDF_KS) : PDF_KS;
          const uint32_t bytes = (uint32_t)(steps * 4 * a.pb * sizeof(double));
          mbar_expect_tx(full + stage, bytes);
          bulk_g2s(Bs + (size_t)stage * SB, a.gp + (size_t)s * SB, bytes, full + stage);
        }
    }
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const int ldf = a.ldf;
  double *fw = Fs + (size_t)warp * PDF_WROWS * ldf;        // this warp's staged rows
  const int nchunk = (a.r0 + 3) >> 2;
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int m0 = tile * PDF_ROWS + warp * PDF_WROWS;
    __syncwarp();
    {
      // the warp's 16 rows are one contiguous block (same pitch in global and shared memory): all loads in flight before
      // the first store, 16 bytes per lane and instruction; rows past the end of the chunk are zero
      const int nvec = (PDF_WROWS * ldf) >> 6;                     // double2 per lane: WROWS ldf / (2 * 32)
      const int live = a.rows - m0 < PDF_WROWS ? (a.rows - m0 > 0 ? a.rows - m0 : 0) : PDF_WROWS;
      const int valid = live * (ldf >> 1);                         // double2 holding real rows
      const double2 *src = reinterpret_cast<const double2 *>(a.F + (size_t)m0 * ldf);
      double2 *dst = reinterpret_cast<double2 *>(fw);
      for (int base = 0; base < nvec; base += 17) {
        double2 v[17];
#pragma unroll
        for (int j = 0; j < 17; j++) {
          const int e = lane + 32 * (base + j);
          v[j] = make_double2(0.0, 0.0);
          if (base + j < nvec && e < valid) v[j] = src[e];
        }
#pragma unroll
        for (int j = 0; j < 17; j++)
          if (base + j < nvec) dst[lane + 32 * (base + j)] = v[j];
      }
    }
    __syncwarp();
    double acc[PDF_MT][NT][2];
#pragma unroll
    for (int mt = 0; mt < PDF_MT; mt++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    double tl[PDF_MT];
#pragma unroll
    for (int mt = 0; mt < PDF_MT; mt++) tl[mt] = 0.0;
    int ks = 0;
    uint32_t stage = it % PDF_STAGES, parity = (it / PDF_STAGES) & 1;
    // one k-step: A values av (the products f[a] f[b] of the lane's pair) against the B fragments of the staged slice
    auto kstep = [&](const double (&av)[PDF_MT], const double *bp) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double b = bp[nt * 8];
#pragma unroll
        for (int mt = 0; mt < PDF_MT; mt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], av[mt], b);
      }
      if (TAIL1) {
        const double bl = bp[8 * NT - g];      // row 4 kk + t, column 8 NT
#pragma unroll
        for (int mt = 0; mt < PDF_MT; mt++) tl[mt] = fma(av[mt], bl, tl[mt]);
      }
    };
    auto release = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      it++;
      stage = it % PDF_STAGES; parity = (it / PDF_STAGES) & 1;
    };
    // off-diagonal part: pairs (a, 4c + t), a < 4c, in groups of four k-steps (4c is a multiple of four, slices hold
    // sixteen): one barrier check and two 32-byte interface loads per group, the four steps unrolled so that the loads of
    // one run under the MMAs of the previous
    for (int c = 1; c < nchunk; c++) {
      double fb[PDF_MT];
#pragma unroll
      for (int mt = 0; mt < PDF_MT; mt++) fb[mt] = fw[(mt * 8 + g) * ldf + 4 * c + t];
      for (int aa = 0; aa < 4 * c; aa += 4) {
        const int kk = ks & (PDF_KS - 1);
        if (kk == 0) mbar_wait(full + stage, parity);
        double f4[PDF_MT][4];
#pragma unroll
        for (int mt = 0; mt < PDF_MT; mt++) {
          const double2 lo = *reinterpret_cast<const double2 *>(fw + (mt * 8 + g) * ldf + aa);
          const double2 hi = *reinterpret_cast<const double2 *>(fw + (mt * 8 + g) * ldf + aa + 2);
          f4[mt][0] = lo.x; f4[mt][1] = lo.y; f4[mt][2] = hi.x; f4[mt][3] = hi.y;
        }
        const double *bp0 = Bs + (size_t)stage * SB + (kk * 4 + t) * a.pb + g;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          double av[PDF_MT];
#pragma unroll
          for (int mt = 0; mt < PDF_MT; mt++) av[mt] = __dmul_rn(f4[mt][u], fb[mt]);
          kstep(av, bp0 + u * 4 * a.pb);
        }
        ks += 4;
        if ((ks & (PDF_KS - 1)) == 0 || ks == a.ksteps) release();
      }
    }
    // diagonal part: lane t of k-step ds holds pair 4 ds + t of the blocks' upper triangles (both factors by its own index)
    for (int ds = 0; ds < ((10 * nchunk + 3) >> 2); ds++) {
      const int kk = ks & (PDF_KS - 1);
      if (kk == 0) mbar_wait(full + stage, parity);
      int pa, pb_;
      sqr_diag_pair(4 * ds + t, pa, pb_);
      if (pb_ >= 4 * nchunk) { pa = 0; pb_ = 0; }          // padding pairs of the last k-step (their k-rows are zero)
      double av[PDF_MT];
#pragma unroll
      for (int mt = 0; mt < PDF_MT; mt++) av[mt] = __dmul_rn(fw[(mt * 8 + g) * ldf + pa], fw[(mt * 8 + g) * ldf + pb_]);
      kstep(av, Bs + (size_t)stage * SB + (kk * 4 + t) * a.pb + g);
      ks++;
      if ((ks & (PDF_KS - 1)) == 0 || ks == a.ksteps) release();
    }
    if (FUSE) {
      // park the tile [row][node] (odd pitch: the row-per-lane reads of the tail are conflict free), then one lane per row
      double *park = reinterpret_cast<double *>(smem_raw) + a.park_off + (size_t)warp * a.park_stride;
      __syncwarp();                          // (when the park aliases this warp's interface rows: every lane is done with them)
#pragma unroll
      for (int mt = 0; mt < PDF_MT; mt++) {
        double *pr = park + (mt * 8 + g) * a.pp;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const int col = nt * 8 + 2 * t;
          if (col < a.n) pr[col] = acc[mt][nt][0];
          if (col + 1 < a.n) pr[col + 1] = acc[mt][nt][1];
        }
        if (TAIL1) {
          double v = tl[mt];
          v += __shfl_xor_sync(FULL, v, 1);
          v += __shfl_xor_sync(FULL, v, 2);
          if (t == 0) pr[8 * NT] = v;
        }
      }
      __syncwarp();
      TailOut mine;
      mine.i0 = 0; mine.wa = 0.0; mine.wb = 0.0;
      for (int row = lane; row < PDF_WROWS; row += 32)
        if (m0 + row < a.rows) mine = fused_tail_row(a.tail, m0 + row, park + row * a.pp, sx, sh, shc, a.tail.hist ? sh_hist : nullptr);
      __syncwarp();
      if (a.upd) {
        // (host side guarantees one row per lane, a park of its own -- the interface rows are intact -- and r_{k+1} <= 32)
        const double *Cs = reinterpret_cast<const double *>(smem_raw) + a.core_off;
        const int r1p = (a.r1 + 3) & ~3, r0p = (a.r0 + 3) & ~3;
        const int LPR = r1p <= 16 ? 16 : 32;          // lanes per row: two rows at a time when the rank allows
        const int l = lane & (LPR - 1), sub = lane / LPR;
        const size_t slab = (size_t)r0p * r1p;
        for (int rb = 0; rb < PDF_WROWS; rb += 32 / LPR) {
          const int rr = rb + sub;
          const int j0 = __shfl_sync(FULL, mine.i0, rr);
          const double wa = __shfl_sync(FULL, mine.wa, rr), wb = __shfl_sync(FULL, mine.wb, rr);
          if (l < r1p && m0 + rr < a.rows) {
            const double *f = fw + rr * ldf;          // (32-byte aligned rows; columns [r0, r0p) are zero)
            const double *s0 = Cs + (size_t)j0 * slab + l, *s1 = s0 + slab;
            double u0a = 0.0, u0b = 0.0, u1a = 0.0, u1b = 0.0;
            for (int aa = 0; aa < r0p; aa += 4) {
              const double2 fl = *reinterpret_cast<const double2 *>(f + aa), fh = *reinterpret_cast<const double2 *>(f + aa + 2);
              u0a = fma(fl.x, s0[0], u0a);       u1a = fma(fl.x, s1[0], u1a);
              u0b = fma(fl.y, s0[r1p], u0b);     u1b = fma(fl.y, s1[r1p], u1b);
              u0a = fma(fh.x, s0[2 * r1p], u0a); u1a = fma(fh.x, s1[2 * r1p], u1a);
              u0b = fma(fh.y, s0[3 * r1p], u0b); u1b = fma(fh.y, s1[3 * r1p], u1b);
              s0 += 4 * r1p; s1 += 4 * r1p;
            }
            const double u0 = __dadd_rn(u0a, u0b), u1 = __dadd_rn(u1a, u1b);
            a.Fout[(size_t)(m0 + rr) * ldf + l] = l < a.r1 ? __dadd_rn(__dmul_rn(u0, wa), __dmul_rn(u1, wb)) : 0.0;   // :205
          }
        }
        __syncwarp();
      }
    } else {
#pragma unroll
      for (int mt = 0; mt < PDF_MT; mt++) {
        const int row = m0 + mt * 8 + g;
        if (row < a.rows) {
#pragma unroll
          for (int nt = 0; nt < NT; nt++) {
            const int col = nt * 8 + 2 * t;
            if (col < a.n) a.pdf[(size_t)col * a.ldp + row] = acc[mt][nt][0];
            if (col + 1 < a.n) a.pdf[(size_t)(col + 1) * a.ldp + row] = acc[mt][nt][1];
          }
        }
      }
      if (TAIL1) {
#pragma unroll
        for (int mt = 0; mt < PDF_MT; mt++) {
          double v = tl[mt];
          v += __shfl_xor_sync(FULL, v, 1);
          v += __shfl_xor_sync(FULL, v, 2);
          const int row = m0 + mt * 8 + g;
          if (t == 0 && row < a.rows) a.pdf[(size_t)(8 * NT) * a.ldp + row] = v;
        }
      }
    }
  }
  if (FUSE && a.tail.hist) {                 // the eight MMA warps (the producer warp has left): flush the CTA's histogram
    asm volatile("bar.sync 1, %0;" ::"n"(32 * PDF_WARPS) : "memory");
    for (int i = tid; i < a.n - 1; i += 32 * PDF_WARPS)
      if (sh_hist[i]) atomicAdd(a.tail.hist + i, sh_hist[i]);
  }
}

constexpr int UPD_WARPS = 8;
constexpr int UPD_TS = 16 * UPD_WARPS;   // samples per tile of the interface update (two 8-row MMA tiles per warp)
constexpr int UPD_THREADS = 32 * UPD_WARPS;

__global__ void sqr_bin_scan_kernel(const int *__restrict__ hist, int nb, int *bin_start, int *bin_tile_start, int *cursor) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int s = 0, t = 0;
  for (int b = 0; b < nb; b++) {
    bin_start[b] = s; bin_tile_start[b] = t; cursor[b] = s;
    const int c = hist[b];
    s += c; t += (c + UPD_TS - 1) / UPD_TS;
  }
  bin_start[nb] = s; bin_tile_start[nb] = t;
}

// counting-sort scatter, one atomic per (warp, distinct interval)
__global__ void sqr_bin_scatter_kernel(const int *__restrict__ idx, int rows, int *cursor, int *perm) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = m < rows;
  const int b = live ? idx[m] : -1;
  const unsigned peers = __match_any_sync(FULL, b);
  const int leader = __ffs(peers) - 1;
  const int rank = __popc(peers & ((1u << lane) - 1));
  int base = 0;
  if (live && lane == leader) base = atomicAdd(cursor + b, __popc(peers));
  base = __shfl_sync(FULL, base, leader);
  if (live) perm[base + rank] = m;
}

struct UpdArgs {
  const double *core; int r0, n, r1;
  const double *Fin; double *Fout; int ldf;
  const int *perm, *bin_start, *bin_tile_start; int nb;
  const double *w1, *w2;
};

// :197-207  fkm1' = (fkm1 core(:, i0, :)) Aq + (fkm1 core(:, i0 + 1, :)) Bq for 128-sample tiles of one interval, as two
// FP64 tensor-core products per tile.  The two core slabs of an interval are staged once per interval and CTA (CTAs own
// contiguous tile ranges) in the layout they have in global memory ([l][a], a fastest), the gathered interface rows as
// [sample][a]; both pitches are 4 (mod 8) doubles, which makes every fragment load bank-conflict free without a swizzle.
// NTU = 8-column output tiles (r_{k+1} <= 8 NTU).
template <int NTU>
__global__ void __launch_bounds__(UPD_THREADS) sqr_update_kernel(UpdArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int r0p = (a.r0 + 3) & ~3, r1p = (a.r1 + 3) & ~3;
  const int P = ((a.r0 + 7) & ~7) + 4;    // pitch of slab rows and of staged interface rows
  double *S0 = sm, *S1 = S0 + (size_t)8 * NTU * P, *Fs = S1 + (size_t)8 * NTU * P;   // Fs: two buffers of UPD_TS rows
  double *Wsm = Fs + (size_t)2 * UPD_TS * P;                                          // [2][UPD_TS][2] interpolation weights
  int *Ids = reinterpret_cast<int *>(Wsm + 4 * UPD_TS);                               // [2][UPD_TS] sample ids
  int *Bts = Ids + 2 * UPD_TS, *Bst = Bts + (a.nb + 1);                               // the two bin tables
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i <= a.nb; i += UPD_THREADS) { Bts[i] = a.bin_tile_start[i]; Bst[i] = a.bin_start[i]; }
  __syncthreads();
  const int total = Bts[a.nb];
  const int t_begin = (int)(((int64_t)total * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((int64_t)total * (blockIdx.x + 1)) / gridDim.x);
  // Each warp gathers and multiplies its own 16 samples: no CTA-wide barrier per tile.  Everything a tile needs from global
  // memory arrives asynchronously while earlier tiles are multiplied: the sample ids two tiles ahead (a register), the rows
  // and the interpolation weights one tile ahead (cp.async: a whole row per instruction, 8 bytes per weight).
  // Rows of the interface buffers are zero in the columns [r, r rounded up to 4), so whole 4-blocks are copied.
  const int chunks = r0p >> 1;            // 16-byte pieces per row
  const int mine = warp * 16 + (lane & 15);
  int cur_p = 0;
  auto load_id = [&](int tile) -> int {
    while (tile >= Bts[cur_p + 1]) cur_p++;
    const int start = Bst[cur_p] + (tile - Bts[cur_p]) * UPD_TS;
    const int cnt = min(UPD_TS, Bst[cur_p + 1] - start);
    return mine < cnt ? a.perm[start + mine] : -1;
  };
  auto issue = [&](int buf, int myid) {
    double *fw = Fs + ((size_t)buf * UPD_TS + warp * 16) * P;
#pragma unroll 4
    for (int row = 0; row < 16; row++) {
      const int id = __shfl_sync(FULL, myid, row);
      if (id >= 0)
        for (int ch = lane; ch < chunks; ch += 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(fw + row * P + 2 * ch)), "l"(a.Fin + (size_t)id * a.ldf + 2 * ch) : "memory");
    }
    if (lane < 16) {
      Ids[buf * UPD_TS + mine] = myid;
      if (myid >= 0) {
        double *w = Wsm + ((size_t)buf * UPD_TS + mine) * 2;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(w)), "l"(a.w1 + myid) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(w + 1)), "l"(a.w2 + myid) : "memory");
      }
    }
  };
  int buf = 0;
  if (t_begin < t_end) issue(0, load_id(t_begin));
  asm volatile("cp.async.commit_group;" ::: "memory");
  int id_after = t_begin + 1 < t_end ? load_id(t_begin + 1) : -1;
  int bin = 0, staged = -1;
  for (int tile = t_begin; tile < t_end; tile++, buf ^= 1) {
    while (tile >= Bts[bin + 1]) bin++;
    __syncwarp();                         // this warp's reads of the other buffer (previous tile) are done
    const bool restage = bin != staged;
    if (restage) {
      __syncthreads();                    // every warp is done with the previous interval's slabs
      // the two slabs of the interval, 8-byte cp.async per element (core columns are r0 contiguous doubles, any parity)
      const int64_t cs = (int64_t)a.r0 * a.n;
      for (int e = tid; e < 8 * NTU * r0p; e += UPD_THREADS) {
        const int aa = e % r0p, l = e / r0p;
        if (l < a.r1 && aa < a.r0) {
          const double *src = a.core + aa + (int64_t)a.r0 * bin + cs * l;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(S0 + l * P + aa)), "l"(src) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(S1 + l * P + aa)), "l"(src + a.r0) : "memory");
        } else {
          S0[l * P + aa] = 0.0; S1[l * P + aa] = 0.0;
        }
      }
      staged = bin;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tile + 1 < t_end) issue(buf ^ 1, id_after);
    asm volatile("cp.async.commit_group;" ::: "memory");
    id_after = tile + 2 < t_end ? load_id(tile + 2) : -1;    // in flight while this tile is multiplied
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // all but the newest group (the next tile's rows) have landed
    if (restage) __syncthreads(); else __syncwarp();
    const double *fw = Fs + ((size_t)buf * UPD_TS + warp * 16) * P;
    double acc0[2][NTU][2], acc1[2][NTU][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < NTU; nt++) { acc0[mt][nt][0] = acc0[mt][nt][1] = 0.0; acc1[mt][nt][0] = acc1[mt][nt][1] = 0.0; }
    for (int ks = 0; ks < (r0p >> 2); ks++) {
      const double a0 = fw[g * P + 4 * ks + t], a1 = fw[(8 + g) * P + 4 * ks + t];
      const double *b0 = S0 + g * P + 4 * ks + t, *b1 = S1 + g * P + 4 * ks + t;
#pragma unroll
      for (int nt = 0; nt < NTU; nt++) {
        const double u0 = b0[nt * 8 * P], u1 = b1[nt * 8 * P];
        dmma884(acc0[0][nt][0], acc0[0][nt][1], a0, u0);
        dmma884(acc0[1][nt][0], acc0[1][nt][1], a1, u0);
        dmma884(acc1[0][nt][0], acc1[0][nt][1], a0, u1);
        dmma884(acc1[1][nt][0], acc1[1][nt][1], a1, u1);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      const int sidx = warp * 16 + mt * 8 + g;
      const int sid = Ids[buf * UPD_TS + sidx];
      if (sid >= 0) {
        const size_t id = (size_t)sid;
        const double wa = Wsm[((size_t)buf * UPD_TS + sidx) * 2], wb = Wsm[((size_t)buf * UPD_TS + sidx) * 2 + 1];
        double *dst = a.Fout + id * a.ldf + 2 * t;
#pragma unroll
        for (int nt = 0; nt < NTU; nt++) {
          const double o0 = __dadd_rn(__dmul_rn(acc0[mt][nt][0], wa), __dmul_rn(acc1[mt][nt][0], wb));   // :205
          const double o1 = __dadd_rn(__dmul_rn(acc0[mt][nt][1], wa), __dmul_rn(acc1[mt][nt][1], wb));
          if (8 * nt + 2 * t < r1p) *reinterpret_cast<double2 *>(dst + 8 * nt) = make_double2(o0, o1);   // zeros up to the next multiple of 4
        }
      }
    }
  }
}

template <int NTU>
static cudaError_t upd_launch(const UpdArgs &a, int64_t max_tiles, int sm_count, cudaStream_t st) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(sqr_update_kernel<NTU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  const int P = ((a.r0 + 7) & ~7) + 4;
  const size_t sm = sizeof(double) * ((size_t)2 * 8 * NTU * P + (size_t)2 * UPD_TS * P + 4 * UPD_TS) + sizeof(int) * (2 * UPD_TS + 2 * (a.nb + 1));
  // one wave of CTAs, each walking a contiguous range of tiles (slabs are restaged only when the interval changes):
  // as many CTAs per SM as the shared memory allows, at most four
  int per_sm = (int)((size_t)(227 * 1024) / (sm + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  const int grid = (int)std::min<int64_t>(max_tiles, (int64_t)per_sm * sm_count);
  sqr_update_kernel<NTU><<<grid, UPD_THREADS, sm, st>>>(a);
  return cudaGetLastError();
}

// tt_dirt_sample.m:36,60  truncated normal -> uniform:  z = erf(z / sqrt(2)) * cdf_factor + 0.5
__global__ void dirt_tn2u_kernel(int64_t M, int d, double cdf_factor, const double *__restrict__ in, int64_t ldi, double *out, int64_t ldo) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M || k >= d) return;
  out[m + ldo * k] = __dadd_rn(__dmul_rn(erf(__ddiv_rn(in[m + ldi * k], 1.4142135623730951)), cdf_factor), 0.5);
}

// tt_dirt_sample.m:51-56  lFapp = lFapp + dlFapp;  with a normal reference  lFapp = lFapp + sum(z.^2, 2) / 2 - log(2 cdf_factor^2 / pi) d / 2
__global__ void dirt_accumulate_kernel(int64_t M, int d, int first, int normal, double logc_half_d, const double *__restrict__ dlf,
                                       const double *__restrict__ z, int64_t ldz, double *lf) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double v = first ? dlf[m] : __dadd_rn(lf[m], dlf[m]);
  if (normal) {
    double s = 0.0;
    for (int k = 0; k < d; k++) { const double t = z[m + ldz * k]; s = __dadd_rn(s, __dmul_rn(t, t)); }
    v = __dsub_rn(__dadd_rn(v, __ddiv_rn(s, 2.0)), logc_half_d);
  }
  lf[m] = v;
}

// tt_dirt_inverse.m:42,55  uniform -> truncated normal:  q = erfinv((q - 0.5) / cdf_factor) * sqrt(2)   (in place)
__global__ void dirt_u2tn_kernel(int64_t M, int d, double cdf_factor, double *q, int64_t ldq) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M || k >= d) return;
  q[m + ldq * k] = __dmul_rn(erfinv(__ddiv_rn(__dsub_rn(q[m + ldq * k], 0.5), cdf_factor)), 1.4142135623730951);
}

// tt_dirt_inverse.m:44,50,57  lFapp = lFapp [+ sum(q.^2, 2) / 2 of the level's input] + dlFapp
__global__ void dirt_inverse_accumulate_kernel(int64_t M, int d, int first, int normal, const double *__restrict__ dlf,
                                               const double *__restrict__ qin, int64_t ldq, double *lf) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double v = first ? 0.0 : lf[m];
  if (normal) {
    double s = 0.0;
    for (int k = 0; k < d; k++) { const double t = qin[m + ldq * k]; s = __dadd_rn(s, __dmul_rn(t, t)); }
    v = __dadd_rn(v, __ddiv_rn(s, 2.0));
  }
  lf[m] = __dadd_rn(v, dlf[m]);
}

}  // namespace

namespace ttirt {
// called by ttirt_cache_clear() (ttirt_engine.cu)
void sqr_pool_clear() {
  int keep = 0;
  cudaGetDevice(&keep);
  for (int g = 0; g < 64; g++) {
    DevPool &pl = g_pool[g];
    std::lock_guard<std::mutex> lock(pl.mu);
    if (pl.idle.empty()) continue;
    cudaSetDevice(g);
    for (auto &kv : pl.idle) cudaFree(kv.second);
    pl.idle.clear();
    pl.idle_bytes = 0;
  }
  cudaSetDevice(keep);
}
}  // namespace ttirt

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct ttirt_sqr_model {
  int device = 0, sm_count = 148;
  int64_t d = 0;
  std::vector<SqrDim> dims;
  int rmax = 1, nmax = 2, ldf = 12, pb = 28, nt = 3, nbpad = 8;
  int64_t sum_x = 0, sum_cin = 0, sum_c = 0, sum_g = 0, sum_r = 0;
  bool extended = false;
  double *d_xs = nullptr, *d_h = nullptr, *d_hc = nullptr, *d_core = nullptr, *d_gp = nullptr, *d_rfac = nullptr;
  // workspace of one chunk
  int64_t cap = 0;
  bool host = false;
  double *F0 = nullptr, *F1 = nullptr, *pdf = nullptr, *cdf = nullptr, *w1 = nullptr, *w2 = nullptr;
  int *idx = nullptr, *perm = nullptr, *hist = nullptr, *bin_start = nullptr, *bin_tile_start = nullptr, *cursor = nullptr;
  // host-buffer mode: two staging slots so that the copies of one chunk overlap the kernels of its neighbours
  double *q[2] = {nullptr, nullptr}, *z[2] = {nullptr, nullptr}, *lf[2] = {nullptr, nullptr};
  int32_t *idx_out[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_flops = 0.0;
  // scratch of the DIRT layer loop (held by the level-0 model): two M x d seed / sample buffers, one length-M log-density
  double *dirt_z[2] = {nullptr, nullptr}, *dirt_lf = nullptr;
  int64_t dirt_cap = 0;
};

static void sqr_ws_free(ttirt_sqr_model *md) {
  sqr_free(md->F0); sqr_free(md->F1); sqr_free(md->pdf); sqr_free(md->cdf); sqr_free(md->w1); sqr_free(md->w2);
  sqr_free(md->idx); sqr_free(md->perm); sqr_free(md->hist); sqr_free(md->bin_start); sqr_free(md->bin_tile_start); sqr_free(md->cursor);
  for (int s = 0; s < 2; s++) {
    sqr_free(md->q[s]); sqr_free(md->z[s]); sqr_free(md->lf[s]); sqr_free(md->idx_out[s]);
    md->q[s] = md->z[s] = md->lf[s] = nullptr; md->idx_out[s] = nullptr;
  }
  md->F0 = md->F1 = md->pdf = md->cdf = md->w1 = md->w2 = nullptr;
  md->idx = md->perm = md->hist = md->bin_start = md->bin_tile_start = md->cursor = nullptr;
  md->cap = 0; md->host = false;
}

static int sqr_ws_alloc(ttirt_sqr_model *md, int64_t cap, bool h) {
  const int64_t d = md->d;
  const size_t fbytes = sizeof(double) * (size_t)cap * md->ldf;
  CKS(sqr_alloc(&md->F0, fbytes));
  CKS(sqr_alloc(&md->F1, fbytes));
  CKS(cudaMemset(md->F0, 0, fbytes));
  CKS(cudaMemset(md->F1, 0, fbytes));
  CKS(sqr_alloc(&md->pdf, sizeof(double) * (size_t)cap * md->nmax));
  CKS(sqr_alloc(&md->cdf, sizeof(double) * (size_t)cap * md->nmax));
  CKS(sqr_alloc(&md->w1, sizeof(double) * cap));
  CKS(sqr_alloc(&md->w2, sizeof(double) * cap));
  CKS(sqr_alloc(&md->idx, sizeof(int) * cap));
  CKS(sqr_alloc(&md->perm, sizeof(int) * cap));
  CKS(sqr_alloc(&md->hist, sizeof(int) * d * md->nbpad));
  CKS(sqr_alloc(&md->bin_start, sizeof(int) * (md->nbpad + 1)));
  CKS(sqr_alloc(&md->bin_tile_start, sizeof(int) * (md->nbpad + 1)));
  CKS(sqr_alloc(&md->cursor, sizeof(int) * (md->nbpad + 1)));
  if (h) {
    for (int s = 0; s < 2; s++) {
      CKS(sqr_alloc(&md->q[s], sizeof(double) * cap * d));
      CKS(sqr_alloc(&md->z[s], sizeof(double) * cap * d));
      CKS(sqr_alloc(&md->lf[s], sizeof(double) * cap));
      CKS(sqr_alloc(&md->idx_out[s], sizeof(int32_t) * cap * d));
    }
  }
  return 0;
}

// Capacity and mode are recorded only after every allocation has succeeded; a failed allocation leaves the workspace
// empty (cap 0), so a later smaller call on a caller-held model allocates afresh instead of running on null scratch.
// Blocks go back to the per-device pool here: work of an earlier device-API call may still be using them on the
// caller's stream, so the device is synchronised first (growing the workspace is rare).
static int sqr_ws_ensure(ttirt_sqr_model *md, int64_t rows, bool host) {
  if (md->cap >= rows && (md->host || !host)) return 0;
  const int64_t cap = std::max(rows, md->cap);
  const bool h = host || md->host;
  if (md->cap > 0) cudaDeviceSynchronize();
  sqr_ws_free(md);
  if (sqr_ws_alloc(md, cap, h) != 0) {
    sqr_ws_free(md);
    cudaGetLastError();
    return -1;
  }
  md->cap = cap; md->host = h;
  return 0;
}

extern "C" void ttirt_sqr_model_destroy(ttirt_sqr_model *md) {
  if (!md) return;
  cudaSetDevice(md->device);
  cudaDeviceSynchronize();
  sqr_ws_free(md);
  sqr_free(md->dirt_z[0]); sqr_free(md->dirt_z[1]); sqr_free(md->dirt_lf);
  sqr_free(md->d_xs); sqr_free(md->d_h); sqr_free(md->d_hc); sqr_free(md->d_core); sqr_free(md->d_gp); sqr_free(md->d_rfac);
  if (md->stream) cudaStreamDestroy(md->stream);
  if (md->copy_stream) cudaStreamDestroy(md->copy_stream);
  for (int s = 0; s < 2; s++) {
    if (md->ev_in[s]) cudaEventDestroy(md->ev_in[s]);
    if (md->ev_done[s]) cudaEventDestroy(md->ev_done[s]);
    if (md->ev_out[s]) cudaEventDestroy(md->ev_out[s]);
  }
  for (auto &p : md->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  delete md;
}

// TTIRT_SQR_FUSED=0 keeps the tail in its own kernel (read at every launch: the tests compare the two paths bit for bit)
static bool sqr_fuse_enabled() {
  const char *e = getenv("TTIRT_SQR_FUSED");
  return e ? atoi(e) != 0 : true;
}

template <int NT, bool TAIL1, int MT, bool FUSE, int STG = 3, int MINB = 1>
static cudaError_t pdf_launch_impl(const PdfArgs &a, size_t bytes, int sm_count, cudaStream_t st) {
  constexpr int PDF_ROWS = PDF_WARPS * 8 * MT;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(sqr_pdf_kernel<NT, TAIL1, MT, FUSE, STG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024 / MINB);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  const int ntiles = (a.rows + PDF_ROWS - 1) / PDF_ROWS;
  const int grid = ntiles < MINB * sm_count ? ntiles : MINB * sm_count;
  sqr_pdf_kernel<NT, TAIL1, MT, FUSE, STG, MINB><<<grid, 32 * (PDF_WARPS + 1), bytes, st>>>(a);
  return cudaGetLastError();
}

// Launches the conditional-pdf kernel; *fused tells the caller whether the tail ran inside it (else sqr_tail_kernel follows),
// *fused_upd whether the interface update did too (the caller asks for it by filling a.core / a.Fout / a.r1 and a.upd = 1).
template <int NT, bool TAIL1, int MT>
static cudaError_t pdf_launch(PdfArgs a, int sm_count, cudaStream_t st, bool *fused, bool *fused_upd) {
  constexpr int PDF_ROWS = PDF_WARPS * 8 * MT, WROWS = 8 * MT;
  const size_t base = pdf_smem_bytes(PDF_ROWS, a.ldf, a.pb, 3);
  *fused = false;
  // Measured on B200: in the two lighter classes (32 rows per warp: every lane walks a row) the fused tail beats the
  // separate kernel by 8 % (r = 32, n = 33) to 14 % (r = 16, n = 17) of the step; in the r <= 64 / n <= 72 class (16 rows
  // per warp, half the lanes idle, its FP64 instructions contending with a saturated DMMA pipe) it loses 2.5 %: not fused
  // there unless TTIRT_SQR_FUSED=2 asks for it.
  const char *fe = getenv("TTIRT_SQR_FUSED");
  const bool want = sqr_fuse_enabled() && (MT >= 4 || (fe && atoi(fe) >= 2));
  const bool want_upd = a.upd != 0;
  a.upd = 0;
  *fused_upd = false;
  if (want) {
    const int pp = (a.n + 1) | 1;                                   // odd pitch >= n + 1
    const size_t tables = pdf_fused_tables_bytes(a.n);
    const size_t park = sizeof(double) * (size_t)PDF_ROWS * pp;
    if (base + tables + park <= (size_t)227 * 1024) {               // a park of its own behind the tables
      a.pp = pp; a.park_off = (int)((base + tables) / sizeof(double)); a.park_stride = WROWS * pp;
      *fused = true;
      const int r1p = (a.r1 + 3) & ~3;
      const size_t corebytes = sizeof(double) * (size_t)((a.r0 + 3) & ~3) * a.n * r1p;
      if (want_upd && WROWS == 32 && r1p <= 32 && base + tables + park + corebytes <= (size_t)227 * 1024 && !(fe && atoi(fe) == 1)) {
        *fused_upd = true;
        a.upd = 1;
        // two CTAs per SM when half-size tiles (16 rows per warp) and a two-stage ring fit twice (and TTIRT_SQR_LIGHT != 0)
        const char *le = getenv("TTIRT_SQR_LIGHT");
        const size_t lbase = pdf_smem_bytes(PDF_WARPS * 16, a.ldf, a.pb, 2), lpark = sizeof(double) * (size_t)PDF_WARPS * 16 * pp;
        if constexpr (NT <= 3) {
          if (lbase + tables + lpark + corebytes <= (size_t)113 * 1024 && !(le && atoi(le) == 0)) {
            a.park_off = (int)((lbase + tables) / sizeof(double)); a.park_stride = 16 * pp;
            a.core_off = (int)((lbase + tables + lpark) / sizeof(double));
            return pdf_launch_impl<NT, TAIL1, 2, true, 2, 2>(a, lbase + tables + lpark + corebytes, sm_count, st);
          }
        }
        a.core_off = (int)((base + tables + park) / sizeof(double));
        return pdf_launch_impl<NT, TAIL1, MT, true>(a, base + tables + park + corebytes, sm_count, st);
      }
      return pdf_launch_impl<NT, TAIL1, MT, true>(a, base + tables + park, sm_count, st);
    }
    if (pp <= a.ldf && base + tables <= (size_t)227 * 1024) {       // alias the warp's interface rows (dead after the k loop)
      a.pp = pp; a.park_off = 0; a.park_stride = WROWS * a.ldf;
      *fused = true;
      return pdf_launch_impl<NT, TAIL1, MT, true>(a, base + tables, sm_count, st);
    }
  }
  return pdf_launch_impl<NT, TAIL1, MT, false>(a, base, sm_count, st);
}

static int sqr_model_build(ttirt_sqr_model *md, const int64_t *n, int64_t nxs, const double *xs, const int64_t *rk, const double *core) {
  const int64_t d = md->d;
  if (rk[0] != 1 || rk[d] != 1) return aux_fail("tt_irt_sqr: ttrank[0] and ttrank[d] must be 1");
  int64_t sn = 0;
  for (int64_t k = 0; k < d; k++) {
    if (n[k] < 1 || rk[k] < 1 || rk[k + 1] < 1) return aux_fail("tt_irt_sqr: non-positive mode size or rank at dimension %lld", (long long)k);
    sn += n[k];
  }
  // tt_irt_sqr.m:33-39
  if (nxs == sn + 2 * d) md->extended = true;
  else if (nxs != sn)
    return aux_fail("tt_irt_sqr: number of grid points (with or without boundaries) in xsf should be sum of mode sizes in f (got %lld, modes sum to %lld)",
                    (long long)nxs, (long long)sn);
  md->dims.resize(d);
  int64_t ox = 0, oci = 0, oc = 0, og = 0, orr = 0;
  for (int64_t k = 0; k < d; k++) {
    SqrDim &di = md->dims[k];
    di.n_in = (int)n[k];
    di.n = (int)n[k] + (md->extended ? 2 : 0);
    if (di.n < 2) return aux_fail("tt_irt_sqr: dimension %lld needs at least 2 grid points", (long long)k);
    if (md->extended && di.n_in < 2) return aux_fail("tt_irt_sqr: boundary extrapolation needs at least 2 interior points (dimension %lld)", (long long)k);
    di.r0 = (int)rk[k]; di.r1 = (int)rk[k + 1];
    di.ksteps = sqr_ksteps(di.r0);
    di.pad = 0;
    di.off_x = ox; di.off_cin = oci; di.off_c = oc; di.off_g = og; di.off_r = orr;
    ox += di.n; oci += (int64_t)di.r0 * di.n_in * di.r1; oc += (int64_t)di.r0 * di.n * di.r1;
    orr += (int64_t)di.r0 * di.r0;
    md->rmax = std::max(md->rmax, std::max(di.r0, di.r1));
    md->nmax = std::max(md->nmax, di.n);
  }
  if (md->rmax > 64 || md->nmax > 72)
    return aux_fail("tt_irt_sqr: shape outside the B200 path (ranks <= 64 and grid sizes <= 72 are supported, got r = %d, n = %d)", md->rmax, md->nmax);
  md->nt = md->nmax <= 24 ? 3 : (md->nmax <= 40 ? 5 : 9);
  md->pb = 8 * md->nt + 4;
  md->ldf = ((md->rmax + 7) & ~7) + 4;
  md->nbpad = (md->nmax + 7) & ~7;
  for (int64_t k = 0; k < d; k++) {
    md->dims[k].off_g = og;
    og += (int64_t)md->dims[k].ksteps * 4 * md->pb;
  }
  // ranks of the factors: s1 of the last core is 1 (:43-44), every QR yields min(rows, columns) rows (:70)
  int s = 1;
  for (int64_t k = d - 1; k >= 0; k--) {
    SqrDim &di = md->dims[k];
    di.s1 = s;
    const int64_t m = (int64_t)di.n * s;
    di.s0 = (int)std::min<int64_t>(m, di.r0);
    s = di.s0;
  }
  md->sum_x = ox; md->sum_cin = oci; md->sum_c = oc; md->sum_g = og; md->sum_r = orr;

  cudaDeviceProp prop;
  CKS(cudaGetDeviceProperties(&prop, md->device));
  if (prop.major < 10) return aux_fail("tt_irt_sqr: device %d (%s, sm_%d%d) is not a Blackwell B200-class GPU", md->device, prop.name, prop.major, prop.minor);
  md->sm_count = prop.multiProcessorCount;
  CKS(cudaStreamCreateWithFlags(&md->stream, cudaStreamNonBlocking));
  CKS(cudaStreamCreateWithFlags(&md->copy_stream, cudaStreamNonBlocking));
  for (int s = 0; s < 2; s++) {
    CKS(cudaEventCreateWithFlags(&md->ev_in[s], cudaEventDisableTiming));
    CKS(cudaEventCreateWithFlags(&md->ev_done[s], cudaEventDisableTiming));
    CKS(cudaEventCreateWithFlags(&md->ev_out[s], cudaEventDisableTiming));
  }

  CKS(sqr_alloc(&md->d_xs, sizeof(double) * ox));
  CKS(sqr_alloc(&md->d_h, sizeof(double) * ox));
  CKS(sqr_alloc(&md->d_hc, sizeof(double) * ox));
  CKS(sqr_alloc(&md->d_core, sizeof(double) * oc));
  CKS(sqr_alloc(&md->d_gp, sizeof(double) * og));
  CKS(sqr_alloc(&md->d_rfac, sizeof(double) * orr));
  CKS(cudaMemset(md->d_gp, 0, sizeof(double) * og));
  CKS(cudaMemset(md->d_rfac, 0, sizeof(double) * orr));
  CKS(cudaMemcpy(md->d_xs, xs, sizeof(double) * ox, cudaMemcpyHostToDevice));

  double *d_cin = nullptr, *d_pm = nullptr, *d_a = nullptr, *d_one = nullptr, *d_stack[2] = {nullptr, nullptr};
  int64_t mmax = 0;
  for (int64_t k = 0; k < d; k++) mmax = std::max(mmax, (int64_t)md->dims[k].n * md->dims[k].s1 * md->dims[k].r0);
  auto cleanup = [&]() { if (md->extended) sqr_free(d_cin); sqr_free(d_pm); sqr_free(d_a); sqr_free(d_one); sqr_free(d_stack[0]); sqr_free(d_stack[1]); };
#define CKB(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) { cleanup(); return aux_fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } \
  } while (0)
  CKB(sqr_alloc(&d_pm, sizeof(double) * mmax));
  CKB(sqr_alloc(&d_a, sizeof(double) * mmax));
  CKB(sqr_alloc(&d_one, sizeof(double)));
  CKB(sqr_alloc(&d_stack[0], sizeof(double) * std::max<int64_t>(mmax, 1)));
  CKB(sqr_alloc(&d_stack[1], sizeof(double) * std::max<int64_t>(mmax, 1)));
  CKB(cudaFuncSetAttribute(sqr_qr_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
  const double one = 1.0;
  CKB(cudaMemcpy(d_one, &one, sizeof(double), cudaMemcpyHostToDevice));
  if (md->extended) {
    CKB(sqr_alloc(&d_cin, sizeof(double) * oci));
    CKB(cudaMemcpy(d_cin, core, sizeof(double) * oci, cudaMemcpyHostToDevice));
  } else {
    CKB(cudaMemcpy(md->d_core, core, sizeof(double) * oc, cudaMemcpyHostToDevice));
  }
  for (int64_t k = d - 1; k >= 0; k--) {
    const SqrDim &di = md->dims[k];
    const double *x = md->d_xs + di.off_x;
    sqr_grid_kernel<<<1, 32>>>(x, md->d_h + di.off_x, md->d_hc + di.off_x, di.n);
    LAUNCHED();
    if (md->extended) {
      sqr_extend_kernel<<<(di.r0 * di.r1 + 127) / 128, 128>>>(d_cin + di.off_cin, x, md->d_core + di.off_c, di.r0, di.n, di.r1);
      LAUNCHED();
    }
    const double *Rp = (k == d - 1) ? d_one : md->d_rfac + md->dims[k + 1].off_r;
    const int64_t m = (int64_t)di.n * di.s1;
    const int64_t tot = m * di.r0;
    sqr_contract_kernel<<<(unsigned)((tot + 127) / 128), 128>>>(md->d_core + di.off_c, Rp, d_pm, di.r0, di.n, di.r1, di.s1);
    LAUNCHED();
    if (k > 0) {
      sqr_weight_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(d_pm, md->d_h + di.off_x, d_a, di.n, (int)m, di.r0);
      LAUNCHED();
      // TSQR levels: row blocks that fit shared memory, stacked R factors ping-pong between d_a's tail and d_pm's
      {
        const int r = di.r0;
        int rb = (int)std::min<int64_t>(m, (int64_t)(200 * 1024) / (8 * r));
        if (rb < 2 * r && m > rb) { cleanup(); return aux_fail("tt_irt_sqr: rank %d too large for the shared-memory QR", r); }
        const double *src = d_a;
        int msrc = (int)m, lda = (int)m, level = 0;
        for (;;) {
          const int nblk = (msrc + rb - 1) / rb;
          const size_t smb = sizeof(double) * (size_t)std::min(rb, msrc) * r;
          if (nblk == 1) {
            sqr_qr_block_kernel<<<1, 1024, smb>>>(src, msrc, lda, r, rb, md->d_rfac + di.off_r, 0, 1, di.s0);
            LAUNCHED();
            break;
          }
          double *dst = d_stack[level & 1];
          sqr_qr_block_kernel<<<nblk, 1024, smb>>>(src, msrc, lda, r, rb, dst, nblk * r, 0, 0);
          LAUNCHED();
          src = dst; msrc = nblk * r; lda = nblk * r; level++;
        }
      }
    }
    const int64_t ge = (int64_t)di.ksteps * 4 * di.n;
    sqr_gram_pack_kernel<<<(unsigned)((ge + 127) / 128), 128>>>(d_pm, di.n, di.s1, di.r0, md->d_gp + di.off_g, md->pb, di.ksteps);
    LAUNCHED();
  }
  CKB(cudaGetLastError());
  CKB(cudaDeviceSynchronize());
#undef CKB
  cleanup();
  return 0;
}

extern "C" ttirt_sqr_model *ttirt_sqr_model_create(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank,
                                                   const double *ttcore, int device) {
  if (d < 1 || !n || !xs || !ttrank || !ttcore) { aux_fail("bad arguments to ttirt_sqr_model_create"); return nullptr; }
  const int cnt = ttirt_device_count();
  if (cnt <= 0) { aux_fail("no CUDA device available (this library has no CPU fallback)"); return nullptr; }
  if (device < 0 || device >= cnt) { aux_fail("device %d out of range (%d visible)", device, cnt); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { aux_fail("cudaSetDevice(%d) failed", device); return nullptr; }
  ttirt_sqr_model *md = new ttirt_sqr_model();
  md->device = device; md->d = d;
  if (sqr_model_build(md, n, nxs, xs, ttrank, ttcore) != 0) { ttirt_sqr_model_destroy(md); return nullptr; }
  return md;
}

extern "C" int64_t ttirt_sqr_model_mode_size(const ttirt_sqr_model *md, int64_t k) {
  if (!md || k < 0 || k >= md->d) return -1;
  return md->dims[k].n;
}

extern "C" int ttirt_sqr_model_get_sweep(const ttirt_sqr_model *md, int64_t k, double *gram_out, double *rr_out) {
  if (!md) return aux_fail("null model");
  if (k < 0 || k >= md->d) return aux_fail("dimension out of range");
  CKS(cudaSetDevice(md->device));
  const SqrDim &di = md->dims[k];
  if (gram_out) {
    std::vector<double> gp((size_t)di.ksteps * 4 * md->pb);
    CKS(cudaMemcpy(gp.data(), md->d_gp + di.off_g, sizeof(double) * gp.size(), cudaMemcpyDeviceToHost));
    // unpack: k-row kr holds w G[a, b, :] for its pair; fill both (a, b) and (b, a)
    for (int kr = 0; kr < 4 * di.ksteps; kr++) {
      int a, b;
      double w;
      sqr_krow_pair(kr, di.r0, a, b, w);
      if (w == 0.0) continue;
      for (int j = 0; j < di.n; j++) {
        const double v = gp[(size_t)kr * md->pb + j] / w;
        gram_out[(size_t)a + (size_t)di.r0 * b + (size_t)di.r0 * di.r0 * j] = v;
        gram_out[(size_t)b + (size_t)di.r0 * a + (size_t)di.r0 * di.r0 * j] = v;
      }
    }
  }
  if (rr_out) {
    if (k == 0) return aux_fail("no factor to the left of the first core");
    std::vector<double> rt((size_t)di.r0 * di.s0);
    CKS(cudaMemcpy(rt.data(), md->d_rfac + di.off_r, sizeof(double) * rt.size(), cudaMemcpyDeviceToHost));
    for (int a = 0; a < di.r0; a++)
      for (int b = 0; b < di.r0; b++) {
        double s = 0.0;
        for (int t = 0; t < di.s0; t++) s += rt[(size_t)a + (size_t)di.r0 * t] * rt[(size_t)b + (size_t)di.r0 * t];
        rr_out[(size_t)a + (size_t)di.r0 * b] = s;
      }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one chunk, device-resident buffers, enqueued on st
// ------------------------------------------------------------------------------------------------
static int sqr_enqueue_chunk(ttirt_sqr_model *md, int64_t rows, int64_t D, const double *q, int64_t ldq, double *z, int64_t ldz,
                             double *lf, int32_t *idx_out, cudaStream_t st, int forward = 0) {
  if (rows <= 0) return 0;
  const int d = (int)md->d;
  CKS(cudaMemsetAsync(md->hist, 0, sizeof(int) * d * md->nbpad, st));
  sqr_init_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(md->F0, md->ldf, (int)rows);
  LAUNCHED();
  double *Fin = md->F0, *Fout = md->F1;
  for (int k = 0; k < (int)D; k++) {
    const SqrDim &di = md->dims[k];
    const bool update = (k + 1 < (int)D);
    PdfArgs pa;
    pa.F = Fin; pa.ldf = md->ldf; pa.rows = (int)rows; pa.gp = md->d_gp + di.off_g; pa.pb = md->pb; pa.ksteps = di.ksteps;
    pa.r0 = di.r0; pa.n = di.n; pa.pdf = md->pdf; pa.ldp = md->cap;
    TailArgs ta;
    ta.pdf = md->pdf; ta.cdf = md->cdf; ta.ldp = md->cap; ta.rows = (int)rows; ta.n = di.n;
    ta.x = md->d_xs + di.off_x; ta.h = md->d_h + di.off_x; ta.hc = md->d_hc + di.off_x;
    ta.q = q + ldq * k; ta.z = z + ldz * k; ta.idx_out = idx_out ? idx_out + ldz * k : nullptr;
    ta.idx = md->idx; ta.w1 = md->w1; ta.w2 = md->w2; ta.lf = lf; ta.first = (k == 0); ta.forward = forward;
    ta.hist = update ? md->hist + (size_t)k * md->nbpad : nullptr;
    pa.tail = ta; pa.pp = 0; pa.park_off = 0; pa.park_stride = 0;
    // ask for the fused interface update too; the launcher grants it when the whole core fits next to the tiles
    pa.upd = update ? 1 : 0; pa.core = md->d_core + di.off_c; pa.Fout = Fout; pa.r1 = di.r1; pa.core_off = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (md->profile) {
      if (md->prof_used == md->prof_events.size()) {
        cudaEvent_t x, y;
        CKS(cudaEventCreate(&x)); CKS(cudaEventCreate(&y));
        md->prof_events.emplace_back(x, y);
      }
      e0 = md->prof_events[md->prof_used].first; e1 = md->prof_events[md->prof_used].second;
      md->prof_used++;
      md->prof_flops += (double)rows * ((double)di.r0 * (di.r0 + 1) * di.n + 0.5 * di.r0 * (di.r0 + 1));
      CKS(cudaEventRecord(e0, st));
    }
    // the dimension's own grid size picks the variant inside the model's class (pb is the class's pitch)
    cudaError_t pe;
    bool fused = false, fused_upd = false;
    const bool t1 = (di.n == 8 * (md->nt - 1) + 1);
    if (md->nt == 3) pe = t1 ? pdf_launch<2, true, 4>(pa, md->sm_count, st, &fused, &fused_upd) : pdf_launch<3, false, 4>(pa, md->sm_count, st, &fused, &fused_upd);
    else if (md->nt == 5) pe = t1 ? pdf_launch<4, true, 4>(pa, md->sm_count, st, &fused, &fused_upd) : pdf_launch<5, false, 4>(pa, md->sm_count, st, &fused, &fused_upd);
    else pe = t1 ? pdf_launch<8, true, 2>(pa, md->sm_count, st, &fused, &fused_upd) : pdf_launch<9, false, 2>(pa, md->sm_count, st, &fused, &fused_upd);
    CKS(pe);
    LAUNCHED();
    if (e1) CKS(cudaEventRecord(e1, st));
    if (!fused) {
      sqr_tail_kernel<<<(unsigned)((rows + 255) / 256), 256, sizeof(int) * di.n, st>>>(ta);
      LAUNCHED();
      CKS(cudaGetLastError());
    }
    if (update && fused_upd) {
      std::swap(Fin, Fout);
    } else if (update) {
      const int nb = di.n - 1;
      sqr_bin_scan_kernel<<<1, 32, 0, st>>>(md->hist + (size_t)k * md->nbpad, nb, md->bin_start, md->bin_tile_start, md->cursor);
      LAUNCHED();
      sqr_bin_scatter_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(md->idx, (int)rows, md->cursor, md->perm);
      LAUNCHED();
      UpdArgs ua;
      ua.core = md->d_core + di.off_c; ua.r0 = di.r0; ua.n = di.n; ua.r1 = di.r1;
      ua.Fin = Fin; ua.Fout = Fout; ua.ldf = md->ldf;
      ua.perm = md->perm; ua.bin_start = md->bin_start; ua.bin_tile_start = md->bin_tile_start; ua.nb = nb;
      ua.w1 = md->w1; ua.w2 = md->w2;
      const int64_t max_tiles = (rows + UPD_TS - 1) / UPD_TS + nb;
      cudaError_t ue;
      if (di.r1 <= 8) ue = upd_launch<1>(ua, max_tiles, md->sm_count, st);
      else if (di.r1 <= 16) ue = upd_launch<2>(ua, max_tiles, md->sm_count, st);
      else if (di.r1 <= 32) ue = upd_launch<4>(ua, max_tiles, md->sm_count, st);
      else ue = upd_launch<8>(ua, max_tiles, md->sm_count, st);
      CKS(ue);
      LAUNCHED();
      std::swap(Fin, Fout);
    }
  }
  return 0;
}

// samples per chunk: TTIRT_SQR_CHUNK, else 2^19 in the r <= 64 / n <= 72 class (measured 3.28 / 3.36 / 3.42 M samples/s at
// 2^17 / 2^18 / 2^19: fewer partial waves of the persistent kernels) and 2^18 in the lighter classes (finer copy / compute overlap)
static int64_t sqr_chunk(const ttirt_sqr_model *md) {
  const char *e = getenv("TTIRT_SQR_CHUNK");
  if (e && atoll(e) > 0) return atoll(e);
  return (int64_t)1 << (md->nt == 9 ? 19 : 18);
}

static int sqr_transform_device(ttirt_sqr_model *md, int64_t M, int64_t D, const double *d_q, int64_t ldq, double *d_z,
                                int64_t ldz, double *d_lf, int32_t *d_idx, void *stream, int forward) {
  if (!md) return aux_fail("null model");
  if (M < 0 || ldq < M || ldz < M) return aux_fail("bad M / leading dimensions");
  if (D < 1 || D > md->d) return aux_fail("tt_irt_sqr: q must have between 1 and d columns (got %lld, d = %lld)", (long long)D, (long long)md->d);
  if (M == 0) return 0;
  CKS(cudaSetDevice(md->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t chunk = std::min<int64_t>(M, sqr_chunk(md));
  if (md->cap < chunk) {
    CKS(cudaStreamSynchronize(st));
    if (sqr_ws_ensure(md, chunk, false) != 0) return -1;
  }
  for (int64_t m0 = 0; m0 < M; m0 += chunk) {
    const int64_t rows = std::min(chunk, M - m0);
    if (sqr_enqueue_chunk(md, rows, D, d_q + m0, ldq, d_z + m0, ldz, d_lf + m0, d_idx ? d_idx + m0 : nullptr, st, forward) != 0) return -1;
  }
  return 0;
}

extern "C" int ttirt_sqr_sample_device(ttirt_sqr_model *md, int64_t M, int64_t D, const double *d_q, int64_t ldq, double *d_z,
                                       int64_t ldz, double *d_lf, int32_t *d_idx, void *stream) {
  return sqr_transform_device(md, M, D, d_q, ldq, d_z, ldz, d_lf, d_idx, stream, 0);
}

extern "C" int ttirt_sqr_forward_device(ttirt_sqr_model *md, int64_t M, int64_t D, const double *d_x, int64_t ldx, double *d_q,
                                        int64_t ldq, double *d_lf, int32_t *d_idx, void *stream) {
  return sqr_transform_device(md, M, D, d_x, ldx, d_q, ldq, d_lf, d_idx, stream, 1);
}

static int sqr_transform_host(ttirt_sqr_model *md, int64_t M, int64_t D, const double *h_q, double *h_z, double *h_lf,
                              int32_t *h_idx, int64_t ld, int forward) {
  if (!md) return aux_fail("null model");
  if (M < 0 || ld < M) return aux_fail("bad M / leading dimension");
  if (D < 1 || D > md->d) return aux_fail("tt_irt_sqr: q must have between 1 and d columns (got %lld, d = %lld)", (long long)D, (long long)md->d);
  if (M == 0) return 0;
  if (!h_q || !h_z || !h_lf) return aux_fail("null host buffer");
  CKS(cudaSetDevice(md->device));
  const int64_t chunk = std::min<int64_t>(M, sqr_chunk(md));
  if (sqr_ws_ensure(md, chunk, true) != 0) return -1;
  cudaStream_t st = md->stream, cs = md->copy_stream;
  const int64_t cap = md->cap;
  // chunk i: copy in (copy stream) -> kernels (compute stream) -> copy out (copy stream, issued after chunk i + 1 has been
  // enqueued, so that a host-blocking copy of pageable memory never leaves the compute stream empty)
  auto copy_out = [&](int64_t j) -> int {
    const int s = (int)(j & 1);
    const int64_t m0 = j * chunk, rows = std::min(chunk, M - m0);
    CKS(cudaStreamWaitEvent(cs, md->ev_done[s], 0));
    CKS(cudaMemcpy2DAsync(h_z + m0, sizeof(double) * ld, md->z[s], sizeof(double) * cap, sizeof(double) * rows, (size_t)D, cudaMemcpyDeviceToHost, cs));
    CKS(cudaMemcpyAsync(h_lf + m0, md->lf[s], sizeof(double) * rows, cudaMemcpyDeviceToHost, cs));
    if (h_idx)
      CKS(cudaMemcpy2DAsync(h_idx + m0, sizeof(int32_t) * ld, md->idx_out[s], sizeof(int32_t) * cap, sizeof(int32_t) * rows, (size_t)D, cudaMemcpyDeviceToHost, cs));
    CKS(cudaEventRecord(md->ev_out[s], cs));
    return 0;
  };
  const int64_t nchunks = (M + chunk - 1) / chunk;
  for (int64_t i = 0; i < nchunks; i++) {
    const int s = (int)(i & 1);
    const int64_t m0 = i * chunk, rows = std::min(chunk, M - m0);
    CKS(cudaMemcpy2DAsync(md->q[s], sizeof(double) * cap, h_q + m0, sizeof(double) * ld, sizeof(double) * rows, (size_t)D, cudaMemcpyHostToDevice, cs));
    CKS(cudaEventRecord(md->ev_in[s], cs));
    CKS(cudaStreamWaitEvent(st, md->ev_in[s], 0));
    if (i >= 2) CKS(cudaStreamWaitEvent(st, md->ev_out[s], 0));
    if (sqr_enqueue_chunk(md, rows, D, md->q[s], cap, md->z[s], cap, md->lf[s], h_idx ? md->idx_out[s] : nullptr, st, forward) != 0) return -1;
    CKS(cudaEventRecord(md->ev_done[s], st));
    if (i >= 1 && copy_out(i - 1) != 0) return -1;
  }
  if (copy_out(nchunks - 1) != 0) return -1;
  CKS(cudaStreamSynchronize(cs));
  CKS(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int ttirt_sqr_sample_host(ttirt_sqr_model *md, int64_t M, int64_t D, const double *h_q, double *h_z, double *h_lf,
                                     int32_t *h_idx, int64_t ld) {
  return sqr_transform_host(md, M, D, h_q, h_z, h_lf, h_idx, ld, 0);
}

extern "C" int ttirt_sqr_forward_host(ttirt_sqr_model *md, int64_t M, int64_t D, const double *h_x, double *h_q, double *h_lf,
                                      int32_t *h_idx, int64_t ld) {
  return sqr_transform_host(md, M, D, h_x, h_q, h_lf, h_idx, ld, 1);
}

// Whole call on host buffers.  Samples are independent (tt_irt_sqr.m:94-208 walks them in blocks), so the M rows are cut
// into contiguous ranges, one per device; every device gets its own copy of the cores and redoes the (small) sweep; no
// collective.  One host thread per device, joined before returning.
static int sqr_run_rows(int device, int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank,
                        const double *ttcore, int64_t rows, int64_t D, const double *h_q, double *h_z, double *h_lf, int64_t ld, int forward) {
  const bool trace = getenv("TTIRT_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  ttirt_sqr_model *md = ttirt_sqr_model_create(d, n, nxs, xs, ttrank, ttcore, device);
  if (!md) return -1;
  const auto t1 = std::chrono::steady_clock::now();
  const int rc = sqr_transform_host(md, rows, D, h_q, h_z, h_lf, nullptr, ld, forward);
  const auto t2 = std::chrono::steady_clock::now();
  ttirt_sqr_model_destroy(md);
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "tt_irt_sqr[b200]: device %d rows %lld: model %.2f ms, sample %.2f ms, release %.2f ms\n", device, (long long)rows,
            ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
  return rc;
}

static int sqr_run_host(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank, const double *ttcore,
                        int64_t M, int64_t D, const double *h_q, double *h_z, double *h_lf, int first_device, int n_devices, int forward) {
  if (M <= 0) return M == 0 ? 0 : aux_fail("negative M");
  const int cnt = ttirt_device_count();
  if (cnt <= 0) return aux_fail("no CUDA device available (this library has no CPU fallback)");
  if (n_devices < 1) n_devices = 1;
  if (first_device < 0 || first_device + n_devices > cnt) return aux_fail("devices %d..%d requested, %d visible", first_device, first_device + n_devices - 1, cnt);
  if (n_devices == 1) return sqr_run_rows(first_device, d, n, nxs, xs, ttrank, ttcore, M, D, h_q, h_z, h_lf, M, forward);
  std::vector<std::thread> th;
  std::vector<int> rc((size_t)n_devices, 0);
  for (int g = 0; g < n_devices; g++) {
    const int64_t m0 = M * g / n_devices, m1 = M * (g + 1) / n_devices;
    th.emplace_back([=, &rc]() {
      rc[(size_t)g] = m1 > m0 ? sqr_run_rows(first_device + g, d, n, nxs, xs, ttrank, ttcore, m1 - m0, D, h_q + m0, h_z + m0, h_lf + m0, M, forward) : 0;
    });
  }
  for (auto &t : th) t.join();
  for (int g = 0; g < n_devices; g++) if (rc[(size_t)g] != 0) return -1;
  return 0;
}

extern "C" int ttirt_sqr_run_host(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank, const double *ttcore,
                                  int64_t M, int64_t D, const double *h_q, double *h_z, double *h_lf, int first_device, int n_devices) {
  return sqr_run_host(d, n, nxs, xs, ttrank, ttcore, M, D, h_q, h_z, h_lf, first_device, n_devices, 0);
}

extern "C" int ttirt_sqr_run_forward_host(int64_t d, const int64_t *n, int64_t nxs, const double *xs, const int64_t *ttrank,
                                          const double *ttcore, int64_t M, int64_t D, const double *h_x, double *h_q, double *h_lf,
                                          int first_device, int n_devices) {
  return sqr_run_host(d, n, nxs, xs, ttrank, ttcore, M, D, h_x, h_q, h_lf, first_device, n_devices, 1);
}

extern "C" void ttirt_sqr_profile_enable(ttirt_sqr_model *md, int on) {
  if (!md) return;
  md->profile = on != 0;
  md->prof_used = 0; md->prof_flops = 0.0;
}

extern "C" int ttirt_sqr_profile_read(ttirt_sqr_model *md, double *ms_total, int64_t *launches, double *flops_total) {
  if (!md) return aux_fail("null model");
  CKS(cudaSetDevice(md->device));
  CKS(cudaDeviceSynchronize());
  double ms = 0.0;
  for (size_t i = 0; i < md->prof_used; i++) {
    float t = 0.f;
    CKS(cudaEventElapsedTime(&t, md->prof_events[i].first, md->prof_events[i].second));
    ms += t;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = (int64_t)md->prof_used;
  if (flops_total) *flops_total = md->prof_flops;
  return 0;
}


// ------------------------------------------------------------------------------------------------
// DIRT sampler loop (reference matlab/samplers/tt_dirt_sample.m:17-73, spline branch): the caller of tt_irt_sqr
// ------------------------------------------------------------------------------------------------
extern "C" int ttirt_dirt_sample_device(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *d_q,
                                        int64_t ldq, double *d_z, int64_t ldz, double *d_lf, void *stream) {
  if (nlevels < 1 || !models) return aux_fail("ttirt_dirt_sample: need at least the level-0 model");
  for (int64_t j = 0; j < nlevels; j++) {
    if (!models[j]) return aux_fail("ttirt_dirt_sample: null model at level %lld", (long long)j);
    if (models[j]->d != models[0]->d || models[j]->device != models[0]->device)
      return aux_fail("ttirt_dirt_sample: all levels must have the same dimension and live on the same device");
  }
  if (M < 0 || ldq < M || ldz < M) return aux_fail("bad M / leading dimensions");
  if (M == 0) return 0;
  ttirt_sqr_model *m0 = models[0];
  const int d = (int)m0->d;
  CKS(cudaSetDevice(m0->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (m0->dirt_cap < M) {
    CKS(cudaStreamSynchronize(st));
    sqr_free(m0->dirt_z[0]); sqr_free(m0->dirt_z[1]); sqr_free(m0->dirt_lf);
    m0->dirt_z[0] = m0->dirt_z[1] = m0->dirt_lf = nullptr; m0->dirt_cap = 0;
    CKS(sqr_alloc(&m0->dirt_z[0], sizeof(double) * M * d));
    CKS(sqr_alloc(&m0->dirt_z[1], sizeof(double) * M * d));
    CKS(sqr_alloc(&m0->dirt_lf, sizeof(double) * M));
    m0->dirt_cap = M;
  }
  const bool normal = sigma > 0.0;
  const double cdf_factor = normal ? 0.5 / erf(sigma / sqrt(2.0)) : 0.0;                     // :30
  const double logc_half_d = normal ? log(2.0 * cdf_factor * cdf_factor / 3.14159265358979323846) * d / 2.0 : 0.0;
  const dim3 g2((unsigned)((M + 255) / 256), (unsigned)d);
  const double *cur = d_q;
  int64_t ldc = ldq;
  int pp = 0;
  bool first = true;
  for (int64_t j = nlevels - 1; j >= 0; j--) {                                                // :34 and level 0 (:59-73)
    if (normal) {
      dirt_tn2u_kernel<<<g2, 256, 0, st>>>(M, d, cdf_factor, cur, ldc, m0->dirt_z[pp], M);    // :36, :60
      LAUNCHED();
      cur = m0->dirt_z[pp]; ldc = M; pp ^= 1;
    }
    double *out = j == 0 ? d_z : m0->dirt_z[pp];
    const int64_t ldo = j == 0 ? ldz : M;
    if (ttirt_sqr_sample_device(models[j], M, d, cur, ldc, out, ldo, m0->dirt_lf, nullptr, stream) != 0) return -1;   // :46, :71
    dirt_accumulate_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(M, d, first ? 1 : 0, (normal && j > 0) ? 1 : 0, logc_half_d,
                                                                       m0->dirt_lf, out, ldo, d_lf);      // :51-56, :73
    LAUNCHED();
    first = false;
    if (j > 0) { cur = out; ldc = M; pp ^= 1; }
  }
  CKS(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_dirt_sample_host(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *h_q,
                                      double *h_z, double *h_lf, int64_t ld) {
  if (nlevels < 1 || !models || !models[0]) return aux_fail("ttirt_dirt_sample: need at least the level-0 model");
  if (M < 0 || ld < M) return aux_fail("bad M / leading dimension");
  if (M == 0) return 0;
  if (!h_q || !h_z || !h_lf) return aux_fail("null host buffer");
  ttirt_sqr_model *m0 = models[0];
  const int64_t d = m0->d;
  CKS(cudaSetDevice(m0->device));
  const int64_t chunk = std::min<int64_t>(M, (int64_t)1 << 20);
  double *dq = nullptr, *dz = nullptr, *dl = nullptr;
  if (sqr_alloc(&dq, sizeof(double) * chunk * d) != cudaSuccess || sqr_alloc(&dz, sizeof(double) * chunk * d) != cudaSuccess ||
      sqr_alloc(&dl, sizeof(double) * chunk) != cudaSuccess) {
    cudaGetLastError();
    sqr_free(dq); sqr_free(dz); sqr_free(dl);   // the blocks that did come through go back to the pool
    return aux_fail("out of device memory for %lld x %lld staging blocks", (long long)chunk, (long long)d);
  }
  cudaStream_t st = m0->stream;
  int rc = 0;
  for (int64_t b = 0; b < M && rc == 0; b += chunk) {
    const int64_t rows = std::min(chunk, M - b);
    if (cudaMemcpy2DAsync(dq, sizeof(double) * chunk, h_q + b, sizeof(double) * ld, sizeof(double) * rows, (size_t)d, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = aux_fail("copy in failed"); break; }
    rc = ttirt_dirt_sample_device(nlevels, models, sigma, rows, dq, chunk, dz, chunk, dl, st);
    if (rc != 0) break;
    if (cudaMemcpy2DAsync(h_z + b, sizeof(double) * ld, dz, sizeof(double) * chunk, sizeof(double) * rows, (size_t)d, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_lf + b, dl, sizeof(double) * rows, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = aux_fail("copy out failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaStreamSynchronize(st);
  sqr_free(dq); sqr_free(dz); sqr_free(dl);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// inverse of the DIRT (reference matlab/samplers/tt_dirt_inverse.m:24-59): the levels from 0 upwards with tt_rt_sqr
// ------------------------------------------------------------------------------------------------
extern "C" int ttirt_dirt_inverse_device(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *d_x,
                                         int64_t ldx, double *d_q, int64_t ldq, double *d_lf, void *stream) {
  if (nlevels < 1 || !models) return aux_fail("ttirt_dirt_inverse: need at least the level-0 model");
  for (int64_t j = 0; j < nlevels; j++) {
    if (!models[j]) return aux_fail("ttirt_dirt_inverse: null model at level %lld", (long long)j);
    if (models[j]->d != models[0]->d || models[j]->device != models[0]->device)
      return aux_fail("ttirt_dirt_inverse: all levels must have the same dimension and live on the same device");
  }
  if (M < 0 || ldx < M || ldq < M) return aux_fail("bad M / leading dimensions");
  if (M == 0) return 0;
  ttirt_sqr_model *m0 = models[0];
  const int d = (int)m0->d;
  CKS(cudaSetDevice(m0->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (m0->dirt_cap < M) {
    CKS(cudaStreamSynchronize(st));
    sqr_free(m0->dirt_z[0]); sqr_free(m0->dirt_z[1]); sqr_free(m0->dirt_lf);
    m0->dirt_z[0] = m0->dirt_z[1] = m0->dirt_lf = nullptr; m0->dirt_cap = 0;
    CKS(sqr_alloc(&m0->dirt_z[0], sizeof(double) * M * d));
    CKS(sqr_alloc(&m0->dirt_z[1], sizeof(double) * M * d));
    CKS(sqr_alloc(&m0->dirt_lf, sizeof(double) * M));
    m0->dirt_cap = M;
  }
  const bool normal = sigma > 0.0;
  const double cdf_factor = normal ? 0.5 / erf(sigma / sqrt(2.0)) : 0.0;                      // :33
  const dim3 g2((unsigned)((M + 255) / 256), (unsigned)d);
  const double *cur = d_x;
  int64_t ldc = ldx;
  int pp = 0;
  for (int64_t j = 0; j < nlevels; j++) {                                                      // :39 (level 0), :47-58
    double *out = j == nlevels - 1 ? d_q : m0->dirt_z[pp];
    const int64_t ldo = j == nlevels - 1 ? ldq : M;
    if (ttirt_sqr_forward_device(models[j], M, d, cur, ldc, out, ldo, m0->dirt_lf, nullptr, stream) != 0) return -1;   // :39, :52
    dirt_inverse_accumulate_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(M, d, j == 0 ? 1 : 0, (normal && j > 0) ? 1 : 0,
                                                                               m0->dirt_lf, cur, ldc, d_lf);              // :44, :50, :57
    LAUNCHED();
    if (normal) {
      dirt_u2tn_kernel<<<g2, 256, 0, st>>>(M, d, cdf_factor, out, ldo);                        // :42, :55
      LAUNCHED();
    }
    cur = out; ldc = ldo; pp ^= 1;
  }
  CKS(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_dirt_inverse_host(int64_t nlevels, ttirt_sqr_model *const *models, double sigma, int64_t M, const double *h_x,
                                       double *h_q, double *h_lf, int64_t ld) {
  if (nlevels < 1 || !models || !models[0]) return aux_fail("ttirt_dirt_inverse: need at least the level-0 model");
  if (M < 0 || ld < M) return aux_fail("bad M / leading dimension");
  if (M == 0) return 0;
  if (!h_x || !h_q || !h_lf) return aux_fail("null host buffer");
  ttirt_sqr_model *m0 = models[0];
  const int64_t d = m0->d;
  CKS(cudaSetDevice(m0->device));
  const int64_t chunk = std::min<int64_t>(M, (int64_t)1 << 20);
  double *dx = nullptr, *dq = nullptr, *dl = nullptr;
  if (sqr_alloc(&dx, sizeof(double) * chunk * d) != cudaSuccess || sqr_alloc(&dq, sizeof(double) * chunk * d) != cudaSuccess ||
      sqr_alloc(&dl, sizeof(double) * chunk) != cudaSuccess) {
    cudaGetLastError();
    sqr_free(dx); sqr_free(dq); sqr_free(dl);
    return aux_fail("out of device memory for %lld x %lld staging blocks", (long long)chunk, (long long)d);
  }
  cudaStream_t st = m0->stream;
  int rc = 0;
  for (int64_t b = 0; b < M && rc == 0; b += chunk) {
    const int64_t rows = std::min(chunk, M - b);
    if (cudaMemcpy2DAsync(dx, sizeof(double) * chunk, h_x + b, sizeof(double) * ld, sizeof(double) * rows, (size_t)d, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = aux_fail("copy in failed"); break; }
    rc = ttirt_dirt_inverse_device(nlevels, models, sigma, rows, dx, chunk, dq, chunk, dl, st);
    if (rc != 0) break;
    if (cudaMemcpy2DAsync(h_q + b, sizeof(double) * ld, dq, sizeof(double) * chunk, sizeof(double) * rows, (size_t)d, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_lf + b, dl, sizeof(double) * rows, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = aux_fail("copy out failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  cudaStreamSynchronize(st);
  sqr_free(dx); sqr_free(dq); sqr_free(dl);
  return rc;
}
