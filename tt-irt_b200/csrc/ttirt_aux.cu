// The steps either side of tt_irt1, on the device:
//
//   seeds      rank-1 lattice with random shift        reference matlab/samplers/qmcnodes.m:6-13
//              uniform pseudo-random (Philox4x32-10)   reference rand / np.random.random (test_shock_absorber_tt.py:147)
//              truncated-normal reference map          reference matlab/samplers/randref.m:22-34
//   consumers  importance weights and their statistics reference matlab/samplers/iw_prune.m:19-29
//              N / ESS                                 reference matlab/samplers/essinv.m:12-14
//              Hellinger distance                      reference matlab/samplers/hellinger.m:12-16
//              independence Metropolis-Hastings prune  reference matlab/samplers/mcmc_prune.m:24-43,
//                                                                python/test_shock_absorber_tt.py:165-171
//
// All of it is HBM-bound streaming or a short sequential chain; none of it is reshaped into tensor-core work.
// No CPU fallback: without a device every entry point fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../../include/tt_irt1.h"
#include "../../include/tt_irt_sqr.h"
#include "ttirt_common.cuh"

namespace ttirt {
int aux_fail(const char *fmt, ...);   // defined in ttirt_engine.cu: records the thread's last error
void aux_launched();                  // bumps the library's kernel-launch counter
}  // namespace ttirt
using ttirt::aux_fail;

#define CKA(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return aux_fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

namespace {

// ------------------------------------------------------------------------------------------------
// seeds
// ------------------------------------------------------------------------------------------------
// qmcnodes.m:6-13:  Y = (0:N-1)/N;  Y = z(1:d) * Y;  Y = Y + Delta;  Y = Y - floor(Y).
// m/N is exact for N a power of two and z*m/N is exact below 2^53, so the only rounding is the shift's:
// bit-exact against the reference arithmetic.
__global__ void lattice_kernel(int d, int64_t M, int64_t m0, double inv_n, const double *__restrict__ z,
                               const double *__restrict__ shift, double *q, int64_t ldq) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M || k >= d) return;
  const double y = __dmul_rn((double)(m0 + m), inv_n);
  const double t = __dadd_rn(__dmul_rn(z[k], y), shift[k]);
  q[m + ldq * k] = __dsub_rn(t, floor(t));
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011): counter-based, so any shard of the sample range can be
// generated independently.  Counter = (sample index lo, hi, dimension, 0), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// u = (53 high bits of the first 64 output bits) * 2^-53, in [0, 1)
__global__ void uniform_kernel(int d, int64_t M, int64_t m0, uint64_t seed, double *q, int64_t ldq) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M || k >= d) return;
  const uint64_t idx = (uint64_t)(m0 + m);
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)k, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t bits = ((uint64_t)c[1] << 32) | c[0];
  q[m + ldq * k] = (double)(bits >> 11) * 1.1102230246251565e-16;  // 2^-53
}

// randref.m:31-33:  cdf_ifactor = erf(sigma/sqrt(2))/0.5;  y = erfinv((u-0.5)*cdf_ifactor)*sqrt(2)
__global__ void truncnormal_kernel(int64_t n, double cdf_ifactor, const double *__restrict__ u, double *y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = erfinv((u[i] - 0.5) * cdf_ifactor) * 1.4142135623730951;
}

// ------------------------------------------------------------------------------------------------
// consumers: importance-weight statistics
// ------------------------------------------------------------------------------------------------
// Deterministic two-level reductions: every block reduces a fixed slice in a fixed tree, a last single block
// reduces the block partials in index order.  Results do not depend on scheduling.
constexpr int RB = 256;     // threads per reduction block
constexpr int RG = 1024;    // reduction blocks

__device__ __forceinline__ double block_sum(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = RB / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double *sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = RB / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// pass 1: max of dF = lFex - lFapp  (essinv.m:13, hellinger.m:13)
__global__ void iw_max_kernel(int64_t M, const double *__restrict__ lfex, const double *__restrict__ lfapp, double *part) {
  __shared__ double sh[RB];
  double v = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < M; i += (int64_t)RG * RB) v = fmax(v, lfex[i] - lfapp[i]);
  const double r = block_max(v, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}
__global__ void iw_max_final_kernel(const double *part, double *out) {
  __shared__ double sh[RB];
  double v = -INFINITY;
  for (int i = threadIdx.x; i < RG; i += RB) v = fmax(v, part[i]);
  const double r = block_max(v, sh);
  if (threadIdx.x == 0) out[0] = r;
}

// pass 2: sums of exp(dF) [iw_prune.m:19-20], exp(dF - max), exp(2 (dF - max)) [essinv.m:14], and the max of exp(dF)
__global__ void iw_sums_kernel(int64_t M, const double *__restrict__ lfex, const double *__restrict__ lfapp,
                               const double *dmax, double *part) {
  __shared__ double sh[RB];
  const double mx = dmax[0];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < M; i += (int64_t)RG * RB) {
    const double dF = lfex[i] - lfapp[i];
    s0 += exp(dF);
    const double e = exp(dF - mx);
    s1 += e;
    s2 += exp((dF - mx) * 2.0);
  }
  const double r0 = block_sum(s0, sh), r1 = block_sum(s1, sh), r2 = block_sum(s2, sh);
  if (threadIdx.x == 0) { part[blockIdx.x] = r0; part[RG + blockIdx.x] = r1; part[2 * RG + blockIdx.x] = r2; }
}
// sums[j] = sum of part[j*RG .. j*RG+RG) for nsum streams
__global__ void iw_sums_final_kernel(const double *part, int nsum, double *sums) {
  __shared__ double sh[RB];
  for (int j = 0; j < nsum; j++) {
    double v = 0.0;
    for (int i = threadIdx.x; i < RG; i += RB) v += part[j * RG + i];
    const double r = block_sum(v, sh);
    if (threadIdx.x == 0) sums[j] = r;
  }
}

// pass 3: with renorm = mean(exp(dF)) and lZex = log(mean(exp(dF - max))) known:
//   weights w = exp(dF)/renorm                              (iw_prune.m:19-21)
//   sum (w - 1)^2                                           (iw_prune.m:29)
//   sum |exp(lFex - log renorm) - exp(lFapp)| / exp(lFapp)  (iw_prune.m:26)
//   sum (exp(0.5 (dF - max - lZex)) - 1)^2                  (hellinger.m:15)
__global__ void iw_pass3_kernel(int64_t M, const double *__restrict__ lfex, const double *__restrict__ lfapp, double renorm,
                                double log_renorm, double mx, double lzex, double *weights, double *part) {
  __shared__ double sh[RB];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < M; i += (int64_t)RG * RB) {
    const double fe = lfex[i], fa = lfapp[i];
    const double w = exp(fe - fa) / renorm;
    if (weights) weights[i] = w;
    s0 += (w - 1.0) * (w - 1.0);
    const double ea = exp(fa);
    s1 += fabs(exp(fe - log_renorm) - ea) / ea;
    const double h = exp(0.5 * ((fe - fa) - mx - lzex)) - 1.0;
    s2 += h * h;
  }
  const double r0 = block_sum(s0, sh), r1 = block_sum(s1, sh), r2 = block_sum(s2, sh);
  if (threadIdx.x == 0) { part[blockIdx.x] = r0; part[RG + blockIdx.x] = r1; part[2 * RG + blockIdx.x] = r2; }
}

// ------------------------------------------------------------------------------------------------
// consumers: independence Metropolis-Hastings prune (mcmc_prune.m:24-43)
// ------------------------------------------------------------------------------------------------
// The reference chain is sequential in the index c of the last ACCEPTED sample: proposal j replaces c iff
//   exp(((lFex(j) - lFex(c)) - lFapp(j)) + lFapp(c)) >= u(j-1)            (reference rounding order, :25-27)
// Parallel formulation with the reference's arithmetic untouched:
//   1. nxt[c] = first j > c that WOULD be accepted if c were the current sample (M if none): independent per c,
//      a short forward scan (expected length 1 / acceptance rate; long scans are finished by the whole warp);
//   2. the chain is the path 0 -> nxt[0] -> nxt[nxt[0]] -> ...; its nodes are marked by pointer doubling
//      (log2 M rounds of "mark J[c] for every marked c, then J = J o J");
//   3. src[i] = last marked node <= i (inclusive max-scan); rejections = unmarked positions; a completed run of L
//      rejections ends at every marked node whose predecessor on the path lies L + 1 back.
__device__ __forceinline__ bool mh_accept(double fe_j, double fa_j, double u_jm1, double fe_c, double fa_c) {
  double al = __dsub_rn(fe_j, fe_c);
  al = __dsub_rn(al, fa_j);
  al = __dadd_rn(al, fa_c);
  return !(exp(al) < u_jm1);
}

__global__ void mh_next_kernel(int M, const double *__restrict__ lfex, const double *__restrict__ lfapp,
                               const double *__restrict__ u, int *nxt) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < M;
  double fe_c = 0.0, fa_c = 0.0;
  int found = M;
  if (live) {
    fe_c = lfex[c]; fa_c = lfapp[c];
    const int jend = min(c + 32, M - 1);
    for (int j = c + 1; j <= jend; j++)
      if (mh_accept(lfex[j], lfapp[j], u[j - 1], fe_c, fa_c)) { found = j; break; }
  }
  // scans that did not finish within 32 proposals are completed by the whole warp, 32 proposals per step
  unsigned pending = __ballot_sync(FULL, live && found == M && c + 32 < M - 1);
  while (pending) {
    const int src_lane = __ffs(pending) - 1;
    pending &= pending - 1;
    const int cc = __shfl_sync(FULL, c, src_lane);
    const double fe = __shfl_sync(FULL, fe_c, src_lane), fa = __shfl_sync(FULL, fa_c, src_lane);
    int res = M;
    for (int j0 = cc + 33; j0 < M; j0 += 32) {
      const int j = j0 + lane;
      const bool acc = j < M && mh_accept(lfex[j], lfapp[j], u[j - 1], fe, fa);
      const unsigned m = __ballot_sync(FULL, acc);
      if (m) { res = j0 + __ffs(m) - 1; break; }
    }
    if (lane == src_lane) found = res;
  }
  if (live) nxt[c] = found;
}

__global__ void mh_init_kernel(int M, const int *__restrict__ nxt, int *J, unsigned char *onpath) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < M) { J[c] = nxt[c]; onpath[c] = c == 0; }
  if (c == M) J[M] = M;
}

// one doubling round: mark the 2^r-th successor of every marked node, then square the jump table
__global__ void mh_double_kernel(int M, const int *__restrict__ J, int *Jn, unsigned char *onpath) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > M) return;
  if (c == M) { Jn[M] = M; return; }
  const int t = J[c];
  if (onpath[c] && t < M) onpath[t] = 1;
  Jn[c] = J[t];
}

constexpr int SCAN_B = 1024;          // threads per scan block
constexpr int SCAN_E = 4;             // elements per thread
constexpr int SCAN_SEG = SCAN_B * SCAN_E;

// inclusive max-scan of v[i] = onpath[i] ? i : 0 (the path starts at 0 and its indices increase)
__global__ void mh_segmax_kernel(int M, const unsigned char *__restrict__ onpath, int *segmax) {
  __shared__ int sh[SCAN_B];
  const int base = blockIdx.x * SCAN_SEG;
  int v = 0;
  for (int e = 0; e < SCAN_E; e++) {
    const int i = base + e * SCAN_B + threadIdx.x;
    if (i < M && onpath[i]) v = max(v, i);
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = SCAN_B / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = max(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) segmax[blockIdx.x] = sh[0];
}
// exclusive max-scan of the segment maxima, in place, by one block: each thread owns a contiguous slice
__global__ void mh_segscan_kernel(int nseg, int *segmax) {
  __shared__ int sh[SCAN_B];
  const int per = (nseg + SCAN_B - 1) / SCAN_B;
  const int b0 = threadIdx.x * per, b1 = min(nseg, b0 + per);
  int run = 0;
  for (int b = b0; b < b1; b++) run = max(run, segmax[b]);
  sh[threadIdx.x] = run;
  __syncthreads();
  for (int off = 1; off < SCAN_B; off <<= 1) {
    const int t = (int)threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] = max(sh[threadIdx.x], t);
    __syncthreads();
  }
  run = threadIdx.x > 0 ? sh[threadIdx.x - 1] : 0;
  for (int b = b0; b < b1; b++) { const int v = segmax[b]; segmax[b] = run; run = max(run, v); }
}
__global__ void mh_src_kernel(int M, const unsigned char *__restrict__ onpath, const int *__restrict__ segmax, int32_t *src,
                              unsigned long long *counters, unsigned long long *rej_hist, int rej_hist_len) {
  __shared__ int sh[SCAN_B];
  const int base = blockIdx.x * SCAN_SEG + threadIdx.x * SCAN_E;   // each thread owns SCAN_E consecutive positions
  int v[SCAN_E];
  int run = 0;
  for (int e = 0; e < SCAN_E; e++) {
    const int i = base + e;
    if (i < M && onpath[i]) run = max(run, i);
    v[e] = run;
  }
  sh[threadIdx.x] = run;
  __syncthreads();
  // Hillis-Steele inclusive max-scan over the thread totals
  for (int off = 1; off < SCAN_B; off <<= 1) {
    const int t = (int)threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
    __syncthreads();
    sh[threadIdx.x] = max(sh[threadIdx.x], t);
    __syncthreads();
  }
  const int before = max(segmax[blockIdx.x], threadIdx.x > 0 ? sh[threadIdx.x - 1] : 0);
  unsigned long long rejects = 0;
  for (int e = 0; e < SCAN_E; e++) {
    const int i = base + e;
    if (i >= M) break;
    const int s = max(v[e], before);
    src[i] = s;
    if (s != i) {
      rejects++;
    } else if (i > 0 && rej_hist_len > 0) {
      const int prev = e > 0 ? max(v[e - 1], before) : before;   // last path node <= i - 1
      const int len = i - prev - 1;
      if (len > 0) atomicAdd(rej_hist + (min(len, rej_hist_len) - 1), 1ULL);
    }
  }
  if (rejects) atomicAdd(counters, rejects);
}

// host-side helpers ---------------------------------------------------------------------------------
// scratch from the stream-ordered pool: cudaMalloc / cudaFree per call would cost more than the kernels
// ------------------------------------------------------------------------------------------------
// tracemult (reference matlab/utils/tracemult.c, real case): the indexed batched product and the column pick that
// tt_irt_sqr.m is written in.  Inside the squared-density path both are fused into its kernels (ttirt_sqr.cu); these are the
// standalone operators for callers that use the MEX directly (tt_irt_sqr.m:77,109,139,146-149,205, tt_rt_sqr.m, tt_irt_fourier.m).
// ------------------------------------------------------------------------------------------------
// C(:,:,i) = A(:,:,i) * B(:,:,j(i))   tracemult.c:103-112;  A p x m x n, B m x k x s, C p x k x n, j 1-based doubles (:106-107).
// One CTA per group of slices, one thread per output element and slice; a thread walks a contiguous column of B (whole
// cache lines, reused by the p threads of that column) and a stride-p row of A.  Slices that pick the same j share B in L1 / L2.
__global__ void tracemult_kernel(int p, int m, int k, int64_t n, int64_t s, const double *__restrict__ A, const double *__restrict__ j,
                                 const double *__restrict__ B, double *C, int *bad) {
  const int64_t per = (int64_t)p * k;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per * n) return;
  const int64_t i = e / per;
  const int o = (int)(e - i * per);
  const int pp = o % p, kk = o / p;
  const int64_t jj = (int64_t)j[i] - 1;                    // matlab -> C indexing (:107)
  if (jj < 0 || jj >= s) { if (bad) atomicExch(bad, 1); C[e] = __longlong_as_double(0x7ff8000000000000LL); return; }
  const double *a = A + pp + (int64_t)p * m * i;
  const double *b = B + (int64_t)m * kk + (int64_t)m * k * jj;
  double acc = 0.0;
  for (int l = 0; l < m; l++) acc = fma(a[(int64_t)p * l], b[l], acc);
  C[e] = acc;
}

// C(i) = A(i, j(i))   tracemult.c:131-136;  A n x s
__global__ void tracepick_kernel(int64_t n, int64_t s, const double *__restrict__ A, const double *__restrict__ j, double *C, int *bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t jj = (int64_t)j[i] - 1;
  if (jj < 0 || jj >= s) { if (bad) atomicExch(bad, 1); C[i] = __longlong_as_double(0x7ff8000000000000LL); return; }
  C[i] = A[i + jj * n];
}

struct DevBuf {
  void *p = nullptr;
  cudaStream_t st = nullptr;
  ~DevBuf() { if (p) cudaFreeAsync(p, st); }
  int alloc(size_t bytes, cudaStream_t stream = nullptr) {
    static thread_local int pool_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pool_dev != dev) {   // keep freed scratch in the pool instead of returning it to the driver at every sync
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_dev = dev;
    }
    st = stream;
    return cudaMallocAsync(&p, bytes ? bytes : 1, st) == cudaSuccess ? 0 : -1;
  }
  template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

int need_device() {
  if (ttirt_device_count() <= 0) return aux_fail("no CUDA device available (this library has no CPU fallback)");
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" int ttirt_seeds_lattice_device(int64_t d, int64_t M, int64_t m0, int64_t N, const double *d_genvec,
                                          const double *d_shift, double *d_q, int64_t ldq, void *stream) {
  if (d < 1 || M < 0 || N < 1 || ldq < M || !d_genvec || !d_shift || (M > 0 && !d_q)) return aux_fail("bad arguments to ttirt_seeds_lattice_device");
  if (M == 0) return 0;
  const dim3 grid((unsigned)((M + 255) / 256), (unsigned)d);
  lattice_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((int)d, M, m0, 1.0 / (double)N, d_genvec, d_shift, d_q, ldq);
  ttirt::aux_launched();
  CKA(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_seeds_uniform_device(int64_t d, int64_t M, int64_t m0, uint64_t seed, double *d_q, int64_t ldq, void *stream) {
  if (d < 1 || M < 0 || ldq < M || (M > 0 && !d_q)) return aux_fail("bad arguments to ttirt_seeds_uniform_device");
  if (M == 0) return 0;
  const dim3 grid((unsigned)((M + 255) / 256), (unsigned)d);
  uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((int)d, M, m0, seed, d_q, ldq);
  ttirt::aux_launched();
  CKA(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_truncnormal_map_device(int64_t n, double sigma, const double *d_u, double *d_y, void *stream) {
  if (n < 0 || !(sigma > 0.0) || (n > 0 && (!d_u || !d_y))) return aux_fail("bad arguments to ttirt_truncnormal_map_device");
  if (n == 0) return 0;
  const double cdf_ifactor = erf(sigma / sqrt(2.0)) / 0.5;   // randref.m:31
  truncnormal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, cdf_ifactor, d_u, d_y);
  ttirt::aux_launched();
  CKA(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_iw_stats_device(int64_t M, const double *d_lfex, const double *d_lfapp, double *d_weights,
                                     double *out, void *stream) {
  if (M < 1 || !d_lfex || !d_lfapp || !out) return aux_fail("bad arguments to ttirt_iw_stats_device");
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf part, scal;
  if (part.alloc(sizeof(double) * 3 * RG, st) || scal.alloc(sizeof(double) * 8, st)) return aux_fail("out of device memory");
  double *dp = part.as<double>(), *ds = scal.as<double>();
  iw_max_kernel<<<RG, RB, 0, st>>>(M, d_lfex, d_lfapp, dp);
  iw_max_final_kernel<<<1, RB, 0, st>>>(dp, ds);
  iw_sums_kernel<<<RG, RB, 0, st>>>(M, d_lfex, d_lfapp, ds, dp);
  iw_sums_final_kernel<<<1, RB, 0, st>>>(dp, 3, ds + 1);
  for (int i = 0; i < 4; i++) ttirt::aux_launched();
  double h[4];
  CKA(cudaMemcpyAsync(h, ds, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  CKA(cudaStreamSynchronize(st));
  const double mx = h[0], s_exp = h[1], s1 = h[2], s2 = h[3];
  const double renorm = s_exp / (double)M;                  // iw_prune.m:20
  const double log_renorm = log(renorm);                    // :25
  const double lzex = log(s1 / (double)M);                  // hellinger.m:14
  iw_pass3_kernel<<<RG, RB, 0, st>>>(M, d_lfex, d_lfapp, renorm, log_renorm, mx, lzex, d_weights, dp);
  iw_sums_final_kernel<<<1, RB, 0, st>>>(dp, 3, ds + 4);
  ttirt::aux_launched(); ttirt::aux_launched();
  double g[3];
  CKA(cudaMemcpyAsync(g, ds + 4, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
  CKA(cudaStreamSynchronize(st));
  out[0] = sqrt(g[0] / (double)M);                          // isstd      iw_prune.m:29
  out[1] = exp(mx) / renorm;                                // max_ratio  iw_prune.m:24 (exp is monotone: max of the ratios)
  out[2] = g[1] / (double)M;                                // err1       iw_prune.m:26
  out[3] = (double)M * s2 / (s1 * s1);                      // tau        essinv.m:14
  out[4] = sqrt(g[2] / (double)M / 2.0);                    // H          hellinger.m:15-16
  out[5] = log_renorm;                                      // log of the IS normalisation constant
  return 0;
}

extern "C" int ttirt_mcmc_prune_device(int64_t M, const double *d_lfex, const double *d_lfapp, const double *d_u,
                                       int32_t *d_src, int64_t *num_rejects, int64_t *rej_hist, int64_t rej_hist_len, void *stream) {
  if (M < 1 || M > 2147483000LL || !d_lfex || !d_lfapp || !d_src || (M > 1 && !d_u) || rej_hist_len < 0 || (rej_hist_len > 0 && !rej_hist))
    return aux_fail("bad arguments to ttirt_mcmc_prune_device");
  cudaStream_t st = (cudaStream_t)stream;
  const int m = (int)M;
  const int nseg = (m + SCAN_SEG - 1) / SCAN_SEG;
  DevBuf cnt, hist, nxt, j0, j1, onp, seg;
  if (cnt.alloc(sizeof(unsigned long long) * 2, st) || hist.alloc(sizeof(unsigned long long) * (size_t)rej_hist_len, st) || nxt.alloc(sizeof(int) * (size_t)m, st) ||
      j0.alloc(sizeof(int) * ((size_t)m + 1), st) || j1.alloc(sizeof(int) * ((size_t)m + 1), st) || onp.alloc((size_t)m, st) || seg.alloc(sizeof(int) * (size_t)nseg, st))
    return aux_fail("out of device memory");
  CKA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long) * 2, st));
  if (rej_hist_len > 0) CKA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned long long) * (size_t)rej_hist_len, st));
  const unsigned g = (unsigned)((m + 1 + 255) / 256);
  mh_next_kernel<<<g, 256, 0, st>>>(m, d_lfex, d_lfapp, d_u, nxt.as<int>());
  mh_init_kernel<<<g, 256, 0, st>>>(m, nxt.as<int>(), j0.as<int>(), onp.as<unsigned char>());
  int launches = 2;
  int *ja = j0.as<int>(), *jb = j1.as<int>();
  for (int64_t reach = 1; reach < M; reach <<= 1) {   // after round r every path node within 2^(r+1) hops of node 0 is marked
    mh_double_kernel<<<g, 256, 0, st>>>(m, ja, jb, onp.as<unsigned char>());
    std::swap(ja, jb);
    launches++;
  }
  mh_segmax_kernel<<<nseg, SCAN_B, 0, st>>>(m, onp.as<unsigned char>(), seg.as<int>());
  mh_segscan_kernel<<<1, SCAN_B, 0, st>>>(nseg, seg.as<int>());
  mh_src_kernel<<<nseg, SCAN_B, 0, st>>>(m, onp.as<unsigned char>(), seg.as<int>(), d_src, cnt.as<unsigned long long>(),
                                         hist.as<unsigned long long>(), (int)rej_hist_len);
  launches += 3;
  for (int i = 0; i < launches; i++) ttirt::aux_launched();
  CKA(cudaGetLastError());
  unsigned long long h[2];
  CKA(cudaMemcpyAsync(h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
  if (rej_hist_len > 0) CKA(cudaMemcpyAsync(rej_hist, hist.p, sizeof(int64_t) * (size_t)rej_hist_len, cudaMemcpyDeviceToHost, st));
  CKA(cudaStreamSynchronize(st));
  if (num_rejects) *num_rejects = (int64_t)h[0];
  return 0;
}

// ---- host-buffer forms (device 0 unless TTIRT_DEVICE says otherwise): allocate, copy, run, copy back ----
static int pick_device() {
  const char *e = getenv("TTIRT_DEVICE");
  const int dev = e ? atoi(e) : 0;
  return cudaSetDevice(dev) == cudaSuccess ? 0 : aux_fail("cudaSetDevice(%d) failed", dev);
}

extern "C" int ttirt_seeds_lattice_host(int64_t d, int64_t M, int64_t m0, int64_t N, const int64_t *genvec, const double *shift,
                                        double *h_q, int64_t ld) {
  if (need_device() || pick_device()) return -1;
  if (d < 1 || M < 0 || ld < M || !genvec || !shift || (M > 0 && !h_q)) return aux_fail("bad arguments to ttirt_seeds_lattice_host");
  if (M == 0) return 0;
  std::vector<double> z(d);
  for (int64_t k = 0; k < d; k++) z[k] = (double)genvec[k];
  DevBuf dz, ds, dq;
  if (dz.alloc(sizeof(double) * d) || ds.alloc(sizeof(double) * d) || dq.alloc(sizeof(double) * M * d)) return aux_fail("out of device memory");
  CKA(cudaMemcpy(dz.p, z.data(), sizeof(double) * d, cudaMemcpyHostToDevice));
  CKA(cudaMemcpy(ds.p, shift, sizeof(double) * d, cudaMemcpyHostToDevice));
  if (ttirt_seeds_lattice_device(d, M, m0, N, dz.as<double>(), ds.as<double>(), dq.as<double>(), M, nullptr) != 0) return -1;
  CKA(cudaMemcpy2D(h_q, sizeof(double) * ld, dq.p, sizeof(double) * M, sizeof(double) * M, d, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ttirt_seeds_uniform_host(int64_t d, int64_t M, int64_t m0, uint64_t seed, double *h_q, int64_t ld) {
  if (need_device() || pick_device()) return -1;
  if (d < 1 || M < 0 || ld < M || (M > 0 && !h_q)) return aux_fail("bad arguments to ttirt_seeds_uniform_host");
  if (M == 0) return 0;
  DevBuf dq;
  if (dq.alloc(sizeof(double) * M * d)) return aux_fail("out of device memory");
  if (ttirt_seeds_uniform_device(d, M, m0, seed, dq.as<double>(), M, nullptr) != 0) return -1;
  CKA(cudaMemcpy2D(h_q, sizeof(double) * ld, dq.p, sizeof(double) * M, sizeof(double) * M, d, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ttirt_truncnormal_map_host(int64_t n, double sigma, const double *h_u, double *h_y) {
  if (need_device() || pick_device()) return -1;
  if (n < 0 || (n > 0 && (!h_u || !h_y))) return aux_fail("bad arguments to ttirt_truncnormal_map_host");
  if (n == 0) return 0;
  DevBuf du, dy;
  if (du.alloc(sizeof(double) * n) || dy.alloc(sizeof(double) * n)) return aux_fail("out of device memory");
  CKA(cudaMemcpy(du.p, h_u, sizeof(double) * n, cudaMemcpyHostToDevice));
  if (ttirt_truncnormal_map_device(n, sigma, du.as<double>(), dy.as<double>(), nullptr) != 0) return -1;
  CKA(cudaMemcpy(h_y, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ttirt_iw_stats_host(int64_t M, const double *h_lfex, const double *h_lfapp, double *h_weights, double *out) {
  if (need_device() || pick_device()) return -1;
  if (M < 1 || !h_lfex || !h_lfapp || !out) return aux_fail("bad arguments to ttirt_iw_stats_host");
  DevBuf de, da, dw;
  if (de.alloc(sizeof(double) * M) || da.alloc(sizeof(double) * M) || (h_weights && dw.alloc(sizeof(double) * M))) return aux_fail("out of device memory");
  CKA(cudaMemcpy(de.p, h_lfex, sizeof(double) * M, cudaMemcpyHostToDevice));
  CKA(cudaMemcpy(da.p, h_lfapp, sizeof(double) * M, cudaMemcpyHostToDevice));
  if (ttirt_iw_stats_device(M, de.as<double>(), da.as<double>(), h_weights ? dw.as<double>() : nullptr, out, nullptr) != 0) return -1;
  if (h_weights) CKA(cudaMemcpy(h_weights, dw.p, sizeof(double) * M, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ttirt_mcmc_prune_host(int64_t M, const double *h_lfex, const double *h_lfapp, const double *h_u, int32_t *h_src,
                                     int64_t *num_rejects, int64_t *rej_hist, int64_t rej_hist_len) {
  if (need_device() || pick_device()) return -1;
  if (M < 1 || !h_lfex || !h_lfapp || !h_src || (M > 1 && !h_u)) return aux_fail("bad arguments to ttirt_mcmc_prune_host");
  DevBuf de, da, du, dsrc;
  if (de.alloc(sizeof(double) * M) || da.alloc(sizeof(double) * M) || du.alloc(sizeof(double) * M) || dsrc.alloc(sizeof(int32_t) * M))
    return aux_fail("out of device memory");
  CKA(cudaMemcpy(de.p, h_lfex, sizeof(double) * M, cudaMemcpyHostToDevice));
  CKA(cudaMemcpy(da.p, h_lfapp, sizeof(double) * M, cudaMemcpyHostToDevice));
  if (M > 1) CKA(cudaMemcpy(du.p, h_u, sizeof(double) * (M - 1), cudaMemcpyHostToDevice));
  if (ttirt_mcmc_prune_device(M, de.as<double>(), da.as<double>(), du.as<double>(), dsrc.as<int32_t>(), num_rejects, rej_hist, rej_hist_len, nullptr) != 0)
    return -1;
  CKA(cudaMemcpy(h_src, dsrc.p, sizeof(int32_t) * M, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int ttirt_tracemult_device(int64_t p, int64_t m, int64_t k, int64_t n, int64_t s, const double *d_A, const double *d_j,
                                      const double *d_B, double *d_C, void *stream) {
  if (n < 0 || (n > 0 && (!d_A || !d_j || !d_C))) return aux_fail("bad arguments to ttirt_tracemult_device");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_B) {                                               // two-argument form: C(i) = A(i, j(i)), A n x s
    if (s < 1) return aux_fail("bad arguments to ttirt_tracemult_device");
    tracepick_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, s, d_A, d_j, d_C, nullptr);
  } else {
    if (p < 1 || m < 1 || k < 1 || s < 1 || p > (1 << 20) || k > (1 << 20) || m > (1 << 24)) return aux_fail("bad shapes for ttirt_tracemult_device");
    const int64_t tot = p * k * n;
    if ((tot + 255) / 256 > 0x7fffffffLL) return aux_fail("ttirt_tracemult_device: too many output elements");
    tracemult_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((int)p, (int)m, (int)k, n, s, d_A, d_j, d_B, d_C, nullptr);
  }
  ttirt::aux_launched();
  CKA(cudaGetLastError());
  return 0;
}

extern "C" int ttirt_tracemult_host(int64_t p, int64_t m, int64_t k, int64_t n, int64_t s, const double *h_A, const double *h_j,
                                    const double *h_B, double *h_C) {
  if (need_device() || pick_device()) return -1;
  if (n < 0 || (n > 0 && (!h_A || !h_j || !h_C))) return aux_fail("bad arguments to ttirt_tracemult_host");
  if (n == 0) return 0;
  for (int64_t i = 0; i < n; i++)                          // the MEX would read out of bounds here; fail instead
    if (!(h_j[i] >= 1.0 && h_j[i] <= (double)s)) return aux_fail("tracemult: j(%lld) = %g outside 1..%lld", (long long)(i + 1), h_j[i], (long long)s);
  const size_t na = h_B ? (size_t)p * m * n : (size_t)n * s, nb = h_B ? (size_t)m * k * s : 0, nc = h_B ? (size_t)p * k * n : (size_t)n;
  DevBuf da, dj, db, dc;
  if (da.alloc(sizeof(double) * na) || dj.alloc(sizeof(double) * n) || (h_B && db.alloc(sizeof(double) * nb)) || dc.alloc(sizeof(double) * nc))
    return aux_fail("out of device memory");
  CKA(cudaMemcpy(da.p, h_A, sizeof(double) * na, cudaMemcpyHostToDevice));
  CKA(cudaMemcpy(dj.p, h_j, sizeof(double) * n, cudaMemcpyHostToDevice));
  if (h_B) CKA(cudaMemcpy(db.p, h_B, sizeof(double) * nb, cudaMemcpyHostToDevice));
  if (ttirt_tracemult_device(p, m, k, n, s, da.as<double>(), dj.as<double>(), h_B ? db.as<double>() : nullptr, dc.as<double>(), nullptr) != 0) return -1;
  CKA(cudaMemcpy(h_C, dc.p, sizeof(double) * nc, cudaMemcpyDeviceToHost));
  return 0;
}
