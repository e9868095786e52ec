// Fast path of tt_irt1 for small uniform-rank TT densities (r = 8 or 16 on 17-point grids: the shapes of the reference's
// MCMC / importance-weighting drivers, BASELINE.json configs[0] and [1]): ONE persistent kernel walks all d dimensions.
//
// For these shapes the per-dimension "sort + transition" pipeline (ttirt_fast.cu) is bound by everything except arithmetic:
// 2d launches per chunk, the left-interface rows round-tripping through HBM every dimension (256 B per sample-dimension
// against 16 algorithmic bytes), a counting sort per dimension, and 8x8x4 DMMA tiles whose per-row bookkeeping does not
// shrink with r^2.  Here a whole core (r x n x r doubles, 35 KB at r = 16) fits in shared memory, so
//   * one lane owns one sample for the whole walk: its left interface f (r doubles) never leaves registers
//     (reference tt_irt1_int32.c:90-91, 167-177), q is read and z written once, coalesced (:135,159);
//   * the per-dimension operands -- P_k with node-weighted columns, the grid tables, core_k transposed to [node][a][b] --
//     are packed once per model (walk_pack_kernel) and stream through a three-stage ring of shared-memory buffers, one
//     1-D TMA bulk copy per dimension, issued by a producer warp and handed over on mbarriers (full / empty);
//   * conditional pdf (:103-105) and interface update (:167-177) are FP64 FMA chains per lane: the pdf operand is a
//     warp-wide broadcast, the two core slabs of a lane's interval are read from the transposed core with a node stride of
//     r*r + 1 doubles, so the 16 intervals of a 17-point grid map to 16 different bank pairs and lanes in the same interval
//     share a broadcast: one shared-memory wavefront per load whatever the mix of intervals in the warp -- no sort;
//   * CDF (:107-113), search (:134-142), inversion (:146-159) and the log-density (:161-165) use the same scaled
//     formulation as the transition kernel's tail (weighted running sums, unnormalised compare, power-of-two scaling,
//     split-product log-density), one row per lane.
// Dimension 0 (left rank 1) is the shared table of the stage-0 kernel, in the reference's operation order.
// Served: every n = 17, every inner rank <= 16 (interface width 8 or 16, smaller ranks zero-padded), d >= 2; everything else
// uses the per-dimension path.
#include "ttirt_common.cuh"

namespace ttirt {

namespace {

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

#ifndef TTIRT_WALK_WARPS
#define TTIRT_WALK_WARPS 15
#endif
constexpr int WALK_WARPS = TTIRT_WALK_WARPS;    // consumer warps per CTA (one sample per lane); warp WALK_WARPS is the producer
constexpr int WALK_STAGES = 3;
constexpr int WALK_THREADS = 32 * (WALK_WARPS + 1);

constexpr int even_up(int x) { return (x + 1) & ~1; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// One dimension's operands as one contiguous, 16-byte aligned block of doubles (the unit of the TMA ring).
template <int R, int N>
struct WalkLayout {
  static constexpr int NS = R * R + 1;              // node stride of the transposed core: odd, see the file header
  // dimensions k >= 1
  static constexpr int PW = 0;                      // Pw[j][a] = node_weight(j) * P_k[a, j]
  static constexpr int X = N * R;                   // grid
  static constexpr int IH = X + N;                  // 1 / (x[j+1] - x[j])
  static constexpr int RW = IH + N;                 // 1 / node_weight(j)
  static constexpr int HR = RW + N;                 // h_{j-1} / node_weight(j)
  static constexpr int CORE = even_up(HR + N);      // core_k as [node][a * R + b], node stride NS (absent in the last dimension)
  static constexpr int SIZE_LAST = CORE;
  static constexpr int SIZE_MID = even_up(CORE + N * NS);
  // dimension 0 (left rank 1): the stage-0 tables and core_0 as [node][b]
  static constexpr int P0 = 0, C0 = N, X0 = 2 * N, CORE0 = even_up(3 * N);
  static constexpr int SIZE0 = even_up(CORE0 + N * R);
  static constexpr int STAGE = cmax(SIZE_MID, SIZE0);   // doubles per ring stage and per dimension in the packed buffer
  static constexpr size_t smem_bytes = sizeof(double) * WALK_STAGES * STAGE + sizeof(uint64_t) * 2 * WALK_STAGES;
};

// bisection with strict '>' (reference :134-142), as the stage-0 kernel of the per-dimension path
__device__ __forceinline__ int walk_search0(const double *cdf, int nk, double qk) {
  int lo = 0, hi = nk - 1;
  while (hi - lo > 1) {
    const int mid = (int)((double)(lo + hi) * 0.5);
    if (qk > cdf[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

template <int R, int N>
__global__ void walk_pack_kernel(const DimInfo *__restrict__ dims, int d, const double *__restrict__ xs, const double *__restrict__ core,
                                 const double *__restrict__ pk, const double *__restrict__ p0, const double *__restrict__ cdf0,
                                 double *pack) {
  using L = WalkLayout<R, N>;
  const int k = blockIdx.x;
  const DimInfo di = dims[k];
  double *blk = pack + (size_t)k * L::STAGE;
  const double *x = xs + di.off_x, *ck = core + di.off_c;
  // ranks below R are padded with zeros: the padded interface entries stay exactly zero through every update and add
  // exact zeros to every sum, so results do not depend on the padding
  const int r0 = di.r0, r1 = di.r1;
  if (k == 0) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) { blk[L::P0 + i] = p0[i]; blk[L::C0 + i] = cdf0[i]; blk[L::X0 + i] = x[i]; }
    for (int e = threadIdx.x; e < N * R; e += blockDim.x) {
      const int i = e / R, b = e - i * R;
      blk[L::CORE0 + e] = b < r1 ? ck[i + (int64_t)b * N] : 0.0;          // core_0[0, i, b], r_0 = 1
    }
    return;
  }
  for (int e = threadIdx.x; e < N * R; e += blockDim.x) {
    const int j = e / R, a = e - j * R;
    blk[L::PW + e] = a < r0 ? pk[di.off_p + a + j * r0] * node_weight(x, j, N) : 0.0;
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double w = node_weight(x, i, N);
    const double hl = i >= 1 ? 0.5 * (x[i] - x[i - 1]) : 0.0;
    blk[L::X + i] = x[i];
    blk[L::IH + i] = i + 1 < N ? 1.0 / (x[i + 1] - x[i]) : 0.0;
    blk[L::RW + i] = w > 0.0 ? 1.0 / w : 0.0;
    blk[L::HR + i] = w > 0.0 ? hl / w : 0.0;
  }
  if (k < d - 1) {
    for (int e = threadIdx.x; e < N * L::NS; e += blockDim.x) {
      const int i = e / L::NS, ab = e - i * L::NS;
      double v = 0.0;
      if (ab < R * R) {
        const int a = ab / R, b = ab - a * R;
        if (a < r0 && b < r1) v = ck[a + i * r0 + (int64_t)b * r0 * N];
      }
      blk[L::CORE + e] = v;
    }
  }
}

template <int R, int N>
__global__ void __launch_bounds__(WALK_THREADS, 1) walk_kernel(const WalkArgs a) {
  using L = WalkLayout<R, N>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *ring = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + WALK_STAGES * L::STAGE), *empty = full + WALK_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < WALK_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, WALK_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int d = a.d;
  const int groups = (a.rows + 31) >> 5;                       // 32 samples per warp pass
  const int slots = gridDim.x * WALK_WARPS;
  const int iters = (groups + slots - 1) / slots;
  const int steps = iters * d;                                 // every warp of the CTA walks the same sequence of ring stages

  if (warp == WALK_WARPS) {
    // ---- producer: dimension t mod d into stage t mod WALK_STAGES, as soon as all consumer warps have released it ----
    if (lane == 0) {
      for (int t = 0; t < steps; t++) {
        const int s = t % WALK_STAGES, k = t % d;
        if (t >= WALK_STAGES) mbar_wait(empty + s, ((t / WALK_STAGES) - 1) & 1);
        const uint32_t bytes = 8u * (uint32_t)(k == 0 ? L::SIZE0 : (k == d - 1 ? L::SIZE_LAST : L::SIZE_MID));
        mbar_expect_tx(full + s, bytes);
        bulk_g2s(ring + s * L::STAGE, a.pack + (size_t)k * L::STAGE, bytes, full + s);
      }
    }
    return;
  }

  int t = 0;
  for (int it = 0; it < iters; it++) {
    const int gidx = blockIdx.x + gridDim.x * (warp + WALK_WARPS * it);   // groups spread over the SMs first
    const int64_t m = (int64_t)gidx * 32 + lane;
    const bool work = gidx < groups;                           // warp-uniform
    const bool live = work && m < a.rows;
    double f[R];
    double lpN = 1.0, lpD = 1.0;
    int lpE = 0;
    double qk = live ? a.q[m] : 0.5;
    for (int k = 0; k < d; k++, t++) {
      const int s = t % WALK_STAGES;
      mbar_wait(full + s, (t / WALK_STAGES) & 1);
      if (work) {
        const double *blk = ring + s * L::STAGE;
        const double qn = (live && k + 1 < d) ? a.q[m + a.ldq * (k + 1)] : 0.5;   // next seed, a whole dimension ahead
        if (k == 0) {
          // ---- dimension 0: the conditional is the same for every sample (stage-0 tables, reference order) ----
          const double *p0 = blk + L::P0, *c0 = blk + L::C0, *x0 = blk + L::X0, *cr = blk + L::CORE0;
          const int lo = walk_search0(c0, N, qk);
          const CellOut o = invert_cell(qk, c0[lo], p0[lo], p0[lo + 1], x0[lo], x0[lo + 1]);
          if (live) {
            a.z[m] = o.xk;
            if (a.idx_out) a.idx_out[m] = lo;
          }
          lp_accumulate(lpN, lpD, lpE, fabs(__dadd_rn(__dmul_rn(p0[lo], o.w1), __dmul_rn(p0[lo + 1], o.w2))), 1.0);  // p0 is normalised
#pragma unroll
          for (int b = 0; b < R; b++) f[b] = fma(o.w2, cr[(lo + 1) * R + b], o.w1 * cr[lo * R + b]);
        } else {
          // ---- conditional pdf on the grid, weighted by the trapezoid node weights: v_j = w_j |sum_a f_a P[a, j]| ----
          double v[N];
#pragma unroll
          for (int j = 0; j < N; j++) v[j] = 0.0;
#pragma unroll
          for (int a2 = 0; a2 < R; a2 += 2) {
#pragma unroll
            for (int j = 0; j < N; j++) {
              const double2 pw = *reinterpret_cast<const double2 *>(blk + L::PW + j * R + a2);   // warp-wide broadcast
              v[j] = fma(f[a2], pw.x, v[j]);
              v[j] = fma(f[a2 + 1], pw.y, v[j]);
            }
          }
          double total = 0.0;                                  // mass of the conditional: cdf at the last node
#pragma unroll
          for (int j = 0; j < N; j++) { v[j] = fabs(v[j]); total += v[j]; }
          // search: largest i0 <= N-2 with cdf_{i0} < q * mass,  cdf_j = R_j + (h_{j-1}/w_j) v_j,  R_j = sum_{i<j} v_i
          const double qt = qk * total;
          int i0 = 0;
          double dq = qt, va = v[0], vb = v[1], Rj = v[0];
          const double *hr = blk + L::HR;
#pragma unroll
          for (int j = 1; j <= N - 2; j++) {
            const double dj = qt - fma(hr[j], v[j], Rj);
            if (__double_as_longlong(dj) > 0) { i0 = j; dq = dj; va = v[j]; vb = v[j + 1]; }
            Rj += v[j];
          }
          const double s2 = pow2_scale(total);                 // exact power-of-two normalisation instead of 1 / mass
          double c1 = va * blk[L::RW + i0] * s2, c2 = vb * blk[L::RW + i0 + 1] * s2;
          double mass = total * s2;
          dq *= s2;
          if (total == 0.0) {
            // zero-mass conditional: uniform in index space (reference tt_irt1_int32.c:116-125)
            const double u = 1.0 / (double)(N - 1);
            const double sf = 1.0 / ((double)(N - 1) * u);
            int k0 = 0;
            for (int j = 1; j <= N - 2; j++) k0 += (qk > ((double)j * u) * sf) ? 1 : 0;
            i0 = k0; dq = qk - ((double)k0 * u) * sf; c1 = u * sf; c2 = u * sf; mass = 1.0;
          }
          const CellFast o = invert_cell_fast(dq, c1, c2, blk[L::X + i0], blk[L::X + i0 + 1], blk[L::IH + i0]);
          lp_accumulate(lpN, lpD, lpE, o.dens, mass);
          if (live) {
            a.z[m + a.ldz * k] = o.xk;
            if (a.idx_out) a.idx_out[m + a.ldz * k] = i0;
          }
          if (k + 1 < d) {
            // ---- interface update: f' = (w1 f) A_{i0} + (w2 f) A_{i0+1}, slabs of this lane's interval ----
            const double *A = blk + L::CORE + i0 * L::NS, *B = A + L::NS;
            double fn[R];
#pragma unroll
            for (int b = 0; b < R; b++) fn[b] = 0.0;
#pragma unroll
            for (int aa = 0; aa < R; aa++) {
              const double f1 = o.w1 * f[aa], f2 = o.w2 * f[aa];
#pragma unroll
              for (int b = 0; b < R; b++) {
                fn[b] = fma(f1, A[aa * R + b], fn[b]);
                fn[b] = fma(f2, B[aa * R + b], fn[b]);
              }
            }
#pragma unroll
            for (int b = 0; b < R; b++) f[b] = fn[b];
          }
        }
        qk = qn;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    if (live) a.lpz[m] = lp_finish(lpN, lpD, lpE);
  }
}

template <int R, int N>
cudaError_t walk_init_one() {
  return cudaFuncSetAttribute(walk_kernel<R, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WalkLayout<R, N>::smem_bytes);
}

template <int R, int N>
cudaError_t walk_launch_one(const WalkArgs &a, int sm_count, cudaStream_t st) {
  const int groups = (a.rows + 31) >> 5;
  int grid = sm_count < groups ? sm_count : groups;
  if (grid < 1) grid = 1;
  walk_kernel<R, N><<<grid, WALK_THREADS, WalkLayout<R, N>::smem_bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

// 0 / 1: walk class (interface width 8 / 16) for a TT whose inner ranks are all <= rmax on n-point grids in every dimension;
// -1: not served by the walk kernel.  Ranks below the class width are zero-padded by walk_pack.
int walk_class_for(int rmax, int n) {
  if (n != 17) return -1;
  if (rmax <= 8) return 0;
  if (rmax <= 16) return 1;
  return -1;
}

int64_t walk_pack_doubles(int cls, int d) {
  return (int64_t)d * (cls == 0 ? WalkLayout<8, 17>::STAGE : WalkLayout<16, 17>::STAGE);
}

cudaError_t walk_init(int) {
  cudaError_t e;
  if ((e = walk_init_one<8, 17>()) != cudaSuccess) return e;
  return walk_init_one<16, 17>();
}

cudaError_t walk_pack(int cls, const DimInfo *d_dims, int d, const double *xs, const double *core, const double *pk,
                      const double *p0, const double *cdf0, double *pack, cudaStream_t st) {
  if (cls == 0) walk_pack_kernel<8, 17><<<d, 256, 0, st>>>(d_dims, d, xs, core, pk, p0, cdf0, pack);
  else walk_pack_kernel<16, 17><<<d, 256, 0, st>>>(d_dims, d, xs, core, pk, p0, cdf0, pack);
  return cudaGetLastError();
}

cudaError_t launch_walk(int cls, const WalkArgs &a, int sm_count, cudaStream_t st) {
  return cls == 0 ? walk_launch_one<8, 17>(a, sm_count, st) : walk_launch_one<16, 17>(a, sm_count, st);
}

}  // namespace ttirt
