// Fast path of tt_irt1 on B200 (sm_100a): the fused per-dimension "transition" kernel.
//
// One launch advances every sample of a chunk from dimension k to dimension k+1:
//   (1) gathered, interpolated left-interface update (reference tt_irt1_int32.c:167-177):
//         F' = (w1 F) A_b + (w2 F) A_{b+1},  A_i = core_k[:, i, :],  b = interval chosen in dimension k
//       Samples arrive ordered by b (counting sort between launches), so a CTA stages the two r x r
//       slabs once in shared memory and runs a dense FP64 tensor-core contraction (DMMA, mma.sync
//       m8n8k4.f64) over all of its samples in that bin.
//   (2) conditional pdf on the grid of dimension k+1 (reference :103-105): p = |F' P_{k+1}|, a second
//       DMMA contraction whose A operand is the accumulator of (1) used in place (the contraction
//       index is permuted identically on both operands, so no register shuffles are needed).
//   (3) trapezoid CDF as a 4-lane prefix straight from the accumulator registers (:107-113),
//       normalisation (:116-130), interval search (:134-142), closed-form quadratic inversion
//       (:146-159), log-density accumulation (:161-165), emission of (interval, w1, w2) and the
//       histogram that drives the next counting sort.
//
// Work decomposition: a warp owns 16 samples end to end (two 8-row MMA tiles); warps never synchronise
// with each other except when their CTA moves to the next interval bin and restages a slab.
#include "ttirt_common.cuh"

namespace ttirt {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Shared-memory image of a B operand (K x NC, column-major in global memory with column stride cs).
// Column c lives at c*KP.  Within a column the contraction index a is permuted so that the four
// values a thread quad needs for one k-step are adjacent: chunk ch = 2*(a/8) + (a&1), lane (a&7)/2;
// chunks are XOR-swizzled with the column's low bits, which makes the quad-strided 8-byte loads of a
// half-warp hit 16 distinct bank pairs (no padding needed).  Entries outside K x NC are zero.
__device__ __forceinline__ int b_phys(int c, int a, int KP) {
  const int ch = ((a >> 3) << 1) | (a & 1);
  return c * KP + (((ch ^ (c & 3)) << 2) | ((a & 7) >> 1));
}

__device__ void stage_b(double *dst, const double *__restrict__ src, int K, int NC, int64_t cs, int KP, int NCP,
                        int tid, int nthr) {
  const int total = NCP * KP;
  for (int e = tid; e < total; e += nthr) {
    const int c = e / KP, a = e - c * KP;
    const double v = (a < K && c < NC) ? __ldg(src + a + (int64_t)c * cs) : 0.0;
    dst[b_phys(c, a, KP)] = v;
  }
}

template <int RT, int NT>
struct SmemLayout {
  static constexpr int KPMAX = (8 * RT + 15) & ~15;
  static constexpr int SLAB = 8 * RT * KPMAX;      // doubles per slab buffer
  static constexpr int PN = 8 * NT * KPMAX;        // doubles for P_{k+1}
  static constexpr int NBMAX = 8 * NT;             // >= n - 1 intervals
  static constexpr size_t bytes = sizeof(double) * (2 * SLAB + PN + 2 * 8 * NT) + sizeof(int) * (2 * (NBMAX + 1) + NBMAX);
};

// EXACT: r0 == r1 == 8*RT and ceil(n1/8) == NT, so every tile loop runs its full static trip count and no
// guard branches are compiled in (the steady state of a uniform-rank TT).  TAIL1 (EXACT only): n1 == 8*(NT-1)+1,
// the usual 2^p+1 grid; the lone last grid column is then a 4-lane DFMA dot product instead of a whole DMMA
// column tile that would be 7/8 padding.
template <int RT, int NT, int WARPS, bool EXACT, bool TAIL1>
__global__ void __launch_bounds__(WARPS * 32, 1) transition_kernel(const TransArgs a) {
  static_assert(EXACT || !TAIL1, "TAIL1 needs EXACT");
  using L = SmemLayout<RT, NT>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *slab0 = reinterpret_cast<double *>(smem_raw);
  double *slab1 = slab0 + L::SLAB;
  double *Ps = slab1 + L::SLAB;
  double *hh = Ps + L::PN;           // half grid steps of dimension k+1, zero beyond n1-2
  double *xg = hh + 8 * NT;          // grid of dimension k+1
  int *bts = reinterpret_cast<int *>(xg + 8 * NT);  // bin -> first CTA tile
  int *bst = bts + (L::NBMAX + 1);                   // bin -> first sorted row
  int *hist = bst + (L::NBMAX + 1);                  // histogram of the intervals chosen in dimension k+1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NTHR = WARPS * 32, ROWS_CTA = WARPS * 16;

  const int r0 = a.r0, r1 = a.r1, n1 = a.n1, nb0 = a.n0 - 1;
  constexpr int KP0 = L::KPMAX, KP1 = L::KPMAX;  // compile-time column pitch: B-fragment offsets fold into immediates
  constexpr int NTD = TAIL1 ? NT - 1 : NT;       // grid column tiles computed by DMMA
  const int ks0 = EXACT ? RT : (r0 + 7) >> 3, rt_act = EXACT ? RT : (r1 + 7) >> 3, nt_act = EXACT ? NT : (n1 + 7) >> 3;

  for (int i = tid; i <= nb0; i += NTHR) {
    bts[i] = a.bin_tile_start[i];
    bst[i] = a.bin_start[i];
  }
  for (int i = tid; i < 8 * NT; i += NTHR) {
    hh[i] = (i + 1 < n1) ? 0.5 * (a.xnext[i + 1] - a.xnext[i]) : 0.0;
    xg[i] = (i < n1) ? a.xnext[i] : 0.0;
    if (i < L::NBMAX) hist[i] = 0;
  }
  stage_b(Ps, a.pnext, r1, n1, r1, KP1, 8 * NT, tid, NTHR);
  __syncthreads();

  const int total_tiles = bts[nb0];
  const int t_begin = (int)(((int64_t)blockIdx.x * total_tiles) / gridDim.x);
  const int t_end = (int)(((int64_t)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int64_t slab_cs = (int64_t)r0 * a.n0;

  int cur0 = -1, cur1 = -1;  // interval slab held by slab0 / slab1
  int b = 0;
  for (int tile = t_begin; tile < t_end; ++tile) {
    while (tile >= bts[b + 1]) ++b;
    const double *sl_lo, *sl_hi;
    if (cur0 == b && cur1 == b + 1) {
      sl_lo = slab0; sl_hi = slab1;
    } else if (cur1 == b && cur0 == b + 1) {
      sl_lo = slab1; sl_hi = slab0;
    } else {
      __syncthreads();  // every warp is done with the previous bin's slabs
      if (cur0 == b) {
        stage_b(slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur1 = b + 1;
      } else if (cur1 == b) {
        stage_b(slab0, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur0 = b + 1;
      } else if (cur0 == b + 1) {
        stage_b(slab1, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur1 = b;
      } else if (cur1 == b + 1) {
        stage_b(slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur0 = b;
      } else {
        stage_b(slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur0 = b;
        stage_b(slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP0, 8 * RT, tid, NTHR); cur1 = b + 1;
      }
      __syncthreads();
      if (cur0 == b) { sl_lo = slab0; sl_hi = slab1; } else { sl_lo = slab1; sl_hi = slab0; }
    }

    const int row0 = bst[b] + (tile - bts[b]) * ROWS_CTA + warp * 16;
    const int nvalid = min(16, bst[b + 1] - row0);
    if (nvalid <= 0) continue;  // warp-uniform; no barrier below

    // ---- rows of this warp: g and g+8 of its 16-sample tile --------------------------------------
    const bool vA = g < nvalid, vB = (g + 8) < nvalid;
    const int mA = a.perm[row0 + (vA ? g : 0)], mB = a.perm[row0 + (vB ? g + 8 : 0)];
    const double w1A = a.w1[mA], w2A = a.w2[mA], w1B = a.w1[mB], w2B = a.w2[mB];
    const double qA = a.q[mA], qB = a.q[mB];
    double *FA = a.F + (size_t)mA * a.ldf + 2 * t, *FB = a.F + (size_t)mB * a.ldf + 2 * t;
    double2 fa[RT], fb[RT];
#pragma unroll
    for (int j = 0; j < RT; j++)
      if (EXACT || j < ks0) {
        fa[j] = *reinterpret_cast<const double2 *>(FA + 8 * j);
        fb[j] = *reinterpret_cast<const double2 *>(FB + 8 * j);
      }

    // ---- (1) interface update ---------------------------------------------------------------------
    double acc[2][RT][2];
#pragma unroll
    for (int j = 0; j < RT; j++) { acc[0][j][0] = acc[0][j][1] = acc[1][j][0] = acc[1][j][1] = 0.0; }
#pragma unroll
    for (int j = 0; j < RT; j++) {
      if (EXACT || j < ks0) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double xa = e ? fa[j].y : fa[j].x, xb = e ? fb[j].y : fb[j].x;
          const double a1A = w1A * xa, a2A = w2A * xa, a1B = w1B * xb, a2B = w2B * xb;
          const int sw = (((2 * j + e) ^ (g & 3)) << 2) | t;
#pragma unroll
          for (int jj = 0; jj < RT; jj++) {
            if (EXACT || jj < rt_act) {
              const int off = (8 * jj + g) * KP0 + sw;
              const double b1 = sl_lo[off], b2 = sl_hi[off];
              dmma884(acc[0][jj][0], acc[0][jj][1], a1A, b1);
              dmma884(acc[1][jj][0], acc[1][jj][1], a1B, b1);
              dmma884(acc[0][jj][0], acc[0][jj][1], a2A, b2);
              dmma884(acc[1][jj][0], acc[1][jj][1], a2B, b2);
            }
          }
        }
      }
    }
    if (!a.last) {
#pragma unroll
      for (int jj = 0; jj < RT; jj++)
        if (EXACT || jj < rt_act) {
          if (vA) *reinterpret_cast<double2 *>(FA + 8 * jj) = make_double2(acc[0][jj][0], acc[0][jj][1]);
          if (vB) *reinterpret_cast<double2 *>(FB + 8 * jj) = make_double2(acc[1][jj][0], acc[1][jj][1]);
        }
    }

    // ---- (2) conditional pdf on the grid of dimension k+1 ----------------------------------------
    double c[2][NT + 1][2];
#pragma unroll
    for (int j = 0; j <= NT; j++) { c[0][j][0] = c[0][j][1] = c[1][j][0] = c[1][j][1] = 0.0; }
#pragma unroll
    for (int jj = 0; jj < RT; jj++) {
      if (EXACT || jj < rt_act) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double a0 = acc[0][jj][e], a1 = acc[1][jj][e];
          const int sw = (((2 * jj + e) ^ (g & 3)) << 2) | t;
#pragma unroll
          for (int jn = 0; jn < NTD; jn++) {
            if (EXACT || jn < nt_act) {
              const double bv = Ps[(8 * jn + g) * KP1 + sw];
              dmma884(c[0][jn][0], c[0][jn][1], a0, bv);
              dmma884(c[1][jn][0], c[1][jn][1], a1, bv);
            }
          }
        }
      }
    }

    if (TAIL1) {
      // last grid column (node 8*(NT-1)): quad-distributed dot product, result kept on lane t == 0
      double tA = 0.0, tB = 0.0;
#pragma unroll
      for (int jj = 0; jj < RT; jj++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double pv = Ps[(8 * (NT - 1)) * KP1 + ((2 * jj + e) << 2) + t];
          tA = fma(acc[0][jj][e], pv, tA);
          tB = fma(acc[1][jj][e], pv, tB);
        }
      }
      tA += __shfl_xor_sync(FULL, tA, 1); tA += __shfl_xor_sync(FULL, tA, 2);
      tB += __shfl_xor_sync(FULL, tB, 1); tB += __shfl_xor_sync(FULL, tB, 2);
      c[0][NT - 1][0] = (t == 0) ? tA : 0.0;
      c[1][NT - 1][0] = (t == 0) ? tB : 0.0;
    }

    // ---- (3) CDF, search: both 8-row tiles, all lanes ----------------------------------------------
    double cdf_lo[2], c1v[2], c2v[2];
    int i0v[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const double qv = i ? qB : qA;
#pragma unroll
      for (int jn = 0; jn < NT; jn++) { c[i][jn][0] = fabs(c[i][jn][0]); c[i][jn][1] = fabs(c[i][jn][1]); }
      double S0[NT], S1[NT];
      double carry = 0.0;
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        S0[jn] = 0.0; S1[jn] = 0.0;
        if (TAIL1 && jn == NT - 1) continue;  // the lone last node starts no cell and is never a search candidate
        if (EXACT || jn < nt_act) {
          const double p0 = c[i][jn][0], p1 = c[i][jn][1];
          const double var = (t == 0) ? c[i][jn + 1][0] : p0;
          const double nxt = __shfl_sync(FULL, var, (t == 3) ? lane - 3 : lane + 1);
          const double Ta = hh[8 * jn + 2 * t] * (p0 + p1), Tb = hh[8 * jn + 2 * t + 1] * (p1 + nxt);
          double incl = Ta + Tb;
          double v = __shfl_up_sync(FULL, incl, 1, 4);
          if (t >= 1) incl += v;
          v = __shfl_up_sync(FULL, incl, 2, 4);
          if (t >= 2) incl += v;
          double excl = __shfl_up_sync(FULL, incl, 1, 4);
          if (t == 0) excl = 0.0;
          const double tot = __shfl_sync(FULL, incl, 3, 4);
          S0[jn] = carry + excl;
          S1[jn] = S0[jn] + Ta;
          carry += tot;
        }
      }
      const double total = carry;
      const double sc = 1.0 / total;
      int cnt = 0;
#pragma unroll
      for (int jn = 0; jn < NTD; jn++) {
        if (EXACT || jn < nt_act) {
          const int node0 = 8 * jn + 2 * t;
          cnt += (node0 >= 1 && node0 <= n1 - 2 && qv > S0[jn] * sc) ? 1 : 0;
          cnt += (node0 + 1 <= n1 - 2 && qv > S1[jn] * sc) ? 1 : 0;
        }
      }
      cnt += __shfl_xor_sync(FULL, cnt, 1);
      cnt += __shfl_xor_sync(FULL, cnt, 2);
      const int i0 = cnt, i1 = cnt + 1;
      const int js = i0 >> 3, ts = (i0 & 7) >> 1, es = i0 & 1;
      const int jq = i1 >> 3, tq = (i1 & 7) >> 1, eq = i1 & 1;
      double selS = 0.0, selP = 0.0, selQ = 0.0;
#pragma unroll
      for (int jn = 0; jn < NT; jn++) {
        if (jn == js) { selS = es ? S1[jn] : S0[jn]; selP = es ? c[i][jn][1] : c[i][jn][0]; }
        if (jn == jq) { selQ = eq ? c[i][jn][1] : c[i][jn][0]; }
      }
      const int base = lane & ~3;
      selS = __shfl_sync(FULL, selS, base | ts);
      selP = __shfl_sync(FULL, selP, base | ts);
      selQ = __shfl_sync(FULL, selQ, base | tq);
      cdf_lo[i] = selS * sc; c1v[i] = selP * sc; c2v[i] = selQ * sc; i0v[i] = i0;
      if (total == 0.0) {
        // zero-mass conditional: uniform in index space (reference tt_irt1_int32.c:116-125)
        const double u = 1.0 / (double)(n1 - 1);
        const double s2 = 1.0 / ((double)(n1 - 1) * u);
        int k0 = 0;
        for (int j = 1; j <= n1 - 2; j++) k0 += (qv > ((double)j * u) * s2) ? 1 : 0;
        i0v[i] = k0; cdf_lo[i] = ((double)k0 * u) * s2; c1v[i] = u * s2; c2v[i] = u * s2;
      }
    }

    // ---- inversion tail: lane t=0 of a quad finishes row g, lane t=1 row g+8 ---------------------
    {
      const int sel = t & 1;
      const bool valid = (t < 2) && (sel ? vB : vA);
      const int m = sel ? mB : mA;
      const int i0 = sel ? i0v[1] : i0v[0];
      const CellOut o = invert_cell(sel ? qB : qA, sel ? cdf_lo[1] : cdf_lo[0], sel ? c1v[1] : c1v[0],
                                    sel ? c2v[1] : c2v[0], xg[i0], xg[i0 + 1]);
      if (valid) {
        a.z[m] = o.xk;
        if (a.idx_out) a.idx_out[m] = i0;
        if (!a.last) {
          a.idx[m] = i0; a.w1[m] = o.w1; a.w2[m] = o.w2;
          a.lp[m] = a.lp[m] + o.logp;
          atomicAdd(&hist[i0], 1);
        } else {
          a.lpz[m] = a.lp[m] + o.logp;
        }
      }
    }
  }

  if (!a.last) {
    __syncthreads();
    for (int i = tid; i < n1 - 1; i += NTHR)
      if (hist[i]) atomicAdd(a.hist_next + i, hist[i]);
  }
}

template <int RT, int NT, int WARPS, bool EXACT, bool TAIL1>
cudaError_t launch_variant(const TransArgs &a, int sm_count, cudaStream_t st) {
  using L = SmemLayout<RT, NT>;
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, transition_kernel<RT, NT, WARPS, EXACT, TAIL1>, WARPS * 32, L::bytes);
    if (e != cudaSuccess) return e;
    occ = o > 0 ? o : 1;
  }
  const int rows_cta = WARPS * 16;
  int64_t max_tiles = ((int64_t)a.rows + rows_cta - 1) / rows_cta + (a.n0 - 1);
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  transition_kernel<RT, NT, WARPS, EXACT, TAIL1><<<(unsigned)grid, WARPS * 32, L::bytes, st>>>(a);
  return cudaGetLastError();
}

template <int RT, int NT, int WARPS>
cudaError_t launch_one(const TransArgs &a, int sm_count, cudaStream_t st) {
  const bool exact = a.r0 == 8 * RT && a.r1 == 8 * RT && (a.n1 + 7) / 8 == NT;
  if (exact && a.n1 == 8 * (NT - 1) + 1) return launch_variant<RT, NT, WARPS, true, true>(a, sm_count, st);
  if (exact) return launch_variant<RT, NT, WARPS, true, false>(a, sm_count, st);
  return launch_variant<RT, NT, WARPS, false, false>(a, sm_count, st);
}

template <int RT, int NT, int WARPS>
cudaError_t init_one() {
  using L = SmemLayout<RT, NT>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
}

constexpr int kWarps = 8;

}  // namespace

int fast_class_for(int rmax, int nmax) {
  if (rmax <= 16 && nmax <= 24) return 0;
  if (rmax <= 32 && nmax <= 40) return 1;
  if (rmax <= 64 && nmax <= 72) return 2;
  return -1;
}

int fast_rows_per_cta(int) { return kWarps * 16; }

cudaError_t fast_init(int) {
  cudaError_t e;
  if ((e = init_one<2, 3, kWarps>()) != cudaSuccess) return e;
  if ((e = init_one<4, 5, kWarps>()) != cudaSuccess) return e;
  if ((e = init_one<8, 9, kWarps>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_transition(int cls, const TransArgs &a, int sm_count, cudaStream_t st) {
  switch (cls) {
    case 0: return launch_one<2, 3, kWarps>(a, sm_count, st);
    case 1: return launch_one<4, 5, kWarps>(a, sm_count, st);
    case 2: return launch_one<8, 9, kWarps>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ttirt
