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

// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a): global rows -> shared memory, completion by transaction bytes ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// Shared-memory image of a B operand (K x NC, column-major in global memory with column stride cs).
// Column c lives at c*KP.  Rows keep their natural order inside groups of 8 and the groups are XOR-swizzled
// with the column's parity: a thread quad reads, per pair of k-steps, 16 bytes per lane (LDS.128: rows
// 8j+2t, 8j+2t+1 of column 8jj+g) and the two columns served by one quarter-warp wavefront fall into
// different 64-byte halves of the bank window, so the loads are conflict-free without padding.
// The contraction index is consumed in the permuted order (8j+2t+e for k-step 2j+e, lane t) on BOTH MMA
// operands, which is what lets phase (2) use phase (1)'s accumulators as its A fragments in place.
// Entries outside K x NC are zero.
__device__ __forceinline__ int b_phys(int c, int a, int KP) { return c * KP + ((((a >> 3) ^ (c & 1)) << 3) | (a & 7)); }

__device__ void stage_b(double *dst, const double *__restrict__ src, int K, int NC, int64_t cs, int KP, int NCP,
                        int tid, int nthr) {
  const int total = NCP * KP;
  for (int e = tid; e < total; e += nthr) {
    const int c = e / KP, a = e - c * KP;
    const double v = (a < K && c < NC) ? __ldg(src + a + (int64_t)c * cs) : 0.0;
    dst[b_phys(c, a, KP)] = v;
  }
}

template <int RT, int NT, int WARPS>
struct SmemLayout {
  static constexpr int KPMAX = (8 * RT + 15) & ~15;
  static constexpr int SLAB = 8 * RT * KPMAX;      // doubles per slab buffer
  static constexpr int PN = 8 * NT * KPMAX;        // doubles for P_{k+1}
  static constexpr int NBMAX = 8 * NT;             // >= n - 1 intervals
  static constexpr int FPITCH = 8 * RT + 8;        // doubles per staged left-interface row (+64 B: rows g, g+1 hit different bank halves)
  static constexpr int FTILE = 16 * FPITCH;        // doubles per warp
  static constexpr size_t bytes = sizeof(double) * (2 * SLAB + PN + 3 * 8 * NT + WARPS * FTILE) + sizeof(uint64_t) * WARPS +
                                  sizeof(int) * (2 * (NBMAX + 1) + NBMAX);
};

// EXACT: r0 == r1 == 8*RT and ceil(n1/8) == NT, so every tile loop runs its full static trip count and no
// guard branches are compiled in (the steady state of a uniform-rank TT).  TAIL1 (EXACT only): n1 == 8*(NT-1)+1,
// the usual 2^p+1 grid; the lone last grid column is then a 4-lane DFMA dot product instead of a whole DMMA
// column tile that would be 7/8 padding.
//
// Row gather: the 16 left-interface rows of a warp's NEXT tile are fetched by TMA bulk copies (cp.async.bulk,
// one 8*r0-byte row per lane, completion counted on a per-warp mbarrier) into a per-warp shared tile as soon
// as the current tile's update phase has consumed that tile, i.e. a whole pdf + inversion phase ahead of use.
// The dependent perm -> row latency (two DRAM round trips) therefore never sits in front of the DMMA stream.
template <int RT, int NT, int WARPS, bool EXACT, bool TAIL1>
__global__ void __launch_bounds__(WARPS * 32, 1) transition_kernel(const TransArgs a) {
  static_assert(EXACT || !TAIL1, "TAIL1 needs EXACT");
  using L = SmemLayout<RT, NT, WARPS>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *slab0 = reinterpret_cast<double *>(smem_raw);
  double *slab1 = slab0 + L::SLAB;
  double *Ps = slab1 + L::SLAB;
  double *ft_all = Ps + L::PN;       // per-warp staged left-interface rows
  double *hh = ft_all + WARPS * L::FTILE;  // half grid steps of dimension k+1, zero beyond n1-2
  double *xg = hh + 8 * NT;          // grid of dimension k+1
  double *ihs = xg + 8 * NT;         // reciprocal cell widths of dimension k+1
  uint64_t *bars = reinterpret_cast<uint64_t *>(ihs + 8 * NT);
  int *bts = reinterpret_cast<int *>(bars + WARPS);  // bin -> first CTA tile
  int *bst = bts + (L::NBMAX + 1);                    // bin -> first sorted row
  int *hist = bst + (L::NBMAX + 1);                   // histogram of the intervals chosen in dimension k+1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NTHR = WARPS * 32, ROWS_CTA = WARPS * 16;
  constexpr int FP = L::FPITCH;
  double *ft = ft_all + warp * L::FTILE;
  uint64_t *bar = bars + warp;

  const int r0 = a.r0, r1 = a.r1, n1 = a.n1, nb0 = a.n0 - 1;
  constexpr int KP = L::KPMAX;                   // compile-time column pitch: B-fragment offsets fold into immediates
  constexpr int NTD = TAIL1 ? NT - 1 : NT;       // grid column tiles computed by DMMA
  const int ks0 = EXACT ? RT : (r0 + 7) >> 3, rt_act = EXACT ? RT : (r1 + 7) >> 3, nt_act = EXACT ? NT : (n1 + 7) >> 3;
  const uint32_t row_bytes = (uint32_t)(8 * ks0) * 8u;   // bytes of a left-interface row that the update reads

  for (int i = tid; i <= nb0; i += NTHR) {
    bts[i] = a.bin_tile_start[i];
    bst[i] = a.bin_start[i];
  }
  for (int i = tid; i < 8 * NT; i += NTHR) {
    hh[i] = (i + 1 < n1) ? 0.5 * (a.xnext[i + 1] - a.xnext[i]) : 0.0;
    xg[i] = (i < n1) ? a.xnext[i] : 0.0;
    ihs[i] = (i + 1 < n1) ? 1.0 / (a.xnext[i + 1] - a.xnext[i]) : 0.0;
    if (i < L::NBMAX) hist[i] = 0;
  }
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  stage_b(Ps, a.pnext, r1, n1, r1, KP, 8 * NT, tid, NTHR);
  __syncthreads();

  const int total_tiles = bts[nb0];
  const int t_begin = (int)(((int64_t)blockIdx.x * total_tiles) / gridDim.x);
  const int t_end = (int)(((int64_t)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int64_t slab_cs = (int64_t)r0 * a.n0;

  int cur0 = -1, cur1 = -1;  // interval slab held by slab0 / slab1
  int b = 0;                 // bin of the current tile
  int bh = 0;                // bin hint of the row look-ahead (monotone)
  const double *sl_lo = slab0, *sl_hi = slab1;

  // rows of this warp in CTA tile `tile`: first sorted row and number of valid rows (0 past the end)
  auto rows_of = [&](int tile, int &nv) -> int {
    if (tile >= t_end) { nv = 0; return 0; }
    while (tile >= bts[bh + 1]) ++bh;
    const int row0 = bst[bh] + (tile - bts[bh]) * ROWS_CTA + warp * 16;
    nv = max(0, min(16, bst[bh + 1] - row0));
    return row0;
  };
  // sample ids of a tile: lane l (and l+16) holds the id of row l & 15; rows past nv repeat row 0
  auto load_ids = [&](int row0, int nv) -> int { return nv > 0 ? a.perm[row0 + ((lane & 15) < nv ? (lane & 15) : 0)] : 0; };

  uint32_t phase = 0;
  int nvC = 0, nvN = 0;      // valid rows: current tile / next tile
  int idC = 0, idN = 0;      // row ids (lane-distributed) of the current / next tile
  // TMA gather of the rows whose ids are idN into this warp's shared tile
  auto issue_gather = [&](int ids, int nv) {
    if (nv > 0) {
      if (lane == 0) mbar_arrive_expect_tx(bar, 16u * row_bytes);
      __syncwarp();
      if (lane < 16) bulk_g2s(ft + lane * FP, a.F + (size_t)ids * a.ldf, row_bytes, bar);
    }
  };
  double w1A = 0, w2A = 0, w1B = 0, w2B = 0, qA = 0, qB = 0, lpA = 0, lpB = 0;       // current tile
  double w1An = 0, w2An = 0, w1Bn = 0, w2Bn = 0, qAn = 0, qBn = 0, lpAn = 0, lpBn = 0;  // next tile (in flight)
  int mA = 0, mB = 0, mAn = 0, mBn = 0;
  {
    const int rowC = rows_of(t_begin, nvC);
    idC = load_ids(rowC, nvC);
    const int rowN = rows_of(t_begin + 1, nvN);
    idN = load_ids(rowN, nvN);
    issue_gather(idC, nvC);
    mA = __shfl_sync(FULL, idC, g); mB = __shfl_sync(FULL, idC, g + 8);
    if (nvC > 0) {
      w1A = a.w1[mA]; w2A = a.w2[mA]; w1B = a.w1[mB]; w2B = a.w2[mB];
      qA = a.q[mA]; qB = a.q[mB]; lpA = a.lp[mA]; lpB = a.lp[mB];
    }
  }

  // Warps w and w+4 share an SM sub-partition (one FP64/DMMA pipe).  Started together and sharing the pipe
  // fairly they stay in lock-step, so both sit in the latency-bound inversion phase at the same time with the
  // pipe idle; a start offset of about half a tile (preserved by the same fairness) interleaves their phases.
  if (warp >= WARPS / 2 && a.stagger_ns) __nanosleep(a.stagger_ns);

  for (int tile = t_begin; tile < t_end; ++tile) {
    while (tile >= bts[b + 1]) ++b;
    if (cur0 == b && cur1 == b + 1) {
      sl_lo = slab0; sl_hi = slab1;
    } else if (cur1 == b && cur0 == b + 1) {
      sl_lo = slab1; sl_hi = slab0;
    } else {
      __syncthreads();  // every warp is done with the previous bin's slabs
      if (cur0 == b) {
        stage_b(slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur1 = b + 1;
      } else if (cur1 == b) {
        stage_b(slab0, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur0 = b + 1;
      } else if (cur0 == b + 1) {
        stage_b(slab1, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur1 = b;
      } else if (cur1 == b + 1) {
        stage_b(slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur0 = b;
      } else {
        stage_b(slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur0 = b;
        stage_b(slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, tid, NTHR); cur1 = b + 1;
      }
      __syncthreads();
      if (cur0 == b) { sl_lo = slab0; sl_hi = slab1; } else { sl_lo = slab1; sl_hi = slab0; }
      if (warp >= WARPS / 2 && a.stagger_ns) __nanosleep(a.stagger_ns);
    }

    const int nvalid = nvC;
    const bool vA = g < nvalid, vB = (g + 8) < nvalid;
    double c[2][NT + 1][2];
#pragma unroll
    for (int j = 0; j <= NT; j++) { c[0][j][0] = c[0][j][1] = c[1][j][0] = c[1][j][1] = 0.0; }

    if (nvalid > 0) {
      // ---- (1) interface update: A fragments from the TMA-staged rows ------------------------------
      while (!mbar_try_wait(bar, phase)) {}
      phase ^= 1;
      double acc[2][RT][2];
#pragma unroll
      for (int j = 0; j < RT; j++) { acc[0][j][0] = acc[0][j][1] = acc[1][j][0] = acc[1][j][1] = 0.0; }
      const double *fra = ft + g * FP + 2 * t, *frb = ft + (g + 8) * FP + 2 * t;
#pragma unroll
      for (int j = 0; j < RT; j++) {
        if (EXACT || j < ks0) {
          const double2 fa = *reinterpret_cast<const double2 *>(fra + 8 * j);
          const double2 fb = *reinterpret_cast<const double2 *>(frb + 8 * j);
          const double a1A0 = w1A * fa.x, a2A0 = w2A * fa.x, a1B0 = w1B * fb.x, a2B0 = w2B * fb.x;
          const double a1A1 = w1A * fa.y, a2A1 = w2A * fa.y, a1B1 = w1B * fb.y, a2B1 = w2B * fb.y;
          const int sw = ((j ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
          for (int jj = 0; jj < RT; jj++) {
            if (EXACT || jj < rt_act) {
              const int off = (8 * jj + g) * KP + sw;
              const double2 b1 = *reinterpret_cast<const double2 *>(sl_lo + off);
              const double2 b2 = *reinterpret_cast<const double2 *>(sl_hi + off);
              dmma884(acc[0][jj][0], acc[0][jj][1], a1A0, b1.x);
              dmma884(acc[1][jj][0], acc[1][jj][1], a1B0, b1.x);
              dmma884(acc[0][jj][0], acc[0][jj][1], a2A0, b2.x);
              dmma884(acc[1][jj][0], acc[1][jj][1], a2B0, b2.x);
              dmma884(acc[0][jj][0], acc[0][jj][1], a1A1, b1.y);
              dmma884(acc[1][jj][0], acc[1][jj][1], a1B1, b1.y);
              dmma884(acc[0][jj][0], acc[0][jj][1], a2A1, b2.y);
              dmma884(acc[1][jj][0], acc[1][jj][1], a2B1, b2.y);
            }
          }
        }
      }
      // ---- the staged tile is consumed: launch the next tile's gather and scalar loads now ---------
      __syncwarp();
      issue_gather(idN, nvN);
      mAn = __shfl_sync(FULL, idN, g); mBn = __shfl_sync(FULL, idN, g + 8);
      if (nvN > 0) {
        w1An = a.w1[mAn]; w2An = a.w2[mAn]; w1Bn = a.w1[mBn]; w2Bn = a.w2[mBn];
        qAn = a.q[mAn]; qBn = a.q[mBn]; lpAn = a.lp[mAn]; lpBn = a.lp[mBn];
      }
      if (!a.last) {
        double *FA = a.F + (size_t)mA * a.ldf + 2 * t, *FB = a.F + (size_t)mB * a.ldf + 2 * t;
#pragma unroll
        for (int jj = 0; jj < RT; jj++)
          if (EXACT || jj < rt_act) {
            if (vA) *reinterpret_cast<double2 *>(FA + 8 * jj) = make_double2(acc[0][jj][0], acc[0][jj][1]);
            if (vB) *reinterpret_cast<double2 *>(FB + 8 * jj) = make_double2(acc[1][jj][0], acc[1][jj][1]);
          }
      }
      // ---- (2) conditional pdf on the grid of dimension k+1 ----------------------------------------
#pragma unroll
      for (int jj = 0; jj < RT; jj++) {
        if (EXACT || jj < rt_act) {
          const int sw = ((jj ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
          for (int jn = 0; jn < NTD; jn++) {
            if (EXACT || jn < nt_act) {
              const double2 bv = *reinterpret_cast<const double2 *>(Ps + (8 * jn + g) * KP + sw);
              dmma884(c[0][jn][0], c[0][jn][1], acc[0][jj][0], bv.x);
              dmma884(c[1][jn][0], c[1][jn][1], acc[1][jj][0], bv.x);
              dmma884(c[0][jn][0], c[0][jn][1], acc[0][jj][1], bv.y);
              dmma884(c[1][jn][0], c[1][jn][1], acc[1][jj][1], bv.y);
            }
          }
        }
      }
      if (TAIL1) {
        // last grid column (node 8*(NT-1), even column: no swizzle): quad-distributed dot product
        double tA = 0.0, tB = 0.0;
#pragma unroll
        for (int jj = 0; jj < RT; jj++) {
          const double2 pv = *reinterpret_cast<const double2 *>(Ps + (8 * (NT - 1)) * KP + (jj << 3) + (t << 1));
          tA = fma(acc[0][jj][0], pv.x, tA); tB = fma(acc[1][jj][0], pv.x, tB);
          tA = fma(acc[0][jj][1], pv.y, tA); tB = fma(acc[1][jj][1], pv.y, tB);
        }
        tA += __shfl_xor_sync(FULL, tA, 1); tA += __shfl_xor_sync(FULL, tA, 2);
        tB += __shfl_xor_sync(FULL, tB, 1); tB += __shfl_xor_sync(FULL, tB, 2);
        c[0][NT - 1][0] = (t == 0) ? tA : 0.0;
        c[1][NT - 1][0] = (t == 0) ? tB : 0.0;
      }
    } else {
      // this warp has no rows in this tile: keep the pipeline moving
      issue_gather(idN, nvN);
      mAn = __shfl_sync(FULL, idN, g); mBn = __shfl_sync(FULL, idN, g + 8);
      if (nvN > 0) {
        w1An = a.w1[mAn]; w2An = a.w2[mAn]; w1Bn = a.w1[mBn]; w2Bn = a.w2[mBn];
        qAn = a.q[mAn]; qBn = a.q[mBn]; lpAn = a.lp[mAn]; lpBn = a.lp[mBn];
      }
    }
    // ids of the tile after next
    int nvNN = 0;
    const int rowNN = rows_of(tile + 2, nvNN);
    const int idNN = load_ids(rowNN, nvNN);

    if (nvalid > 0) {
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
      const double qt = qv * total;   // q > S/total  <=>  q*total > S up to one rounding: decided on the unnormalised CDF
      int cnt = 0;
#pragma unroll
      for (int jn = 0; jn < NTD; jn++) {
        if (EXACT || jn < nt_act) {
          const int node0 = 8 * jn + 2 * t;
          cnt += (node0 >= 1 && node0 <= n1 - 2 && qt > S0[jn]) ? 1 : 0;
          cnt += (node0 + 1 <= n1 - 2 && qt > S1[jn]) ? 1 : 0;
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
      const CellOut o = invert_cell_fast(sel ? qB : qA, sel ? cdf_lo[1] : cdf_lo[0], sel ? c1v[1] : c1v[0],
                                         sel ? c2v[1] : c2v[0], xg[i0], xg[i0 + 1], ihs[i0]);
      if (valid) {
        a.z[m] = o.xk;
        if (a.idx_out) a.idx_out[m] = i0;
        if (!a.last) {
          a.idx[m] = i0; a.w1[m] = o.w1; a.w2[m] = o.w2;
          a.lp[m] = (sel ? lpB : lpA) + o.logp;
          atomicAdd(&hist[i0], 1);
        } else {
          a.lpz[m] = (sel ? lpB : lpA) + o.logp;
        }
      }
    }
    }  // nvalid > 0

    // ---- rotate the row pipeline --------------------------------------------------------------------
    nvC = nvN; idC = idN; mA = mAn; mB = mBn;
    w1A = w1An; w2A = w2An; w1B = w1Bn; w2B = w2Bn; qA = qAn; qB = qBn; lpA = lpAn; lpB = lpBn;
    nvN = nvNN; idN = idNN;
  }

  if (!a.last) {
    __syncthreads();
    for (int i = tid; i < n1 - 1; i += NTHR)
      if (hist[i]) atomicAdd(a.hist_next + i, hist[i]);
  }
}

template <int RT, int NT, int WARPS, bool EXACT, bool TAIL1>
cudaError_t launch_variant(const TransArgs &a, int sm_count, cudaStream_t st) {
  using L = SmemLayout<RT, NT, WARPS>;
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
  using L = SmemLayout<RT, NT, WARPS>;
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
