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

#ifndef TTIRT_WARPS
#define TTIRT_WARPS 8    // warps per CTA (one CTA per SM)
#endif
#ifndef TTIRT_MT
#define TTIRT_MT 2       // 8-row MMA tiles per warp
#endif

namespace ttirt {

#ifdef TTIRT_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8 + 32];
#define PT_DECL unsigned long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_t = clock64(); const long long pt_start = pt_t;
#define PT_MARK(k) { const long long n_ = clock64(); pt_[k] += (unsigned long long)(n_ - pt_t); pt_t = n_; }
#define PT_FLUSH if (lane == 0) { for (int k_ = 0; k_ < 8; k_++) atomicAdd(&g_phase_cycles[k_], pt_[k_]); atomicAdd(&g_phase_cycles[8 + warp], (unsigned long long)(clock64() - pt_start)); }
#else
#define PT_DECL
#define PT_MARK(k)
#define PT_FLUSH
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- cp.async (LDGSTS): 16 bytes per lane, global -> shared, per-thread completion groups ---------------------
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

template <int RT, int NT, int WARPS, int MT>
struct SmemLayout {
  static constexpr int KPMAX = (8 * RT + 15) & ~15;
  static constexpr int SLAB = 8 * RT * KPMAX;      // doubles per slab buffer
  static constexpr int PN = 8 * NT * KPMAX;        // doubles for P_{k+1}
  static constexpr int NBMAX = 8 * NT;             // >= n - 1 intervals
  static constexpr int FPITCH = 8 * RT + 8;        // doubles per staged left-interface row (+64 B: rows g, g+1 hit different bank halves)
  static constexpr int FTILE = 8 * MT * FPITCH;    // doubles per warp
  static constexpr size_t bytes = sizeof(double) * (2 * SLAB + PN + 3 * 8 * NT + WARPS * FTILE) +
                                  sizeof(int) * (2 * (NBMAX + 1) + NBMAX);
};

// EXACT: r0 == r1 == 8*RT and ceil(n1/8) == NT, so every tile loop runs its full static trip count and no
// guard branches are compiled in (the steady state of a uniform-rank TT).  TAIL1 (EXACT only): n1 == 8*(NT-1)+1,
// the usual 2^p+1 grid; the lone last grid column is then a 4-lane DFMA dot product instead of a whole DMMA
// column tile that would be 7/8 padding.
//
// MT: 8-row MMA tiles per warp.  A warp owns 8*MT samples end to end.  MT = 1 with 16 warps (128 registers per
// thread) puts four warps on every SM sub-partition: while one warp walks the latency-bound CDF / inversion
// phase three others feed the DMMA pipe.  MT = 2 with 8 warps halves the B-operand shared-memory traffic per
// DMMA but leaves only two warps per sub-partition.
//
// Row gather: the left-interface rows of a warp's NEXT tile are fetched asynchronously (cp.async / LDGSTS, one
// coalesced 8*r0-byte row per instruction) into a per-warp shared tile as soon as the current tile's update
// phase has consumed that tile, i.e. a whole pdf + inversion phase ahead of use.  The dependent perm -> row
// latency (two DRAM round trips) therefore never sits in front of the DMMA stream.
template <int RT, int NT, int WARPS, int MT, bool EXACT, bool TAIL1>
__global__ void __launch_bounds__(WARPS * 32, 1) transition_kernel(const TransArgs a) {
  static_assert(EXACT || !TAIL1, "TAIL1 needs EXACT");
  static_assert(MT == 1 || MT == 2, "one or two 8-row tiles per warp");
  using L = SmemLayout<RT, NT, WARPS, MT>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *slab0 = reinterpret_cast<double *>(smem_raw);
  double *slab1 = slab0 + L::SLAB;
  double *Ps = slab1 + L::SLAB;
  double *ft_all = Ps + L::PN;       // per-warp staged left-interface rows
  double *hh = ft_all + WARPS * L::FTILE;  // half grid steps of dimension k+1, zero beyond n1-2
  double *xg = hh + 8 * NT;          // grid of dimension k+1
  double *ihs = xg + 8 * NT;         // reciprocal cell widths of dimension k+1
  int *bts = reinterpret_cast<int *>(ihs + 8 * NT);  // bin -> first CTA tile
  int *bst = bts + (L::NBMAX + 1);                    // bin -> first sorted row
  int *hist = bst + (L::NBMAX + 1);                   // histogram of the intervals chosen in dimension k+1

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NTHR = WARPS * 32, WROWS = 8 * MT, ROWS_CTA = WARPS * WROWS;
  constexpr int FP = L::FPITCH;
  double *ft = ft_all + warp * L::FTILE;

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
    const int row0 = bst[bh] + (tile - bts[bh]) * ROWS_CTA + warp * WROWS;
    nv = max(0, min(WROWS, bst[bh + 1] - row0));
    return row0;
  };
  // sample ids of a tile: lane l holds the id of row l % WROWS; rows past nv repeat row 0
  auto load_ids = [&](int row0, int nv) -> int {
    const int r = lane & (WROWS - 1);
    return nv > 0 ? a.perm[row0 + (r < nv ? r : 0)] : 0;
  };

  int nvC = 0, nvN = 0;      // valid rows: current tile / next tile
  int idN = 0;               // row ids (lane-distributed) of the next tile
  // Gather of the rows whose ids are `ids` into this warp's shared tile: one cp.async (LDGSTS) instruction per row,
  // 16 bytes per lane, so a row is one coalesced 8*r0-byte segment.  (Sixteen single-row cp.async.bulk copies per
  // tile were measured at ~1000 cycles of issue time per tile; this costs ~150.)
  auto issue_gather = [&](int ids, int nv) {
    if (nv > 0) {
#pragma unroll
      for (int r = 0; r < WROWS; r++) {
        const int id = __shfl_sync(FULL, ids, r);
        if (16u * lane < row_bytes)
          cp_async16(reinterpret_cast<char *>(ft + r * FP) + 16 * lane, reinterpret_cast<const char *>(a.F + (size_t)id * a.ldf) + 16 * lane);
      }
      cp_async_commit();
    }
  };
  // per-row scalars, m-tile i = rows 8i+g: current tile and next tile (in flight)
  int mrow[MT], mrow_n[MT];
  double w1[MT], w2[MT], qv_[MT], lp_[MT], w1n[MT], w2n[MT], qn[MT], lpn[MT];
#pragma unroll
  for (int i = 0; i < MT; i++) { mrow[i] = mrow_n[i] = 0; w1[i] = w2[i] = qv_[i] = lp_[i] = w1n[i] = w2n[i] = qn[i] = lpn[i] = 0.0; }
  auto load_scalars_next = [&]() {
#pragma unroll
    for (int i = 0; i < MT; i++) mrow_n[i] = __shfl_sync(FULL, idN, 8 * i + g);
    if (nvN > 0) {
#pragma unroll
      for (int i = 0; i < MT; i++) {
        w1n[i] = a.w1[mrow_n[i]]; w2n[i] = a.w2[mrow_n[i]]; qn[i] = a.q[mrow_n[i]]; lpn[i] = a.lp[mrow_n[i]];
      }
    }
  };
  {
    const int rowC = rows_of(t_begin, nvC);
    const int idC = load_ids(rowC, nvC);
    issue_gather(idC, nvC);
    idN = idC; nvN = nvC;
    load_scalars_next();
#pragma unroll
    for (int i = 0; i < MT; i++) { mrow[i] = mrow_n[i]; w1[i] = w1n[i]; w2[i] = w2n[i]; qv_[i] = qn[i]; lp_[i] = lpn[i]; }
    const int rowN = rows_of(t_begin + 1, nvN);
    idN = load_ids(rowN, nvN);
  }

  // optional start offset for a subset of the warps (breaks the lock-step of warps sharing a pipe)
  if (a.stagger_ns && (warp & a.stagger_mask)) __nanosleep(a.stagger_ns);
  PT_DECL
  for (int tile = t_begin; tile < t_end; ++tile) {
    PT_MARK(7)
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
    }

    PT_MARK(0)
    const int nvalid = nvC;
    double c[MT][NT + 1][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j <= NT; j++) c[i][j][0] = c[i][j][1] = 0.0;

    if (nvalid > 0) {
      // ---- (1) interface update: A fragments from the TMA-staged rows ------------------------------
      cp_async_wait_all();
      __syncwarp();
      PT_MARK(1)
      double acc[MT][RT][2];
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < RT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
      for (int j = 0; j < RT; j++) {
        if (EXACT || j < ks0) {
          double a1[MT][2], a2[MT][2];
#pragma unroll
          for (int i = 0; i < MT; i++) {
            const double2 f = *reinterpret_cast<const double2 *>(ft + (8 * i + g) * FP + 2 * t + 8 * j);
            a1[i][0] = w1[i] * f.x; a2[i][0] = w2[i] * f.x; a1[i][1] = w1[i] * f.y; a2[i][1] = w2[i] * f.y;
          }
          const int sw = ((j ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
          for (int jj = 0; jj < RT; jj++) {
            if (EXACT || jj < rt_act) {
              const int off = (8 * jj + g) * KP + sw;
              const double2 b1 = *reinterpret_cast<const double2 *>(sl_lo + off);
              const double2 b2 = *reinterpret_cast<const double2 *>(sl_hi + off);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(acc[i][jj][0], acc[i][jj][1], a1[i][0], b1.x);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(acc[i][jj][0], acc[i][jj][1], a2[i][0], b2.x);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(acc[i][jj][0], acc[i][jj][1], a1[i][1], b1.y);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(acc[i][jj][0], acc[i][jj][1], a2[i][1], b2.y);
            }
          }
        }
      }
      PT_MARK(2)
      // ---- the staged tile is consumed: launch the next tile's gather and scalar loads now ---------
      __syncwarp();
      issue_gather(idN, nvN);
      PT_MARK(7)
      load_scalars_next();
      // F' rows go out one 64-byte segment per pdf k-pair (below), so the stores trickle out under the DMMA stream
      double *Fo[MT];
#pragma unroll
      for (int i = 0; i < MT; i++) Fo[i] = a.F + (size_t)mrow[i] * a.ldf + 2 * t;
      const bool do_store = !a.last;
      PT_MARK(3)
      // ---- (2) conditional pdf on the grid of dimension k+1 ----------------------------------------
#pragma unroll
      for (int jj = 0; jj < RT; jj++) {
        if (EXACT || jj < rt_act) {
          const int sw = ((jj ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
          for (int jn = 0; jn < NTD; jn++) {
            if (EXACT || jn < nt_act) {
              const double2 bv = *reinterpret_cast<const double2 *>(Ps + (8 * jn + g) * KP + sw);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(c[i][jn][0], c[i][jn][1], acc[i][jj][0], bv.x);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(c[i][jn][0], c[i][jn][1], acc[i][jj][1], bv.y);
            }
          }
          if (do_store) {
#pragma unroll
            for (int i = 0; i < MT; i++)
              if (8 * i + g < nvalid) *reinterpret_cast<double2 *>(Fo[i] + 8 * jj) = make_double2(acc[i][jj][0], acc[i][jj][1]);
          }
        }
      }
      if (TAIL1) {
        // last grid column (node 8*(NT-1), even column: no swizzle): quad-distributed dot product
        double tl[MT];
#pragma unroll
        for (int i = 0; i < MT; i++) tl[i] = 0.0;
#pragma unroll
        for (int jj = 0; jj < RT; jj++) {
          const double2 pv = *reinterpret_cast<const double2 *>(Ps + (8 * (NT - 1)) * KP + (jj << 3) + (t << 1));
#pragma unroll
          for (int i = 0; i < MT; i++) { tl[i] = fma(acc[i][jj][0], pv.x, tl[i]); tl[i] = fma(acc[i][jj][1], pv.y, tl[i]); }
        }
#pragma unroll
        for (int i = 0; i < MT; i++) {
          tl[i] += __shfl_xor_sync(FULL, tl[i], 1); tl[i] += __shfl_xor_sync(FULL, tl[i], 2);
          c[i][NT - 1][0] = (t == 0) ? tl[i] : 0.0;
        }
      }
    } else {
      // this warp has no rows in this tile: keep the pipeline moving
      issue_gather(idN, nvN);
      load_scalars_next();
    }
    PT_MARK(4)
    // ids of the tile after next
    int nvNN = 0;
    const int rowNN = rows_of(tile + 2, nvNN);
    const int idNN = load_ids(rowNN, nvNN);

    if (nvalid > 0) {
      // ---- (3) CDF, search: all lanes; the MT row tiles are advanced together (independent chains interleave) ----
      double cdf_lo[MT], c1v[MT], c2v[MT];
      int i0v[MT];
      {
        double S0[MT][NT], S1[MT][NT], carry[MT];
#pragma unroll
        for (int i = 0; i < MT; i++) {
          carry[i] = 0.0;
#pragma unroll
          for (int jn = 0; jn < NT; jn++) { c[i][jn][0] = fabs(c[i][jn][0]); c[i][jn][1] = fabs(c[i][jn][1]); S0[i][jn] = 0.0; S1[i][jn] = 0.0; }
        }
#pragma unroll
        for (int jn = 0; jn < NT; jn++) {
          if (TAIL1 && jn == NT - 1) continue;  // the lone last node starts no cell and is never a search candidate
          if (EXACT || jn < nt_act) {
            const double h0 = hh[8 * jn + 2 * t], h1 = hh[8 * jn + 2 * t + 1];
            double Ta[MT], incl[MT];
#pragma unroll
            for (int i = 0; i < MT; i++) {
              const double p0 = c[i][jn][0], p1 = c[i][jn][1];
              const double var = (t == 0) ? c[i][jn + 1][0] : p0;
              const double nxt = __shfl_sync(FULL, var, (t == 3) ? lane - 3 : lane + 1);
              Ta[i] = h0 * (p0 + p1);
              incl[i] = Ta[i] + h1 * (p1 + nxt);
            }
#pragma unroll
            for (int i = 0; i < MT; i++) { const double v = __shfl_up_sync(FULL, incl[i], 1, 4); if (t >= 1) incl[i] += v; }
#pragma unroll
            for (int i = 0; i < MT; i++) { const double v = __shfl_up_sync(FULL, incl[i], 2, 4); if (t >= 2) incl[i] += v; }
#pragma unroll
            for (int i = 0; i < MT; i++) {
              double excl = __shfl_up_sync(FULL, incl[i], 1, 4);
              if (t == 0) excl = 0.0;
              const double tot = __shfl_sync(FULL, incl[i], 3, 4);
              S0[i][jn] = carry[i] + excl;
              S1[i][jn] = S0[i][jn] + Ta[i];
              carry[i] += tot;
            }
          }
        }
        double sc[MT], qt[MT];
        int cnt[MT];
#pragma unroll
        for (int i = 0; i < MT; i++) { sc[i] = 1.0 / carry[i]; qt[i] = qv_[i] * carry[i]; cnt[i] = 0; }  // q > S/total <=> q*total > S
#pragma unroll
        for (int jn = 0; jn < NTD; jn++) {
          if (EXACT || jn < nt_act) {
            const int node0 = 8 * jn + 2 * t;
#pragma unroll
            for (int i = 0; i < MT; i++) {
              cnt[i] += (node0 >= 1 && node0 <= n1 - 2 && qt[i] > S0[i][jn]) ? 1 : 0;
              cnt[i] += (node0 + 1 <= n1 - 2 && qt[i] > S1[i][jn]) ? 1 : 0;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < MT; i++) { cnt[i] += __shfl_xor_sync(FULL, cnt[i], 1); }
#pragma unroll
        for (int i = 0; i < MT; i++) { cnt[i] += __shfl_xor_sync(FULL, cnt[i], 2); }
#pragma unroll
        for (int i = 0; i < MT; i++) {
          const int i0 = cnt[i], i1 = cnt[i] + 1;
          const int js = i0 >> 3, ts = (i0 & 7) >> 1, es = i0 & 1;
          const int jq = i1 >> 3, tq = (i1 & 7) >> 1, eq = i1 & 1;
          double selS = 0.0, selP = 0.0, selQ = 0.0;
#pragma unroll
          for (int jn = 0; jn < NT; jn++) {
            if (jn == js) { selS = es ? S1[i][jn] : S0[i][jn]; selP = es ? c[i][jn][1] : c[i][jn][0]; }
            if (jn == jq) { selQ = eq ? c[i][jn][1] : c[i][jn][0]; }
          }
          const int base = lane & ~3;
          selS = __shfl_sync(FULL, selS, base | ts);
          selP = __shfl_sync(FULL, selP, base | ts);
          selQ = __shfl_sync(FULL, selQ, base | tq);
          cdf_lo[i] = selS * sc[i]; c1v[i] = selP * sc[i]; c2v[i] = selQ * sc[i]; i0v[i] = i0;
          if (carry[i] == 0.0) {
            // zero-mass conditional: uniform in index space (reference tt_irt1_int32.c:116-125)
            const double u = 1.0 / (double)(n1 - 1);
            const double s2 = 1.0 / ((double)(n1 - 1) * u);
            int k0 = 0;
            for (int j = 1; j <= n1 - 2; j++) k0 += (qv_[i] > ((double)j * u) * s2) ? 1 : 0;
            i0v[i] = k0; cdf_lo[i] = ((double)k0 * u) * s2; c1v[i] = u * s2; c2v[i] = u * s2;
          }
        }
      }

      PT_MARK(5)
      // ---- inversion tail: lane t = i of a quad finishes row 8i + g ---------------------------------
      {
        const int sel = (MT == 2) ? (t & 1) : 0;
        const bool valid = (t < MT) && (8 * sel + g < nvalid);
        const int m = mrow[sel];
        const int i0 = i0v[sel];
        const CellOut o = invert_cell_fast(qv_[sel], cdf_lo[sel], c1v[sel], c2v[sel], xg[i0], xg[i0 + 1], ihs[i0]);
        if (valid) {
          a.z[m] = o.xk;
          if (a.idx_out) a.idx_out[m] = i0;
          if (!a.last) {
            a.idx[m] = i0; a.w1[m] = o.w1; a.w2[m] = o.w2;
            a.lp[m] = lp_[sel] + o.logp;
            atomicAdd(&hist[i0], 1);
          } else {
            a.lpz[m] = lp_[sel] + o.logp;
          }
        }
      }
    }  // nvalid > 0

    PT_MARK(6)
    // ---- rotate the row pipeline --------------------------------------------------------------------
    nvC = nvN;
#pragma unroll
    for (int i = 0; i < MT; i++) { mrow[i] = mrow_n[i]; w1[i] = w1n[i]; w2[i] = w2n[i]; qv_[i] = qn[i]; lp_[i] = lpn[i]; }
    nvN = nvNN; idN = idNN;
  }

  PT_FLUSH
  if (!a.last) {
    __syncthreads();
    for (int i = tid; i < n1 - 1; i += NTHR)
      if (hist[i]) atomicAdd(a.hist_next + i, hist[i]);
  }
}

template <int RT, int NT, int WARPS, int MT, bool EXACT, bool TAIL1>
cudaError_t launch_variant(const TransArgs &a, int sm_count, cudaStream_t st) {
  using L = SmemLayout<RT, NT, WARPS, MT>;
  static int occ = 0;
  if (occ == 0) {
    int o = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, transition_kernel<RT, NT, WARPS, MT, EXACT, TAIL1>, WARPS * 32, L::bytes);
    if (e != cudaSuccess) return e;
    occ = o > 0 ? o : 1;
  }
  const int rows_cta = WARPS * 8 * MT;
  int64_t max_tiles = ((int64_t)a.rows + rows_cta - 1) / rows_cta + (a.n0 - 1);
  int64_t grid = (int64_t)sm_count * occ;
  if (grid > max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  transition_kernel<RT, NT, WARPS, MT, EXACT, TAIL1><<<(unsigned)grid, WARPS * 32, L::bytes, st>>>(a);
  return cudaGetLastError();
}

template <int RT, int NT, int WARPS, int MT>
cudaError_t launch_one(const TransArgs &a, int sm_count, cudaStream_t st) {
  const bool exact = a.r0 == 8 * RT && a.r1 == 8 * RT && (a.n1 + 7) / 8 == NT;
  if (exact && a.n1 == 8 * (NT - 1) + 1) return launch_variant<RT, NT, WARPS, MT, true, true>(a, sm_count, st);
  if (exact) return launch_variant<RT, NT, WARPS, MT, true, false>(a, sm_count, st);
  return launch_variant<RT, NT, WARPS, MT, false, false>(a, sm_count, st);
}

template <int RT, int NT, int WARPS, int MT>
cudaError_t init_one() {
  using L = SmemLayout<RT, NT, WARPS, MT>;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, MT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, MT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(transition_kernel<RT, NT, WARPS, MT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
}

constexpr int kWarps = TTIRT_WARPS;
constexpr int kMT = TTIRT_MT;

}  // namespace

int fast_class_for(int rmax, int nmax) {
  if (rmax <= 16 && nmax <= 24) return 0;
  if (rmax <= 32 && nmax <= 40) return 1;
  if (rmax <= 64 && nmax <= 72) return 2;
  return -1;
}

int fast_rows_per_cta(int) { return kWarps * 8 * kMT; }

cudaError_t fast_init(int) {
  cudaError_t e;
  if ((e = init_one<2, 3, kWarps, kMT>()) != cudaSuccess) return e;
  if ((e = init_one<4, 5, kWarps, kMT>()) != cudaSuccess) return e;
  if ((e = init_one<8, 9, kWarps, kMT>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_transition(int cls, const TransArgs &a, int sm_count, cudaStream_t st) {
  switch (cls) {
    case 0: return launch_one<2, 3, kWarps, kMT>(a, sm_count, st);
    case 1: return launch_one<4, 5, kWarps, kMT>(a, sm_count, st);
    case 2: return launch_one<8, 9, kWarps, kMT>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
  }
}

#ifdef TTIRT_PHASE_TIMING
void phase_cycles_read(unsigned long long *out) {
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 40);
  unsigned long long z[40] = {0};
  cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
}
#endif

}  // namespace ttirt
