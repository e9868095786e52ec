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
//   (3) trapezoid CDF of the weighted pdf (:107-113), interval search (:134-142), closed-form quadratic
//       inversion (:146-159), log-density accumulation (:161-165), emission of (interval, w1, w2) and the
//       histogram that drives the next counting sort.
//
// Work decomposition: (1) and (2) run in eight MMA warps, each owning 16 samples per tile (two 8-row MMA
// tiles); (3) runs in four tail warps that receive the pdf tiles through shared memory.  The MMA warps only meet
// when their CTA moves to the next interval bin and restages a slab.
#include <cstdlib>

#include "ttirt_common.cuh"

namespace ttirt {

#ifdef TTIRT_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8 + 32];
__device__ long long g_trace[2 * 24 * 8];   // CTA 0, the two MMA warps of sub-partition 0: clock at every PT_MARK of the first 24 tiles
#define PT_DECL unsigned long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_t = clock64(); const long long pt_start = pt_t;
#define PT_MARK(k) { const long long n_ = clock64(); pt_[k] += (unsigned long long)(n_ - pt_t); pt_t = n_; \
  if (blockIdx.x == 0 && lane == 0 && (mw & 3) == 0 && tile - t_begin < 24) g_trace[((mw >> 2) * 24 + (tile - t_begin)) * 8 + (k)] = n_; }
#define PT_FLUSH if (lane == 0) { for (int k_ = 0; k_ < 8; k_++) atomicAdd(&g_phase_cycles[k_], pt_[k_]); atomicAdd(&g_phase_cycles[8 + (warp & 7)], (unsigned long long)(clock64() - pt_start)); }
// tail warps: phases land in g_phase_cycles[24 + k].  BAR.SYNC defers its blocking to the next dependent
// instruction, so the time waiting for a parked tile shows up in the phase after the barrier.
#define TT_DECL unsigned long long tt_[6] = {0, 0, 0, 0, 0, 0}; long long tt_t = clock64();
#define TT_MARK(k) { const long long n_ = clock64(); tt_[k] += (unsigned long long)(n_ - tt_t); tt_t = n_; }
#define TT_FLUSH if (lane == 0) { for (int k_ = 0; k_ < 6; k_++) atomicAdd(&g_phase_cycles[24 + k_], tt_[k_]); }
#else
#define PT_DECL
#define PT_MARK(k)
#define PT_FLUSH
#define TT_DECL
#define TT_MARK(k)
#define TT_FLUSH
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
  // four independent loads in flight per thread and pass (a dependent load-store loop pays one L2 latency per element)
  for (int e0 = tid; e0 < total; e0 += 4 * nthr) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * nthr;
      const int c = e / KP, a = e - c * KP;
      v[u] = (e < total && a < K && c < NC) ? __ldg(src + a + (int64_t)c * cs) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * nthr;
      if (e < total) { const int c = e / KP, a = e - c * KP; dst[b_phys(c, a, KP)] = v[u]; }
    }
  }
}

// The same image, written by cp.async (LDGSTS) in 16-byte pieces: nothing passes through registers and every piece of
// the slab is in flight at once, so a restage costs one L2 round trip instead of one per element.  Needs K even, cs even
// and a 16-byte aligned src (rows 2i, 2i+1 of a column are one piece and stay adjacent under the swizzle); pieces outside
// K x NC are zero-filled (src-size 0).  The caller commits / waits.
__device__ __forceinline__ void cp_async16_zfill(void *dst, const void *src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ void stage_b_async(double *dst, const double *__restrict__ src, int K, int NC, int64_t cs, int KP, int NCP,
                              int tid, int nthr) {
  const int hp = KP >> 1, total = NCP * hp;
  for (int e = tid; e < total; e += nthr) {
    const int c = e / hp, a = (e - c * hp) << 1;
    const bool valid = a < K && c < NC;
    cp_async16_zfill(dst + b_phys(c, a, KP), valid ? src + a + (int64_t)c * cs : src, valid);
  }
}
__device__ __forceinline__ void stage_slab(bool async_ok, double *dst, const double *__restrict__ src, int K, int NC, int64_t cs,
                                           int KP, int NCP, int tid, int nthr) {
  if (async_ok) stage_b_async(dst, src, K, NC, cs, KP, NCP, tid, nthr);
  else stage_b(dst, src, K, NC, cs, KP, NCP, tid, nthr);
}

#ifndef TTIRT_ALT
#define TTIRT_ALT 1     // partner alternation of the update phases: 0 off, 1 the r <= 32 class (+2 % there; -2 % at r <= 16 and r <= 64), 2 every class
#endif
#ifndef TTIRT_SWPIPE
#define TTIRT_SWPIPE 0
#endif
// Warps per CTA (template parameters MW, TW): MMA warps and tail warps, in multiples of the four SM sub-partitions.
//   MW = 8, TW = 4   the r <= 64 class: two MMA warps per sub-partition share its FP64 tensor pipe, one tail warp serves both
//                    (the MMA side is the long one there; a third MMA warp fits neither the registers nor shared memory)
//   MW = 8, TW = 8   lighter classes, round-1 layout: every MMA warp has a tail warp of its own
//   MW = 12, TW = 4  lighter classes: three MMA warps per sub-partition (their accumulators are small enough for 128
//                    registers), so that two are inside their DMMA loops while the third gathers, parks and waits
constexpr int MT = 2;           // 8-row MMA tiles per MMA warp
constexpr int WROWS = 8 * MT;   // samples per warp tile
__host__ __device__ constexpr int nthr_of(int mw, int tw) { return 32 * (mw + tw); }
// register file split (setmaxnreg, per warpgroup of four warps): the launch allocates 65536 / threads registers per thread
// (168 for 12 warps, 128 for 16), the tail warpgroup(s) shrink to tail_regs and the MMA warpgroups grow to mma_regs;
// 8 * 192 + 4 * 120 = 12 * 168, 8 * 152 + 8 * 104 = 16 * 128, 12 * 128 + 4 * 128 = 16 * 128: exactly the pool of the launch
#ifndef TTIRT_LIGHT_TAIL_REGS
#define TTIRT_LIGHT_TAIL_REGS 104   // 120 (MMA warps at 136) measured: no difference
#endif
__host__ __device__ constexpr int tail_regs_of(int mw, int tw) { return mw == 12 ? 128 : (tw == 4 ? 120 : TTIRT_LIGHT_TAIL_REGS); }
__host__ __device__ constexpr int mma_regs_of(int mw, int tw) { return mw == 12 ? 128 : (tw == 4 ? 192 : 256 - tail_regs_of(mw, tw)); }

// GD: depth of the row gather.  1: the rows of tile t+1 are requested once tile t's update phase has consumed the staged
// tile (a whole pdf phase ahead: enough at r = 64, where that phase lasts ~2 us).  2 (the two lighter classes, where the pdf
// phase is a quarter of that and a request would still be in flight when it is needed): two staged tiles per MMA warp,
// the rows of tile t+2 are requested at that point.
template <int RT, int NT, bool TAIL1, int MW, int TW, int GD>
struct SmemLayout {
  static constexpr int KPMAX = (8 * RT + 15) & ~15;
  static constexpr int SLAB = 8 * RT * KPMAX;      // doubles per slab buffer
  static constexpr int PN = 8 * NT * KPMAX;        // doubles for P_{k+1}
  static constexpr int NBMAX = 8 * NT;             // >= n - 1 intervals
  static constexpr int FPITCH = 8 * RT + 8;        // doubles per staged left-interface row (+64 B: rows g, g+1 hit different bank halves)
  static constexpr int FTILE = WROWS * FPITCH;     // doubles of staged rows per MMA warp
  // parked |pdf| tile of one warp tile, [node][row]: pitch 18 makes the MMA warps' fragment stores (rows g, nodes 2t)
  // and the tail warp's row-per-lane reads both bank-conflict free
  static constexpr int NPD = TAIL1 ? 8 * (NT - 1) + 1 : 8 * NT;
  static constexpr int RS = WROWS + 2;
  static constexpr int PB = NPD * RS;
  static constexpr int HS = TAIL1 ? 4 * (NT - 1) : 4 * NT;   // cells walked by each of the two lanes of a row
  static constexpr int HB = HS / 4;                // ... as four blocks of HB consecutive cells, walked side by side
  static constexpr int NHH = 2 * HS + 8;           // entries of the per-node tables rw / hr
  static constexpr size_t bytes = sizeof(double) * (2 * SLAB + PN + MW * GD * FTILE + TW * PB + 2 * NHH + 2 * 8 * NT) +
                                  sizeof(int) * (2 * (NBMAX + 1) + NBMAX + TW * (WROWS + 2) + MW);
};

// Named barriers (id 0 is __syncthreads).  Tile hand-over between an MMA warp and its tail warp is a 64-thread
// barrier on which one side only arrives: shared-memory ordering without a memory fence, which would also wait
// for the F' stores and the row gather still in flight.
//   1                 the eight MMA warps (slab restaging at a bin change)
//   2 + tw            FULL : a producer of buffer tw arrives after parking a tile, tail warp tw waits
//   6 + 4*p + tw      EMPTY (MW = 8, TW = 4): tail warp tw arrives when producer p (0 / 1) may overwrite the buffer, producer p waits.
//                     With eight tail warps or three producers per buffer the ids would not fit (16 in all): there the
//                     producers of a buffer poll a counter in shared memory that the tail warp bumps after its last read of
//                     a tile (same warp, same memory pipe: the store follows the loads); use number PRODS * i + p of the
//                     buffer belongs to producer p's tile i, so the counter also keeps the producers in turn.
template <int MW>
__device__ __forceinline__ void bar_mma_warps() { asm volatile("bar.sync 1, %0;" ::"n"(32 * MW) : "memory"); }
__device__ __forceinline__ void bar_pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ int ld_volatile_s(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_volatile_s(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

// EXACT: r0 == r1 == 8*RT and ceil(n1/8) == NT, so every tile loop runs its full static trip count and no
// guard branches are compiled in (the steady state of a uniform-rank TT).  TAIL1 (EXACT only): n1 == 8*(NT-1)+1,
// the usual 2^p+1 grid; the lone last grid column is then a 4-lane DFMA dot product instead of a whole DMMA
// column tile that would be 7/8 padding.
//
// Warp roles.  On B200 the scalar FP64 instructions of the CDF / inversion phase go through the same pipe as DMMA:
// while a warp walks that dependent chain its DMMA stream stops, and a partner warp streaming DMMAs stretches the
// chain ~3x.  So the chain is taken off the MMA warps altogether.  Eight MMA warps (two per SM sub-partition) do
// nothing but gather -> interface update -> pdf contraction and park the |pdf| tile (16 rows x n1 nodes) in shared
// memory; four tail warps (one per sub-partition) turn parked tiles into samples.  Each tail warp owns one tile
// buffer that its two MMA warps fill in strict alternation; hand-over is by named barriers (below).
//
// Row gather: the left-interface rows of an MMA warp's NEXT tile are fetched asynchronously (cp.async / LDGSTS, one
// coalesced 8*r0-byte row per instruction) into a per-warp shared tile as soon as the current tile's update
// phase has consumed that tile, i.e. a whole pdf phase ahead of use.
template <int RT, int NT, bool EXACT, bool TAIL1, int MW, int TW, int GD>
__global__ void __launch_bounds__(nthr_of(MW, TW), 1) transition_kernel(const TransArgs a) {
  static_assert(EXACT || !TAIL1, "TAIL1 needs EXACT");
  static_assert(GD == 1 || GD == 2, "gather depth 1 or 2");
  static_assert(MW % TW == 0 && (TW == 4 || TW == 8), "tail warp tw serves the MMA warps mw with mw % TW == tw (same sub-partition)");
  constexpr bool ALT = MW == 8 && ((TTIRT_ALT == 2) || (TTIRT_ALT == 1 && TW == 8 && RT == 4));   // pairs mw, mw ^ 4
  using L = SmemLayout<RT, NT, TAIL1, MW, TW, GD>;
  constexpr int MMA_WARPS = MW, ROWS_CTA = MW * WROWS;
  constexpr int TAIL_WARPS = TW, NTHR = nthr_of(MW, TW), PRODS = MMA_WARPS / TW;
  constexpr bool EMPTY_BARS = MW == 8 && TW == 4;   // hand-back by named barriers (ids 6 .. 13); else by the polled counter
  constexpr int LAUNCH_REGS = (65536 / NTHR) & ~7;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *slab0 = reinterpret_cast<double *>(smem_raw);
  double *slab1 = slab0 + L::SLAB;
  double *Ps = slab1 + L::SLAB;
  double *ft_all = Ps + L::PN;                     // per-MMA-warp staged left-interface rows
  double *pb_all = ft_all + MMA_WARPS * GD * L::FTILE;  // per-tail-warp parked |pdf| tile
  double *rw = pb_all + TAIL_WARPS * L::PB;        // 1 / node_weight(j) of dimension k+1's grid (0 beyond node n1-1)
  double *hr = rw + L::NHH;                        // h_{j-1} / node_weight(j): share of node j's weight left of it
  double *xg = hr + L::NHH;                        // grid of dimension k+1
  double *ihs = xg + 8 * NT;                       // reciprocal cell widths of dimension k+1
  int *bts = reinterpret_cast<int *>(ihs + 8 * NT);  // bin -> first CTA tile
  int *bst = bts + (L::NBMAX + 1);                    // bin -> first sorted row
  int *hist = bst + (L::NBMAX + 1);                   // histogram of the intervals chosen in dimension k+1
  int *ids_all = hist + L::NBMAX;                     // per tail warp: sample ids of the parked tile
  int *nv_all = ids_all + TAIL_WARPS * WROWS;         // per tail warp: valid rows of the parked tile
  int *consumed = nv_all + TAIL_WARPS;                // per tail warp: parked tiles released so far (TW == 8 hand-back)
  int *upd_done = consumed + TAIL_WARPS;              // per MMA warp: tiles whose update phase is finished (partner alternation)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int FP = L::FPITCH, RS = L::RS;

  const int r0 = a.r0, r1 = a.r1, n1 = a.n1, nb0 = a.n0 - 1;
  constexpr int KP = L::KPMAX;                   // compile-time column pitch: B-fragment offsets fold into immediates
  constexpr int NTD = TAIL1 ? NT - 1 : NT;       // grid column tiles computed by DMMA

  // ---- prologue: everything the CTA needs is requested at once (cp.async), the small tables are computed underneath ----
  // 16-byte pieces need even leading dimensions and aligned bases (the host checks the bases: TransArgs::async_ok)
  const bool p_async = !(r1 & 1);                    // pnext itself is aligned by construction (DimInfo::off_pw)
  const bool slab_async = a.async_ok && !(r0 & 1);   // slab b starts at core + b * r0, column stride r0 * n0
  stage_slab(p_async, Ps, a.pnext, r1, n1, r1, KP, 8 * NT, tid, NTHR);
  if (warp == 0) bin_offsets_warp(a.hist_cur, nb0, ROWS_CTA, bst, bts, lane);
  for (int i = tid; i < L::NHH; i += NTHR) {
    const double w = node_weight(a.xnext, i, n1);
    const double hl = (i >= 1 && i < n1) ? 0.5 * (a.xnext[i] - a.xnext[i - 1]) : 0.0;
    rw[i] = w > 0.0 ? 1.0 / w : 0.0;
    hr[i] = w > 0.0 ? hl / w : 0.0;
  }
  for (int i = tid; i < 8 * NT; i += NTHR) {
    xg[i] = (i < n1) ? a.xnext[i] : 0.0;
    ihs[i] = (i + 1 < n1) ? 1.0 / (a.xnext[i + 1] - a.xnext[i]) : 0.0;
    if (i < L::NBMAX) hist[i] = 0;
  }
  if (tid < TAIL_WARPS) consumed[tid] = 0;
  if (tid < MMA_WARPS) upd_done[tid] = 0;
  __syncthreads();

  const int total_tiles = bts[nb0];
  const int t_begin = (int)(((int64_t)blockIdx.x * total_tiles) / gridDim.x);
  const int t_end = (int)(((int64_t)(blockIdx.x + 1) * total_tiles) / gridDim.x);
  // the two slabs of the first tile's interval, by all threads
  int b_first = -1;
  if (t_begin < t_end) {
    b_first = 0;
    while (t_begin >= bts[b_first + 1]) ++b_first;
    const int64_t slab_cs0 = (int64_t)r0 * a.n0;
    stage_slab(slab_async, slab0, a.core + (int64_t)b_first * r0, r0, r1, slab_cs0, KP, 8 * RT, tid, NTHR);
    stage_slab(slab_async, slab1, a.core + (int64_t)(b_first + 1) * r0, r0, r1, slab_cs0, KP, 8 * RT, tid, NTHR);
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  const bool is_tail = warp < TAIL_WARPS;   // tail warps first, then the MMA warps (role changes at warpgroup granularity for setmaxnreg)
  if (is_tail) {
    // =========================================== tail warps ===========================================
    // Lanes l and l+16 share row l & 15 of the parked tile, lane l the lower half of the grid and lane l+16 the upper
    // half (reference tt_irt1_int32.c:105-165 for one sample).
    //
    // The MMA warps deliver v_j = w_j p_j (P_{k+1} is staged with its columns scaled by the trapezoid node weights, see
    // node_weight()), so with R_j = sum_{i<j} |v_i|
    //   cdf_j = R_j + (h_{j-1} / w_j) |v_j|,   R_j <= cdf_j <= R_{j+1},   mass = R_{n1},
    // and the CDF pass is a plain running sum.  Next to a DMMA stream every instruction of another warp waits for a
    // gap between two DMMAs (~20 cycles, whatever its type), so the tail is written for instruction count: one pass,
    // a two-level search on integer bit patterns, one inversion per pair of tiles.
    if (tail_regs_of(MW, TW) < LAUNCH_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(tail_regs_of(MW, TW)));
    const int tw = warp, row = lane & 15, hf = lane >> 4;
    const double *pbr = pb_all + tw * L::PB + row;
    const int *ids = ids_all + tw * WROWS;
    constexpr int HS = L::HS, HB = L::HB;
    const int c0 = hf * HS, nlast = n1 - 1;
    const int uses = PRODS * (t_end - t_begin);
    // The inversion runs once per PAIR of tiles: the first tile of a pair waits in registers (st_*), then lanes 0-15
    // finish its rows while lanes 16-31 finish the second tile's.
    int st_nv = 0, st_m = 0, st_i0 = 0, st_E = 0;
    double st_dq = 0.0, st_c1 = 1.0, st_c2 = 1.0, st_mass = 1.0, st_N = 1.0, st_D = 1.0;
    TT_DECL
    // hand the buffer back: to the other producer through its EMPTY barrier, or to whoever is next through the counter
    auto release = [&](int seq) {
      if (EMPTY_BARS) {
        if (seq + 1 < uses) bar_pair_arrive(6 + 4 * ((seq + 1) % PRODS) + tw);
      } else {
        __syncwarp();
        if (lane == 0) st_volatile_s(consumed + tw, seq + 1);
      }
    };
    for (int seq = 0; seq < uses + (uses & 1); ++seq) {   // an odd count gets one empty pass that finishes the waiting tile
      TT_MARK(4)
      const bool real = seq < uses;
      if (real) bar_pair_sync(2 + tw);              // FULL: the next tile is parked
      TT_MARK(0)
      int nv = real ? nv_all[tw] : 0;
      int m = 0, i0 = 0, lpE = 0;
      double dq = 0.0, c1 = 1.0, c2 = 1.0, mass = 1.0, lpN = 1.0, lpD = 1.0;
      if (nv == 0) {
        if (real) release(seq);
      } else {
        m = ids[row];
        const double qv = a.q[m];
        lpN = a.lp[m]; lpD = a.lpd[m];
        lpE = a.lpe[m];
        // running sums of this lane's half row: four independent chains (blocks of HB nodes), kept in registers
        double Rl[4][HB];
#pragma unroll
        for (int c = 0; c < HB; ++c) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const double v = fabs(pbr[(c0 + k * HB + c) * RS]);   // nodes beyond n1-1 are parked as zeros
            Rl[k][c] = c ? Rl[k][c - 1] + v : v;
          }
        }
        TT_MARK(1)
        const double th = (Rl[0][HB - 1] + Rl[1][HB - 1]) + (Rl[2][HB - 1] + Rl[3][HB - 1]);
        const double ot = __shfl_xor_sync(FULL, th, 16);
        const double tot0 = hf ? ot : th, tot1 = hf ? th : ot;
        // mass of the row's conditional (reference :107-113, cdf[n-1]); TAIL1: the last node sits outside the two halves
        const double total = TAIL1 ? (tot0 + tot1) + fabs(pbr[nlast * RS]) : tot0 + tot1;
        const double qt = qv * total;                // q > cdf/mass  <=>  q*mass > cdf (unnormalised compare)
        // search (reference :134-142), first on R: js = last node <= n1-2 with R_js < q*mass.  Two levels: the block of
        // this lane's half row whose first node is still below, then the node inside it.  Non-negative doubles order
        // like their bit patterns, so the inner compares are integer ones.
        double base[4];
        base[0] = hf ? tot0 : 0.0;
#pragma unroll
        for (int k = 1; k < 4; k++) base[k] = base[k - 1] + Rl[k - 1][HB - 1];
        int ks = 0;
#pragma unroll
        for (int k = 1; k < 4; k++) ks += (qt > base[k] && (TAIL1 || c0 + k * HB <= n1 - 2)) ? 1 : 0;
        const bool k1 = ks & 1, k2 = ks & 2;
        const double bsel = k2 ? (k1 ? base[3] : base[2]) : (k1 ? base[1] : base[0]);
        const double qk = qt - bsel;                 // q*mass - R at the block's first node
        const long long qb = __double_as_longlong(qk);
        int cs = (qb > 0 && (TAIL1 || c0 + ks * HB <= n1 - 2)) ? 1 : 0;   // the block's first node itself
        double Rsel = 0.0;
#pragma unroll
        for (int c = 1; c < HB; ++c) {
          const double Rc = k2 ? (k1 ? Rl[3][c - 1] : Rl[2][c - 1]) : (k1 ? Rl[1][c - 1] : Rl[0][c - 1]);   // R - base at node c of the block
          const bool below = qb > __double_as_longlong(Rc) && (TAIL1 || c0 + ks * HB + c <= n1 - 2);
          if (below) { cs = c + 1; Rsel = Rc; }
        }
        int js = cs > 0 ? c0 + ks * HB + cs - 1 : -1;
        double dR = qk - Rsel;                       // q*mass - R_js
        {
          const int js_o = __shfl_xor_sync(FULL, js, 16);
          const double dR_o = __shfl_xor_sync(FULL, dR, 16);
          const int js_hi = hf ? js : js_o, js_lo = hf ? js_o : js;
          const double dR_hi = hf ? dR : dR_o, dR_lo = hf ? dR_o : dR;
          js = js_hi >= 0 ? js_hi : js_lo;
          dR = js_hi >= 0 ? dR_hi : dR_lo;
        }
        // nodes before js are below q, nodes after it are not; js itself is decided by its own CDF value
        if (js < 1) {
          i0 = 0; dq = qt;                           // cdf_0 = 0
          c1 = fabs(pbr[0]) * rw[0]; c2 = fabs(pbr[RS]) * rw[1];
        } else {
          const double vB = fabs(pbr[(js - 1) * RS]), vA = fabs(pbr[js * RS]), vC = fabs(pbr[(js + 1) * RS]);
          const double dA = fma(-hr[js], vA, dR);                  // q*mass - cdf_js
          if (__double_as_longlong(dA) > 0) {
            i0 = js; dq = dA; c1 = vA * rw[js]; c2 = vC * rw[js + 1];
          } else {
            i0 = js - 1; c1 = vB * rw[js - 1]; c2 = vA * rw[js];
            dq = fma(1.0 - hr[js - 1], vB, dR);                    // q*mass - cdf_{js-1} = dR + (w_{js-1} - h_{js-2}) p_{js-1}
          }
        }
        const double s2 = pow2_scale(total);         // exact power-of-two normalisation instead of 1/mass
        c1 *= s2; c2 *= s2;                          // (consumes the loads before the buffer is handed back)
        release(seq);                                // the (other) producer may park its next tile
        dq *= s2;
        mass = total * s2;
        if (total == 0.0) {
          // zero-mass conditional: uniform in index space (reference tt_irt1_int32.c:116-125)
          const double u = 1.0 / (double)(n1 - 1);
          const double sf = 1.0 / ((double)(n1 - 1) * u);
          int k0 = 0;
          for (int j = 1; j <= n1 - 2; j++) k0 += (qv > ((double)j * u) * sf) ? 1 : 0;
          i0 = k0; dq = qv - ((double)k0 * u) * sf; c1 = u * sf; c2 = u * sf; mass = 1.0;
        }
      }
      TT_MARK(2)
      if ((seq & 1) == 0) {
        st_nv = nv; st_m = m; st_i0 = i0; st_E = lpE;
        st_dq = dq; st_c1 = c1; st_c2 = c2; st_mass = mass; st_N = lpN; st_D = lpD;
        continue;
      }
      if (hf == 0) {
        nv = st_nv; m = st_m; i0 = st_i0; lpE = st_E;
        dq = st_dq; c1 = st_c1; c2 = st_c2; mass = st_mass; lpN = st_N; lpD = st_D;
      }
      const CellFast o = invert_cell_fast(dq, c1, c2, xg[i0], xg[i0 + 1], ihs[i0]);
      lp_accumulate(lpN, lpD, lpE, o.dens, mass);
      if (row < nv) {
        a.z[m] = o.xk;
        if (a.idx_out) a.idx_out[m] = i0;
        if (!a.last) {
          a.idx[m] = i0; a.w1[m] = o.w1; a.w2[m] = o.w2;
          a.lp[m] = lpN; a.lpd[m] = lpD; a.lpe[m] = lpE;
          atomicAdd(&hist[i0], 1);
        } else {
          a.lpz[m] = lp_finish(lpN, lpD, lpE);
        }
      }
      TT_MARK(3)
    }
    TT_FLUSH
  } else {
    // =========================================== MMA warps ============================================
    if (mma_regs_of(MW, TW) > LAUNCH_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(mma_regs_of(MW, TW)));
    const int g = lane >> 2, t = lane & 3;
    const int mw = warp - TAIL_WARPS;          // index among the MMA warps
    const int mtid = 32 * mw + lane;
    const int tw = mw & (TAIL_WARPS - 1), prod = mw / TAIL_WARPS;  // tail warp served (same SM sub-partition); first / second producer of its buffer
    double *ft_base = ft_all + mw * GD * L::FTILE;
    double *pbw = pb_all + tw * L::PB;
    const int ks0 = EXACT ? RT : (r0 + 7) >> 3, rt_act = EXACT ? RT : (r1 + 7) >> 3, nt_act = EXACT ? NT : (n1 + 7) >> 3;
    const uint32_t row_bytes = (uint32_t)(8 * ks0) * 8u;   // bytes of a left-interface row that the update reads
    const int64_t slab_cs = (int64_t)r0 * a.n0;

    int cur0 = b_first, cur1 = b_first < 0 ? -1 : b_first + 1;  // interval slab held by slab0 / slab1 (the prologue staged the first pair)
    int b = b_first < 0 ? 0 : b_first;   // bin of the current tile
    int bh = 0;                // bin hint of the row look-ahead (monotone)
    const double *sl_lo = slab0, *sl_hi = slab1;

    // rows of this warp in CTA tile `tile`: first sorted row and number of valid rows (0 past the end)
    auto rows_of = [&](int tile, int &nv) -> int {
      if (tile >= t_end) { nv = 0; return 0; }
      while (tile >= bts[bh + 1]) ++bh;
      const int row0 = bst[bh] + (tile - bts[bh]) * ROWS_CTA + mw * WROWS;
      nv = max(0, min(WROWS, bst[bh + 1] - row0));
      return row0;
    };
    // sample ids of a tile: lane l holds the id of row l % WROWS; rows past nv repeat row 0
    auto load_ids = [&](int row0, int nv) -> int {
      const int r = lane & (WROWS - 1);
      return nv > 0 ? a.perm[row0 + (r < nv ? r : 0)] : 0;
    };

    // Row pipeline of this warp.  Stage 0 is the current tile, stage j the tile j ahead; ids are known GD + 1 tiles ahead,
    // rows and per-row scalars are requested GD tiles ahead.
    int nvC = 0, idC = 0;            // current tile: valid rows, row ids (lane-distributed)
    int nvQ[GD + 1], idQ[GD + 1];    // [j]: tile j ahead, j = 1 .. GD
    // Gather of the rows whose ids are `ids` into a staged tile of this warp: one cp.async (LDGSTS) instruction per row,
    // 16 bytes per lane, so a row is one coalesced 8*r0-byte segment.  Always commits (an empty group when there are no
    // rows), so that "all but the latest GD - 1 groups" below always means "the current tile's rows".
    auto issue_gather = [&](int ids, int nv, double *ft) {
      if (nv > 0) {
#pragma unroll
        for (int r = 0; r < WROWS; r++) {
          const int id = __shfl_sync(FULL, ids, r);
          if (16u * lane < row_bytes)
            cp_async16(reinterpret_cast<char *>(ft + r * FP) + 16 * lane, reinterpret_cast<const char *>(a.F + (size_t)id * a.ldf) + 16 * lane);
        }
      }
      cp_async_commit();
    };
    // per-row scalars, m-tile i = rows 8i+g: [0] current tile, [j] the tile j ahead (in flight)
    int mrow[GD + 1][MT];
    double w1s[GD + 1][MT], w2s[GD + 1][MT];
#pragma unroll
    for (int j = 0; j <= GD; j++)
#pragma unroll
      for (int i = 0; i < MT; i++) { mrow[j][i] = 0; w1s[j][i] = w2s[j][i] = 0.0; }
    auto load_scalars = [&](int j, int ids, int nv) {
#pragma unroll
      for (int i = 0; i < MT; i++) mrow[j][i] = __shfl_sync(FULL, ids, 8 * i + g);
      if (nv > 0) {
#pragma unroll
        for (int i = 0; i < MT; i++) { w1s[j][i] = a.w1[mrow[j][i]]; w2s[j][i] = a.w2[mrow[j][i]]; }
      }
    };
    {
      const int rowC = rows_of(t_begin, nvC);
      idC = load_ids(rowC, nvC);
#pragma unroll
      for (int j = 1; j <= GD; j++) {
        const int rowj = rows_of(t_begin + j, nvQ[j]);
        idQ[j] = load_ids(rowj, nvQ[j]);
      }
      issue_gather(idC, nvC, ft_base);
      load_scalars(0, idC, nvC);
      if (GD == 2) {
        issue_gather(idQ[1], nvQ[1], ft_base + L::FTILE);
        load_scalars(1, idQ[1], nvQ[1]);
      }
    }
    double (&w1)[MT] = w1s[0];
    double (&w2)[MT] = w2s[0];

    PT_DECL
    for (int tile = t_begin; tile < t_end; ++tile) {
      double *ft = ft_base + (GD == 2 ? ((tile - t_begin) & 1) * L::FTILE : 0);   // this tile's staged rows
      PT_MARK(7)
      while (tile >= bts[b + 1]) ++b;
      if (cur0 == b && cur1 == b + 1) {
        sl_lo = slab0; sl_hi = slab1;
      } else if (cur1 == b && cur0 == b + 1) {
        sl_lo = slab1; sl_hi = slab0;
      } else {
        bar_mma_warps<MW>();  // every MMA warp is done with the previous bin's slabs
        constexpr int NM = 32 * MMA_WARPS;
        if (cur0 == b) {
          stage_slab(slab_async, slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur1 = b + 1;
        } else if (cur1 == b) {
          stage_slab(slab_async, slab0, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur0 = b + 1;
        } else if (cur0 == b + 1) {
          stage_slab(slab_async, slab1, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur1 = b;
        } else if (cur1 == b + 1) {
          stage_slab(slab_async, slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur0 = b;
        } else {
          stage_slab(slab_async, slab0, a.core + (int64_t)b * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur0 = b;
          stage_slab(slab_async, slab1, a.core + (int64_t)(b + 1) * r0, r0, r1, slab_cs, KP, 8 * RT, mtid, NM); cur1 = b + 1;
        }
        cp_async_commit();
        cp_async_wait_all();   // (also the row gather in flight: it is needed right after anyway)
        bar_mma_warps<MW>();
        if (cur0 == b) { sl_lo = slab0; sl_hi = slab1; } else { sl_lo = slab1; sl_hi = slab0; }
      }

      PT_MARK(0)
      const int nvalid = nvC;
      double c[MT][NT][2];
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) c[i][j][0] = c[i][j][1] = 0.0;

      // Partner alternation.  The two MMA warps of an SM sub-partition (mw, mw ^ 4) run the same instruction stream on the one
      // FP64 pipe they share; left alone they fall into lock-step in the lighter classes (timeline in
      // profiles/r02_phase_timing_r32.log: both leave their DMMA loops at the same time and the pipe idles for a quarter of
      // every tile).  So their update phases take turns: the first warp of a pair starts tile i's update when its partner has
      // finished tile i-1's, the second when the first has finished tile i's.  Each warp's pdf phase, parking and
      // bookkeeping then run under the partner's update phase.
      if (ALT) {
        const int need = (mw & 4) ? (tile - t_begin) + 1 : (tile - t_begin);
        if (lane == 0)
          while (ld_volatile_s(upd_done + (mw ^ 4)) < need) __nanosleep(64);
        __syncwarp();
      }
      if (nvalid > 0) {
        // ---- (1) interface update: A fragments from the staged rows ------------------------------------
        if (GD == 1) cp_async_wait_all(); else asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        PT_MARK(1)
        double acc[MT][RT][2];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < RT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#if TTIRT_SWPIPE
        if (EXACT) {
          // Software-pipelined operand loads: the fragments of step (j, jj + 1) are requested before the DMMAs of step
          // (j, jj) issue, so a warp that has the sub-partition to itself (its partner is between tiles) does not wait
          // for shared memory in front of every group of DMMAs.
          double2 fq[2][MT], bl[2], bh[2];
#pragma unroll
          for (int i = 0; i < MT; i++) fq[0][i] = *reinterpret_cast<const double2 *>(ft + (8 * i + g) * FP + 2 * t);
          {
            const int off0 = g * KP + (((0 ^ (g & 1)) << 3) | (t << 1));
            bl[0] = *reinterpret_cast<const double2 *>(sl_lo + off0);
            bh[0] = *reinterpret_cast<const double2 *>(sl_hi + off0);
          }
#pragma unroll
          for (int j = 0; j < RT; j++) {
            double a1[MT][2], a2[MT][2];
#pragma unroll
            for (int i = 0; i < MT; i++) {
              const double2 f = fq[j & 1][i];
              a1[i][0] = w1[i] * f.x; a2[i][0] = w2[i] * f.x; a1[i][1] = w1[i] * f.y; a2[i][1] = w2[i] * f.y;
            }
            if (j + 1 < RT) {
#pragma unroll
              for (int i = 0; i < MT; i++) fq[(j + 1) & 1][i] = *reinterpret_cast<const double2 *>(ft + (8 * i + g) * FP + 2 * t + 8 * (j + 1));
            }
#pragma unroll
            for (int jj = 0; jj < RT; jj++) {
              const int cur = (j * RT + jj) & 1;
              if (jj + 1 < RT || j + 1 < RT) {
                const int jn = jj + 1 < RT ? j : j + 1, jjn = jj + 1 < RT ? jj + 1 : 0;
                const int offn = (8 * jjn + g) * KP + (((jn ^ (g & 1)) << 3) | (t << 1));
                bl[cur ^ 1] = *reinterpret_cast<const double2 *>(sl_lo + offn);
                bh[cur ^ 1] = *reinterpret_cast<const double2 *>(sl_hi + offn);
              }
              const double2 b1 = bl[cur], b2 = bh[cur];
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
        } else
#endif
        {
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
        }
        PT_MARK(2)
        // ---- the staged tile is consumed: launch the next tile's gather and scalar loads now ---------
        __syncwarp();
        if (ALT && lane == 0) st_volatile_s(upd_done + mw, (tile - t_begin) + 1);
        issue_gather(idQ[GD], nvQ[GD], ft);
        load_scalars(GD, idQ[GD], nvQ[GD]);
        // F' rows go out one 64-byte segment per pdf k-pair (below), so the stores trickle out under the DMMA stream
        double *Fo[MT];
#pragma unroll
        for (int i = 0; i < MT; i++) Fo[i] = a.F + (size_t)mrow[0][i] * a.ldf + 2 * t;
        const bool do_store = !a.last;
        PT_MARK(3)
        // ---- (2) conditional pdf on the grid of dimension k+1 ----------------------------------------
#if TTIRT_SWPIPE
        if (EXACT) {
          double2 bq[2];
          bq[0] = *reinterpret_cast<const double2 *>(Ps + g * KP + (((0 ^ (g & 1)) << 3) | (t << 1)));
#pragma unroll
          for (int jj = 0; jj < RT; jj++) {
#pragma unroll
            for (int jn = 0; jn < NTD; jn++) {
              const int cur = (jj * NTD + jn) & 1;
              if (jn + 1 < NTD || jj + 1 < RT) {
                const int jjn = jn + 1 < NTD ? jj : jj + 1, jnn = jn + 1 < NTD ? jn + 1 : 0;
                bq[cur ^ 1] = *reinterpret_cast<const double2 *>(Ps + (8 * jnn + g) * KP + (((jjn ^ (g & 1)) << 3) | (t << 1)));
              }
              const double2 bv = bq[cur];
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(c[i][jn][0], c[i][jn][1], acc[i][jj][0], bv.x);
#pragma unroll
              for (int i = 0; i < MT; i++) dmma884(c[i][jn][0], c[i][jn][1], acc[i][jj][1], bv.y);
            }
            if (do_store) {
#pragma unroll
              for (int i = 0; i < MT; i++)
                if (8 * i + g < nvalid) *reinterpret_cast<double2 *>(Fo[i] + 8 * jj) = make_double2(acc[i][jj][0], acc[i][jj][1]);
            }
          }
        } else
#endif
        {
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
            c[i][NT - 1][0] = tl[i];
          }
        }
      } else {
        // this warp has no rows in this tile: keep the pipeline moving
        if (ALT && lane == 0) st_volatile_s(upd_done + mw, (tile - t_begin) + 1);
        issue_gather(idQ[GD], nvQ[GD], ft);
        load_scalars(GD, idQ[GD], nvQ[GD]);
      }
      PT_MARK(4)
      // ids of the tile GD + 1 ahead
      int nvNew = 0;
      const int rowNew = rows_of(tile + GD + 1, nvNew);
      const int idNew = load_ids(rowNew, nvNew);

      // ---- (3) park the signed pdf tile for the tail warp (it takes |.|, reference :105); the two producers of a buffer alternate ----
      {
        // producer 0 waits for the tail warp to have released producer 1's previous tile, and vice versa
        if (EMPTY_BARS) {
          if (prod > 0 || tile > t_begin) bar_pair_sync(6 + 4 * prod + tw);
        } else {
          if (lane == 0)
            while (ld_volatile_s(consumed + tw) < PRODS * (tile - t_begin) + prod) __nanosleep(20);   // every earlier use of the buffer released
          __syncwarp();
        }
        PT_MARK(5)
        if (nvalid > 0) {
#pragma unroll
          for (int i = 0; i < MT; i++) {
#pragma unroll
            for (int jn = 0; jn < NTD; jn++) {
              pbw[(8 * jn + 2 * t) * RS + 8 * i + g] = c[i][jn][0];
              pbw[(8 * jn + 2 * t + 1) * RS + 8 * i + g] = c[i][jn][1];
            }
            if (TAIL1 && t == 0) pbw[(8 * (NT - 1)) * RS + 8 * i + g] = c[i][NT - 1][0];
          }
          if (lane < WROWS) ids_all[tw * WROWS + lane] = idC;
        }
        if (lane == 0) nv_all[tw] = nvalid;
        bar_pair_arrive(2 + tw);                    // FULL
      }

      PT_MARK(6)
      // ---- rotate the row pipeline --------------------------------------------------------------------
      nvC = nvQ[1]; idC = idQ[1];
#pragma unroll
      for (int j = 1; j < GD; j++) { nvQ[j] = nvQ[j + 1]; idQ[j] = idQ[j + 1]; }
      nvQ[GD] = nvNew; idQ[GD] = idNew;
#pragma unroll
      for (int j = 0; j < GD; j++)
#pragma unroll
        for (int i = 0; i < MT; i++) { mrow[j][i] = mrow[j + 1][i]; w1s[j][i] = w1s[j + 1][i]; w2s[j][i] = w2s[j + 1][i]; }
    }
    PT_FLUSH
  }

  if (!a.last) {
    __syncthreads();
    for (int i = tid; i < n1 - 1; i += NTHR)
      if (hist[i]) atomicAdd(a.hist_next + i, hist[i]);
  }
}

template <int RT, int NT, bool EXACT, bool TAIL1, int MW, int TW, int GD>
cudaError_t launch_variant(const TransArgs &a, int sm_count, cudaStream_t st) {
  using L = SmemLayout<RT, NT, TAIL1, MW, TW, GD>;
  constexpr int ROWS_CTA = MW * WROWS;
  int64_t max_tiles = ((int64_t)a.rows + ROWS_CTA - 1) / ROWS_CTA + (a.n0 - 1);
  int64_t grid = sm_count;   // persistent: one CTA per SM (the register file allows no more)
  if (grid > max_tiles) grid = max_tiles;
  if (grid < 1) grid = 1;
  transition_kernel<RT, NT, EXACT, TAIL1, MW, TW, GD><<<(unsigned)grid, nthr_of(MW, TW), L::bytes, st>>>(a);
  return cudaGetLastError();
}

template <int RT, int NT, int MW, int TW, int GD>
cudaError_t launch_one(const TransArgs &a, int sm_count, cudaStream_t st) {
  const bool exact = a.r0 == 8 * RT && a.r1 == 8 * RT && (a.n1 + 7) / 8 == NT;
  if (exact && a.n1 == 8 * (NT - 1) + 1) return launch_variant<RT, NT, true, true, MW, TW, GD>(a, sm_count, st);
  if (exact) return launch_variant<RT, NT, true, false, MW, TW, GD>(a, sm_count, st);
  return launch_variant<RT, NT, false, false, MW, TW, GD>(a, sm_count, st);
}

template <int RT, int NT, int MW, int TW, int GD>
cudaError_t init_one() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, true, true, MW, TW, GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout<RT, NT, true, MW, TW, GD>::bytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(transition_kernel<RT, NT, true, false, MW, TW, GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout<RT, NT, false, MW, TW, GD>::bytes)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(transition_kernel<RT, NT, false, false, MW, TW, GD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemLayout<RT, NT, false, MW, TW, GD>::bytes);
}

#ifndef TTIRT_LIGHT_MW
#define TTIRT_LIGHT_MW 8    // MMA warps of the two lighter classes: 8 with eight tail warps; 12 with four (measured: the one tail warp per sub-partition cannot keep up with three MMA warps, 65 against 101 M samples/s at d=40 n=33 r=32)
#endif
#ifndef TTIRT_LIGHT_TW
#define TTIRT_LIGHT_TW (TTIRT_LIGHT_MW == 12 ? 4 : 8)    // tail warps of the two lighter classes
#endif
constexpr int kLightMW = TTIRT_LIGHT_MW, kLightTW = TTIRT_LIGHT_TW;
#ifndef TTIRT_LIGHT_GD
#define TTIRT_LIGHT_GD 2    // gather depth of the two lighter classes
#endif
constexpr int kLightGD = TTIRT_LIGHT_GD;

}  // namespace

int fast_class_for(int rmax, int nmax) {
  if (rmax <= 16 && nmax <= 24) return 0;
  if (rmax <= 32 && nmax <= 40) return 1;
  if (rmax <= 64 && nmax <= 72) return 2;
  // anything larger: the unfused DMMA path of ttirt_wide.cu (TTIRT_WIDE=0: the strict kernel serves these shapes, as before)
  static const bool wide_on = !(getenv("TTIRT_WIDE") && atoi(getenv("TTIRT_WIDE")) == 0);
  if (wide_on && rmax <= 1024 && nmax <= 1024) return kWideClass;
  return -1;
}


cudaError_t fast_init(int) {
  cudaError_t e;
  if ((e = init_one<2, 3, kLightMW, kLightTW, kLightGD>()) != cudaSuccess) return e;
  if ((e = init_one<4, 5, kLightMW, kLightTW, kLightGD>()) != cudaSuccess) return e;
  if ((e = init_one<8, 9, 8, 4, 1>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_transition(int cls, const TransArgs &a, int sm_count, cudaStream_t st) {
  switch (cls) {
    case 0: return launch_one<2, 3, kLightMW, kLightTW, kLightGD>(a, sm_count, st);
    case 1: return launch_one<4, 5, kLightMW, kLightTW, kLightGD>(a, sm_count, st);
    case 2: return launch_one<8, 9, 8, 4, 1>(a, sm_count, st);
    default: return cudaErrorInvalidValue;
  }
}

#ifdef TTIRT_PHASE_TIMING
void trace_read(long long *out) { cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 2 * 24 * 8); }
void phase_cycles_read(unsigned long long *out) {
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 40);
  unsigned long long z[40] = {0};
  cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
}
#endif

}  // namespace ttirt
