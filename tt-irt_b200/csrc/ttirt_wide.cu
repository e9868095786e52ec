// Fast path of tt_irt1 for WIDE shapes on B200 (sm_100a): ranks or grids beyond what the fused transition kernel
// (ttirt_fast.cu) can hold in shared memory (r > 64 or n > 72).  The reference serves every shape on one code path
// (tt_irt1_int32.c:41-53); here such shapes used to fall to the one-thread-per-sample strict kernel.
//
// At these sizes the two contractions dominate everything else by a wide margin (4 r_k r_{k+1} + 2 r_{k+1} n_{k+1} flops
// per sample and dimension against O(n) for the CDF and the inversion), so the step is NOT fused: per dimension
//   (1) wide_gemm_kernel<true>   gathered, interpolated interface update (reference :167-177) as a grouped GEMM: the
//                                samples arrive ordered by the interval b chosen in dimension k (the counting sort of the
//                                per-dimension path), a CTA tile never straddles two intervals, and
//                                  F' = (w1 F) A_b + (w2 F) A_{b+1},   A_i = core_k[:, i, :]
//                                is one FP64 tensor-core product (DMMA, mma.sync m8n8k4.f64) over K = r_k in two phases.
//                                Different column tiles of F' read the same rows of F, so F is double-buffered in HBM.
//   (2) wide_gemm_kernel<false>  conditional pdf on the grid of dimension k+1 (reference :103-105) with the trapezoid node
//                                weights folded into P_{k+1}'s columns: V = F' Pw_{k+1}, written node-major (n x rows) so
//                                that step (3) reads it coalesced; its epilogue also leaves every row's mass sum_j |v_j| in
//                                a few shares (fixed summation order), so that step (3) reads the nodes once.
//   (3) wide_tail_kernel         one thread per sample: unnormalised search on the running sums (:107-142), closed-form
//                                quadratic inversion (:146-159), log-density in split form (:161-165), interval histogram
//                                for the next counting sort.  Same scaled formulation as the walk kernel's tail
//                                (ttirt_walk.cu), the nodes streamed from HBM instead of held in registers.
// HBM traffic per sample and dimension: F in (the column tiles of a row tile run side by side and share it in L2) + F' out
// + F' in + V out + V in ~ 8 (3 r + 2 n) bytes against 4 r^2 + 2 r n flops: FP64 tensor pipe bound from r ~ 48 on.
// Measured (profiles/r02_wide.md): 23 TFLOP/s = 62 % of the DMMA peak at d=8 n=129 r=128, 28 x the strict kernel.
#include <cstdio>
#include <type_traits>

#include "ttirt_common.cuh"

namespace ttirt {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// CTA tile: 64 rows x up to 16 NJ columns (NJ 8-column groups per warp at most), K in slices of W_KS; four warps in a 2 x 2
// grid, each 32 rows x up to 8 NJ columns (4 x NJ DMMA tiles: 16 - 20 independent accumulator chains).  The operands go
// through a ring of W_STAGES shared-memory stages filled by cp.async one slice ahead, ONE barrier per slice, and three CTAs
// per SM cover each other's barriers.  Measured alternatives (profiles/r02_wide.md): four CTAs per SM, three stages, slices
// of 32 -- all slower: next to a DMMA stream fewer non-DMMA instructions help, more resident warps do not.
#ifndef TTIRT_WIDE_KS
#define TTIRT_WIDE_KS 16
#endif
constexpr int W_TM = 64, W_KS = TTIRT_WIDE_KS, W_THREADS = 128;
constexpr int W_HK = W_KS / 2;            // k per thread and slice of the A tile (two threads per row)
constexpr int W_CPT = W_THREADS / W_KS;   // B columns a pass of the CTA covers (lanes along k)
static_assert(W_KS == 16 || W_KS == 32, "slices of 16 or 32");
constexpr int W_NJ_UPDATE = 4, W_NJ_PDF = 5;
// row pitch of the staged operands in doubles: 24 = 8 (mod 16), so the LDS.128 fragment loads of a quarter warp (rows g, g+1,
// k pairs 2t) fall into eight different 16-byte slots of the 128-byte bank window: conflict-free without a swizzle
constexpr int W_PITCH = W_KS + 8;
#ifndef TTIRT_WIDE_CTAS
#define TTIRT_WIDE_CTAS 3
#endif
#ifndef TTIRT_WIDE_STAGES
#define TTIRT_WIDE_STAGES 2
#endif
constexpr int W_STAGES = TTIRT_WIDE_STAGES, W_CTAS = TTIRT_WIDE_CTAS;
constexpr size_t wide_smem_bytes(int nj) { return sizeof(double) * W_STAGES * (W_TM + 16 * nj) * W_PITCH; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// cp.async (LDGSTS) global -> shared: 16-byte (A rows) and 8-byte (B columns) pieces are issued inline by issue_slice();
// src-size 0 zero-fills a piece (src still a mapped address)
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// -DTTIRT_WIDE_CHECK: every global address the GEMM touches is checked against the operand extents (compute-sanitizer is not
// available on the GPU pool); a violation prints one line and traps
#ifdef TTIRT_WIDE_CHECK
#define W_CHECK(cond, what) do { if (!(cond)) { printf("ttirt_wide: bounds violation (%s) block %u thread %u\n", what, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define W_CHECK(cond, what) do { } while (0)
#endif

struct WideGemmArgs {
  const double *A;        // rows of the A operand: A + id * lda, valid and zero-padded up to K rounded to 8
  int lda;
  const int *perm;        // UPDATE: sorted position -> sample id
  const int *hist;        // UPDATE: samples per interval chosen in dimension k
  int nb;                 // UPDATE: number of intervals, n_k - 1
  const double *w1, *w2;  // UPDATE: per-sample interpolation weights
  int rows;
  const double *B;        // B(kk, c) = B[kk + c * ldb]; UPDATE: phase p of interval b starts at B + (b + p) * K
  int64_t ldb;
  int K, N;               // contraction length per phase, output columns
  int ncol;               // column tiles (grid = row tiles x ncol, column tile fastest)
  int gpt;                // 8-column groups per column tile (<= 2 NJ; the last tile may hold fewer)
  double *C;              // UPDATE: C[id * ldc + c] for c < N rounded to 8; else C[c * ldc + position] for c < N
  int64_t ldc;
  int64_t a_elems, b_elems, c_elems, m_elems;   // extents of A, B, C and mass_part in doubles (checked builds)
  double *mass_part;      // pdf: per-row sums of |C| over the columns of one warp column, slot-major: [(2 tile + wc) * rows + position]
};

// The contraction index is consumed in a permuted order on BOTH operands (lane t of a quad takes k = 8j + 2t and 8j + 2t + 1
// for the two k-steps of an 8-block), so one LDS.128 per operand feeds two DMMAs.
// NJ: 8-column groups per warp at most (a column tile is up to 16 NJ columns wide).  The column groups of a launch are
// spread evenly over its column tiles and the groups of a tile over its two warp columns, so a 2^p + 1 grid costs one
// extra group in one tile (NJ = 5: 129 nodes are tiles of 9 and 8 groups) instead of a column tile of its own.
template <bool UPDATE, int NJ>
__global__ void __launch_bounds__(W_THREADS, W_CTAS) wide_gemm_kernel(const WideGemmArgs a) {
  extern __shared__ __align__(16) unsigned char wide_smem[];
  double *As = reinterpret_cast<double *>(wide_smem);          // W_STAGES stages of the A tile, then of the B tile
  double *Bs = As + W_STAGES * W_TM * W_PITCH;
  constexpr int TN = 16 * NJ;                                  // columns of the B tile
  __shared__ double sc[2][W_TM];
  __shared__ int ids[W_TM];
  __shared__ int tile_info[3];   // interval, first (sorted) row, valid rows (0: no such tile)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // column tiles of one row tile are neighbours in the grid: they run at the same time and share the tile's A rows in L2
  const int T = (int)(blockIdx.x / (unsigned)a.ncol), c0 = (int)(blockIdx.x % (unsigned)a.ncol) * a.gpt * 8;
  const int Nout = UPDATE ? ((a.N + 7) & ~7) : a.N;            // output columns (UPDATE writes the zero padding too)
  const int gt = min(a.gpt, (Nout - c0 + 7) >> 3);             // column groups of this tile

  // ---- which rows: tile T of the chunk.  UPDATE: tiles are numbered interval by interval (none straddles two), the warp
  //      scans the histogram for the interval that holds tile T ----
  if (UPDATE) {
    if (warp == 0) {
      int cs = 0, ct = 0;
      bool found = false;
      for (int b0 = 0; b0 < a.nb && !found; b0 += 32) {
        const int b = b0 + lane;
        const int c = b < a.nb ? a.hist[b] : 0;
        const int t = (c + W_TM - 1) / W_TM;
        int is = c, it = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int us = __shfl_up_sync(FULL, is, o), ut = __shfl_up_sync(FULL, it, o);
          if (lane >= o) { is += us; it += ut; }
        }
        const int first = ct + it - t;                       // first tile of interval b
        const bool hit = t > 0 && T >= first && T < first + t;
        if (__ballot_sync(FULL, hit) != 0u) {
          found = true;
          if (hit) {
            const int local = T - first;
            tile_info[0] = b;
            tile_info[1] = cs + is - c + local * W_TM;
            tile_info[2] = min(W_TM, c - local * W_TM);
          }
        }
        cs += __shfl_sync(FULL, is, 31);
        ct += __shfl_sync(FULL, it, 31);
      }
      if (!found && lane == 0) tile_info[2] = 0;
    }
  } else if (tid == 0) {
    const int64_t r0 = (int64_t)T * W_TM;
    tile_info[0] = 0;
    tile_info[1] = (int)r0;
    tile_info[2] = r0 < a.rows ? (int)min((int64_t)W_TM, (int64_t)a.rows - r0) : 0;
  }
  __syncthreads();
  const int bin = tile_info[0], row0 = tile_info[1], nv = tile_info[2];
  if (nv <= 0) return;

  if (tid < W_TM) {
    int id = -1;
    double s1 = 0.0, s2 = 0.0;
    if (tid < nv) {
      id = UPDATE ? a.perm[row0 + tid] : row0 + tid;
      if (UPDATE) { s1 = a.w1[id]; s2 = a.w2[id]; } else { s1 = 1.0; }
    }
    ids[tid] = id; sc[0][tid] = s1; sc[1][tid] = s2;
  }
  __syncthreads();

  // ---- staging by cp.async (LDGSTS) into a ring of W_STAGES shared-memory stages, one slice ahead of the DMMA loop.
  //      Written for instruction count: next to a DMMA stream every other instruction of the sub-partition waits for a gap
  //      (ttirt_fast.cu), so the per-slice address arithmetic is a handful of adds on bases prepared here.
  //      A: thread (srow, h) moves 8 consecutive k of row srow as four 16-byte pieces (rows are 64-byte aligned);
  //      B: 8-byte pieces (a slab column may start at an odd element), lanes along k, so that a warp instruction reads two
  //         128-byte runs; a thread's eight columns are a constant byte stride apart.
  //      Pieces outside the operands are zero-filled (src-size 0, any mapped address). ----
  const int srow = tid >> 1, h = tid & 1;
  const int my_id = ids[srow];
  const int K8 = (a.K + 7) & ~7;
  const double *a_src = a.A + (size_t)(my_id >= 0 ? my_id : 0) * a.lda + W_HK * h;  // + k0 per slice
  const int a_lim = my_id >= 0 ? K8 - W_HK * h : 0;                                 // the 8-block at k0 + 8 b is valid while k0 + 8 b < a_lim (K8, k0: multiples of 8)
  const uint32_t a_dst = smem_u32(As + srow * W_PITCH + W_HK * h);
  const int bk = tid & (W_KS - 1), bc = tid / W_KS;                                 // B: element k0 + bk of columns c0 + bc + W_CPT e
  const double *b_src = a.B + (UPDATE ? (int64_t)bin * a.K : 0) + bk + (int64_t)min(c0 + bc, a.N - 1) * a.ldb;   // + phase * K + k0 per slice
  const uint32_t b_stride = (uint32_t)(W_CPT * a.ldb * sizeof(double));                 // bytes between a thread's columns (< 2^32: ldb <= 2^20)
  const uint32_t b_dst = smem_u32(Bs + bc * W_PITCH + bk);
  int b_cols = 0;                                                                   // how many of the thread's columns bc + 8 e exist in this tile
#pragma unroll
  for (int e = 0; e < TN / W_CPT; e++) b_cols += (bc + W_CPT * e < 8 * gt && c0 + bc + W_CPT * e < a.N) ? 1 : 0;
  const int b_lim = a.K - bk;                                                       // element valid while k0 < b_lim
  const int nks = (a.K + W_KS - 1) / W_KS;
  const int nsl = (UPDATE ? 2 : 1) * nks;
  auto issue_slice = [&](int s) {
    if (s < nsl) {
      const int p = (UPDATE && s >= nks) ? 1 : 0;
      const int k0 = (s - p * nks) * W_KS;
      const uint32_t st_a = (uint32_t)((s % W_STAGES) * (W_TM * W_PITCH) * sizeof(double)), st_b = (uint32_t)((s % W_STAGES) * (TN * W_PITCH) * sizeof(double));
      {
        const char *src = reinterpret_cast<const char *>(a_src + k0);
#pragma unroll
        for (int b = 0; b < W_HK / 8; b++) {
          const int sz = k0 + 8 * b < a_lim ? 16 : 0;
          W_CHECK(!sz || ((a_src + k0 + 8 * b) - a.A >= 0 && (a_src + k0 + 8 * b + 8) - a.A <= a.a_elems), "A");
#pragma unroll
          for (int u = 0; u < 4; u++)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(a_dst + st_a + 64 * b + 16 * u), "l"(sz ? src + 64 * b + 16 * u : src), "r"(sz) : "memory");
        }
      }
      {
        const int live = k0 < b_lim ? b_cols : 0;
        const char *src = reinterpret_cast<const char *>(b_src + (UPDATE ? p * a.K : 0) + k0);
#ifdef TTIRT_WIDE_CHECK
        for (int e = 0; e < live; e++) {
          const int64_t off = (reinterpret_cast<const double *>(src + (size_t)e * b_stride)) - a.B;
          W_CHECK(off >= 0 && off < a.b_elems, "B");
        }
#endif
#pragma unroll
        for (int e = 0; e < TN / W_CPT; e++)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(b_dst + st_b + (uint32_t)(W_CPT * e * W_PITCH * sizeof(double))),
                       "l"(e < live ? src + (size_t)e * b_stride : src), "r"(e < live ? 8 : 0) : "memory");
      }
    }
    cp_async_commit();   // (an empty group past the last slice keeps the group count in step)
  };

  const int g = lane >> 2, t = lane & 3, wr = warp >> 1, wc = warp & 1;
  // the tile's column groups are split between the two warp columns: the first takes the larger half
  const int j_first = wc ? (gt + 1) >> 1 : 0, jmax = wc ? gt >> 1 : (gt + 1) >> 1;
  const bool warp_has_rows = 32 * wr < nv && jmax > 0;
  double acc[4][NJ][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  // per-row interpolation weights of this lane's four row tiles (UPDATE): applied to the A fragments
  double wsc[2][4];
#pragma unroll
  for (int i = 0; i < 4; i++) { wsc[0][i] = sc[0][32 * wr + 8 * i + g]; wsc[1][i] = sc[1][32 * wr + 8 * i + g]; }
  const double *a_frag = As + (32 * wr + g) * W_PITCH + 2 * t, *b_frag = Bs + (8 * j_first + g) * W_PITCH + 2 * t;

  // one slice of the contraction from ring stage `stage` for a warp with JM column groups (compile time); JM = 0: nothing
  // to do; JM = -1: jmax groups, decided at run time
  auto compute = [&](auto jm_tag, int stage, int ph) {
    constexpr int JM = decltype(jm_tag)::value;
    if (JM == 0) return;
    constexpr int JN = JM <= 0 ? NJ : JM;
    const double *as = a_frag + stage * (W_TM * W_PITCH), *bs = b_frag + stage * (TN * W_PITCH);
#pragma unroll
    for (int jb = 0; jb < W_KS / 8; jb++) {
      double2 av[4], bv[JN];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        av[i] = *reinterpret_cast<const double2 *>(as + 8 * i * W_PITCH + 8 * jb);
        if (UPDATE) { const double w = ph ? wsc[1][i] : wsc[0][i]; av[i].x *= w; av[i].y *= w; }
      }
#pragma unroll
      for (int j = 0; j < JN; j++) bv[j] = *reinterpret_cast<const double2 *>(bs + 8 * j * W_PITCH + 8 * jb);
      if (JM > 0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < JN; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].x, bv[j].x);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < JN; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].y, bv[j].y);
      } else {
#pragma unroll
        for (int j = 0; j < JN; j++) {
          if (j < jmax) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma884(acc[i][j][0], acc[i][j][1], av[i].x, bv[j].x);
#pragma unroll
            for (int i = 0; i < 4; i++) dmma884(acc[i][j][0], acc[i][j][1], av[i].y, bv[j].y);
          }
        }
      }
    }
  };
  auto mainloop = [&](auto jm_tag) {
#pragma unroll
    for (int s = 0; s < W_STAGES - 1; s++) issue_slice(s);
    for (int s = 0; s < nsl; s++) {
      cp_async_wait<W_STAGES - 2>();   // this thread's pieces of slice s have landed
      __syncthreads();                 // ... everybody's have, and every warp is done with slice s - 1 (whose stage is refilled next)
      issue_slice(s + W_STAGES - 1);
      compute(jm_tag, s % W_STAGES, (UPDATE && s >= nks) ? 1 : 0);
    }
  };
  if (!warp_has_rows) mainloop(std::integral_constant<int, 0>());
  else if (jmax == NJ) mainloop(std::integral_constant<int, NJ>());
  else if (jmax == NJ - 1) mainloop(std::integral_constant<int, NJ - 1>());
  else mainloop(std::integral_constant<int, -1>());

  // ---- epilogue: lane (g, t) holds C[8i + g][8j + 2t], C[8i + g][8j + 2t + 1] ----
  if (UPDATE ? !warp_has_rows : 32 * wr >= nv) return;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = 32 * wr + 8 * i + g;
    if (UPDATE) {
      if (r >= nv) continue;
      double *crow = a.C + (size_t)ids[r] * a.ldc;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const int c = c0 + 8 * (j_first + j) + 2 * t;
        W_CHECK(j >= jmax || ((crow + c) - a.C >= 0 && (crow + c + 2) - a.C <= a.c_elems), "C update");
        if (j < jmax) *reinterpret_cast<double2 *>(crow + c) = make_double2(acc[i][j][0], acc[i][j][1]);   // columns N .. N8-1: exact zeros
      }
    } else {
      // the tail needs the row's mass sum_j |v_j| before it can search: every warp column leaves its share here, summed in a
      // fixed order (columns ascending per lane, then the quad's four lanes), so the result does not depend on scheduling
      const bool live = r < nv;
      double *cpos = a.C + (row0 + r);
      double psum = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        const int c = c0 + 8 * (j_first + j) + 2 * t;
        W_CHECK(!(live && j < jmax && c < a.N) || (cpos + (int64_t)(c + (c + 1 < a.N ? 1 : 0)) * a.ldc) - a.C < a.c_elems, "C pdf");
        if (live && j < jmax && c < a.N) { cpos[(int64_t)c * a.ldc] = acc[i][j][0]; psum += fabs(acc[i][j][0]); }
        if (live && j < jmax && c + 1 < a.N) { cpos[(int64_t)(c + 1) * a.ldc] = acc[i][j][1]; psum += fabs(acc[i][j][1]); }
      }
      psum += __shfl_xor_sync(FULL, psum, 1);
      psum += __shfl_xor_sync(FULL, psum, 2);
      W_CHECK(!live || (int64_t)(2 * (blockIdx.x % (unsigned)a.ncol) + wc) * a.rows + row0 + r < a.m_elems, "mass");
      if (live && t == 0) a.mass_part[(size_t)(2 * (blockIdx.x % (unsigned)a.ncol) + wc) * a.rows + row0 + r] = psum;
    }
  }
}

struct WideTailArgs {
  const double *pb;        // weighted signed pdf, node-major: pb[j * ldp + m]
  int64_t ldp;
  const double *mass_part; // nslots shares of the row's mass, slot-major (left by the pdf GEMM): [s * rows + m]
  int nslots;
  const double *x, *ih, *rw, *hr;   // grid of dimension k+1, 1 / cell width, 1 / node weight, h_{j-1} / node weight
  int n1, rows, last;
  const double *q;
  double *z;
  int32_t *idx_out;
  double *lpz;
  int *idx;
  double *w1, *w2, *lp, *lpd;
  int *lpe;
  int *hist_next;
};

// v_j = w_j |p_j| (w_j the trapezoid node weight, see node_weight()):  cdf_j = R_j + (h_{j-1} / w_j) v_j with
// R_j = sum_{i<j} v_i, mass = R_n (summed by the pdf GEMM's epilogue: one pass over the nodes here instead of two).  Largest i0 <= n-2 with cdf_{i0} < q * mass (reference :134-142 on the unnormalised
// CDF; it is monotone, so the last node that passes the test is the answer), then the walk kernel's tail verbatim.
__global__ void __launch_bounds__(256) wide_tail_kernel(const WideTailArgs a) {
  extern __shared__ int sh[];   // n1 - 1 interval counters
  const int n1 = a.n1;
  if (!a.last) {
    for (int i = threadIdx.x; i < n1 - 1; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < a.rows) {
    const double *pv = a.pb + m;
    const int64_t ld = a.ldp;
    const double qv = a.q[m];
    double total = 0.0;
    for (int s = 0; s < a.nslots; s++) total += a.mass_part[(size_t)s * a.rows + m];
    const double qt = qv * total;
    int i0 = 0;
    double dq = qt, Rj = fabs(pv[0]);
    // eight nodes at a time: the loads of a batch are independent of the running sum and all in flight at once
    int j = 1;
    for (; j + 7 <= n1 - 2; j += 8) {
      double v[8], hrj[8];
#pragma unroll
      for (int u = 0; u < 8; u++) { v[u] = fabs(pv[(j + u) * ld]); hrj[u] = a.hr[j + u]; }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double dj = qt - fma(hrj[u], v[u], Rj);
        if (__double_as_longlong(dj) > 0) { i0 = j + u; dq = dj; }
        Rj += v[u];
      }
    }
    for (; j <= n1 - 2; j++) {
      const double v = fabs(pv[j * ld]);
      const double dj = qt - fma(a.hr[j], v, Rj);
      if (__double_as_longlong(dj) > 0) { i0 = j; dq = dj; }
      Rj += v;
    }
    const double va = fabs(pv[i0 * ld]), vb = fabs(pv[(i0 + 1) * ld]);
    const double s2 = pow2_scale(total);                 // exact power-of-two normalisation instead of 1 / mass
    double c1 = va * a.rw[i0] * s2, c2 = vb * a.rw[i0 + 1] * s2;
    double mass = total * s2;
    dq *= s2;
    if (total == 0.0) {
      // zero-mass conditional: uniform in index space (reference tt_irt1_int32.c:116-125)
      const double u = 1.0 / (double)(n1 - 1);
      const double sf = 1.0 / ((double)(n1 - 1) * u);
      int k0 = 0;
      for (int j = 1; j <= n1 - 2; j++) k0 += (qv > ((double)j * u) * sf) ? 1 : 0;
      i0 = k0; dq = qv - ((double)k0 * u) * sf; c1 = u * sf; c2 = u * sf; mass = 1.0;
    }
    const CellFast o = invert_cell_fast(dq, c1, c2, a.x[i0], a.x[i0 + 1], a.ih[i0]);
    double lpN = a.lp[m], lpD = a.lpd[m];
    int lpE = a.lpe[m];
    lp_accumulate(lpN, lpD, lpE, o.dens, mass);
    a.z[m] = o.xk;
    if (a.idx_out) a.idx_out[m] = i0;
    if (!a.last) {
      a.idx[m] = i0; a.w1[m] = o.w1; a.w2[m] = o.w2;
      a.lp[m] = lpN; a.lpd[m] = lpD; a.lpe[m] = lpE;
      atomicAdd(&sh[i0], 1);
    } else {
      a.lpz[m] = lp_finish(lpN, lpD, lpE);
    }
  }
  if (!a.last) {
    __syncthreads();
    for (int i = threadIdx.x; i < n1 - 1; i += blockDim.x)
      if (sh[i]) atomicAdd(a.hist_next + i, sh[i]);
  }
}

// per-dimension grid tables of the tail, in the layout of xs: 1 / cell width, 1 / node weight, h_{j-1} / node weight
__global__ void wide_tables_kernel(const DimInfo *__restrict__ dims, const double *__restrict__ xs, double *ih, double *rw, double *hr) {
  const DimInfo di = dims[blockIdx.x];
  const double *x = xs + di.off_x;
  for (int i = threadIdx.x; i < di.n; i += blockDim.x) {
    const double w = node_weight(x, i, di.n);
    const double hl = i >= 1 ? 0.5 * (x[i] - x[i - 1]) : 0.0;
    ih[di.off_x + i] = i + 1 < di.n ? 1.0 / (x[i + 1] - x[i]) : 0.0;
    rw[di.off_x + i] = w > 0.0 ? 1.0 / w : 0.0;
    hr[di.off_x + i] = w > 0.0 ? hl / w : 0.0;
  }
}

}  // namespace

// column tiling of `groups` 8-column groups with at most 2 nj groups per tile: as few tiles as possible, evenly filled
static void wide_col_tiles(int groups, int nj, int &ncol, int &gpt) {
  ncol = (groups + 2 * nj - 1) / (2 * nj);
  if (ncol < 1) ncol = 1;
  gpt = (groups + ncol - 1) / ncol;
}

cudaError_t wide_tables(const DimInfo *d_dims, int d, const double *xs, double *ih, double *rw, double *hr, cudaStream_t st) {
  wide_tables_kernel<<<d, 128, 0, st>>>(d_dims, xs, ih, rw, hr);
  return cudaGetLastError();
}

// shares of a row's mass the pdf GEMM leaves for the tail (two warp columns per column tile), for grids up to nmax nodes
int wide_mass_slots(int nmax) {
  int ncol, gpt;
  wide_col_tiles((nmax + 7) >> 3, W_NJ_PDF, ncol, gpt);
  return 2 * ncol;
}

cudaError_t wide_init(int) {
  cudaError_t e = cudaFuncSetAttribute(wide_gemm_kernel<true, W_NJ_UPDATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(W_NJ_UPDATE));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(wide_gemm_kernel<false, W_NJ_PDF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem_bytes(W_NJ_PDF));
}

// One dimension step k -> k+1 of the wide path: three launches (update, pdf, tail) on st.
cudaError_t launch_wide_step(const WideArgs &w, cudaStream_t st) {
  const int row_tiles = (w.rows + W_TM - 1) / W_TM;
  int pdf_ncol = 1;
  {
    WideGemmArgs g;
    g.A = w.Fin; g.lda = w.ldf; g.perm = w.perm; g.hist = w.hist_cur; g.nb = w.n0 - 1; g.w1 = w.w1; g.w2 = w.w2; g.rows = w.rows;
    g.B = w.core; g.ldb = (int64_t)w.r0 * w.n0; g.K = w.r0; g.N = w.r1;
    g.C = w.Fout; g.ldc = w.ldf; g.mass_part = nullptr;
    g.a_elems = (int64_t)w.rows * w.ldf; g.b_elems = (int64_t)w.r0 * w.n0 * w.r1; g.c_elems = (int64_t)w.rows * w.ldf; g.m_elems = 0;
    wide_col_tiles((w.r1 + 7) >> 3, W_NJ_UPDATE, g.ncol, g.gpt);
    const unsigned grid = (unsigned)(row_tiles + (w.n0 - 1)) * (unsigned)g.ncol;   // at most rows / 64 + one ragged tile per interval
    wide_gemm_kernel<true, W_NJ_UPDATE><<<grid, W_THREADS, wide_smem_bytes(W_NJ_UPDATE), st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    WideGemmArgs g;
    g.A = w.Fout; g.lda = w.ldf; g.perm = nullptr; g.hist = nullptr; g.nb = 0; g.w1 = g.w2 = nullptr; g.rows = w.rows;
    g.B = w.pnext; g.ldb = w.r1; g.K = w.r1; g.N = w.n1;
    g.C = w.pb; g.ldc = w.rows; g.mass_part = w.mass_part;
    g.a_elems = (int64_t)w.rows * w.ldf; g.b_elems = (int64_t)w.r1 * w.n1; g.c_elems = (int64_t)w.rows * w.n1;
    wide_col_tiles((w.n1 + 7) >> 3, W_NJ_PDF, g.ncol, g.gpt);
    pdf_ncol = g.ncol;
    g.m_elems = (int64_t)2 * g.ncol * w.rows;
    const unsigned grid = (unsigned)row_tiles * (unsigned)g.ncol;
    wide_gemm_kernel<false, W_NJ_PDF><<<grid, W_THREADS, wide_smem_bytes(W_NJ_PDF), st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    WideTailArgs t;
    t.pb = w.pb; t.ldp = w.rows; t.mass_part = w.mass_part; t.nslots = 2 * pdf_ncol; t.x = w.xnext; t.ih = w.ihnext; t.rw = w.rwnext; t.hr = w.hrnext;
    t.n1 = w.n1; t.rows = w.rows; t.last = w.last;
    t.q = w.q; t.z = w.z; t.idx_out = w.idx_out; t.lpz = w.lpz;
    t.idx = w.idx; t.w1 = w.w1; t.w2 = w.w2; t.lp = w.lp; t.lpd = w.lpd; t.lpe = w.lpe; t.hist_next = w.hist_next;
    wide_tail_kernel<<<(unsigned)((w.rows + 255) / 256), 256, sizeof(int) * (size_t)(w.n1 > 1 ? w.n1 - 1 : 1), st>>>(t);
    return cudaGetLastError();
  }
}

}  // namespace ttirt
