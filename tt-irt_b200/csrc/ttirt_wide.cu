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
//                                that step (3) reads it coalesced.
//   (3) wide_tail_kernel         one thread per sample: mass, unnormalised search on the running sums (:107-142), closed-form
//                                quadratic inversion (:146-159), log-density in split form (:161-165), interval histogram
//                                for the next counting sort.  Same scaled formulation as the walk kernel's tail
//                                (ttirt_walk.cu), the nodes streamed from HBM instead of held in registers.
// HBM traffic per sample and dimension: F in (once per column tile) + F' out + V out + V in twice ~ 8 (3 r + 3 n) bytes
// against 4 r^2 + 2 r n flops: FP64 tensor pipe bound from r ~ 48 on.
#include "ttirt_common.cuh"

namespace ttirt {

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// CTA tile 64 rows x 64 columns, K in slices of 16; four warps, each 32 x 32 (4 x 4 DMMA tiles: 16 independent
// accumulator chains).  One shared-memory stage; the next slice's global loads are in flight (registers) while the current
// slice is multiplied, and three CTAs per SM cover each other's barriers.
constexpr int W_TM = 64, W_TN = 64, W_KS = 16, W_THREADS = 128;
// row pitch of the staged operands in doubles: 24 = 8 (mod 16), so the LDS.128 fragment loads of a quarter warp (rows g, g+1,
// k pairs 2t) fall into eight different 16-byte slots of the 128-byte bank window: conflict-free without a swizzle
constexpr int W_PITCH = W_KS + 8;

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
  double *C;              // UPDATE: C[id * ldc + c] for c < N rounded to 8; else C[c * ldc + position] for c < N
  int64_t ldc;
};

// The contraction index is consumed in a permuted order on BOTH operands (lane t of a quad takes k = 8j + 2t and 8j + 2t + 1
// for the two k-steps of an 8-block), so one LDS.128 per operand feeds two DMMAs.
template <bool UPDATE>
__global__ void __launch_bounds__(W_THREADS, 3) wide_gemm_kernel(const WideGemmArgs a) {
  __shared__ __align__(16) double As[W_TM * W_PITCH];
  __shared__ __align__(16) double Bs[W_TN * W_PITCH];
  __shared__ double sc[2][W_TM];
  __shared__ int ids[W_TM];
  __shared__ int tile_info[3];   // interval, first (sorted) row, valid rows (0: no such tile)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = blockIdx.x, c0 = blockIdx.y * W_TN;

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

  // ---- staging: thread (srow, h) moves 8 consecutive k of row srow of the A tile and of column srow of the B tile, as four
  //      16-byte pieces in the rotated order (u + srow) & 3, which makes the shared-memory stores of a quarter warp
  //      conflict-free as well ----
  const int srow = tid >> 1, h = tid & 1;
  const int my_id = ids[srow];
  const double *arow = my_id >= 0 ? a.A + (size_t)my_id * a.lda : nullptr;
  const int K8 = (a.K + 7) & ~7;
  const int bcol = c0 + srow;
  const bool bcol_ok = bcol < a.N;
  const int nks = (a.K + W_KS - 1) / W_KS;
  const int nsl = (UPDATE ? 2 : 1) * nks;
  // The loaded values stay untouched in registers until they are stored (the interpolation weight is applied there): an
  // instruction that consumes them here would make the warp wait for the loads before its DMMA loop instead of after it.
  double2 ra[4], rb[4];
  auto load_slice = [&](int s) {
    const int p = (UPDATE && s >= nks) ? 1 : 0;
    const int k0 = (s - p * nks) * W_KS + 8 * h;
    const double *bp = a.B + (UPDATE ? (int64_t)(bin + p) * a.K : 0) + (int64_t)bcol * a.ldb;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int kk = k0 + 2 * ((u + srow) & 3);
      double2 v = make_double2(0.0, 0.0);
      if (arow != nullptr && kk < K8) v = *reinterpret_cast<const double2 *>(arow + kk);
      ra[u] = v;
      double bx = 0.0, by = 0.0;
      if (bcol_ok) {
        if (kk < a.K) bx = __ldg(bp + kk);
        if (kk + 1 < a.K) by = __ldg(bp + kk + 1);
      }
      rb[u] = make_double2(bx, by);
    }
  };
  auto store_slice = [&](int s) {
    const double scale = UPDATE ? sc[s >= nks ? 1 : 0][srow] : 1.0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int off = srow * W_PITCH + 8 * h + 2 * ((u + srow) & 3);
      *reinterpret_cast<double2 *>(As + off) = UPDATE ? make_double2(ra[u].x * scale, ra[u].y * scale) : ra[u];
      *reinterpret_cast<double2 *>(Bs + off) = rb[u];
    }
  };

  const int g = lane >> 2, t = lane & 3, wr = warp >> 1, wc = warp & 1;
  bool warp_has_rows = 32 * wr < nv;
  // 8-column groups of this warp that hold output columns at all (the last column tile of a 2^p + 1 grid holds one column)
  const int jmax = min(4, max(0, ((UPDATE ? ((a.N + 7) & ~7) : a.N) - (c0 + 32 * wc) + 7) >> 3));
  if (jmax == 0) warp_has_rows = false;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  load_slice(0);
  for (int s = 0; s < nsl; s++) {
    __syncthreads();          // every warp is done with the previous slice
    store_slice(s);
    __syncthreads();
    if (s + 1 < nsl) load_slice(s + 1);
    if (warp_has_rows) {
#pragma unroll
      for (int jb = 0; jb < W_KS / 8; jb++) {
        double2 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) av[i] = *reinterpret_cast<const double2 *>(As + (32 * wr + 8 * i + g) * W_PITCH + 8 * jb + 2 * t);
#pragma unroll
        for (int j = 0; j < 4; j++) bv[j] = *reinterpret_cast<const double2 *>(Bs + (32 * wc + 8 * j + g) * W_PITCH + 8 * jb + 2 * t);
        if (jmax == 4) {
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].x, bv[j].x);
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].y, bv[j].y);
        } else {
#pragma unroll
          for (int j = 0; j < 3; j++) {
            if (j < jmax) {
#pragma unroll
              for (int i = 0; i < 4; i++) dmma884(acc[i][j][0], acc[i][j][1], av[i].x, bv[j].x);
#pragma unroll
              for (int i = 0; i < 4; i++) dmma884(acc[i][j][0], acc[i][j][1], av[i].y, bv[j].y);
            }
          }
        }
      }
    }
  }

  // ---- epilogue: lane (g, t) holds C[8i + g][8j + 2t], C[8i + g][8j + 2t + 1] ----
  if (!warp_has_rows) return;
  const int N8 = (a.N + 7) & ~7;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = 32 * wr + 8 * i + g;
    if (r >= nv) continue;
    if (UPDATE) {
      double *crow = a.C + (size_t)ids[r] * a.ldc;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = c0 + 32 * wc + 8 * j + 2 * t;
        if (c < N8) *reinterpret_cast<double2 *>(crow + c) = make_double2(acc[i][j][0], acc[i][j][1]);   // columns N .. N8-1: exact zeros
      }
    } else {
      double *cpos = a.C + (row0 + r);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = c0 + 32 * wc + 8 * j + 2 * t;
        if (c < a.N) cpos[(int64_t)c * a.ldc] = acc[i][j][0];
        if (c + 1 < a.N) cpos[(int64_t)(c + 1) * a.ldc] = acc[i][j][1];
      }
    }
  }
}

struct WideTailArgs {
  const double *pb;        // weighted signed pdf, node-major: pb[j * ldp + m]
  int64_t ldp;
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
// R_j = sum_{i<j} v_i, mass = R_n.  Largest i0 <= n-2 with cdf_{i0} < q * mass (reference :134-142 on the unnormalised
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
#pragma unroll 4
    for (int j = 0; j < n1; j++) total += fabs(pv[j * ld]);
    const double qt = qv * total;
    int i0 = 0;
    double dq = qt, Rj = fabs(pv[0]);
#pragma unroll 4
    for (int j = 1; j <= n1 - 2; j++) {
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

cudaError_t wide_tables(const DimInfo *d_dims, int d, const double *xs, double *ih, double *rw, double *hr, cudaStream_t st) {
  wide_tables_kernel<<<d, 128, 0, st>>>(d_dims, xs, ih, rw, hr);
  return cudaGetLastError();
}

// One dimension step k -> k+1 of the wide path: three launches (update, pdf, tail) on st.
cudaError_t launch_wide_step(const WideArgs &w, cudaStream_t st) {
  const int row_tiles = (w.rows + W_TM - 1) / W_TM;
  {
    WideGemmArgs g;
    g.A = w.Fin; g.lda = w.ldf; g.perm = w.perm; g.hist = w.hist_cur; g.nb = w.n0 - 1; g.w1 = w.w1; g.w2 = w.w2; g.rows = w.rows;
    g.B = w.core; g.ldb = (int64_t)w.r0 * w.n0; g.K = w.r0; g.N = w.r1;
    g.C = w.Fout; g.ldc = w.ldf;
    const int n8 = (w.r1 + 7) & ~7;
    const dim3 grid((unsigned)(row_tiles + (w.n0 - 1)), (unsigned)((n8 + W_TN - 1) / W_TN));   // at most rows / 64 + one ragged tile per interval
    wide_gemm_kernel<true><<<grid, W_THREADS, 0, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    WideGemmArgs g;
    g.A = w.Fout; g.lda = w.ldf; g.perm = nullptr; g.hist = nullptr; g.nb = 0; g.w1 = g.w2 = nullptr; g.rows = w.rows;
    g.B = w.pnext; g.ldb = w.r1; g.K = w.r1; g.N = w.n1;
    g.C = w.pb; g.ldc = w.rows;
    const dim3 grid((unsigned)row_tiles, (unsigned)((w.n1 + W_TN - 1) / W_TN));
    wide_gemm_kernel<false><<<grid, W_THREADS, 0, st>>>(g);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    WideTailArgs t;
    t.pb = w.pb; t.ldp = w.rows; t.x = w.xnext; t.ih = w.ihnext; t.rw = w.rwnext; t.hr = w.hrnext;
    t.n1 = w.n1; t.rows = w.rows; t.last = w.last;
    t.q = w.q; t.z = w.z; t.idx_out = w.idx_out; t.lpz = w.lpz;
    t.idx = w.idx; t.w1 = w.w1; t.w2 = w.w2; t.lp = w.lp; t.lpd = w.lpd; t.lpe = w.lpe; t.hist_next = w.hist_next;
    wide_tail_kernel<<<(unsigned)((w.rows + 255) / 256), 256, sizeof(int) * (size_t)(w.n1 > 1 ? w.n1 - 1 : 1), st>>>(t);
    return cudaGetLastError();
  }
}

}  // namespace ttirt
