// B200-native engine behind the tt_irt1 C-ABI (include/tt_irt1.h).
//
//   model_create : upload grid + cores, right-to-left marginalisation sweep on the device
//                  (reference tt_irt1_int32.c:59-82), stage-0 tables.
//   sample       : per chunk of samples
//                    fast   : stage-0 kernel, then per dimension {bin scan, bin scatter, fused transition}
//                    strict : one thread per sample, every operation in the reference's order
//                  host-buffer mode adds a chunked H2D / compute / D2H pipeline over several streams.
// There is no CPU fallback anywhere in this file: without a device every entry point fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tt_irt1.h"
#include "ttirt_common.cuh"

using namespace ttirt;

// ------------------------------------------------------------------------------------------------
// errors, counters, options
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int64_t> g_chunk{0};
// launches enqueued while a chunk graph is being captured do not run: they are counted when the graph is launched
static thread_local bool t_capturing = false;

static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (getenv("TTIRT_QUIET") == nullptr) fprintf(stderr, "tt_irt1[b200]: %s\n", g_err);
  return -1;
}

namespace ttirt {
// shared with ttirt_aux.cu
int aux_fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (getenv("TTIRT_QUIET") == nullptr) fprintf(stderr, "tt_irt1[b200]: %s\n", g_err);
  return -1;
}
void aux_launched() { if (!t_capturing) g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ttirt

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define LAUNCHED() do { if (!t_capturing) g_launches.fetch_add(1, std::memory_order_relaxed); } while (0)

static bool trace_on() {
  static int v = -1;
  if (v < 0) v = getenv("TTIRT_TRACE") != nullptr;
  return v != 0;
}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// samples per chunk; never beyond 2^30 (kernels index a chunk's rows with int)
static int64_t default_chunk() {
  const int64_t cap = (int64_t)1 << 30;
  int64_t c = g_chunk.load();
  if (c > 0) return std::min(c, cap);
  const char *e = getenv("TTIRT_CHUNK");
  if (e && atoll(e) > 0) return std::min<int64_t>(atoll(e), cap);
  return (int64_t)1 << 20;
}

// wide path: per-sample scratch is two interface rows and one pdf column (16 r + 8 n bytes and change): chunks of at most
// ~1.5 GB of it, at least 2^14 rows (the GEMMs bound this path, not launches)
static int64_t wide_chunk_for(int64_t rmax, int64_t nmax) {
  const int64_t per_row = 16 * ((rmax + 7) & ~(int64_t)7) + 8 * nmax + 64;
  int64_t c = (int64_t)1 << 20;
  while (c > ((int64_t)1 << 14) && c * per_row > ((int64_t)3 << 29)) c >>= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
constexpr int kSlots = 4;
constexpr size_t kImageBytes = 512 << 10;   // models up to this size keep a host image for the bit-identical-reload check

struct Workspace {
  int64_t cap = 0;       // samples
  bool strict = false, host = false, want_idx = false;
  double *F = nullptr;   // cap x ldf
  int *idx = nullptr, *perm = nullptr;
  double *w1 = nullptr, *w2 = nullptr, *lp = nullptr, *lpd = nullptr;
  int *lpe = nullptr;
  int *hist = nullptr;   // d x nbpad
  double *F2 = nullptr;  // wide path: second interface buffer (cap x ldf), weighted pdf (nmax x cap)
  double *pw = nullptr;
  double *mp = nullptr;  // wide path: shares of the rows' masses (wide_mass_slots x cap)
  // strict scratch
  double *left = nullptr, *pbuf = nullptr, *cbuf = nullptr;
  // host-mode staging
  double *q = nullptr, *z = nullptr, *lpz = nullptr;
  int32_t *idx_out = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  // CUDA graphs of a chunk's whole kernel sequence (memset, stage 0, per dimension scan / scatter / transition):
  // small and medium shapes are bound by those ~3d launches, a replayed graph costs one
  struct ChunkGraph {
    int64_t rows = 0, ldq = 0, ldz = 0;
    const double *q = nullptr; double *z = nullptr, *lpz = nullptr; int32_t *idx_out = nullptr;
    int mode = 0;
    cudaGraphExec_t exec = nullptr;
    uint64_t last_use = 0;
  };
  std::vector<ChunkGraph> graphs;
  uint64_t graph_clock = 0;
};

struct ttirt_model {
  int device = 0, sm_count = 148;
  int64_t d = 0;
  std::vector<int64_t> n, r;
  std::vector<DimInfo> dims;
  int64_t rmax = 1, nmax = 2, nbpad = 0, sum_pk = 0, sum_mg = 0, sum_x = 0, sum_c = 0;
  int ldf = 8;
  int fast_cls = -1;
  double *d_xs = nullptr, *d_core = nullptr, *d_pk = nullptr, *d_pkw = nullptr, *d_marg = nullptr;
  int64_t sum_pkw = 0;
  double *d_p0 = nullptr, *d_cdf0 = nullptr;
  double *d_wtab = nullptr;      // wide path: 1 / cell width, 1 / node weight, h_{j-1} / node weight (3 x sum_x, the layout of xs)
  DimInfo *d_dims = nullptr;
  // model_load() enqueues upload, sweep and operand packing on load_stream and records `loaded`; the host pipeline lets its
  // slot streams wait for that event instead of synchronising the device (a small call is all latency)
  cudaStream_t load_stream = nullptr;
  cudaEvent_t loaded = nullptr;
  double *core_stage = nullptr;   // page-locked staging buffer for large cores arriving in pageable memory
  int64_t core_stage_cap = 0;
  std::vector<double> image;      // host copy of (xs, cores) the model was loaded from, small models only
  bool image_valid = false;
  int walk_cls = -1;             // >= 0: the all-dimensions walk kernel serves this model's fast path (ttirt_walk.cu)
  double *d_walk = nullptr;      // its packed per-dimension operand blocks
  Workspace ws[kSlots];      // host pipeline slots (own streams)
  // device-pointer API: two workspaces of its own.  A call of several chunks alternates them on two internal streams
  // forked from / joined into the caller's stream, so that one chunk's sort, kernel prologues and drains run under the
  // other chunk's transition kernel; a one-chunk call runs on the caller's stream directly.  `done` of each workspace is
  // recorded after its last use and waited for before its next one, whichever stream that is on.
  Workspace dws[2];
  bool dws_used[2] = {false, false};
  cudaEvent_t fork_ev = nullptr;
  // page-locked bounce buffers of the host pipeline, one set per slot: used when the caller's arrays are ordinary
  // pageable memory (numpy, mxArray), which cudaMemcpyAsync would otherwise stage synchronously at a fraction of PCIe rate
  struct HostStage { double *q = nullptr, *z = nullptr, *lpz = nullptr; int64_t cap = 0; } stage[kSlots];
  // optional per-launch timing of the dominant (transition) kernel, for bench.py's roofline
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_flops = 0.0;
};

static void ws_drop_graphs(Workspace &w) {
  for (auto &g : w.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  w.graphs.clear();
}

static void ws_free(Workspace &w) {
  ws_drop_graphs(w);
  cudaFree(w.F); cudaFree(w.idx); cudaFree(w.perm); cudaFree(w.w1); cudaFree(w.w2); cudaFree(w.lp); cudaFree(w.lpd); cudaFree(w.lpe);
  cudaFree(w.hist); cudaFree(w.F2); cudaFree(w.pw); cudaFree(w.mp);
  cudaFree(w.left); cudaFree(w.pbuf); cudaFree(w.cbuf);
  cudaFree(w.q); cudaFree(w.z); cudaFree(w.lpz); cudaFree(w.idx_out);
  if (w.stream) cudaStreamDestroy(w.stream);
  if (w.done) cudaEventDestroy(w.done);
  w = Workspace();
}

static int ws_alloc(ttirt_model *md, Workspace &w, int64_t cap, bool s, bool h, bool wi) {
  const int64_t d = md->d;
  // per-sample state of the per-dimension path; the walk kernel keeps all of it in registers
  if (md->walk_cls < 0) {
    CK(cudaMalloc(&w.F, sizeof(double) * cap * md->ldf));
    CK(cudaMalloc(&w.idx, sizeof(int) * cap));
    CK(cudaMalloc(&w.perm, sizeof(int) * cap));
    CK(cudaMalloc(&w.w1, sizeof(double) * cap));
    CK(cudaMalloc(&w.w2, sizeof(double) * cap));
    CK(cudaMalloc(&w.lp, sizeof(double) * cap));
    CK(cudaMalloc(&w.lpd, sizeof(double) * cap));
    CK(cudaMalloc(&w.lpe, sizeof(int) * cap));
    CK(cudaMalloc(&w.hist, sizeof(int) * 2 * d * md->nbpad));   // d interval histograms, then d sets of scatter cursors
    if (md->fast_cls == kWideClass) {
      CK(cudaMalloc(&w.F2, sizeof(double) * cap * md->ldf));
      CK(cudaMalloc(&w.pw, sizeof(double) * cap * md->nmax));
      CK(cudaMalloc(&w.mp, sizeof(double) * cap * wide_mass_slots((int)md->nmax)));
    }
  }
  if (s) {
    CK(cudaMalloc(&w.left, sizeof(double) * 2 * md->rmax * cap));
    CK(cudaMalloc(&w.pbuf, sizeof(double) * md->nmax * cap));
    CK(cudaMalloc(&w.cbuf, sizeof(double) * md->nmax * cap));
  }
  if (h) {
    CK(cudaMalloc(&w.q, sizeof(double) * cap * d));
    CK(cudaMalloc(&w.z, sizeof(double) * cap * d));
    CK(cudaMalloc(&w.lpz, sizeof(double) * cap));
    if (wi) CK(cudaMalloc(&w.idx_out, sizeof(int32_t) * cap * d));
  }
  if (!w.stream) CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
  if (!w.done) CK(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
  return 0;
}

// Grow a workspace.  The capacity and the mode flags are recorded only once every allocation has succeeded: after a
// failed cudaMalloc the workspace is empty again (cap 0), so a later, smaller call on a caller-held model allocates
// afresh instead of launching kernels on null scratch pointers.
static int ws_ensure(ttirt_model *md, Workspace &w, int64_t rows, bool strict, bool host, bool want_idx) {
  if (w.cap >= rows && (w.strict || !strict) && (w.host || !host) && (w.want_idx || !(host && want_idx))) return 0;
  cudaStream_t keep_s = w.stream; cudaEvent_t keep_e = w.done;
  w.stream = nullptr; w.done = nullptr;
  const bool s = strict || w.strict, h = host || w.host, wi = want_idx || w.want_idx;
  const int64_t cap = rows > w.cap ? rows : w.cap;
  if (keep_s) cudaStreamSynchronize(keep_s);   // work still using the old scratch
  ws_free(w);   // (drops the chunk graphs: they hold the old pointers)
  w.stream = keep_s; w.done = keep_e;
  if (ws_alloc(md, w, cap, s, h, wi) != 0) {
    keep_s = w.stream; keep_e = w.done;
    w.stream = nullptr; w.done = nullptr;
    ws_free(w);                                // cap = 0, every pointer null
    w.stream = keep_s; w.done = keep_e;
    cudaGetLastError();                        // an out-of-memory error is not sticky: clear it
    return -1;
  }
  w.cap = cap; w.strict = s; w.host = h; w.want_idx = wi;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernels: sweep
// ------------------------------------------------------------------------------------------------
// P[i] = sum_l core[i + l*rows] * marg[l], one sequential chain per output element in l order and
// without FMA contraction: the netlib dgemm order the oracle pins (reference :72).
__global__ void sweep_contract_kernel(const double *__restrict__ core, const double *__restrict__ marg, double *P,
                                      int rows, int rr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s = 0.0;
  for (int l = 0; l < rr; l++) s = __dadd_rn(s, __dmul_rn(marg[l], core[i + (int64_t)l * rows]));
  P[i] = s;
}

// C_{k-1}[a] = sum_j (P[a,j] + P[a,j+1]) * h_j * 0.5, accumulated in j order (reference :76-81).
__global__ void sweep_integrate_kernel(const double *__restrict__ P, const double *__restrict__ x, double *marg_out,
                                       int rk, int nk) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= rk) return;
  double s = 0.0;
  for (int j = 0; j + 1 < nk; j++) {
    const double h = __dsub_rn(x[j + 1], x[j]);
    s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dadd_rn(P[a + j * rk], P[a + (j + 1) * rk]), h), 0.5));
  }
  marg_out[a] = s;
}

// The whole right-to-left sweep in one launch (one CTA walks the dimensions; the per-element operation order is that
// of the two kernels above, so the results are bit-identical).  The sweep is ~16 MFLOP even at the metric shape:
// what it costs is launches, and a call of the drop-in symbol pays for it every time.
__device__ void strict_cdf(double *p, double *cdf, const double *x, int nk, int64_t st);
__global__ void sweep_fused_kernel(const DimInfo *__restrict__ dims, int d, const double *__restrict__ core,
                                   const double *__restrict__ xs, double *pk, double *marg, double *p0, double *cdf0) {
  if (threadIdx.x == 0) marg[dims[d - 1].off_m] = 1.0;   // C_{d-1} = {1} (reference :60-61)
  __syncthreads();
  for (int k = d - 1; k >= 0; k--) {
    const DimInfo di = dims[k];
    const int rows = di.r0 * di.n, rr = di.r1;
    const double *ck = core + di.off_c, *mg = marg + di.off_m;
    double *P = pk + di.off_p;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
      double s = 0.0;
      for (int l = 0; l < rr; l++) s = __dadd_rn(s, __dmul_rn(mg[l], ck[i + (int64_t)l * rows]));
      P[i] = s;
    }
    __syncthreads();
    if (k > 0) {
      const double *x = xs + di.off_x;
      double *mo = marg + dims[k - 1].off_m;
      for (int a = threadIdx.x; a < di.r0; a += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j + 1 < di.n; j++) {
          const double h = __dsub_rn(x[j + 1], x[j]);
          s = __dadd_rn(s, __dmul_rn(__dmul_rn(__dadd_rn(P[a + j * di.r0], P[a + (j + 1) * di.r0]), h), 0.5));
        }
        mo[a] = s;
      }
      __syncthreads();
    }
  }
  // stage-0 tables (the body of stage0_table_kernel): the first conditional is the same for every sample
  if (threadIdx.x == 0) {
    const int n0 = dims[0].n;
    const double *P0 = pk + dims[0].off_p, *x = xs + dims[0].off_x;
    for (int j = 0; j < n0; j++) p0[j] = fabs(__dadd_rn(0.0, __dmul_rn(P0[j], 1.0)));
    strict_cdf(p0, cdf0, x, n0, 1);
  }
}

// One sample's conditional on a grid, strict arithmetic (reference :105-130): abs, trapezoid prefix,
// zero-mass fallback, reciprocal normalisation.  p and cdf are strided arrays (stride st).
__device__ void strict_cdf(double *p, double *cdf, const double *x, int nk, int64_t st) {
  cdf[0] = 0.0;
  for (int j = 1; j < nk; j++) {
    const double hq = __dmul_rn(__dsub_rn(x[j], x[j - 1]), 0.5);
    double c = cdf[(j - 1) * st];
    c = __dadd_rn(c, __dmul_rn(hq, p[(j - 1) * st]));
    c = __dadd_rn(c, __dmul_rn(hq, p[j * st]));
    cdf[j * st] = c;
  }
  if (cdf[(nk - 1) * st] == 0.0) {
    const double u = __ddiv_rn(1.0, (double)(nk - 1));
    for (int j = 0; j < nk; j++) { p[j * st] = __dmul_rn(1.0, u); cdf[j * st] = __dmul_rn((double)j, u); }
  }
  const double s = __ddiv_rn(1.0, cdf[(nk - 1) * st]);
  for (int j = 0; j < nk; j++) { cdf[j * st] = __dmul_rn(cdf[j * st], s); p[j * st] = __dmul_rn(p[j * st], s); }
}

// bisection with strict '>' (reference :134-142)
__device__ __forceinline__ int strict_search(const double *cdf, int nk, int64_t st, double qk) {
  int lo = 0, hi = nk - 1;
  while (hi - lo > 1) {
    const int mid = (int)((double)(lo + hi) * 0.5);
    if (qk > cdf[mid * st]) lo = mid; else hi = mid;
  }
  return lo;
}

// Fast path: P_k with column j scaled by the trapezoid node weight of grid node j (ttirt_common.cuh node_weight), laid out
// per dimension at 16-byte aligned offsets so that the transition kernel stages it with cp.async.
__global__ void weight_p_kernel(const DimInfo *__restrict__ dims, int d, const double *__restrict__ xs, const double *__restrict__ pk,
                                double *pkw) {
  const int k = blockIdx.y;
  if (k >= d) return;
  const DimInfo di = dims[k];
  const int total = di.r0 * di.n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e / di.r0;
    pkw[di.off_pw + e] = pk[di.off_p + e] * node_weight(xs + di.off_x, c, di.n);
  }
}

// Stage-0 tables: the first conditional is the same for every sample (r_0 = 1, left interface {1}).
__global__ void stage0_table_kernel(const double *__restrict__ P0, const double *__restrict__ x, double *p0, double *cdf0, int n0) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int j = 0; j < n0; j++) p0[j] = fabs(__dadd_rn(0.0, __dmul_rn(P0[j], 1.0)));
  strict_cdf(p0, cdf0, x, n0, 1);
}

// ------------------------------------------------------------------------------------------------
// kernels: strict path (any shape).  One thread per sample, all dimensions, reference operation order.
// Scratch is sample-minor so that neighbouring threads touch neighbouring addresses.
// ------------------------------------------------------------------------------------------------
__global__ void strict_kernel(const DimInfo *__restrict__ dims, int d, const double *__restrict__ xs,
                              const double *__restrict__ core, const double *__restrict__ pk, int rows,
                              const double *__restrict__ q, int64_t ldq, double *z, int64_t ldz, double *lpz,
                              int32_t *idx_out, double *left, double *pbuf, double *cbuf, int rmax) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  const int64_t st = rows;
  double *la = left + m, *lb = left + (int64_t)rmax * rows + m;
  double *p = pbuf + m, *cdf = cbuf + m;
  la[0] = 1.0;
  double lp = 0.0;
  for (int k = 0; k < d; k++) {
    const DimInfo di = dims[k];
    const int rk = di.r0, nk = di.n, rn = di.r1;
    const double *x = xs + di.off_x, *P = pk + di.off_p, *ck = core + di.off_c;
    const double qk = q[m + ldq * k];
    for (int j = 0; j < nk; j++) {
      double s = 0.0;
      for (int a = 0; a < rk; a++) s = __dadd_rn(s, __dmul_rn(P[a + j * rk], la[a * st]));
      p[j * st] = fabs(s);
    }
    strict_cdf(p, cdf, x, nk, st);
    const int lo = strict_search(cdf, nk, st, qk);
    const double c1 = p[lo * st], c2 = p[(lo + 1) * st];
    const CellOut o = invert_cell(qk, cdf[lo * st], c1, c2, x[lo], x[lo + 1]);
    z[m + ldz * k] = o.xk;
    if (idx_out) idx_out[m + ldz * k] = lo;
    lp = __dadd_rn(lp, o.logp);
    if (k < d - 1) {
      for (int b = 0; b < rn; b++) {
        const double *s1 = ck + (int64_t)lo * rk + (int64_t)b * rk * nk, *s2 = s1 + rk;
        double s = 0.0;
        for (int a = 0; a < rk; a++) {
          double tv = __dmul_rn(o.w1, s1[a]);
          tv = __dadd_rn(tv, __dmul_rn(o.w2, s2[a]));
          s = __dadd_rn(s, __dmul_rn(tv, la[a * st]));
        }
        lb[b * st] = s;
      }
      double *tmp = la; la = lb; lb = tmp;
    }
  }
  lpz[m] = lp;
}

// ------------------------------------------------------------------------------------------------
// kernels: fast path, stage 0 and binning
// ------------------------------------------------------------------------------------------------
__global__ void stage0_kernel(const double *__restrict__ p0, const double *__restrict__ cdf0, const double *__restrict__ x,
                              int n0, int rows, const double *__restrict__ q, double *z, int32_t *idx_out, int *idx,
                              double *w1, double *w2, double *lp, double *lpd, int *lpe, double *lpz, double *F, int ldf, int *hist, int last) {
  extern __shared__ double sm[];
  double *sp = sm, *sc = sm + n0, *sx = sm + 2 * n0;
  int *sh = reinterpret_cast<int *>(sm + 3 * n0);
  for (int i = threadIdx.x; i < n0; i += blockDim.x) { sp[i] = p0[i]; sc[i] = cdf0[i]; sx[i] = x[i]; sh[i] = 0; }
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < rows) {
    const double qk = q[m];
    const int lo = strict_search(sc, n0, 1, qk);
    const CellOut o = invert_cell(qk, sc[lo], sp[lo], sp[lo + 1], sx[lo], sx[lo + 1]);
    z[m] = o.xk;
    if (idx_out) idx_out[m] = lo;
    const double l0 = __dadd_rn(0.0, o.logp);
    if (last) {
      lpz[m] = l0;
    } else {
      idx[m] = lo; w1[m] = o.w1; w2[m] = o.w2;
      double Pn = 1.0, Pd = 1.0;
      int E = 0;
      lp_accumulate(Pn, Pd, E, fabs(__dadd_rn(__dmul_rn(sp[lo], o.w1), __dmul_rn(sp[lo + 1], o.w2))), 1.0);  // sp is normalised
      lp[m] = Pn; lpd[m] = Pd; lpe[m] = E;
      double2 *f = reinterpret_cast<double2 *>(F + (size_t)m * ldf);
      f[0] = make_double2(1.0, 0.0); f[1] = make_double2(0.0, 0.0);
      f[2] = make_double2(0.0, 0.0); f[3] = make_double2(0.0, 0.0);
      atomicAdd(&sh[lo], 1);
    }
  }
  if (!last) {
    __syncthreads();
    for (int i = threadIdx.x; i < n0 - 1; i += blockDim.x)
      if (sh[i]) atomicAdd(hist + i, sh[i]);
  }
}

// counting-sort scatter: perm[position] = sample, positions grouped by interval.  The bin offsets are scanned from the
// histogram by every CTA (no scan launch); cursor[] starts at zero and hands out ranges inside a bin.
__global__ void bin_scatter_kernel(const int *__restrict__ idx, int rows, int nb, const int *__restrict__ hist, int *cursor, int *perm) {
  extern __shared__ int sh[];  // nb counts, nb bases, nb + 1 bin starts
  int *cnt = sh, *base = sh + nb, *bst = sh + 2 * nb;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) cnt[i] = 0;
  if (threadIdx.x < 32) bin_offsets_warp(hist, nb, 1, bst, nullptr, threadIdx.x);
  __syncthreads();
  constexpr int PER = 4;
  int myb[PER], myr[PER];
  const int m0 = (blockIdx.x * blockDim.x) * PER + threadIdx.x;
#pragma unroll
  for (int u = 0; u < PER; u++) {
    const int m = m0 + u * blockDim.x;
    myb[u] = -1;
    if (m < rows) { myb[u] = idx[m]; myr[u] = atomicAdd(&cnt[myb[u]], 1); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) base[i] = cnt[i] ? bst[i] + atomicAdd(cursor + i, cnt[i]) : 0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PER; u++) {
    const int m = m0 + u * blockDim.x;
    if (myb[u] >= 0) perm[base[myb[u]] + myr[u]] = m;
  }
}

__global__ void fill_nan_kernel(double *p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// ------------------------------------------------------------------------------------------------
// model create / destroy
// ------------------------------------------------------------------------------------------------
static int physical_device_count() {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}
// Test hook: TTIRT_VIRTUAL_DEVICES=K lets one physical GPU stand in for K devices of a multi-device call (logical device g
// runs on physical device g mod count, with its own engine slot, host thread and row shard), so that the sharding, the
// fan-out of the cores and the per-device pipelines are exercised on a one-GPU box.  Never set in production.
static int virtual_devices() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("TTIRT_VIRTUAL_DEVICES"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
  return v;
}
static int physical_of(int logical) {
  const int c = physical_device_count();
  return (virtual_devices() > 0 && c > 0) ? logical % c : logical;
}
extern "C" int ttirt_device_count(void) {
  const int c = physical_device_count();
  return (c > 0 && virtual_devices() > c) ? virtual_devices() : c;
}

// Row range of shard `shard` out of `n_shards` for a batch of M samples: contiguous, balanced to one row, in order.
// (Samples are independent, reference tt_irt1_int32.c:88-181; this is the whole multi-GPU decomposition.)
extern "C" int ttirt_shard_rows(int64_t M, int n_shards, int shard, int64_t *m0, int64_t *m1) {
  if (M < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || !m0 || !m1) return -1;
  *m0 = (int64_t)(((__int128)M * shard) / n_shards);
  *m1 = (int64_t)(((__int128)M * (shard + 1)) / n_shards);
  return 0;
}

// Default device count of the drop-in call (TTIRT_DEVICES unset / "auto"): one device per 2^22 seed points (a full
// four-chunk pipeline each), at most `visible`, at least one.
extern "C" int ttirt_auto_devices(int64_t M, int visible) {
  const int64_t want = M >> 22;
  if (visible < 1 || want < 1) return 1;
  return (int)(want > visible ? visible : want);
}

// Which kernels serve a shape in fast mode: the same arithmetic model_build() applies (walk kernel for small TTs on 17-point
// grids, else the fused class by largest rank / grid, else the wide path, else the strict kernel).  No device needed.
static int walk_class_of_shape(int64_t d, const int64_t *n, int64_t rmax) {
  static const bool walk_on = !(getenv("TTIRT_WALK") && atoi(getenv("TTIRT_WALK")) == 0);
  bool uniform = d >= 2;                                      // one grid size throughout; ranks may differ (zero-padded)
  for (int64_t k = 0; k < d && uniform; k++) uniform = n[k] == n[0];
  return (walk_on && uniform) ? walk_class_for((int)rmax, (int)n[0]) : -1;
}
extern "C" int ttirt_path_for_shape(int64_t d, const int64_t *n, const int64_t *ttrank) {
  if (d < 1 || !n || !ttrank) return TTIRT_PATH_STRICT;
  int64_t rmax = 1, nmax = 2;
  for (int64_t k = 0; k < d; k++) {
    if (n[k] < 2 || ttrank[k + 1] < 1 || n[k] > (1 << 20) || ttrank[k + 1] > (1 << 14)) return TTIRT_PATH_STRICT;
    rmax = std::max(rmax, ttrank[k + 1]); nmax = std::max(nmax, n[k]);
  }
  if (walk_class_of_shape(d, n, rmax) >= 0) return TTIRT_PATH_WALK;
  return fast_class_for((int)rmax, (int)nmax);
}

extern "C" void ttirt_model_destroy(ttirt_model *md) {
  if (!md) return;
  cudaSetDevice(md->device);
  for (auto &w : md->ws) ws_free(w);
  for (auto &w : md->dws) ws_free(w);
  if (md->fork_ev) cudaEventDestroy(md->fork_ev);
  if (md->load_stream) { cudaStreamSynchronize(md->load_stream); cudaStreamDestroy(md->load_stream); }
  if (md->loaded) cudaEventDestroy(md->loaded);
  for (auto &h : md->stage) { cudaFreeHost(h.q); cudaFreeHost(h.z); cudaFreeHost(h.lpz); }
  cudaFreeHost(md->core_stage);
  for (auto &p : md->prof_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  cudaFree(md->d_xs); cudaFree(md->d_core); cudaFree(md->d_pk); cudaFree(md->d_pkw); cudaFree(md->d_marg);
  cudaFree(md->d_p0); cudaFree(md->d_cdf0); cudaFree(md->d_dims); cudaFree(md->d_walk); cudaFree(md->d_wtab);
  delete md;
}

// Replication of grid + cores in a multi-device call ("one H2D + cudaMemcpyPeerAsync fan-out over
// NVSwitch"): the first device (root) uploads from the host and publishes its device copy; the other devices pull it over
// NVLink instead of reading the caller's pageable arrays eight times, and fall back to the host upload if that fails.
// Plain copies, no collective.  The root keeps its buffers untouched until every peer has reported (pending == 0).
struct Fanout {
  std::mutex mu;
  std::condition_variable cv;
  int state = 0;                 // 0 pending, 1 root copy resident, -1 root failed (peers upload from the host)
  int src_device = 0;
  const double *d_xs = nullptr, *d_core = nullptr;
  int pending = 0;               // peers that have not finished with the root's copy yet
  bool reported[64] = {false};   // per peer: peer_done() counts once
  void publish(int st, int dev, const double *x, const double *c) {
    { std::lock_guard<std::mutex> l(mu); if (state == 0) { state = st; src_device = dev; d_xs = x; d_core = c; } }
    cv.notify_all();
  }
  void peer_done(int peer) {
    { std::lock_guard<std::mutex> l(mu); if (!reported[peer & 63]) { reported[peer & 63] = true; --pending; } }
    cv.notify_all();
  }
  void wait_peers() {
    std::unique_lock<std::mutex> l(mu);
    cv.wait(l, [&] { return pending <= 0; });
  }
};
struct CoreSource {
  const double *xs = nullptr, *core = nullptr;   // host arrays (always valid)
  Fanout *fan = nullptr;
  bool root = false;
  int peer = 0;                                  // index among the devices of the call
};

static bool fanout_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("TTIRT_FANOUT"); v = e ? atoi(e) != 0 : 1; }
  return v != 0;
}

static bool is_pageable(const void *p);
// Cores from the caller's array to the device.  Large cores in ordinary pageable memory (numpy, mxArray: 64 MB at the metric
// shape) go through a page-locked staging buffer of the model in eight pieces: host threads copy the pieces, each piece is
// sent as soon as it is there (the driver's own staging of pageable memory runs at ~10 GB/s: 6 ms against ~3).
static cudaError_t upload_cores(ttirt_model *md, const double *core, cudaStream_t ls) {
  const size_t bytes = sizeof(double) * (size_t)md->sum_c;
  static const bool no_stage = getenv("TTIRT_NO_STAGING") != nullptr;
  if (no_stage || bytes < ((size_t)8 << 20) || !is_pageable(core)) return cudaMemcpyAsync(md->d_core, core, bytes, cudaMemcpyHostToDevice, ls);
  if (md->core_stage_cap < md->sum_c) {
    cudaFreeHost(md->core_stage); md->core_stage = nullptr; md->core_stage_cap = 0;
    if (cudaHostAlloc(&md->core_stage, bytes, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return cudaMemcpyAsync(md->d_core, core, bytes, cudaMemcpyHostToDevice, ls);
    }
    md->core_stage_cap = md->sum_c;
  }
  constexpr int kPieces = 8;
  std::atomic<int> ready[kPieces];
  for (auto &r : ready) r.store(0);
  const int64_t per = (md->sum_c + kPieces - 1) / kPieces;
  std::vector<std::thread> th;
  for (int p = 0; p < kPieces; p++)
    th.emplace_back([&, p]() {
      const int64_t a = p * per, b = std::min<int64_t>(md->sum_c, a + per);
      if (b > a) memcpy(md->core_stage + a, core + a, sizeof(double) * (size_t)(b - a));
      ready[p].store(1, std::memory_order_release);
    });
  cudaError_t e = cudaSuccess;
  for (int p = 0; p < kPieces; p++) {
    while (!ready[p].load(std::memory_order_acquire)) std::this_thread::yield();
    const int64_t a = p * per, b = std::min<int64_t>(md->sum_c, a + per);
    if (b > a && e == cudaSuccess)
      e = cudaMemcpyAsync(md->d_core + a, md->core_stage + a, sizeof(double) * (size_t)(b - a), cudaMemcpyHostToDevice, ls);
  }
  for (auto &t : th) t.join();
  return e;
}

// Upload grid and cores into an allocated model of matching shape and run the right-to-left sweep:
// C_{d-1} = {1}; P_k = core_k x_3 C_k; C_{k-1} = trapezoid(P_k)   (reference tt_irt1_int32.c:59-82)
static int model_load_body(ttirt_model *md, const CoreSource &src);
static int model_load(ttirt_model *md, const CoreSource &src) {
  const int rc = model_load_body(md, src);
  // a root that fails after publishing its copy must not let the caller free it under the peers' copies
  if (rc != 0 && src.fan && src.root) { src.fan->publish(-1, 0, nullptr, nullptr); src.fan->wait_peers(); }
  return rc;
}
static int model_load_body(ttirt_model *md, const CoreSource &src) {
  const int64_t d = md->d;
  const double *xs = src.xs, *core = src.core;
  bool have = false;
  if (src.fan && !src.root) {
    Fanout &f = *src.fan;
    {
      std::unique_lock<std::mutex> l(f.mu);
      f.cv.wait(l, [&] { return f.state != 0; });
    }
    if (f.state == 1) {
      cudaError_t e = cudaSuccess;
      cudaStream_t ps = md->load_stream;
      if (f.src_device == md->device) {   // only under the virtual-device test hook: two logical devices on one GPU
        e = cudaMemcpyAsync(md->d_xs, f.d_xs, sizeof(double) * md->sum_x, cudaMemcpyDeviceToDevice, ps);
        if (e == cudaSuccess) e = cudaMemcpyAsync(md->d_core, f.d_core, sizeof(double) * md->sum_c, cudaMemcpyDeviceToDevice, ps);
      } else {
        e = cudaDeviceEnablePeerAccess(f.src_device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(md->d_xs, md->device, f.d_xs, f.src_device, sizeof(double) * md->sum_x, ps);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(md->d_core, md->device, f.d_core, f.src_device, sizeof(double) * md->sum_c, ps);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(ps);   // the root may reuse its buffers once this peer has reported
      if (e == cudaSuccess) have = true; else cudaGetLastError();
    }
    f.peer_done(src.peer);
  }
  cudaStream_t ls = md->load_stream;
  const double tl0 = now_s();
  if (!have) {
    // pageable sources are staged by the driver before cudaMemcpyAsync returns, so the caller's arrays may change afterwards
    cudaError_t e = cudaMemcpyAsync(md->d_xs, xs, sizeof(double) * md->sum_x, cudaMemcpyHostToDevice, ls);
    if (e == cudaSuccess) e = upload_cores(md, core, ls);
    if (src.fan && src.root) {
      if (e == cudaSuccess) e = cudaStreamSynchronize(ls);   // the peers copy from this device's buffers
      src.fan->publish(e == cudaSuccess ? 1 : -1, md->device, md->d_xs, md->d_core);
    }
    CK(e);
  }
  // small TTs (the MH / IW use of the reference: r ~ 8..16, n = 17): one CTA walks the whole sweep (and builds the stage-0
  // tables) in less time than two launches take; larger cores use one launch per step so that the contraction spreads
  // over the SMs
  if (md->sum_c <= (1 << 18)) {
    sweep_fused_kernel<<<1, 1024, 0, ls>>>(md->d_dims, (int)d, md->d_core, md->d_xs, md->d_pk, md->d_marg, md->d_p0, md->d_cdf0);
    LAUNCHED();
  } else {
    static const double one = 1.0;
    CK(cudaMemcpyAsync(md->d_marg + md->dims[d - 1].off_m, &one, sizeof(double), cudaMemcpyHostToDevice, ls));
    for (int64_t k = d - 1; k >= 0; k--) {
      const DimInfo &di = md->dims[k];
      const int rows = di.r0 * di.n;
      sweep_contract_kernel<<<(rows + 127) / 128, 128, 0, ls>>>(md->d_core + di.off_c, md->d_marg + di.off_m, md->d_pk + di.off_p, rows, di.r1);
      LAUNCHED();
      if (k > 0) {
        sweep_integrate_kernel<<<(di.r0 + 127) / 128, 128, 0, ls>>>(md->d_pk + di.off_p, md->d_xs + di.off_x,
                                                                    md->d_marg + md->dims[k - 1].off_m, di.r0, di.n);
        LAUNCHED();
      }
    }
    stage0_table_kernel<<<1, 32, 0, ls>>>(md->d_pk, md->d_xs, md->d_p0, md->d_cdf0, md->dims[0].n);
    LAUNCHED();
  }
  if (md->walk_cls >= 0) {
    CK(walk_pack(md->walk_cls, md->d_dims, (int)d, md->d_xs, md->d_core, md->d_pk, md->d_p0, md->d_cdf0, md->d_walk, ls));
    LAUNCHED();
  }
  if (md->fast_cls >= 0 && d > 1 && md->walk_cls < 0) {
    weight_p_kernel<<<dim3(8, (unsigned)d), 256, 0, ls>>>(md->d_dims, (int)d, md->d_xs, md->d_pk, md->d_pkw);
    LAUNCHED();
  }
  if (md->fast_cls == kWideClass && d > 1) {
    CK(wide_tables(md->d_dims, (int)d, md->d_xs, md->d_wtab, md->d_wtab + md->sum_x, md->d_wtab + 2 * md->sum_x, ls));
    LAUNCHED();
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(md->loaded, ls));
  // The model is complete when this returns.  (Letting the pipeline's slot streams wait for `loaded` instead was measured
  // and lost: 417 against 405 ms per call at the metric shape, 61 against 53 ms at d=40 n=33 r=32 -- all four slots then
  // start at once instead of staggered behind their uploads.)
  const double tl1 = now_s();
  CK(cudaStreamSynchronize(ls));
  if (trace_on())
    fprintf(stderr, "tt_irt1[b200] trace: model_load device %d: enqueue %.3f ms (uploads are staged inside), + %.3f ms until loaded\n",
            md->device, 1e3 * (tl1 - tl0), 1e3 * (now_s() - tl1));
  // small models: keep the host-side image, so that the next drop-in call with bit-identical grid and cores (an MH / IW
  // driver calling the sampler again on the same TT, reference test_shock_absorber_tt.py:138-153) skips upload and sweep
  md->image_valid = false;
  if ((size_t)(md->sum_x + md->sum_c) * sizeof(double) <= kImageBytes) {
    md->image.resize((size_t)(md->sum_x + md->sum_c));
    memcpy(md->image.data(), xs, sizeof(double) * md->sum_x);
    memcpy(md->image.data() + md->sum_x, core, sizeof(double) * md->sum_c);
    md->image_valid = true;
  }
  return 0;
}

// true when the model was loaded from exactly these bytes (exact comparison, no hashing) and reuse is enabled
static bool model_image_matches(const ttirt_model *md, const double *xs, const double *core) {
  static const bool reuse = !(getenv("TTIRT_MODEL_REUSE") && atoi(getenv("TTIRT_MODEL_REUSE")) == 0);
  if (!reuse || !md->image_valid) return false;
  return memcmp(md->image.data(), xs, sizeof(double) * md->sum_x) == 0 &&
         memcmp(md->image.data() + md->sum_x, core, sizeof(double) * md->sum_c) == 0;
}

static int model_build(ttirt_model *md, const int64_t *n, const int64_t *rk, const CoreSource &src) {
  const int64_t d = md->d;
  if (rk[0] != 1 || rk[d] != 1) return fail("ttrank[0] and ttrank[d] must be 1 (got %lld, %lld)", (long long)rk[0], (long long)rk[d]);
  md->n.assign(n, n + d);
  md->r.assign(rk, rk + d + 1);
  md->dims.resize(d);
  int64_t ox = 0, oc = 0, op = 0, om = 0, opw = 0;
  for (int64_t k = 0; k < d; k++) {
    if (n[k] < 2) return fail("n[%lld] = %lld: every grid needs at least 2 points", (long long)k, (long long)n[k]);
    if (rk[k] < 1 || rk[k + 1] < 1) return fail("non-positive TT rank at %lld", (long long)k);
    if (n[k] > (1 << 20) || rk[k] > (1 << 14)) return fail("shape too large at dimension %lld", (long long)k);
    DimInfo &di = md->dims[k];
    di.n = (int)n[k]; di.r0 = (int)rk[k]; di.r1 = (int)rk[k + 1]; di.pad = 0;
    di.off_x = ox; di.off_c = oc; di.off_p = op; di.off_m = om; di.off_pw = opw;
    ox += n[k]; oc += rk[k] * n[k] * rk[k + 1]; op += rk[k] * n[k]; om += rk[k + 1];
    opw += (rk[k] * n[k] + 1) & ~(int64_t)1;
    if (rk[k + 1] > md->rmax) md->rmax = rk[k + 1];
    if (n[k] > md->nmax) md->nmax = n[k];
  }
  md->sum_pk = op; md->sum_mg = om; md->sum_pkw = opw;
  md->nbpad = (md->nmax + 7) & ~(int64_t)7;
  md->fast_cls = fast_class_for((int)md->rmax, (int)md->nmax);
  md->ldf = (int)((md->rmax + 7) & ~(int64_t)7);
  if (md->ldf < 8) md->ldf = 8;

  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, md->device));
  if (prop.major < 10) return fail("device %d (%s, sm_%d%d) is not a Blackwell B200-class GPU", md->device, prop.name, prop.major, prop.minor);
  md->sm_count = prop.multiProcessorCount;
  CK(fast_init(md->device));
  if (md->fast_cls == kWideClass) CK(wide_init(md->device));
  // small uniform-rank TTs: one persistent kernel for the whole walk (TTIRT_WALK=0: per-dimension path, for comparison)
  {
    md->walk_cls = walk_class_of_shape(d, n, md->rmax);
    if (md->walk_cls >= 0) {
      CK(walk_init(md->device));
      CK(cudaMalloc(&md->d_walk, sizeof(double) * walk_pack_doubles(md->walk_cls, (int)d)));
    }
  }

  CK(cudaMalloc(&md->d_xs, sizeof(double) * ox));
  CK(cudaMalloc(&md->d_core, sizeof(double) * oc));
  CK(cudaMalloc(&md->d_pk, sizeof(double) * op));
  CK(cudaMalloc(&md->d_pkw, sizeof(double) * opw));
  CK(cudaMalloc(&md->d_marg, sizeof(double) * om));
  CK(cudaMalloc(&md->d_p0, sizeof(double) * n[0]));
  CK(cudaMalloc(&md->d_cdf0, sizeof(double) * n[0]));
  CK(cudaMalloc(&md->d_dims, sizeof(DimInfo) * d));
  if (md->fast_cls == kWideClass) CK(cudaMalloc(&md->d_wtab, sizeof(double) * 3 * ox));
  CK(cudaMemcpy(md->d_dims, md->dims.data(), sizeof(DimInfo) * d, cudaMemcpyHostToDevice));
  md->sum_x = ox; md->sum_c = oc;
  CK(cudaStreamCreateWithFlags(&md->load_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&md->loaded, cudaEventDisableTiming));
  return model_load(md, src);
}

static ttirt_model *model_create_from(int64_t d, const int64_t *n, const int64_t *ttrank, const CoreSource &src, int device) {
  int cnt = physical_device_count();
  if (cnt <= 0) { fail("no CUDA device available (this library has no CPU fallback)"); return nullptr; }
  if (device < 0 || device >= cnt) { fail("device %d out of range (%d visible)", device, cnt); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", device); return nullptr; }
  ttirt_model *md = new ttirt_model();
  md->device = device; md->d = d;
  if (model_build(md, n, ttrank, src) != 0) { ttirt_model_destroy(md); return nullptr; }
  return md;
}

extern "C" ttirt_model *ttirt_model_create(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank,
                                           const double *ttcore, int device) {
  g_err[0] = 0;
  if (d < 1 || !n || !xs || !ttrank || !ttcore) { fail("bad arguments to ttirt_model_create"); return nullptr; }
  CoreSource src;
  src.xs = xs; src.core = ttcore;
  ttirt_model *md = model_create_from(d, n, ttrank, src, device);
  // a caller-held model is complete when this returns (the drop-in call lets its pipeline wait for `loaded` instead)
  if (md && cudaStreamSynchronize(md->load_stream) != cudaSuccess) {
    fail("model upload / sweep failed: %s", cudaGetErrorString(cudaGetLastError()));
    ttirt_model_destroy(md);
    return nullptr;
  }
  return md;
}

extern "C" int ttirt_model_get_sweep(const ttirt_model *md, double *pk_out, double *marg_out) {
  if (!md) return fail("null model");
  CK(cudaSetDevice(md->device));
  CK(cudaStreamSynchronize(md->load_stream));
  if (pk_out) CK(cudaMemcpy(pk_out, md->d_pk, sizeof(double) * md->sum_pk, cudaMemcpyDeviceToHost));
  if (marg_out) CK(cudaMemcpy(marg_out, md->d_marg, sizeof(double) * md->sum_mg, cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// one chunk, device-resident buffers, enqueued on st
// ------------------------------------------------------------------------------------------------
// kernels in one chunk's sequence (what a graph replay launches)
static int64_t g_launches_per_graph(const ttirt_model *md, int mode) {
  if (mode == TTIRT_MODE_STRICT || md->fast_cls < 0 || md->walk_cls >= 0) return 1;
  return 1 + (md->fast_cls == kWideClass ? 4 : 2) * (md->d - 1);
}

static int enqueue_chunk(ttirt_model *md, Workspace &w, int64_t rows, const double *q, int64_t ldq, double *z, int64_t ldz,
                         double *lpz, int32_t *idx_out, int mode, cudaStream_t st) {
  const int d = (int)md->d;
  if (rows <= 0) return 0;
  if (mode == TTIRT_MODE_STRICT || md->fast_cls < 0) {
    strict_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(md->d_dims, d, md->d_xs, md->d_core, md->d_pk, (int)rows, q, ldq,
                                                                  z, ldz, lpz, idx_out, w.left, w.pbuf, w.cbuf, (int)md->rmax);
    LAUNCHED();
    CK(cudaGetLastError());
    return 0;
  }
  if (md->walk_cls >= 0) {
    WalkArgs wa;
    wa.pack = md->d_walk; wa.d = d; wa.rows = (int)rows; wa.q = q; wa.ldq = ldq; wa.z = z; wa.ldz = ldz; wa.lpz = lpz; wa.idx_out = idx_out;
    CK(launch_walk(md->walk_cls, wa, md->sm_count, st));
    LAUNCHED();
    return 0;
  }
  const int nbpad = (int)md->nbpad;
  if (d > 1) CK(cudaMemsetAsync(w.hist, 0, sizeof(int) * 2 * d * nbpad, st));   // histograms and scatter cursors
  {
    const DimInfo &d0 = md->dims[0];
    const size_t sm = sizeof(double) * 3 * d0.n + sizeof(int) * d0.n;
    stage0_kernel<<<(unsigned)((rows + 255) / 256), 256, sm, st>>>(md->d_p0, md->d_cdf0, md->d_xs + d0.off_x, d0.n, (int)rows, q, z,
                                                                   idx_out, w.idx, w.w1, w.w2, w.lp, w.lpd, w.lpe, lpz, w.F, md->ldf, w.hist, d == 1);
    LAUNCHED();
    CK(cudaGetLastError());
  }
  for (int k = 0; k + 1 < d; k++) {
    const DimInfo &dk = md->dims[k], &dn = md->dims[k + 1];
    const int nb = dk.n - 1;
    bin_scatter_kernel<<<(unsigned)((rows + 1023) / 1024), 256, sizeof(int) * (3 * nb + 1), st>>>(
        w.idx, (int)rows, nb, w.hist + (size_t)k * nbpad, w.hist + (size_t)(d + k) * nbpad, w.perm);
    LAUNCHED();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (md->profile) {
      if (md->prof_used == md->prof_events.size()) {
        cudaEvent_t x, y;
        CK(cudaEventCreate(&x)); CK(cudaEventCreate(&y));
        md->prof_events.emplace_back(x, y);
      }
      e0 = md->prof_events[md->prof_used].first; e1 = md->prof_events[md->prof_used].second;
      md->prof_used++;
      md->prof_flops += (double)rows * (4.0 * dk.r0 * dk.r1 + 2.0 * dk.r1 * dn.n);
      CK(cudaEventRecord(e0, st));
    }
    if (md->fast_cls == kWideClass) {
      // wide shapes: update GEMM (F -> F2 or back), pdf GEMM, tail (ttirt_wide.cu); stage 0 left the first rows in w.F
      WideArgs wa;
      wa.core = md->d_core + dk.off_c; wa.pnext = md->d_pkw + dn.off_pw;
      wa.xnext = md->d_xs + dn.off_x; wa.ihnext = md->d_wtab + dn.off_x; wa.rwnext = md->d_wtab + md->sum_x + dn.off_x;
      wa.hrnext = md->d_wtab + 2 * md->sum_x + dn.off_x;
      wa.r0 = dk.r0; wa.n0 = dk.n; wa.r1 = dk.r1; wa.n1 = dn.n;
      wa.last = (k + 1 == d - 1); wa.rows = (int)rows;
      wa.Fin = (k & 1) ? w.F2 : w.F; wa.Fout = (k & 1) ? w.F : w.F2; wa.ldf = md->ldf;
      wa.pb = w.pw; wa.mass_part = w.mp; wa.perm = w.perm; wa.hist_cur = w.hist + (size_t)k * nbpad;
      wa.idx = w.idx; wa.w1 = w.w1; wa.w2 = w.w2; wa.lp = w.lp; wa.lpd = w.lpd; wa.lpe = w.lpe;
      wa.q = q + ldq * (k + 1); wa.z = z + ldz * (k + 1);
      wa.idx_out = idx_out ? idx_out + ldz * (k + 1) : nullptr;
      wa.lpz = lpz; wa.hist_next = w.hist + (size_t)(k + 1) * nbpad;
      CK(launch_wide_step(wa, st));
      LAUNCHED(); LAUNCHED(); LAUNCHED();
      if (e1) CK(cudaEventRecord(e1, st));
      continue;
    }
    TransArgs a;
    a.core = md->d_core + dk.off_c; a.pnext = md->d_pkw + dn.off_pw; a.xnext = md->d_xs + dn.off_x;
    a.r0 = dk.r0; a.n0 = dk.n; a.r1 = dk.r1; a.n1 = dn.n;
    a.async_ok = (reinterpret_cast<uintptr_t>(a.core) & 15) == 0;
    a.last = (k + 1 == d - 1); a.rows = (int)rows; a.F = w.F; a.ldf = md->ldf;
    a.perm = w.perm; a.hist_cur = w.hist + (size_t)k * nbpad;
    a.idx = w.idx; a.w1 = w.w1; a.w2 = w.w2; a.lp = w.lp; a.lpd = w.lpd; a.lpe = w.lpe;
    a.q = q + ldq * (k + 1); a.z = z + ldz * (k + 1);
    a.idx_out = idx_out ? idx_out + ldz * (k + 1) : nullptr;
    a.lpz = lpz; a.hist_next = w.hist + (size_t)(k + 1) * nbpad;
    CK(launch_transition(md->fast_cls, a, md->sm_count, st));
    LAUNCHED();
    if (e1) CK(cudaEventRecord(e1, st));
  }
  return 0;
}

static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("TTIRT_GRAPHS"); v = e ? atoi(e) != 0 : 1; }
  return v != 0;
}

// One chunk on stream st: replay the chunk's CUDA graph when there is one for exactly these buffers, else capture it
// while enqueueing (the legacy default stream cannot be captured, and per-launch profiling needs its events outside
// a graph: both enqueue directly).
static int run_chunk(ttirt_model *md, Workspace &w, int64_t rows, const double *q, int64_t ldq, double *z, int64_t ldz,
                     double *lpz, int32_t *idx_out, int mode, cudaStream_t st) {
  if (rows <= 0) return 0;
  // large chunks are not launch-bound, and the device-resident API would need one graph per chunk offset
  const bool one_launch = mode == TTIRT_MODE_STRICT || md->fast_cls < 0 || md->walk_cls >= 0;   // nothing for a graph to save
  if (!graphs_enabled() || st == nullptr || md->profile || rows > (1 << 19) || one_launch) return enqueue_chunk(md, w, rows, q, ldq, z, ldz, lpz, idx_out, mode, st);
  for (auto &g : w.graphs) {
    if (g.rows == rows && g.q == q && g.ldq == ldq && g.z == z && g.ldz == ldz && g.lpz == lpz && g.idx_out == idx_out && g.mode == mode) {
      g.last_use = ++w.graph_clock;
      CK(cudaGraphLaunch(g.exec, st));
      g_launches.fetch_add(g_launches_per_graph(md, mode), std::memory_order_relaxed);
      return 0;
    }
  }
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return enqueue_chunk(md, w, rows, q, ldq, z, ldz, lpz, idx_out, mode, st);
  }
  t_capturing = true;    // nothing runs yet: this thread's launches are counted at cudaGraphLaunch
  const int rc = enqueue_chunk(md, w, rows, q, ldq, z, ldz, lpz, idx_out, mode, st);
  t_capturing = false;
  const cudaError_t ec = cudaStreamEndCapture(st, &graph);
  if (rc != 0 || ec != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc != 0 ? rc : fail("CUDA graph capture of a chunk failed: %s", cudaGetErrorString(ec));
  }
  Workspace::ChunkGraph g;
  g.rows = rows; g.q = q; g.ldq = ldq; g.z = z; g.ldz = ldz; g.lpz = lpz; g.idx_out = idx_out; g.mode = mode;
  const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) { cudaGetLastError(); return enqueue_chunk(md, w, rows, q, ldq, z, ldz, lpz, idx_out, mode, st); }
  if (w.graphs.size() >= 8) {   // least recently used out
    size_t lru = 0;
    for (size_t i = 1; i < w.graphs.size(); i++) if (w.graphs[i].last_use < w.graphs[lru].last_use) lru = i;
    cudaGraphExecDestroy(w.graphs[lru].exec);
    w.graphs.erase(w.graphs.begin() + lru);
  }
  g.last_use = ++w.graph_clock;
  w.graphs.push_back(g);
  CK(cudaGraphLaunch(g.exec, st));
  g_launches.fetch_add(g_launches_per_graph(md, mode), std::memory_order_relaxed);
  return 0;
}

extern "C" int ttirt_sample_device(ttirt_model *md, int64_t M, const double *d_q, int64_t ldq, double *d_z, int64_t ldz,
                                   double *d_lpz, int32_t *d_idx, int mode, void *stream) {
  g_err[0] = 0;
  if (!md) return fail("null model");
  if (M < 0 || ldq < M || ldz < M) return fail("bad M / leading dimensions");
  if (M == 0) return 0;
  CK(cudaSetDevice(md->device));
  cudaStream_t st = (cudaStream_t)stream;
  const bool strict = (mode == TTIRT_MODE_STRICT) || md->fast_cls < 0;
  // chunk size of the device-pointer API when none is set: 2^22 rows in the r <= 64 class (the transition kernel's prologue
  // and drain are paid once per launch: 43.2 -> 43.8 M samples/s at the metric shape against 2^20), one launch for the
  // walk kernel (no per-sample scratch), 2^20 otherwise
  int64_t want = default_chunk();
  if (g_chunk.load() <= 0 && getenv("TTIRT_CHUNK") == nullptr && !strict) {
    if (md->walk_cls >= 0) want = (int64_t)1 << 24;
    else if (md->fast_cls == 2) want = (int64_t)1 << 22;
    else if (md->fast_cls == kWideClass) want = wide_chunk_for(md->rmax, md->nmax);
  }
  const int64_t chunk = std::min<int64_t>(M, want);
  const int64_t nchunks = (M + chunk - 1) / chunk;
  static const bool two_streams = !(getenv("TTIRT_DEVICE_STREAMS") && atoi(getenv("TTIRT_DEVICE_STREAMS")) < 2);
  // per-launch profiling needs the kernels serialised (events around overlapping kernels measure the overlap too)
  const int nws = (nchunks >= 2 && !strict && !md->profile && two_streams) ? 2 : 1;
  for (int i = 0; i < nws; i++) {
    Workspace &w = md->dws[i];
    if (w.cap < chunk || (strict && !w.strict)) {
      CK(cudaStreamSynchronize(st));  // a previous call on this stream may still use the old scratch (ws_ensure syncs w.stream)
      if (md->dws_used[i]) CK(cudaEventSynchronize(w.done));
      if (ws_ensure(md, w, chunk, strict, false, false) != 0) return -1;
    }
  }
  if (nws == 1) {
    Workspace &w = md->dws[0];
    if (md->dws_used[0]) CK(cudaStreamWaitEvent(st, w.done, 0));
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
      const int64_t rows = std::min(chunk, M - m0);
      if (run_chunk(md, w, rows, d_q + m0, ldq, d_z + m0, ldz, d_lpz + m0, d_idx ? d_idx + m0 : nullptr, mode, st) != 0) return -1;
    }
    CK(cudaEventRecord(w.done, st));
    md->dws_used[0] = true;
    return 0;
  }
  if (!md->fork_ev) CK(cudaEventCreateWithFlags(&md->fork_ev, cudaEventDisableTiming));
  CK(cudaEventRecord(md->fork_ev, st));
  for (int i = 0; i < 2; i++) {
    CK(cudaStreamWaitEvent(md->dws[i].stream, md->fork_ev, 0));
    if (md->dws_used[i]) CK(cudaStreamWaitEvent(md->dws[i].stream, md->dws[i].done, 0));
  }
  int64_t c = 0;
  for (int64_t m0 = 0; m0 < M; m0 += chunk, ++c) {
    const int64_t rows = std::min(chunk, M - m0);
    Workspace &w = md->dws[c & 1];
    if (run_chunk(md, w, rows, d_q + m0, ldq, d_z + m0, ldz, d_lpz + m0, d_idx ? d_idx + m0 : nullptr, mode, w.stream) != 0) return -1;
  }
  for (int i = 0; i < 2; i++) {
    CK(cudaEventRecord(md->dws[i].done, md->dws[i].stream));
    md->dws_used[i] = true;
    CK(cudaStreamWaitEvent(st, md->dws[i].done, 0));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// host buffers: chunked H2D -> kernels -> D2H over kSlots streams
// ------------------------------------------------------------------------------------------------
// Page-lock a caller buffer for the duration of a call when it is ordinary pageable memory and large enough for
// the registration to pay for itself (asynchronous, full-rate PCIe copies instead of staged synchronous ones).
// Off by default (measured on B200: cudaHostRegister of 8.6 GB per call costs ~1 s, more than the staged copies
// it replaces); TTIRT_PIN=1 enables it for callers that reuse the same buffers.  Returns true when registered here.
static bool pin_if_pageable(const void *p, size_t bytes) {
  static int mode = -1;
  if (mode < 0) { const char *e = getenv("TTIRT_PIN"); mode = e ? atoi(e) + 1 : 1; }   // 1 off (default: registering 8 GB per call costs more than it saves), 2 on
  if (!p || bytes == 0 || mode == 1) return false;
  if (bytes < ((size_t)32 << 20)) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeUnregistered) return false;
  if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}
static void unpin(const void *p, bool pinned) { if (pinned) { cudaHostUnregister(const_cast<void *>(p)); } }

// Work queue of one call: chunks of `chunk` rows of [base, end), handed out in order to whoever asks next.  A one-device call
// walks a private queue; a multi-device call shares one between its device threads, so that a device that moves its data
// faster takes more chunks (on the 8-GPU boxes of this pool four of the GPUs reach host memory ~25 % slower than the
// other four).  Chunk boundaries are fixed by the queue, never by who takes a chunk: results do not depend on the assignment.
struct ChunkQueue {
  std::atomic<int64_t> next{0};
  int64_t base = 0, end = 0, chunk = 1;
  // ramp (one-device calls of many chunks): the first and the last `levels` chunks are 1 / 2^levels ... 1 / 2 of the regular
  // size, so that the first upload and the last download -- which nothing can overlap -- are short
  std::vector<int64_t> starts;   // empty: uniform chunks
  void ramp(int levels) {
    const int64_t M = end - base;
    if (levels < 1 || (chunk >> levels) < 1024 || M < 8 * chunk) return;
    int64_t m = base, tail_rows = 0;
    for (int l = levels; l >= 1; l--) { starts.push_back(m); m += chunk >> l; tail_rows += chunk >> l; }
    while (end - m - tail_rows >= chunk) { starts.push_back(m); m += chunk; }
    if (end - m > tail_rows) { starts.push_back(m); m = end - tail_rows; }   // a shorter regular chunk takes up the remainder
    for (int l = 1; l <= levels; l++) { starts.push_back(m); m += chunk >> l; }
    starts.push_back(end);
  }
  bool claim(int64_t &m0, int64_t &rows) {
    const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
    if (!starts.empty()) {
      if (i + 1 >= (int64_t)starts.size()) return false;
      m0 = starts[(size_t)i]; rows = starts[(size_t)i + 1] - m0;
      return true;
    }
    m0 = base + i * chunk;
    if (m0 >= end) return false;
    rows = std::min(chunk, end - m0);
    return true;
  }
};

// Where a chunk's seeds come from: the caller's q (host memory), or generated on the device (csrc/ttirt_aux.cu), which
// removes the q upload altogether.
struct SeedSpec {
  int kind = 0;                 // 0: host q, 1: shifted rank-1 lattice, 2: Philox uniforms
  int64_t N = 0, m_base = 0;    // lattice size; global index of row 0 of this call
  const double *d_genvec = nullptr, *d_shift = nullptr;
  uint64_t seed = 0;
};

static bool is_pageable(const void *p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

static int copy_threads() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TTIRT_COPY_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    v = e ? atoi(e) : std::min(8, std::max(2, hw / 2));   // 8 on the 16-core boxes: 38.6 -> 39.6 M samples/s on pageable arrays
    if (v < 1) v = 1;
    if (v > 32) v = 32;
  }
  return v;
}

// d column segments of `rows` doubles, split over a few host threads (a single core tops out near 10 GB/s)
static void copy_columns(double *dst, int64_t dst_ld, const double *src, int64_t src_ld, int64_t rows, int d) {
  const int nt = std::min(copy_threads(), d);
  auto work = [&](int t) {
    for (int k = t; k < d; k += nt) memcpy(dst + (int64_t)k * dst_ld, src + (int64_t)k * src_ld, sizeof(double) * (size_t)rows);
  };
  if (nt <= 1 || (int64_t)rows * d < (1 << 17)) { for (int t = 0; t < nt; t++) work(t); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
}

static int stage_ensure(ttirt_model *md, int slot, int64_t rows, bool need_q, bool need_z) {
  auto &h = md->stage[slot];
  if (h.cap >= rows && (!need_q || h.q) && (!need_z || h.z)) return 0;
  const int64_t cap = std::max(rows, h.cap);
  if (need_q && (h.cap < cap || !h.q)) { cudaFreeHost(h.q); h.q = nullptr; CK(cudaHostAlloc(&h.q, sizeof(double) * cap * md->d, cudaHostAllocDefault)); }
  if (need_z && (h.cap < cap || !h.z)) {
    cudaFreeHost(h.z); cudaFreeHost(h.lpz); h.z = h.lpz = nullptr;
    CK(cudaHostAlloc(&h.z, sizeof(double) * cap * md->d, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h.lpz, sizeof(double) * cap, cudaHostAllocDefault));
  }
  if (h.cap < cap) {   // the other buffer, if present, is too small now
    if (!need_q && h.q) { cudaFreeHost(h.q); h.q = nullptr; }
    if (!need_z && h.z) { cudaFreeHost(h.z); cudaFreeHost(h.lpz); h.z = h.lpz = nullptr; }
  }
  h.cap = cap;
  return 0;
}

// Rows [m_begin, m_end) of one call on this model's device: chunks of rows go round kSlots slots (stream + device
// scratch + optional pinned bounce buffers).  Per chunk: seeds in (async H2D from pinned memory, or a host-thread copy
// into the slot's pinned buffer first when the caller's q is pageable, or generated on the device) -> kernels -> D2H
// (straight into pinned caller memory, or into the slot's pinned buffer from which a drain thread copies into the
// caller's pageable arrays while the next chunks compute).
// chunk size of the host pipeline for a model of fast-path class `cls` and `rows` rows per device
static int64_t host_chunk_for(int cls, bool strict, bool walk, int64_t rows, int64_t rmax, int64_t nmax) {
  int64_t chunk = default_chunk();
  if (g_chunk.load() <= 0 && getenv("TTIRT_CHUNK") == nullptr && !strict && cls == kWideClass) chunk = wide_chunk_for(rmax, nmax);
  // light shapes (r <= 16: 2^17 rows, r <= 32: 2^19 rows per chunk): a chunk computes in about a millisecond, so the pipeline is cut finer
  // (shorter fill and drain, copies and kernels of neighbouring chunks overlap) and every chunk is a graph replay
  if (g_chunk.load() <= 0 && getenv("TTIRT_CHUNK") == nullptr && !strict && cls >= 0 && cls <= 1)
    chunk = cls == 0 ? (1 << 17) : (1 << 19);
  // a walk-kernel chunk is one launch: cut small calls finer, so that upload, kernel and download of neighbouring chunks overlap
  if (rows < chunk * kSlots) {
    const int64_t fine = std::max<int64_t>((rows + kSlots - 1) / kSlots, std::min<int64_t>(rows, walk ? (1 << 12) : (1 << 16)));
    chunk = (cls == kWideClass && !strict) ? std::min(chunk, fine) : fine;   // (the wide path's chunk bounds its per-sample scratch)
  }
  return chunk;
}

static int sample_host_rows(ttirt_model *md, int64_t m_begin, int64_t m_end, const double *h_q, double *h_z, double *h_lpz,
                            int32_t *h_idx, int64_t ld, int mode, const SeedSpec &seeds = SeedSpec(), ChunkQueue *shared = nullptr) {
  const int64_t M = m_end - m_begin;
  if (M <= 0) return 0;
  const double ts0 = now_s();
  CK(cudaSetDevice(md->device));
  const int d = (int)md->d;
  const bool strict = (mode == TTIRT_MODE_STRICT) || md->fast_cls < 0;
  const bool walk = !strict && md->walk_cls >= 0;
  // the chunks of this call: a private queue over [m_begin, m_end), or the queue shared by the devices of the call
  ChunkQueue own;
  own.base = m_begin; own.end = m_end; own.chunk = host_chunk_for(md->fast_cls, strict, walk, M, md->rmax, md->nmax);
  // (metric shape, pinned arrays: 41.35 -> 42.05 M samples/s with two levels; TTIRT_RAMP=<levels>, 0 turns it off)
  static const int ramp_levels = getenv("TTIRT_RAMP") ? atoi(getenv("TTIRT_RAMP")) : 2;
  if (!shared && !strict && md->fast_cls == 2) own.ramp(ramp_levels);   // (r <= 32 class measured: 52.6 against 50.8 ms per call, left uniform)
  ChunkQueue &queue = shared ? *shared : own;
  const int64_t chunk = queue.chunk;
  const int64_t nchunks = shared ? (int64_t)kSlots : (M + chunk - 1) / chunk;   // (shared: unknown in advance)
  const int nslots = (int)std::min<int64_t>(kSlots, nchunks);
  static const bool no_stage = getenv("TTIRT_NO_STAGING") != nullptr;
  const bool big = M * d >= (1 << 20);   // tiny calls: the driver's own staging is as good
  const bool stage_q = !no_stage && big && seeds.kind == 0 && is_pageable(h_q + m_begin);
  const bool stage_z = !no_stage && big && is_pageable(h_z + m_begin);
  for (int s = 0; s < nslots; s++) {
    if (ws_ensure(md, md->ws[s], chunk, strict, true, h_idx != nullptr) != 0) return -1;
    if ((stage_q || stage_z) && stage_ensure(md, s, chunk, stage_q, stage_z) != 0) return -1;
  }

  const double ts1 = now_s();
  // drain thread: waits for a chunk's event, copies its outputs from the slot's pinned buffer into the caller's arrays
  std::mutex mu;
  std::condition_variable cv;
  int64_t submitted = 0, drained = 0;
  std::vector<std::pair<int64_t, int64_t>> claimed;   // (first row, rows) of this device's chunks, in submission order
  bool abort_drain = false, all_submitted = false;
  int drain_rc = 0;
  std::thread drain;
  if (stage_z) {
    drain = std::thread([&]() {
      if (cudaSetDevice(md->device) != cudaSuccess) { std::lock_guard<std::mutex> l(mu); drain_rc = -1; drained = INT64_MAX / 2; cv.notify_all(); return; }
      for (int64_t c = 0;; c++) {
        int64_t m0, rows;
        {
          std::unique_lock<std::mutex> l(mu);
          cv.wait(l, [&] { return submitted > c || all_submitted || abort_drain; });
          if (abort_drain || submitted <= c) break;
          m0 = claimed[(size_t)c].first; rows = claimed[(size_t)c].second;
        }
        const int s = (int)(c % kSlots);
        if (cudaEventSynchronize(md->ws[s].done) != cudaSuccess) { std::lock_guard<std::mutex> l(mu); drain_rc = -1; }
        copy_columns(h_z + m0, ld, md->stage[s].z, rows, rows, d);
        memcpy(h_lpz + m0, md->stage[s].lpz, sizeof(double) * (size_t)rows);
        { std::lock_guard<std::mutex> l(mu); drained = c + 1; }
        cv.notify_all();
      }
    });
  }
  auto finish = [&](int rc) -> int {
    if (drain.joinable()) {
      { std::lock_guard<std::mutex> l(mu); if (rc != 0) abort_drain = true; all_submitted = true; }
      cv.notify_all();
      drain.join();
    }
    return rc != 0 ? rc : drain_rc;
  };
#define CKF(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) { fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return finish(-1); } \
  } while (0)

  int64_t taken = 0;
  for (int64_t c = 0;; c++) {
    const int s = (int)(c % kSlots);
    Workspace &w = md->ws[s];
    if (c < kSlots) CKF(cudaStreamWaitEvent(w.stream, md->loaded, 0));   // upload + sweep of this call's model (model_load)
    // first a free slot, then a chunk: a device that is waiting for its own pipeline does not sit on work another could do
    if (c >= kSlots) {
      if (stage_z) {   // the slot's pinned output buffer must have been drained
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return drained > c - kSlots; });
      } else {
        CKF(cudaEventSynchronize(w.done));
      }
    }
    int64_t m0 = 0, rows = 0;
    if (!queue.claim(m0, rows)) break;
    taken++;
    if (seeds.kind == 1) {
      if (ttirt_seeds_lattice_device(d, rows, seeds.m_base + (m0 - m_begin), seeds.N, seeds.d_genvec, seeds.d_shift, w.q, rows, w.stream) != 0) return finish(-1);
    } else if (seeds.kind == 2) {
      if (ttirt_seeds_uniform_device(d, rows, seeds.m_base + (m0 - m_begin), seeds.seed, w.q, rows, w.stream) != 0) return finish(-1);
    } else if (stage_q) {
      if (c >= kSlots) CKF(cudaEventSynchronize(w.done));   // the previous H2D out of this pinned buffer is long done; make it certain
      copy_columns(md->stage[s].q, rows, h_q + m0, ld, rows, d);
      CKF(cudaMemcpyAsync(w.q, md->stage[s].q, sizeof(double) * rows * d, cudaMemcpyHostToDevice, w.stream));
    } else {
      CKF(cudaMemcpy2DAsync(w.q, sizeof(double) * rows, h_q + m0, sizeof(double) * ld, sizeof(double) * rows, d,
                            cudaMemcpyHostToDevice, w.stream));
    }
    if (run_chunk(md, w, rows, w.q, rows, w.z, rows, w.lpz, h_idx ? w.idx_out : nullptr, mode, w.stream) != 0) return finish(-1);
    if (seeds.kind != 0 && h_q)   // hand the generated seeds back when the caller wants them
      CKF(cudaMemcpy2DAsync(const_cast<double *>(h_q) + m0, sizeof(double) * ld, w.q, sizeof(double) * rows, sizeof(double) * rows, d,
                            cudaMemcpyDeviceToHost, w.stream));
    if (stage_z) {
      CKF(cudaMemcpyAsync(md->stage[s].z, w.z, sizeof(double) * rows * d, cudaMemcpyDeviceToHost, w.stream));
      CKF(cudaMemcpyAsync(md->stage[s].lpz, w.lpz, sizeof(double) * rows, cudaMemcpyDeviceToHost, w.stream));
    } else {
      CKF(cudaMemcpy2DAsync(h_z + m0, sizeof(double) * ld, w.z, sizeof(double) * rows, sizeof(double) * rows, d,
                            cudaMemcpyDeviceToHost, w.stream));
      CKF(cudaMemcpyAsync(h_lpz + m0, w.lpz, sizeof(double) * rows, cudaMemcpyDeviceToHost, w.stream));
    }
    if (h_idx)
      CKF(cudaMemcpy2DAsync(h_idx + m0, sizeof(int32_t) * ld, w.idx_out, sizeof(int32_t) * rows, sizeof(int32_t) * rows, d,
                            cudaMemcpyDeviceToHost, w.stream));
    CKF(cudaEventRecord(w.done, w.stream));
    if (stage_z) {
      { std::lock_guard<std::mutex> l(mu); claimed.emplace_back(m0, rows); submitted = c + 1; }
      cv.notify_all();
    }
  }
  const double te = now_s();
  for (int s = 0; s < nslots; s++) CKF(cudaStreamSynchronize(md->ws[s].stream));
#undef CKF
  if (trace_on())
    fprintf(stderr, "tt_irt1[b200] trace: pipeline device %d: %lld chunks of %lld rows, setup %.3f ms, enqueue %.3f ms, wait %.3f ms (staged q %d, z %d)\n",
            md->device, (long long)taken, (long long)chunk, 1e3 * (ts1 - ts0), 1e3 * (te - ts1), 1e3 * (now_s() - te), (int)stage_q, (int)stage_z);
  return finish(0);
}

extern "C" int ttirt_sample_host(ttirt_model *md, int64_t M, const double *h_q, double *h_z, double *h_lpz, int32_t *h_idx,
                                 int64_t ld, int mode) {
  g_err[0] = 0;
  if (!md) return fail("null model");
  if (M < 0 || ld < M) return fail("bad M / leading dimension");
  CK(cudaSetDevice(md->device));
  const bool pq = pin_if_pageable(h_q, sizeof(double) * ld * md->d), pz = pin_if_pageable(h_z, sizeof(double) * ld * md->d),
             pl = pin_if_pageable(h_lpz, sizeof(double) * M);
  const int rc = sample_host_rows(md, 0, M, h_q, h_z, h_lpz, h_idx, ld, mode);
  unpin(h_q, pq); unpin(h_z, pz); unpin(h_lpz, pl);
  return rc;
}

// Sampling with the seeds generated on the device: no q upload.  h_q may be NULL;
// when given it receives the seeds that were used (column-major, leading dimension ld).
extern "C" int ttirt_sample_lattice_host(ttirt_model *md, int64_t M, int64_t m0, int64_t N, const int64_t *genvec, const double *shift,
                                         double *h_q, double *h_z, double *h_lpz, int64_t ld, int mode) {
  g_err[0] = 0;
  if (!md) return fail("null model");
  if (M < 0 || ld < M || N < 1 || !genvec || !shift || (M > 0 && (!h_z || !h_lpz))) return fail("bad arguments to ttirt_sample_lattice_host");
  if (M == 0) return 0;
  CK(cudaSetDevice(md->device));
  const int64_t d = md->d;
  std::vector<double> z(d);
  for (int64_t k = 0; k < d; k++) z[k] = (double)genvec[k];
  double *dz = nullptr;
  CK(cudaMalloc(&dz, sizeof(double) * 2 * d));
  SeedSpec sp;
  sp.kind = 1; sp.N = N; sp.m_base = m0; sp.d_genvec = dz; sp.d_shift = dz + d;
  int rc = 0;
  if (cudaMemcpy(dz, z.data(), sizeof(double) * d, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(dz + d, shift, sizeof(double) * d, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail("seed upload failed");
  if (rc == 0) rc = sample_host_rows(md, 0, M, h_q, h_z, h_lpz, nullptr, ld, mode, sp);
  cudaFree(dz);
  return rc;
}

extern "C" int ttirt_sample_uniform_host(ttirt_model *md, int64_t M, int64_t m0, uint64_t seed, double *h_q, double *h_z, double *h_lpz,
                                         int64_t ld, int mode) {
  g_err[0] = 0;
  if (!md) return fail("null model");
  if (M < 0 || ld < M || (M > 0 && (!h_z || !h_lpz))) return fail("bad arguments to ttirt_sample_uniform_host");
  if (M == 0) return 0;
  SeedSpec sp;
  sp.kind = 2; sp.m_base = m0; sp.seed = seed;
  return sample_host_rows(md, 0, M, h_q, h_z, h_lpz, nullptr, ld, mode, sp);
}

// ------------------------------------------------------------------------------------------------
// The drop-in call.  The reference's tt_irt1 keeps no state between calls and neither does this one as far as
// the caller can tell, but device allocations (the model's buffers and the pipeline workspaces, ~4.6 GB at
// the metric shape) are kept per device and reused by the next call of the same shape: MH / IW drivers call the
// sampler repeatedly on one TT (reference test_shock_absorber_tt.py:138-153).  Cores and grid are uploaded and
// the sweep is redone on every call, so results never depend on a previous call.  TTIRT_CACHE=0 disables the
// reuse; ttirt_cache_clear() releases everything.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kMaxDevices = 64;
struct EngineSlot {
  std::mutex mu;
  ttirt_model *md = nullptr;
};
EngineSlot g_slots[kMaxDevices];

bool cache_enabled() {
  const char *e = getenv("TTIRT_CACHE");
  return !(e && atoi(e) == 0);
}

bool same_shape(const ttirt_model *md, int64_t d, const int64_t *n, const int64_t *rk) {
  if (!md || md->d != d) return false;
  for (int64_t k = 0; k < d; k++) if (md->n[k] != n[k]) return false;
  for (int64_t k = 0; k <= d; k++) if (md->r[k] != rk[k]) return false;
  return true;
}

// rows [m0, m1) of one call on one device, through that device's cached engine
int run_on_device(int device, int64_t d, const int64_t *n, const int64_t *rk, const CoreSource &src,
                  int64_t m0, int64_t m1, const double *h_q, double *h_z, double *h_lpz, int32_t *h_idx, int64_t ld, int mode,
                  const SeedSpec &seeds, ChunkQueue *queue = nullptr) {
  // whatever happens below, a root must publish (so that peers never wait for ever) and a peer must report (so that the
  // root never does); both are idempotent
  struct FanGuard {
    const CoreSource &s;
    ~FanGuard() {
      if (!s.fan) return;
      if (s.root) { s.fan->publish(-1, 0, nullptr, nullptr); s.fan->wait_peers(); }
      else s.fan->peer_done(s.peer);
    }
  } guard{src};
  if (device < 0 || device >= kMaxDevices) return fail("device %d out of range", device);
  EngineSlot &sl = g_slots[device];   // per LOGICAL device (equal to the physical one outside the virtual-device test hook)
  std::lock_guard<std::mutex> lock(sl.mu);
  const double t0 = now_s();
  device = physical_of(device);
  if (same_shape(sl.md, d, n, rk)) {
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice(%d) failed", device);
    if (!src.fan && model_image_matches(sl.md, src.xs, src.core)) {
      // same bytes as the resident model: nothing to upload, nothing to recompute
    } else if (model_load(sl.md, src) != 0) {
      if (src.fan && src.root) src.fan->wait_peers();
      ttirt_model_destroy(sl.md); sl.md = nullptr;
      return -1;
    }
  } else {
    ttirt_model_destroy(sl.md);
    sl.md = model_create_from(d, n, rk, src, device);
    if (!sl.md) return -1;
  }
  const double t1 = now_s();
  const int rc = sample_host_rows(sl.md, m0, m1, h_q, h_z, h_lpz, h_idx, ld, mode, seeds, queue);
  const double t2 = now_s();
  if (src.fan && src.root) src.fan->wait_peers();   // peers copy from this model's buffers
  if (!cache_enabled() || rc != 0) { ttirt_model_destroy(sl.md); sl.md = nullptr; }
  if (trace_on())
    fprintf(stderr, "tt_irt1[b200] trace: device %d rows %lld: model %.1f ms, pipeline %.1f ms, release %.1f ms\n", device,
            (long long)(m1 - m0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (now_s() - t2));
  return rc;
}
}  // namespace

namespace ttirt { void sqr_pool_clear(); }   // ttirt_sqr.cu: idle device blocks of the squared-density path

extern "C" void ttirt_cache_clear(void) {
  ttirt::sqr_pool_clear();
  for (int g = 0; g < kMaxDevices; g++) {
    std::lock_guard<std::mutex> lock(g_slots[g].mu);
    if (g_slots[g].md) { ttirt_model_destroy(g_slots[g].md); g_slots[g].md = nullptr; }
  }
}

static int run_host_impl(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank, const double *ttcore,
                         int64_t M, const double *h_q, double *h_z, double *h_lpz, int32_t *h_idx, int mode,
                         int first_device, int n_devices, const SeedSpec &seeds) {
  g_err[0] = 0;
  if (M < 0) return fail("negative M");
  if (d < 1 || !n || !xs || !ttrank || !ttcore) return fail("bad arguments to ttirt_run_host");
  const int cnt = ttirt_device_count();
  if (cnt <= 0) return fail("no CUDA device available (this library has no CPU fallback)");
  if (n_devices <= 0 || first_device < 0 || first_device + n_devices > cnt)
    return fail("device range [%d, %d) not available (%d visible)", first_device, first_device + n_devices, cnt);
  if (M == 0) return 0;
  if (cudaSetDevice(physical_of(first_device)) != cudaSuccess) return fail("cudaSetDevice(%d) failed", first_device);
  const bool pq = pin_if_pageable(h_q, sizeof(double) * M * d), pz = pin_if_pageable(h_z, sizeof(double) * M * d),
             pl = pin_if_pageable(h_lpz, sizeof(double) * M);
  struct Unpin {
    const void *q, *z, *l; bool pq, pz, pl;
    ~Unpin() { unpin(q, pq); unpin(z, pz); unpin(l, pl); }
  } unpin_guard{h_q, h_z, h_lpz, pq, pz, pl};
  CoreSource src;
  src.xs = xs; src.core = ttcore;
  if (n_devices == 1) return run_on_device(first_device, d, n, ttrank, src, 0, M, h_q, h_z, h_lpz, h_idx, M, mode, seeds);
  // Samples are independent (reference tt_irt1_int32.c:88-181): one host thread and one pipeline per device, no collective.
  // The rows go to the devices chunk by chunk from a shared queue (default), or as contiguous equal shards
  // (TTIRT_BALANCE=static: ttirt_shard_rows).  Grid and cores reach the first device from the host and the others from
  // there (Fanout); the tiny sweep is redone on every device.
  static const bool balance = !(getenv("TTIRT_BALANCE") && strcmp(getenv("TTIRT_BALANCE"), "static") == 0);
  Fanout fan;
  fan.pending = n_devices - 1;
  int64_t core_elems = 0, rmax = 1, nmax = 2;
  for (int64_t k = 0; k < d; k++) {
    core_elems += ttrank[k] * n[k] * ttrank[k + 1];
    rmax = std::max(rmax, ttrank[k + 1]); nmax = std::max(nmax, n[k]);
  }
  const bool use_fan = fanout_enabled() && n_devices <= 64 && core_elems >= (1 << 16);   // small cores: eight host uploads are as quick
  ChunkQueue queue;
  queue.base = 0; queue.end = M;
  {
    const int cls = fast_class_for((int)rmax, (int)nmax);
    queue.chunk = host_chunk_for(cls, mode == TTIRT_MODE_STRICT || cls < 0, false, M / n_devices, rmax, nmax);
    // four or more devices share the host's memory path: finer chunks even out what the faster and the slower devices
    // take and shorten the tail (8 x B200, M = 2^26: 237 -> 246 M samples/s)
    if (balance && n_devices >= 4 && cls == 2 && mode != TTIRT_MODE_STRICT && g_chunk.load() <= 0 && getenv("TTIRT_CHUNK") == nullptr &&
        queue.chunk > (1 << 19))
      queue.chunk = 1 << 19;
  }
  std::vector<int> rcs(n_devices, 0);
  std::vector<std::string> errs(n_devices);
  std::vector<std::thread> th;
  for (int g = 0; g < n_devices; g++) {
    th.emplace_back([&, g]() {
      int64_t m0 = 0, m1 = M;
      if (!balance) ttirt_shard_rows(M, n_devices, g, &m0, &m1);
      CoreSource s = src;
      if (use_fan) { s.fan = &fan; s.root = g == 0; s.peer = g; }
      SeedSpec sp = seeds;
      sp.m_base = seeds.m_base + m0;
      rcs[g] = run_on_device(first_device + g, d, n, ttrank, s, m0, m1, h_q, h_z, h_lpz, h_idx, M, mode, sp, balance ? &queue : nullptr);
      if (rcs[g] != 0) errs[g] = g_err;
    });
  }
  for (auto &t : th) t.join();
  for (int g = 0; g < n_devices; g++)
    if (rcs[g] != 0) return fail("device %d: %s", first_device + g, errs[g].c_str());
  return 0;
}

extern "C" int ttirt_run_host(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank, const double *ttcore,
                              int64_t M, const double *h_q, double *h_z, double *h_lpz, int32_t *h_idx, int mode,
                              int first_device, int n_devices) {
  if (M > 0 && (!h_q || !h_z || !h_lpz)) return fail("bad arguments to ttirt_run_host");
  return run_host_impl(d, n, xs, ttrank, ttcore, M, h_q, h_z, h_lpz, h_idx, mode, first_device, n_devices, SeedSpec());
}

// The whole call with the seeds generated on the devices (Philox4x32-10, counter = global sample index, so the result does
// not depend on the number of devices): no q upload at all.  h_q may be NULL; when given it receives the seeds.
extern "C" int ttirt_run_uniform_host(int64_t d, const int64_t *n, const double *xs, const int64_t *ttrank, const double *ttcore,
                                      int64_t M, int64_t m0, uint64_t seed, double *h_q, double *h_z, double *h_lpz, int mode,
                                      int first_device, int n_devices) {
  if (M > 0 && (!h_z || !h_lpz)) return fail("bad arguments to ttirt_run_uniform_host");
  SeedSpec sp;
  sp.kind = 2; sp.m_base = m0; sp.seed = seed;
  return run_host_impl(d, n, xs, ttrank, ttcore, M, h_q, h_z, h_lpz, nullptr, mode, first_device, n_devices, sp);
}

extern "C" void ttirt_profile_enable(ttirt_model *md, int on) {
  if (!md) return;
  md->profile = on != 0; md->prof_used = 0; md->prof_flops = 0.0;
}

extern "C" int ttirt_profile_read(ttirt_model *md, double *ms_total, int64_t *launches, double *flops_total) {
  if (!md) return fail("null model");
  CK(cudaSetDevice(md->device));
  double ms = 0.0;
  for (size_t i = 0; i < md->prof_used; i++) {
    CK(cudaEventSynchronize(md->prof_events[i].second));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, md->prof_events[i].first, md->prof_events[i].second));
    ms += t;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = (int64_t)md->prof_used;
  if (flops_total) *flops_total = md->prof_flops;
  return 0;
}

#ifdef TTIRT_PHASE_TIMING
namespace ttirt { void phase_cycles_read(unsigned long long *out); void trace_read(long long *out); }
extern "C" __attribute__((visibility("default"))) void ttirt_debug_trace(long long *out) { cudaDeviceSynchronize(); ttirt::trace_read(out); }
extern "C" __attribute__((visibility("default"))) void ttirt_debug_phase_cycles(unsigned long long *out) { cudaDeviceSynchronize(); ttirt::phase_cycles_read(out); }
#endif

extern "C" int64_t ttirt_kernel_launches(void) { return g_launches.load(); }
extern "C" const char *ttirt_last_error(void) { return g_err; }
extern "C" void ttirt_set_chunk(int64_t samples) { g_chunk.store(samples > 0 ? samples : 0); }
