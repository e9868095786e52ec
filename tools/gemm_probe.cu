// Probe: how close can the transition kernel's two DMMA phases (interface update, conditional pdf) get to the
// DMMA peak when fed from shared memory, as a function of warps per SM and B-fragment load width?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_probe gemm_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int RT = 8, NTD = 8, KP = 64;

// MODE 0: LDS.64 per B fragment, A scaled inside the k loop (the shipped structure)
// MODE 1: LDS.128 per pair of k-steps (paired layout)
// MODE 2: MODE 1 + scaling hoisted (two k-steps share the DMULs issued up front)
template <int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(double *out, int tiles, double w1, double w2) {
  extern __shared__ __align__(16) double sm[];
  double *sl_lo = sm, *sl_hi = sm + 64 * KP, *Ps = sm + 2 * 64 * KP;
  for (int i = threadIdx.x; i < (2 * 64 + 72) * KP; i += blockDim.x) sm[i] = 1.0 / (1 + (i & 15));
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double2 fa[RT], fb[RT];
#pragma unroll
  for (int j = 0; j < RT; j++) { fa[j] = make_double2(1e-3 * lane, 1e-3 * j); fb[j] = make_double2(2e-3 * lane, 1e-3 * j); }
  double keep = 0.0;
  for (int tile = 0; tile < tiles; tile++) {
    asm volatile("" ::: "memory");  // shared memory is re-read every tile, as in the real kernel
    double acc[2][RT][2];
#pragma unroll
    for (int j = 0; j < RT; j++) { acc[0][j][0] = acc[0][j][1] = acc[1][j][0] = acc[1][j][1] = 0.0; }
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < RT; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double xa = e ? fa[j].y : fa[j].x, xb = e ? fb[j].y : fb[j].x;
          const double a1A = w1 * xa, a2A = w2 * xa, a1B = w1 * xb, a2B = w2 * xb;
          const int sw = (((2 * j + e) ^ (g & 3)) << 2) | t;
#pragma unroll
          for (int jj = 0; jj < RT; jj++) {
            const int off = (8 * jj + g) * KP + sw;
            const double b1 = sl_lo[off], b2 = sl_hi[off];
            dmma884(acc[0][jj][0], acc[0][jj][1], a1A, b1);
            dmma884(acc[1][jj][0], acc[1][jj][1], a1B, b1);
            dmma884(acc[0][jj][0], acc[0][jj][1], a2A, b2);
            dmma884(acc[1][jj][0], acc[1][jj][1], a2B, b2);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < RT; j++) {
        const double a1A0 = w1 * fa[j].x, a2A0 = w2 * fa[j].x, a1B0 = w1 * fb[j].x, a2B0 = w2 * fb[j].x;
        const double a1A1 = w1 * fa[j].y, a2A1 = w2 * fa[j].y, a1B1 = w1 * fb[j].y, a2B1 = w2 * fb[j].y;
        const int sw = ((j ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
        for (int jj = 0; jj < RT; jj++) {
          const int off = (8 * jj + g) * KP + sw;
          const double2 b1 = *reinterpret_cast<const double2 *>(sl_lo + off), b2 = *reinterpret_cast<const double2 *>(sl_hi + off);
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
    double c[2][NTD][2];
#pragma unroll
    for (int j = 0; j < NTD; j++) { c[0][j][0] = c[0][j][1] = c[1][j][0] = c[1][j][1] = 0.0; }
    if (MODE == 0) {
#pragma unroll
      for (int jj = 0; jj < RT; jj++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const double a0 = acc[0][jj][e], a1 = acc[1][jj][e];
          const int sw = (((2 * jj + e) ^ (g & 3)) << 2) | t;
#pragma unroll
          for (int jn = 0; jn < NTD; jn++) {
            const double bv = Ps[(8 * jn + g) * KP + sw];
            dmma884(c[0][jn][0], c[0][jn][1], a0, bv);
            dmma884(c[1][jn][0], c[1][jn][1], a1, bv);
          }
        }
      }
    } else {
#pragma unroll
      for (int jj = 0; jj < RT; jj++) {
        const int sw = ((jj ^ (g & 1)) << 3) | (t << 1);
#pragma unroll
        for (int jn = 0; jn < NTD; jn++) {
          const double2 bv = *reinterpret_cast<const double2 *>(Ps + (8 * jn + g) * KP + sw);
          dmma884(c[0][jn][0], c[0][jn][1], acc[0][jj][0], bv.x);
          dmma884(c[1][jn][0], c[1][jn][1], acc[1][jj][0], bv.x);
          dmma884(c[0][jn][0], c[0][jn][1], acc[0][jj][1], bv.y);
          dmma884(c[1][jn][0], c[1][jn][1], acc[1][jj][1], bv.y);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NTD; j++) keep += c[0][j][0] + c[0][j][1] + c[1][j][0] + c[1][j][1];
#pragma unroll
    for (int j = 0; j < RT; j++) { fa[j].x += 1e-9 * keep; fb[j].y += 1e-9; }
  }
  if (keep == 123.456) out[0] = keep;
}

template <int MODE, int WARPS>
void run(double *out, int sms) {
  const size_t smem = sizeof(double) * (2 * 64 + 72) * KP;
  CK(cudaFuncSetAttribute(probe<MODE, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = 400;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  probe<MODE, WARPS><<<sms, WARPS * 32, smem>>>(out, tiles, 0.3, 0.7);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 3; r++) probe<MODE, WARPS><<<sms, WARPS * 32, smem>>>(out, tiles, 0.3, 0.7);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
  const double dmma = (double)tiles * (512 + 256) * WARPS * sms;
  printf("mode %d warps %2d: %.2f TFLOP/s (%.3f ms)\n", MODE, WARPS, dmma * 512 / (ms * 1e-3) / 1e12, ms);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double *out; CK(cudaMalloc(&out, 8));
  run<0, 4>(out, p.multiProcessorCount); run<0, 8>(out, p.multiProcessorCount); run<0, 12>(out, p.multiProcessorCount);
  run<1, 4>(out, p.multiProcessorCount); run<1, 8>(out, p.multiProcessorCount); run<1, 12>(out, p.multiProcessorCount);
  return 0;
}
