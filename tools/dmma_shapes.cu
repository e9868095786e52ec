// Microbenchmark: FP64 tensor-core MMA shapes on B200.  mma.sync m8n8k4 (DMMA.8x8x4) against the larger sm_90+ shapes
// m16n8k4 / m16n8k8 / m16n8k16: TFLOP/s with W warps per SM sub-partition and D independent accumulator chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_shapes tools/dmma_shapes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__device__ __forceinline__ void mma16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  if (K == 4)
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  else if (K == 8)
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int K, int D>
__global__ void k16(double *out, int iters, long long *cycles) {
  double c[D][4], a[8], b[4];
  for (int d = 0; d < D; d++) for (int i = 0; i < 4; i++) c[d][i] = 0.0;
  for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; i++) b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int d = 0; d < D; d++) mma16<K>(c[d], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
  for (int d = 0; d < D; d++) for (int i = 0; i < 4; i++) s += c[d][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K, int D>
void run(int w, double *out, long long *cyc) {
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k16<K, D><<<148, 128 * w>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k16<K, D><<<148, 128 * w>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long h = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double n_mma = (double)iters * 4 * D;
  const double flops = n_mma * 2.0 * 16 * 8 * K * 148 * 4 * w;
  printf("{\"shape\": \"m16n8k%d\", \"chains\": %d, \"warps_per_smsp\": %d, \"cycles_per_mma_per_warp\": %.2f, \"tflops\": %.2f}\n", K, D, w,
         (double)h / n_mma, flops / (ms * 1e-3) / 1e12);
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  for (int w = 1; w <= 2; w++) {
    run<4, 2>(w, out, cyc); run<4, 4>(w, out, cyc);
    run<8, 2>(w, out, cyc); run<8, 4>(w, out, cyc);
    run<16, 1>(w, out, cyc); run<16, 2>(w, out, cyc); run<16, 4>(w, out, cyc);
  }
  return 0;
}
