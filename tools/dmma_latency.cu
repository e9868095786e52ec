// Microbenchmark: issue rate of FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) from ONE warp as a
// function of the number of independent accumulator chains D, with W warps per SM sub-partition.
// Answers: how far apart must dependent DMMAs be for a lone warp to saturate the pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int D>
__global__ void chain_kernel(double *out, int iters, long long *cycles) {
  double c[D][2];
#pragma unroll
  for (int d = 0; d < D; d++) c[d][0] = c[d][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int d = 0; d < D; d++) dmma884(c[d][0], c[d][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < D; d++) s += c[d][0] + c[d][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int D>
void run(int warps_per_smsp, double *out, long long *cyc) {
  const int iters = 2000;
  const int threads = 128 * warps_per_smsp;
  chain_kernel<D><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  chain_kernel<D><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * 8 * D);
  printf("{\"chains\": %d, \"warps_per_smsp\": %d, \"cycles_per_dmma_per_warp\": %.2f, \"pipe_share\": %.3f}\n", D, warps_per_smsp, per,
         16.0 * warps_per_smsp / per);
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  for (int w = 1; w <= 2; w++) {
    run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); run<4>(w, out, cyc); run<6>(w, out, cyc); run<8>(w, out, cyc); run<16>(w, out, cyc);
  }
  return 0;
}
