// FP64 peak microbenchmark for B200 (sm_100a): DFMA pipe vs DMMA (mma.sync.m8n8k4.f64).
// MEASURED_PEAKS.json has no FP64 entry; this program supplies the roofline
// denominator used by bench.py (the larger of the two sustained figures).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char **argv) {
  bool quick = argc > 1 && std::string(argv[1]) == "--quick";
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, 8));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  double best_dfma = 0, best_dmma = 0;
  // burst (short) and sustained (seconds long) variants
  for (int pass = quick ? 1 : 0; pass < 2; pass++) {
    int iters = pass == 0 ? 20000 : (quick ? 100000 : 200000);
    int reps = pass == 0 ? 5 : (quick ? 3 : 10);
    for (int wpb = 4; wpb <= 8; wpb *= 2) {
      int blocks = sms * (wpb == 8 ? 2 : 4);
      int threads = wpb * 32;
      double ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, reps);
      double tf = 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
      if (pass == 1 && tf > best_dfma) best_dfma = tf;
      printf(", \"dfma_p%d_w%d\": %.2f", pass, wpb, tf);
      ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); }, reps);
      tf = 2.0 * 256 * 8 * (double)(iters / 4) * blocks * wpb / (ms * 1e-3) / 1e12;
      if (pass == 1 && tf > best_dmma) best_dmma = tf;
      printf(", \"dmma_p%d_w%d\": %.2f", pass, wpb, tf);
      fflush(stdout);
    }
  }
  // single-warp-per-SMSP DMMA issue rate: 4 warps/SM, 1 block/SM -> cycles per DMMA
  if (!quick) {
    int iters = 50000;
    double ms = time_ms([&] { dmma_kernel<16><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double tf = 2.0 * 256 * 16 * (double)iters * sms * 4 / (ms * 1e-3) / 1e12;
    printf(", \"dmma_1warp_per_smsp\": %.2f", tf);
    ms = time_ms([&] { dmma_kernel<2><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 5);
    tf = 2.0 * 256 * 2 * (double)iters * sms * 4 / (ms * 1e-3) / 1e12;
    printf(", \"dmma_1warp_2acc\": %.2f", tf);
  }
  printf(", \"dfma_sustained_tflops\": %.2f, \"dmma_sustained_tflops\": %.2f}\n", best_dfma, best_dmma);
  return 0;
}
