// FP64 FMA throughput microbenchmark (roofline denominator for the FP64-bound kernels; SURVEY.md 8d asks
// for a measured value).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
double run(int blocks, int threads, int iters) {
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(out);
  return 2.0 * ILP * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  // occupancy / ILP sweep: warps per SM and independent chains per thread
  for (int threads : {32, 64, 128, 256, 512, 1024}) {
    printf("{\"threads_per_sm\": %d, \"ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f, \"ilp8\": %.2f}\n", threads,
           run<1>(sms, threads, iters), run<2>(sms, threads, iters), run<4>(sms, threads, iters),
           run<8>(sms, threads, iters));
  }
  printf("{\"fp64_fma_peak_tflops\": %.3f, \"sms\": %d}\n", run<8>(sms * 2, 1024, iters), sms);
  return 0;
}
