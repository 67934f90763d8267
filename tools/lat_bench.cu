// Dependent-issue latencies on the target GPU (tools/, not part of the library): DFMA / DADD / DMUL chains, LDS,
// cp.async (LDGSTS) round trip for an L2 hit and for a DRAM miss, prefetch.global.L2 effectiveness.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat_bench tools/lat_bench.cu && /tmp/lat_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_dfma(double a, double b, double* out, long long* cyc, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double a, double b, double* out, long long* cyc, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) x = x + b;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) j = idx[j];
  long long t1 = clock64();
  out[threadIdx.x] = j;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// cp.async round trip: each iteration copies 16 B per thread from a new line and waits for it
template <int MODE>  // 0: cp.async.ca 1: cp.async.cg 2: ld.global  3: prefetch.L2 issued `ahead` iterations earlier + cp.async.ca
__global__ void k_cpasync(const double* g, size_t stride, double* out, long long* cyc, int n, int ahead) {
  __shared__ __align__(16) double buf[2 * 256];
  const unsigned s = (unsigned)__cvta_generic_to_shared(buf + 2 * threadIdx.x);
  const double* p = g + 2 * threadIdx.x;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 3) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p + (size_t)ahead * stride));
    if (MODE == 2) {
      acc += *reinterpret_cast<const volatile double*>(p);
    } else {
      if (MODE == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(p) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(p) : "memory");
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
      double v;
      asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(s) : "memory");
      acc += v;
    }
    if (MODE == 3) {  // some dependent work so the prefetch has time: ~`work` cycles
      for (int w = 0; w < 16; ++w) acc = fma(acc, 1.0000001, 1e-9);
    }
    p += stride;
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// issue cost of 16-byte async copies: every iteration issues NI cp.async (one commit group) for lines that are
// L2 resident, does `work` dependent DFMAs, then waits for the group -- the copies have the whole DFMA chain to land,
// so what is left beyond the chain is the cost of issuing them
template <int MODE>  // 0 cp.async.ca 16 B, 1 ld.global.v2.f64 into registers (consumed after the chain), 2 no loads
__global__ void k_issue(const double* g, double* out, long long* cyc, int n, int work) {
  extern __shared__ __align__(16) double buf[];
  const unsigned s = (unsigned)__cvta_generic_to_shared(buf) + threadIdx.x * 48;
  const unsigned sstride = blockDim.x * 48;
  const double* p = g + 6 * threadIdx.x;
  double acc = 1.0, sum = 0.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double2 r[15];
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + u * sstride), "l"(p + u * 1024) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + u * sstride + 16), "l"(p + u * 1024 + 2) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + u * sstride + 32), "l"(p + u * 1024 + 4) : "memory");
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        r[3 * u] = *reinterpret_cast<const double2*>(p + u * 1024);
        r[3 * u + 1] = *reinterpret_cast<const double2*>(p + u * 1024 + 2);
        r[3 * u + 2] = *reinterpret_cast<const double2*>(p + u * 1024 + 4);
      }
    }
    for (int w = 0; w < work; ++w) acc = fma(acc, 1.0000001, 1e-9);
    if (MODE == 0) {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#pragma unroll
      for (int u = 0; u < 15; ++u) {
        double a, b;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a), "=d"(b) : "r"(s + (u / 3) * sstride + (u % 3) * 16) : "memory");
        sum += a + b;
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < 15; ++u) sum += r[u].x + r[u].y;
    }
    p += 8192;
    if (p > g + (1 << 20)) p -= (1 << 20);
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc + sum;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; CK(cudaMalloc(&out, 8 * 1024)); CK(cudaMalloc(&cyc, 64));
  long long h;
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    k_dfma<<<1, 32>>>(1.0, 0.999, out, cyc, n); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    if (rep) printf("{\"op\": \"DFMA dependent\", \"cycles\": %.2f}\n", (double)h / n);
    k_dadd<<<1, 32>>>(1.0, 0.999, out, cyc, n); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    if (rep) printf("{\"op\": \"DADD dependent\", \"cycles\": %.2f}\n", (double)h / n);
    k_lds<<<1, 32>>>(out, cyc, n); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    if (rep) printf("{\"op\": \"LDS dependent\", \"cycles\": %.2f}\n", (double)h / n);
  }
  // memory: 1 GiB buffer, stride 64 KiB apart per iteration -> every access a new DRAM page; second run over a
  // 16 MiB window (L2-resident after the first pass)
  const size_t bytes = (size_t)1 << 30;
  double* g; CK(cudaMalloc(&g, bytes)); CK(cudaMemset(g, 0, bytes));
  const int m = 2048;
  auto run = [&](const char* name, int mode, size_t stride_d, int ahead, bool warm) {
    for (int rep = 0; rep < (warm ? 3 : 1); ++rep) {
      if (!warm) { CK(cudaMemset(g, 0, bytes)); }  // flush L2 with 1 GiB of writes
      if (mode == 0) k_cpasync<0><<<1, 32>>>(g, stride_d, out, cyc, m, ahead);
      if (mode == 1) k_cpasync<1><<<1, 32>>>(g, stride_d, out, cyc, m, ahead);
      if (mode == 2) k_cpasync<2><<<1, 32>>>(g, stride_d, out, cyc, m, ahead);
      if (mode == 3) k_cpasync<3><<<1, 32>>>(g, stride_d, out, cyc, m, ahead);
      CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    }
    printf("{\"op\": \"%s\", \"cycles\": %.1f}\n", name, (double)h / m);
  };
  run("cp.async.ca + wait, DRAM miss", 0, 32768, 0, false);
  run("cp.async.cg + wait, DRAM miss", 1, 32768, 0, false);
  run("ld.global, DRAM miss", 2, 32768, 0, false);
  run("cp.async.ca + wait, L2 hit", 0, 512, 0, true);
  run("cp.async.cg + wait, L2 hit", 1, 512, 0, true);
  run("ld.global, L2 hit", 2, 512, 0, true);
  run("16 DFMA + cp.async.ca, DRAM miss, no prefetch (prefetch of the same line)", 3, 32768, 0, false);
  run("16 DFMA + cp.async.ca, DRAM miss, prefetch.L2 8 ahead", 3, 32768, 8, false);
  run("16 DFMA + cp.async.ca, DRAM miss, prefetch.L2 32 ahead", 3, 32768, 32, false);
  {
    const int it = 2000, work = 50;
    CK(cudaFuncSetAttribute(k_issue<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 192 * 5));
    for (int nthreads : {32, 192}) {
      for (int rep = 0; rep < 2; ++rep) {
        k_issue<2><<<1, nthreads, 48 * 192 * 5>>>(g, out, cyc, it, work); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double base = (double)h / it;
        k_issue<0><<<1, nthreads, 48 * 192 * 5>>>(g, out, cyc, it, work); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double a = (double)h / it;
        k_issue<1><<<1, nthreads, 48 * 192 * 5>>>(g, out, cyc, it, work); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double b = (double)h / it;
        if (rep) printf("{\"op\": \"group of 15 x 16 B per thread behind a %d-DFMA chain, %d threads\", \"chain_only\": %.0f, \"cp_async_ring\": %.0f, \"ld_global_regs\": %.0f}\n", work, nthreads, base, a, b);
      }
    }
  }
  CK(cudaGetLastError());
  return 0;
}
