// Streaming microbenchmark for the projection-solve sweeps (tools/, not part of the library):
// N tiles x 168 threads, every thread consumes one 48-byte record ([qw 16 B][K 32 B]) per step for 125 steps, as in
// constr_sweep, from HBM arrays much larger than L2.  Compares
//   mode 0  per-thread cp.async (16-byte LDGSTS) into a thread-private shared-memory ring   (what the kernels do)
//   mode 1  one cp.async.bulk per array and step for the whole tile (contiguous 2688 + 5376 bytes) + mbarriers
//   mode 2  plain ld.global.v2.f64 in the loop (no staging)
//   mode 3  mode 0 with cp.async.ca; mode 4  per-thread cp.async.cg issued so that every instruction covers whole sectors
// and prints the achieved DRAM bandwidth.  Shared memory per CTA is padded to 47 KB so 3 CTAs are resident per SM, as
// in k_leapfrog.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_bench tools/stream_bench.cu && /tmp/stream_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NT = 168, NSTEP = 125, NSL = 3;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(192, 3) k_stream(const double* __restrict__ qw, const double* __restrict__ K, double* out,
                                                   int pf) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const double* q0 = qw + (size_t)blockIdx.x * NSTEP * NT * 2;
  const double* k0 = K + (size_t)blockIdx.x * NSTEP * NT * 4;
  double acc0 = 0.0, acc1 = 0.0;
  if (MODE == 3 || MODE == 4) {
    // mode 3: as mode 0 with cp.async.ca (the second half of every K sector then hits L1 instead of going back to L2)
    // mode 4: cooperative full-sector copies: slot layout [qw slab][K slab]; lane l of warp w copies the 16-byte
    //         pieces l, l + 32 of the warp's contiguous K slab (every instruction covers whole sectors)
    const int lane = tid & 31, w0 = tid & ~31;
    const int wn = (NT - w0) < 32 ? (NT - w0) : 32;   // threads of this warp
    const unsigned base = smem_u32(smem);
    constexpr unsigned SLOT = NT * 48;
    auto fetch = [&](int n) {
      const unsigned sl = base + (n % NSL) * SLOT;
      if (MODE == 3) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sl + tid * 16), "l"(q0 + (size_t)n * NT * 2 + tid * 2));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sl + NT * 16 + tid * 32), "l"(k0 + (size_t)n * NT * 4 + tid * 4));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sl + NT * 16 + tid * 32 + 16), "l"(k0 + (size_t)n * NT * 4 + tid * 4 + 2));
      } else {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sl + tid * 16), "l"(q0 + (size_t)n * NT * 2 + tid * 2));
        const unsigned kd = sl + NT * 16 + w0 * 32;
        const double* ks = k0 + (size_t)n * NT * 4 + w0 * 4;
        if (lane < wn) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(kd + lane * 16), "l"(ks + lane * 2));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(kd + (lane + wn) * 16), "l"(ks + (lane + wn) * 2));
        }
      }
      asm volatile("cp.async.commit_group;\n");
    };
    for (int i = 0; i < NSL - 1; ++i) fetch(i);
    for (int s = 0; s < NSTEP; ++s) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(NSL - 2) : "memory");
      if (MODE == 4) __syncwarp();
      const unsigned r = base + (s % NSL) * SLOT;
      double v0, v1, a, b, c, d;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v0), "=d"(v1) : "r"(r + tid * 16));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a), "=d"(b) : "r"(r + NT * 16 + tid * 32));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(c), "=d"(d) : "r"(r + NT * 16 + tid * 32 + 16));
      if (MODE == 4) __syncwarp();
      if (s + NSL - 1 < NSTEP) fetch(s + NSL - 1);
      else asm volatile("cp.async.commit_group;\n");
      acc0 = fma(a, v0, fma(b, v1, acc0 * 0.999));
      acc1 = fma(c, v0, fma(d, v1, acc1 * 0.999));
    }
  } else if (MODE == 0) {
    const unsigned ring0 = smem_u32(smem) + tid * 48, sstride = NT * 48;
    for (int i = 0; i < NSL - 1; ++i) {
      const unsigned w = ring0 + i * sstride;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w), "l"(q0 + (size_t)i * NT * 2 + tid * 2));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w + 16), "l"(k0 + (size_t)i * NT * 4 + tid * 4));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w + 32), "l"(k0 + (size_t)i * NT * 4 + tid * 4 + 2));
      asm volatile("cp.async.commit_group;\n");
    }
    for (int s = 0; s < NSTEP; ++s) {
      asm volatile("cp.async.wait_group %0;\n" ::"n"(NSL - 2) : "memory");
      const unsigned r = ring0 + (s % NSL) * sstride;
      double v0, v1, a, b, c, d;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v0), "=d"(v1) : "r"(r));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a), "=d"(b) : "r"(r + 16));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(c), "=d"(d) : "r"(r + 32));
      const int n = s + NSL - 1;
      if (n < NSTEP) {
        const unsigned w = ring0 + (n % NSL) * sstride;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w), "l"(q0 + (size_t)n * NT * 2 + tid * 2));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w + 16), "l"(k0 + (size_t)n * NT * 4 + tid * 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(w + 32), "l"(k0 + (size_t)n * NT * 4 + tid * 4 + 2));
        if (pf > 0 && n + pf < NSTEP) {
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(q0 + (size_t)(n + pf) * NT * 2 + tid * 2));
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(k0 + (size_t)(n + pf) * NT * 4 + tid * 4));
        }
      }
      asm volatile("cp.async.commit_group;\n");
      acc0 = fma(a, v0, fma(b, v1, acc0 * 0.999));
      acc1 = fma(c, v0, fma(d, v1, acc1 * 0.999));
    }
  } else if (MODE == 1) {
    // slot layout: [qw slab NT*16 B][K slab NT*32 B]; barriers behind the slots
    constexpr unsigned SLOT = NT * 48;
    const unsigned base = smem_u32(smem);
    const unsigned full0 = base + NSL * SLOT, empty0 = full0 + NSL * 8;
    if (tid == 0) {
      for (int i = 0; i < NSL; ++i) { mbar_init(full0 + i * 8, 1); mbar_init(empty0 + i * 8, NT); }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < NSL - 1; ++i) {
        mbar_expect_tx(full0 + i * 8, SLOT);
        bulk_g2s(base + i * SLOT, q0 + (size_t)i * NT * 2, NT * 16, full0 + i * 8);
        bulk_g2s(base + i * SLOT + NT * 16, k0 + (size_t)i * NT * 4, NT * 32, full0 + i * 8);
      }
    }
    for (int s = 0; s < NSTEP; ++s) {
      const int sl = s % NSL;
      if (tid == 0) {
        // refill the slot consumed in step s - 1 with step s + NSL - 1
        const int n = s + NSL - 1;
        if (n < NSTEP) {
          const int ws = n % NSL;
          if (s > 0) mbar_wait(empty0 + ws * 8, ((s - 1) / NSL) & 1);
          mbar_expect_tx(full0 + ws * 8, SLOT);
          bulk_g2s(base + ws * SLOT, q0 + (size_t)n * NT * 2, NT * 16, full0 + ws * 8);
          bulk_g2s(base + ws * SLOT + NT * 16, k0 + (size_t)n * NT * 4, NT * 32, full0 + ws * 8);
        }
      }
      mbar_wait(full0 + sl * 8, (s / NSL) & 1);
      const unsigned r = base + sl * SLOT;
      double v0, v1, a, b, c, d;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v0), "=d"(v1) : "r"(r + tid * 16));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(a), "=d"(b) : "r"(r + NT * 16 + tid * 32));
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(c), "=d"(d) : "r"(r + NT * 16 + tid * 32 + 16));
      mbar_arrive(empty0 + sl * 8);
      acc0 = fma(a, v0, fma(b, v1, acc0 * 0.999));
      acc1 = fma(c, v0, fma(d, v1, acc1 * 0.999));
    }
  } else {
    for (int s = 0; s < NSTEP; ++s) {
      const double2 v = *reinterpret_cast<const double2*>(q0 + (size_t)s * NT * 2 + tid * 2);
      const double2 ab = *reinterpret_cast<const double2*>(k0 + (size_t)s * NT * 4 + tid * 4);
      const double2 cd = *reinterpret_cast<const double2*>(k0 + (size_t)s * NT * 4 + tid * 4 + 2);
      acc0 = fma(ab.x, v.x, fma(ab.y, v.y, acc0 * 0.999));
      acc1 = fma(cd.x, v.x, fma(cd.y, v.y, acc1 * 0.999));
    }
  }
  out[(size_t)blockIdx.x * NT + tid] = acc0 + acc1;
}

int main(int argc, char** argv) {
  const int tiles = argc > 1 ? atoi(argv[1]) : 8192;
  const size_t nq = (size_t)tiles * NSTEP * NT * 2, nk = nq * 2;
  double *qw, *K, *out;
  CK(cudaMalloc(&qw, nq * 8)); CK(cudaMalloc(&K, nk * 8)); CK(cudaMalloc(&out, (size_t)tiles * NT * 8));
  CK(cudaMemset(qw, 0, nq * 8)); CK(cudaMemset(K, 0, nk * 8));
  const size_t smem = 47104;
  CK(cudaFuncSetAttribute(k_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double gb = (double)(nq + nk) * 8 / 1e9;
  auto run = [&](int mode, int pf) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      if (mode == 0) k_stream<0><<<tiles, NT, smem>>>(qw, K, out, pf);
      else if (mode == 1) k_stream<1><<<tiles, NT, smem>>>(qw, K, out, pf);
      else if (mode == 2) k_stream<2><<<tiles, NT, smem>>>(qw, K, out, pf);
      else if (mode == 3) k_stream<3><<<tiles, NT, smem>>>(qw, K, out, pf);
      else k_stream<4><<<tiles, NT, smem>>>(qw, K, out, pf);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    printf("{\"mode\": %d, \"l2_prefetch\": %d, \"tiles\": %d, \"GB\": %.2f, \"ms\": %.3f, \"GBps\": %.0f}\n", mode, pf, tiles, gb, best, gb / (best * 1e-3));
  };
  run(0, 0); run(0, 8); run(1, 0); run(2, 0); run(3, 0); run(4, 0);
  return 0;
}
