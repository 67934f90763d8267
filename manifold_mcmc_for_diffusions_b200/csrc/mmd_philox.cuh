// Philox-4x32-10 counter-based generator + Box-Muller in double precision for the on-device
// momentum draws (reference: rng.standard_normal in sample_momentum, mici_extensions.py:1256-1259).
// Counter = (element index, offset), key = seed: the stream is a pure function of
// (seed, offset, element), so draws are reproducible regardless of grid shape or GPU count.
#pragma once
#include <stdint.h>
#include "mmd_common.cuh"

namespace mmd {

MMD_HD static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += W0; k1 += W1;
  }
}

// two standard normals from 128 random bits
MMD_HD static void philox_normal_pair(uint64_t seed, uint64_t offset, uint64_t idx, double* n0, double* n1) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)c[1] << 32) | c[0], b = ((uint64_t)c[3] << 32) | c[2];
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
  const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
  const double r = sqrt(-2.0 * log(u1));
  const double th = 6.283185307179586476925286766559 * u2;
  *n0 = r * cos(th);
  *n1 = r * sin(th);
}

#if defined(__CUDACC__)
static __global__ void k_philox_normal(double* out, long long n, uint64_t seed, uint64_t offset) {
  const long long npair = (n + 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npair;
       i += (long long)gridDim.x * blockDim.x) {
    double a, b;
    philox_normal_pair(seed, offset, (uint64_t)i, &a, &b);
    out[2 * i] = a;
    if (2 * i + 1 < n) out[2 * i + 1] = b;
  }
}
#endif

}  // namespace mmd
