// Common macros for the CHMC kernels (host+device so the model functors can be unit-tested with g++).
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define MMD_HD __host__ __device__ __forceinline__
#define MMD_D __device__ __forceinline__
// The phases of a leapfrog step (projection, projection solve, linearisation) are inlined into the fused kernel by
// default.  -DMMD_NOINLINE_PHASES compiles them once per kernel and calls them instead (k_leapfrog 745 -> 372 KB of
// SASS, a third of the compile time): measured on B200 that is 6 % slower at 8 chains per tile (the tile shape the
// bench uses) and 15-25 % faster at 1-4 chains per tile, where the resident warps are in different phases and the
// instruction cache is the limiter (profiles/r2_icache_tiles.md).
#if defined(MMD_NOINLINE_PHASES)
#define MMD_PHASE __device__ __noinline__
#else
#define MMD_PHASE __device__ __forceinline__
#endif
#else
#define MMD_PHASE inline
#define MMD_HD inline
#define MMD_D inline
#endif
