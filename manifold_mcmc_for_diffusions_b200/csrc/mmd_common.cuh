// Common macros for the CHMC kernels (host+device so the model functors can be unit-tested with g++).
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#define MMD_HD __host__ __device__ __forceinline__
#define MMD_D __device__ __forceinline__
#else
#define MMD_HD inline
#define MMD_D inline
#endif
