// Stand-ins that let the per-pair CUDA kernels (one thread = one pair, no shared memory) compile as plain C++:
// tests/test_oracle_and_math.py builds tests/tools/emu_generic_act.cpp with -DFBP_HOST_EMU and runs the kernels'
// own source on the CPU, one "thread" at a time, each thread being lane 0 of a one-lane warp.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
inline void __syncthreads() {}

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

inline float __shfl_xor_sync(unsigned, float, int) { return 0.0f; }     // the other lanes of the warp do not exist
inline int __shfl_sync(unsigned, int v, int) { return v; }
inline bool __all_sync(unsigned, bool p) { return p; }
inline void __syncwarp() {}
inline float atomicAdd(float* a, float v) { float o = *a; *a += v; return o; }
inline void sincospif(float x, float* s, float* c) { *s = sinf(3.14159265358979323846f * x); *c = cosf(3.14159265358979323846f * x); }
