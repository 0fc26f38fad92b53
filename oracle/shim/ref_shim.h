// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Macro shim that lets g++ compile the reference's CUDA kernel *text* for the host.
// The reference selects its plain-array FETCH path when __DEVICE_EMULATION__ is
// defined (source/CUDA/Params.cuh:5-14) and then skips every texture declaration
// (source/CUDA/System.cu:20-31).  Everything below only supplies what nvcc would
// have supplied: execution-space keywords, the launch index variables, __mul24,
// __syncthreads, and float overloads of min/max/abs (the host branch of
// source/external/cutil_math.h:55-73 only has the int ones, and its fmaxf is a min).
#pragma once
#define __DEVICE_EMULATION__ 1

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>   // vector types + make_float3() etc. (host side only)
// <cuda_runtime.h> re-defines the execution-space keywords for a host compiler
// (to nothing / attributes); make sure they are inert here.
#undef __device__
#undef __global__
#undef __constant__
#undef __shared__
#undef __host__
#define __device__
#define __global__
#define __host__
#define __constant__ static
#define __shared__ static thread_local

struct RefLaunchIdx { unsigned x, y, z; };
static thread_local RefLaunchIdx blockIdx  = {0, 0, 0};
static thread_local RefLaunchIdx threadIdx = {0, 0, 0};
static thread_local RefLaunchIdx blockDim  = {1, 1, 1};

static inline int  __mul24(int a, int b)  { return a * b; }
static inline void __syncthreads()        {}

using std::min;
using std::max;
using std::abs;
static inline float min(float a, float b) { return a < b ? a : b; }
static inline float max(float a, float b) { return a > b ? a : b; }
