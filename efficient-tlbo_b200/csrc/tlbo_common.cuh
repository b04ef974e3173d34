// Shared device helpers for the tlbo_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/tlbo_b200.h"

#define TLBO_HMAX 24
#define SQRT5 2.23606797749979
#define VERY_SMALL_NUMBER 1e-10
#define LOG_2PI 1.8378770664093453

// Input-space description, passed to kernels by value.
struct SpaceDesc {
  int d, d_cont, d_cat;
  int perm[TLBO_HMAX];   // perm[p] = original column of permuted column p (cont first, then cat)
};

void tlbo_set_error(const char* msg);
int tlbo_make_space(int d, const int32_t* h_types, SpaceDesc* sp);

#define TLBO_CUDA_CHECK(expr)                                   \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) {                                    \
      tlbo_set_error(cudaGetErrorString(_e));                   \
      return (int)_e;                                           \
    }                                                           \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `red` is shared scratch of >= 32 doubles.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; i++) t += red[i];
  return t;
}

// Matern-5/2 x Hamming kernel value from the squared scaled distance `s` (sum of
// ((x-x')/l)^2 over continuous columns) and the Hamming exponent `hs` (sum of
// [x != x'] / (2 l^2) over categorical columns); reference gp_kernels.py:433-451,706-718.
__device__ __forceinline__ double matern_hamming(double c, double s, double hs) {
  const double t = sqrt(s) * SQRT5;
  return c * ((1.0 + t + t * t / 3.0) * exp(-t - hs));
}
