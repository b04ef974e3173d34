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
  int has_cond;          // the space has conditions: kernel values carry the activity mask of gp_kernels.py:21-29
  int perm[TLBO_HMAX];   // perm[p] = original column of permuted column p (cont first, then cat)
};

// Activity pattern of a row (reference tlbo/model/gp_kernels.py:21-29, get_conditional_hyperparameters): bit p
// set iff x_p <= -1 (an inactive hyper-parameter after GaussianProcess._impute_inactive, base_gp.py:148-151).
// k(x, x') is multiplied by [pattern(x) == pattern(x')] in every kernel of the stack when the space has conditions.
__device__ __forceinline__ unsigned inactive_pattern(const double* x, int d) {
  unsigned m = 0u;
  for (int p = 0; p < d; p++) m |= (x[p] <= -1.0) ? (1u << p) : 0u;
  return m;
}

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

// Sum over the 32 lanes of each of 16 per-lane values: on return lanes r and r + 16 hold sum_lanes v[r].  Recursive
// halving -- 16 exchanges instead of the 80 of 16 separate butterflies.
__device__ __forceinline__ double warp_transpose_sum16(double (&v)[16]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int r = 0; r < h; r++) {
      const double send = up ? v[r] : v[r + h];
      const double keep = up ? v[r + h] : v[r];
      v[r] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
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

// exp(x) for x <= 0 (the only range the Matern / Hamming kernels need): Cody-Waite reduction
// x = k ln2 + r, |r| <= ln2/2, degree-12 Taylor polynomial in Horner form, exponent insertion.
// Relative error < 4e-16 on [-708, 0]; returns 0 below (the libm result is subnormal there and
// far under the 1e-6 parity budget).  About half the instructions of the libm exp (no special
// cases), which ncu shows to be 40% of the fused predict kernel's instruction stream.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double L2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  if (!(x > -708.0)) return (x != x) ? x : 0.0;
  const double kf = rint(x * L2E);
  const double r = fma(-kf, LN2_LO, fma(-kf, LN2_HI, x));
  double p = 2.08767569878681e-09;                    // 1/12!
  p = fma(p, r, 2.505210838544172e-08);                // 1/11!
  p = fma(p, r, 2.755731922398589e-07);                // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);               // 1/9!
  p = fma(p, r, 2.48015873015873e-05);                 // 1/8!
  p = fma(p, r, 1.984126984126984e-04);                // 1/7!
  p = fma(p, r, 1.388888888888889e-03);                // 1/6!
  p = fma(p, r, 8.333333333333333e-03);                // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);               // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);               // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k = (int)kf;                                // -1022 < k <= 0 here
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(x) for finite normal x > 0: x = m 2^e with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m - 1) / (m + 1)
// (|s| <= 0.1716: eleven terms of the odd series), the reciprocal from the hardware seed + two Newton steps.
// Error <= 2 ulp; ~35 FP64 operations on one dependent chain of ~300 cycles where the libm call (special cases,
// table look-ups, IEEE division) takes 2-3 times as long -- it sits in a phase other warps wait for.
__device__ __forceinline__ double log_pos(double x) {
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, lo);
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double den = m + 1.0;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  double t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  t = fma(-den, r, 1.0);
  r = fma(r, t, r);
  const double num = m - 1.0;
  double s = num * r;
  s = fma(fma(-s, den, num), r, s);                   // one correction step: s = (m - 1) / (m + 1) to < 1 ulp
  const double z = s * s;
  double p = 1.0 / 25.0;
  p = fma(p, z, 1.0 / 23.0);
  p = fma(p, z, 1.0 / 21.0);
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  const double lm = fma(2.0 * s, z * p, 2.0 * s);     // 2 atanh(s)
  const double ef = (double)e;
  return fma(ef, LN2_HI, fma(ef, LN2_LO, lm));
}

// sqrt(s) for s >= 0 from the hardware reciprocal-square-root seed and two Newton steps
// (error < 1 ulp of the result; s == 0 -> 0).
__device__ __forceinline__ double sqrt_nonneg(double s) {
  if (!(s > 0.0)) return (s != s) ? s : 0.0;
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  // one Newton step on r ~ 1/sqrt(s), then g = s r with a residual correction
  double h = 0.5 * r;
  double g = s * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, g, s);
  return fma(e, h, g);
}

// Matern-5/2 x Hamming kernel value from the squared scaled distance `s` (sum of
// ((x-x')/l)^2 over continuous columns) and the Hamming exponent `hs` (sum of
// [x != x'] / (2 l^2) over categorical columns); reference gp_kernels.py:433-451,706-718.
__device__ __forceinline__ double matern_hamming(double c, double s, double hs) {
  const double t = sqrt(s) * SQRT5;
  return c * ((1.0 + t + t * t / 3.0) * exp(-t - hs));
}

// Same value with the range-specialised sqrt / exp above (candidate-pool kernel).
// t^2 / 3 is a multiplication by the rounded 1/3 (<= 1 ulp from the IEEE quotient): the division is a
// ~25-instruction subroutine, a quarter of the instructions of one kernel value.
__device__ __forceinline__ double matern_hamming_fast(double c, double s, double hs) {
  const double t = sqrt_nonneg(s) * SQRT5;
  return c * ((1.0 + t + t * t * (1.0 / 3.0)) * exp_nonpos(-t - hs));
}
