// Latency of the cross-lane primitives the warp optimiser is built from (one warp running alone):
// dependent-chain cycles per iteration of (a) an all-gather through shared memory (STS, __syncwarp,
// N/2 LDS.128, local dot), (b) an all-gather by N double shuffles, (c) a 5-level xor-butterfly sum,
// (d) one broadcast shuffle, (e) IEEE division, rcp-based reciprocal, sqrt, rsqrt-based 1/sqrt.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__device__ __forceinline__ double vdotN(const double (&a)[N], const double (&b)[N]) {
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
  for (int j = 0; j < N; j += 4) { a0 = fma(a[j], b[j], a0); a1 = fma(a[j + 1], b[j + 1], a1); a2 = fma(a[j + 2], b[j + 2], a2); a3 = fma(a[j + 3], b[j + 3], a3); }
  return (a0 + a1) + (a2 + a3);
}
template <int N, int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double scr[4][32];
  const int lane = threadIdx.x;
  double w[N];
#pragma unroll
  for (int j = 0; j < N; j++) w[j] = out[j] * 1e-3;
  double v = out[0] + lane * 1e-3;
  int slot = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double g[N];
    if (MODE == 0) {
      double* b = scr[slot]; slot = (slot + 1) & 3;
      b[lane] = v;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < N; j += 2) { const double2 q = *reinterpret_cast<const double2*>(b + j); g[j] = q.x; g[j + 1] = q.y; }
      v = vdotN<N>(w, g);
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < N; j++) g[j] = __shfl_sync(0xffffffffu, v, j);
      v = vdotN<N>(w, g);
    } else if (MODE == 2) {
      double s = v * w[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      v = s;
    } else if (MODE == 3) {
      v = __shfl_sync(0xffffffffu, v, it & 7) * w[0];
    } else if (MODE == 4) {
      v = w[0] / v;
    } else if (MODE == 5) {
      double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
      double e = fma(-v, r, 1.0); r = fma(r, e, r); e = fma(-v, r, 1.0); r = fma(r, e, r);
      v = r;
    } else if (MODE == 6) {
      v = sqrt(v) + 1.0;
    } else if (MODE == 7) {
      double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
      double h = 0.5 * r, gg = v * r; double e = fma(-h, gg, 0.5); gg = fma(gg, e, gg); h = fma(h, e, h); e = fma(-h, gg, 0.5); h = fma(h, e, h);
      v = 2.0 * h + 1.0;
    } else if (MODE == 8) {   // 16-lane butterfly (4 levels)
      double s = v * w[0];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      v = s;
    } else if (MODE == 9) {   // plain dependent vdot without any exchange (baseline to subtract)
#pragma unroll
      for (int j = 0; j < N; j++) g[j] = v + j;
      v = vdotN<N>(w, g);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[64 + threadIdx.x] = v;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
  double init[32]; for (int i = 0; i < 32; i++) init[i] = 1.0 + 0.01 * i;
  cudaMemcpy(out, init, sizeof(init), cudaMemcpyHostToDevice);
  const int it = 2000;
#define RUN(N, MODE, name) \
  for (int r = 0; r < 2; r++) { k<N, MODE><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); } \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-52s N=%2d: %.1f cycles per dependent iteration\n", name, N, (double)h / it);
  RUN(12, 9, "vdot only (baseline)") RUN(24, 9, "vdot only (baseline)")
  RUN(12, 0, "smem all-gather + vdot") RUN(24, 0, "smem all-gather + vdot")
  RUN(12, 1, "shuffle all-gather + vdot") RUN(24, 1, "shuffle all-gather + vdot")
  RUN(12, 2, "32-lane xor butterfly sum") RUN(12, 8, "16-lane xor butterfly sum")
  RUN(12, 3, "one broadcast shuffle + mul")
  RUN(12, 4, "IEEE division") RUN(12, 5, "rcp.approx + 2 Newton") RUN(12, 6, "sqrt + add") RUN(12, 7, "rsqrt.approx + 2 Newton + add")
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
