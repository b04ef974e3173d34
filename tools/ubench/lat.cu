// Latency micro-benchmarks (cycles) used to model the GP-fit kernel's critical path.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, long long* cyc, int iters) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
__global__ void k_ddiv(double* out, long long* cyc, int iters) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = b / a; a = b / a; a = b / a; a = b / a; }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
__global__ void k_dsqrt(double* out, long long* cyc, int iters) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; a = sqrt(a) + 1.0; }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
__global__ void k_dexp(double* out, long long* cyc, int iters) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = exp(-a); a = exp(-a); a = exp(-a); a = exp(-a); }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
__global__ void k_lds(double* out, long long* cyc, int iters) {
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x + 1) & 255;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = p;
}
__global__ void k_bar(double* out, long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
}
// publish -> barrier -> load -> fma chain, as in the elimination step
__global__ void k_step(double* out, long long* cyc, int iters) {
  __shared__ double buf[2][64];
  double a = out[0] + threadIdx.x;
  buf[0][threadIdx.x & 63] = 1.0; buf[1][threadIdx.x & 63] = 1.0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    double* b = buf[i & 1];
    if ((threadIdx.x & 15) == (i & 15)) b[threadIdx.x >> 4] = a;
    __syncthreads();
    const double f = b[i & 15] * 1.0000001;
    a = fma(-b[(threadIdx.x >> 4)], f, a);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
__global__ void k_shfl(double* out, long long* cyc, int iters) {
  double a = out[0] + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    a += __shfl_xor_sync(0xffffffffu, a, 16); a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 4); a += __shfl_xor_sync(0xffffffffu, a, 2);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0); }
  out[3 + threadIdx.x] = a;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
  double init[3] = {1.000001, 0.999999, 1e-9};
  cudaMemcpy(out, init, 24, cudaMemcpyHostToDevice);
  const int it = 2000;
#define RUN(name, kern, threads, per) \
  for (int r = 0; r < 2; r++) { kern<<<1, threads>>>(out, cyc, it); cudaDeviceSynchronize(); } \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s threads=%4d  %.1f cycles/op\n", name, threads, (double)h / (it * per));
  RUN("dfma dependent", k_dfma, 32, 4)
  RUN("dfma dependent (8 warps)", k_dfma, 256, 4)
  RUN("ddiv dependent", k_ddiv, 32, 4)
  RUN("dsqrt+add dependent", k_dsqrt, 32, 4)
  RUN("dexp dependent", k_dexp, 32, 4)
  RUN("lds pointer chase", k_lds, 32, 4)
  RUN("__syncthreads (8 warps)", k_bar, 256, 4)
  RUN("__syncthreads (1 warp)", k_bar, 32, 4)
  RUN("elim step chain (8 warps)", k_step, 256, 1)
  RUN("shfl+dadd dependent", k_shfl, 32, 4)
  cudaError_t e = cudaGetLastError(); printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
