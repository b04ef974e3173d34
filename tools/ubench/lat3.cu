#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, int iters) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = out[0] + i + threadIdx.x;
  const double b = out[1], c = out[2];
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[3 + threadIdx.x] = s;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
  double init[3] = {1.000001, 0.999999, 1e-9};
  cudaMemcpy(out, init, 24, cudaMemcpyHostToDevice);
  const int it = 2000;
#define RUN(ILP, threads) \
  for (int r = 0; r < 2; r++) { k_dfma_ilp<ILP><<<1, threads>>>(out, cyc, it); cudaDeviceSynchronize(); } \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("ILP=%2d threads=%4d: %.2f cycles per iteration, %.2f cycles per warp-DFMA per SMSP-warp, SM rate %.1f lanes/clk\n", ILP, threads, (double)h / it, (double)h / it / ILP, (double)ILP * threads / ((double)h / it));
  RUN(1, 32) RUN(4, 32) RUN(16, 32) RUN(16, 128) RUN(16, 256) RUN(16, 512) RUN(16, 1024) RUN(4, 256) RUN(8, 256)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
