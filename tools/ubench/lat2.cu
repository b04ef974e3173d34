#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__global__ void k_rcp(double* out, long long* cyc, int iters) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = fast_rcp(a) + 0.5; a = fast_rcp(a) + 0.5; }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[1 + threadIdx.x] = a;
}
// token passed round-robin between W warps: warp w waits on barrier, reads token from smem, adds, writes, arrives.
__global__ void k_pingpong(double* out, long long* cyc, int iters, int W) {
  __shared__ double token[2];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { token[0] = 1.0; token[1] = 1.0; }
  __syncthreads();
  long long t0 = clock64();
  // step s: publisher = warp (s % W) writes token[s&1], arrives on barrier 1+(s&1); all others sync on it
  for (int s = 0; s < iters; s++) {
    const int id = 1 + (s & 1);
    if (wid == (s % W)) {
      if (lane == 0) token[s & 1] = token[(s + 1) & 1] * 1.0000001;
      __syncwarp();
      bar_arrive(id, 32 * W);
    } else {
      bar_sync(id, 32 * W);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[1 + threadIdx.x] = token[0];
}
// same with plain __syncthreads: publisher writes then everyone syncs
__global__ void k_syncpub(double* out, long long* cyc, int iters, int W) {
  __shared__ double token[2];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { token[0] = 1.0; token[1] = 1.0; }
  __syncthreads();
  long long t0 = clock64();
  for (int s = 0; s < iters; s++) {
    if (wid == (s % W) && lane == 0) token[s & 1] = token[(s + 1) & 1] * 1.0000001;
    __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[1 + threadIdx.x] = token[0];
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64);
  double init[1] = {1.7};
  cudaMemcpy(out, init, 8, cudaMemcpyHostToDevice);
  const int it = 4000;
  for (int r = 0; r < 2; r++) { k_rcp<<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("fast_rcp+add dependent: %.1f cycles\n", (double)h / (2 * it));
  for (int W = 2; W <= 8; W *= 2) {
    for (int r = 0; r < 2; r++) { k_pingpong<<<1, 32 * W>>>(out, cyc, it, W); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("arrive/sync token pass, %d warps: %.1f cycles/step\n", W, (double)h / it);
    for (int r = 0; r < 2; r++) { k_syncpub<<<1, 32 * W>>>(out, cyc, it, W); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("__syncthreads token pass, %d warps: %.1f cycles/step\n", W, (double)h / it);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
