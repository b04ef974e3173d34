// Checks and times the PRODUCT's sweep_rot (csrc/gp_nll.cuh) stand-alone against a host Gauss-Jordan sweep.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep2 sweep2.cu && ./sweep2
#include <cstdio>
#include <cmath>
#include "../../efficient-tlbo_b200/csrc/gp_nll.cuh"
void tlbo_set_error(const char*) {}
template <int GX, int TY, int TX>
__global__ void __launch_bounds__(256, 1) k_sweep(const double* K, double* out, long long* cyc, int n, int reps) {
  __shared__ __align__(16) double colbuf[2 * FIT_COLBUF];
  __shared__ double pivs[128];
  const int tid = threadIdx.x;
  constexpr int NTS = 16 * GX;
  const int ty = tid & 15, tx = tid >> 4;
  long long total = 0;
  double a[TY][TX];
  int rot = 0;
  for (int rep = 0; rep < reps; rep++) {
    __syncthreads();
    if (tid < NTS) {
#pragma unroll
      for (int ia = 0; ia < TY; ia++)
#pragma unroll
        for (int jb = 0; jb < TX; jb++) {
          const int i = ty + 16 * ia, j = tx + GX * jb;
          a[ia][jb] = (i <= n && j <= n) ? K[i * 128 + j] : 0.0;
        }
      const long long t0 = clock64();
      ElimCtx ec;
      ec.colbuf = colbuf; ec.pivs = pivs; ec.n = n; ec.tx = tx; ec.ty = ty; ec.bad = false;
      rot = sweep_rot<GX, TY, TX>(a, ec);
      total += clock64() - t0;
      if (ec.bad && rep == 0 && tid == 0) printf("bad pivot\n");
    }
  }
  __syncthreads();
  if (tid < NTS) {
#pragma unroll
    for (int ia = 0; ia < TY; ia++)
#pragma unroll
      for (int jb = 0; jb < TX; jb++) {
        const int i = ty + 16 * ia, j = tx + GX * ((jb + rot) % TX);
        out[i * 128 + j] = a[ia][jb] - ((i == j && i < n) ? 2.0 : 0.0);
      }
  }
  if (tid == 0) cyc[0] = total;
}
static void host_sweep(double* A, int n) {
  static double c[128], f[128];
  for (int k = 0; k < n; k++) {
    const double p = A[k * 128 + k], ip = 1.0 / p;
    for (int i = 0; i < 128; i++) { c[i] = A[i * 128 + k]; f[i] = A[k * 128 + i] * ip; }
    for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) A[i * 128 + j] -= c[i] * f[j];
    for (int i = 0; i < 128; i++) { A[i * 128 + k] = c[i] * ip; A[k * 128 + i] = f[i]; }
    A[k * 128 + k] = -ip;
  }
}
template <int GX, int TY, int TX>
static void run(const char* name, int n) {
  static double hK[128 * 128], ref[128 * 128], ho[128 * 128];
  for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) hK[i * 128 + j] = (i < n && j < n) ? ((i == j) ? 2.0 + 1e-3 * i : 0.5 / (1.0 + abs(i - j))) : 0.0;
  for (int j = 0; j < n; j++) { hK[n * 128 + j] = 0.1 * j - 2.0; hK[j * 128 + n] = 0.1 * j - 2.0; }
  for (int i = 0; i < 128 * 128; i++) ref[i] = hK[i];
  host_sweep(ref, n);
  double *K, *out; long long* cyc;
  cudaMalloc(&K, sizeof(hK)); cudaMalloc(&out, sizeof(hK)); cudaMalloc(&cyc, 64);
  cudaMemcpy(K, hK, sizeof(hK), cudaMemcpyHostToDevice);
  cudaMemset(out, 0, sizeof(hK));
  const int reps = 200;
  for (int r = 0; r < 2; r++) { k_sweep<GX, TY, TX><<<1, 256>>>(K, out, cyc, n, reps); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  double err = 0, mx = 0;
  for (int i = 0; i <= n; i++) for (int j = 0; j <= n; j++) { err = fmax(err, fabs(ho[i * 128 + j] - ref[i * 128 + j])); mx = fmax(mx, fabs(ref[i * 128 + j])); }
  printf("%-36s n=%3d %8.1f cycles per sweep = %6.1f cycles/step   max err %.1e (scale %.1e) [%s]\n", name, n, (double)h / reps,
         (double)h / reps / n, err, mx, cudaGetErrorString(cudaGetLastError()));
  cudaFree(K); cudaFree(out); cudaFree(cyc);
}
int main() {
  run<8, 4, 8>("16x8 threads, 4x8 tile", 60);
  run<8, 4, 8>("16x8 threads, 4x8 tile", 45);
  run<8, 4, 8>("16x8 threads, 4x8 tile", 63);
  run<8, 4, 8>("16x8 threads, 4x8 tile", 7);
  run<8, 5, 10>("16x8 threads, 5x10 tile", 75);
  run<8, 6, 12>("16x8 threads, 6x12 tile", 90);
  run<16, 7, 7>("16x16 threads, 7x7 tile", 100);
  run<16, 8, 8>("16x16 threads, 8x8 tile", 127);
  run<16, 4, 4>("16x16 threads, 4x4 tile", 60);
  return 0;
}
