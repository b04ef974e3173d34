// Stand-alone replica of the symmetric-sweep step of csrc/gp_nll.cuh (sweep_block) for what-if timing of the
// thread-grid / tile shape: GY x GX threads, TY x TX register tile per thread (GY*TY = GX*TX = 64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sweep sweep.cu && ./sweep
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#define COLBUF 392
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ void sts2_if(double* p, double x, double y, bool pred) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(addr), "d"(x), "d"(y), "r"((int)pred) : "memory");
}
template <int NT>
__device__ __forceinline__ void bar() {
  if (NT == 256) __syncthreads();
  else asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

template <int GY, int GX, int TY, int TX, int KB>
__device__ __forceinline__ void sweep_block(double (&a)[TY][TX], double* colbuf, double* pivs, int n, int ty, int tx,
                                            double& ip_next, bool& bad) {
  constexpr int NT = GY * GX;
  constexpr int TP = (TY + 1) & ~1;
  constexpr int RB = (GX * KB) / GY;                       // row block of the pivots of column block KB
  const int kend = (n - GX * KB) < GX ? (n - GX * KB) : GX;
  for (int kt = 0; kt < kend; kt++) {
    const int k = GX * KB + kt;
    double* buf = colbuf + (kt & 1) * COLBUF;
    const bool col_eq = (tx == kt), row_eq = (ty == (k % GY)), own = col_eq && row_eq;
#pragma unroll
    for (int q = 0; q < TP; q += 2)
      sts2_if(buf + ty * TP + q, a[q][KB], (q + 1 < TY) ? a[(q + 1 < TY) ? q + 1 : q][KB] : 0.0, col_eq);
    sts2_if(buf + 384, a[RB][KB], ip_next, own);
    if (own) pivs[k] = a[RB][KB];
    bar<NT>();
    const double2 pp = *reinterpret_cast<const double2*>(buf + 384);
    bad = bad || !(pp.x > 0.0);
    const double ip = pp.y;
    double ck[TP], f[TX];
    {
      const double2* c2 = reinterpret_cast<const double2*>(buf + ty * TP);
#pragma unroll
      for (int q = 0; q < TP; q += 2) { const double2 v = c2[q >> 1]; ck[q] = v.x; ck[q + 1] = v.y; }
#pragma unroll
      for (int jb = 0; jb < TX; jb++) {
        const int j = tx + GX * jb;
        f[jb] = buf[(j % GY) * TP + (j / GY)] * ip;
      }
    }
#pragma unroll
    for (int jb = 0; jb < TX; jb++) {
#pragma unroll
      for (int ia = 0; ia < TY; ia++) {
        double v = fma(-ck[ia], f[jb], a[ia][jb]);
        if (ia == RB) v = row_eq ? f[jb] : v;
        if (jb == KB) v = col_eq ? ck[ia] * ip : v;
        if (ia == RB && jb == KB) v = own ? -ip : v;
        a[ia][jb] = v;
      }
    }
    {
      // next pivot (k+1, k+1): column block KB (or KB+1 when kt == GX-1), row block (k+1)/GY
      constexpr int KN = (KB + 1 < TX) ? KB + 1 : KB;
      constexpr int RN = (GX * KN) / GY;
      const double pv = (kt < GX - 1) ? a[RB][KB] : a[RN][KN];
      // (for GY > GX the row block can also change inside a column block; both candidates are static)
      constexpr int RB2 = (GX * KB + GX - 1) / GY;
      const double pv2 = (RB2 != RB && ((k + 1) / GY) == RB2 && kt < GX - 1) ? a[RB2][KB] : pv;
      ip_next = fast_rcp(pv2);
    }
  }
}
template <int GY, int GX, int TY, int TX, int KB>
__device__ __forceinline__ void sweep_all(double (&a)[TY][TX], double* colbuf, double* pivs, int n, int ty, int tx,
                                          double& ip_next, bool& bad) {
  if (GX * KB < n) sweep_block<GY, GX, TY, TX, KB>(a, colbuf, pivs, n, ty, tx, ip_next, bad);
  if constexpr (KB + 1 < TX) sweep_all<GY, GX, TY, TX, KB + 1>(a, colbuf, pivs, n, ty, tx, ip_next, bad);
}

template <int GY, int GX, int TY, int TX, int LAUNCH>
__global__ void __launch_bounds__(LAUNCH, 1) k_sweep(const double* K, double* out, long long* cyc, int n, int reps) {
  __shared__ __align__(16) double colbuf[2 * COLBUF];
  __shared__ double pivs[128];
  const int tid = threadIdx.x;
  constexpr int NT = GY * GX;
  const int ty = tid % GY, tx = tid / GY;
  long long total = 0;
  double a[TY][TX];
  for (int rep = 0; rep < reps; rep++) {
    __syncthreads();
    if (tid < NT) {
#pragma unroll
      for (int ia = 0; ia < TY; ia++)
#pragma unroll
        for (int jb = 0; jb < TX; jb++) {
          const int i = ty + GY * ia, j = tx + GX * jb;
          a[ia][jb] = (i <= n && j <= n) ? K[i * 64 + j] : 0.0;
        }
      const long long t0 = clock64();
      double ip_next = fast_rcp(a[0][0]);
      bool bad = false;
      sweep_all<GY, GX, TY, TX, 0>(a, colbuf, pivs, n, ty, tx, ip_next, bad);
      total += clock64() - t0;
      if (bad && rep == 0 && tid == 0) printf("bad pivot\n");
    }
  }
  __syncthreads();
  if (tid < NT) {
#pragma unroll
    for (int ia = 0; ia < TY; ia++)
#pragma unroll
      for (int jb = 0; jb < TX; jb++) {
        const int i = ty + GY * ia, j = tx + GX * jb;
        out[i * 64 + j] = a[ia][jb];
      }
  }
  if (tid == 0) cyc[0] = total;
}

static void host_sweep(double* A, int n) {   // A: 64x64 full symmetric augmented
  for (int k = 0; k < n; k++) {
    const double p = A[k * 64 + k], ip = 1.0 / p;
    double c[64], f[64];
    for (int i = 0; i < 64; i++) { c[i] = A[i * 64 + k]; f[i] = A[k * 64 + i] * ip; }
    for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) A[i * 64 + j] -= c[i] * f[j];
    for (int i = 0; i < 64; i++) { A[i * 64 + k] = c[i] * ip; A[k * 64 + i] = f[i]; }
    A[k * 64 + k] = -ip;
  }
}

template <int GY, int GX, int TY, int TX, int LAUNCH>
static void run(const char* name, const double* dK, const double* ref, double* out, long long* cyc, int n) {
  const int reps = 200;
  for (int r = 0; r < 2; r++) { k_sweep<GY, GX, TY, TX, LAUNCH><<<1, LAUNCH>>>(dK, out, cyc, n, reps); cudaDeviceSynchronize(); }
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  static double ho[64 * 64]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  double err = 0, mx = 0;
  for (int i = 0; i <= n; i++) for (int j = 0; j <= n; j++) { err = fmax(err, fabs(ho[i * 64 + j] - ref[i * 64 + j])); mx = fmax(mx, fabs(ref[i * 64 + j])); }
  printf("%-44s %8.1f cycles per sweep (n=%d) = %6.1f cycles/step   max err %.1e (scale %.1e) [%s]\n", name, (double)h / reps, n,
         (double)h / reps / n, err, mx, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = 60;
  static double hK[64 * 64], ref[64 * 64];
  for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) hK[i * 64 + j] = (i < n && j < n) ? ((i == j) ? 2.0 + 1e-3 * i : 0.5 / (1.0 + abs(i - j))) : 0.0;
  for (int j = 0; j < n; j++) { hK[n * 64 + j] = 0.1 * j - 2.0; hK[j * 64 + n] = 0.1 * j - 2.0; }
  for (int i = 0; i < 64 * 64; i++) ref[i] = hK[i];
  host_sweep(ref, n);
  double *K, *out; long long* cyc;
  cudaMalloc(&K, sizeof(hK)); cudaMalloc(&out, sizeof(hK)); cudaMalloc(&cyc, 64);
  cudaMemcpy(K, hK, sizeof(hK), cudaMemcpyHostToDevice);
  run<16, 16, 4, 4, 256>("16x16 threads, 4x4 tile (product)", K, ref, out, cyc, n);
  run<16, 8, 4, 8, 128>("16x8 threads, 4x8 tile, 128 launched", K, ref, out, cyc, n);
  run<16, 8, 4, 8, 256>("16x8 threads, 4x8 tile, 256 launched", K, ref, out, cyc, n);
  run<8, 8, 8, 8, 64>("8x8 threads, 8x8 tile, 64 launched", K, ref, out, cyc, n);
  run<8, 8, 8, 8, 256>("8x8 threads, 8x8 tile, 256 launched", K, ref, out, cyc, n);
  run<16, 4, 4, 16, 64>("16x4 threads, 4x16 tile, 64 launched", K, ref, out, cyc, n);
  run<8, 4, 8, 16, 32>("8x4 threads (one warp), 8x16 tile", K, ref, out, cyc, n);
  return 0;
}
