// Stand-alone replica of the register-panel elimination loop of csrc/gp_fit.cu for quick
// what-if timing (variants selected by VARIANT).
#include <cstdio>
#include <cuda_runtime.h>
#ifndef VARIANT
#define VARIANT 0
#endif
#define T 4
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
__global__ void __launch_bounds__(256, 1) k_elim(const double* K, double* out, long long* cyc, int n, int reps) {
  long long tl[6] = {0, 0, 0, 0, 0, 0};
  __shared__ double colbuf[2 * COLBUF];
  __shared__ double pivs[128];
  const int tid = threadIdx.x;
  const int tx = tid >> 4, ty = tid & 15;
  double acc = 0.0;
  long long total = 0;
  for (int rep = 0; rep < reps; rep++) {
    double a[T][T];
#pragma unroll
    for (int ia = 0; ia < T; ia++)
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int i = ty + 16 * ia, j = tx + 16 * jb;
        a[ia][jb] = (i <= n && j < n && i >= j) ? K[i * 64 + j] : 0.0;
      }
    __syncthreads();
    const long long t0 = clock64();
    const bool diag_lower = ty >= tx;
    double ip_next = 0.0;
    if (tid == 0) ip_next = fast_rcp(a[0][0]);
    for (int k = 0; k < n; k++) {
      double* buf = colbuf + (k & 1) * COLBUF;
      double* bufU = buf + 128;
      const int kb = k >> 4, kt = k & 15;
      const bool rec = false;
      if (rec && k == 20) tl[0] = clock64();
      if (rec && k == 21) tl[5] = clock64();
#if VARIANT == 5
      if (tx == kt) {
#pragma unroll
        for (int q = 0; q < T; q++) {
          if (q == kb) {
            double2* b2 = reinterpret_cast<double2*>(buf + ty * T);
#pragma unroll
            for (int ia = 0; ia < T; ia += 2) b2[ia >> 1] = make_double2(a[ia][q], a[ia + 1][q]);
            if (ty == kt) *reinterpret_cast<double2*>(buf + 384) = make_double2(a[q][q], ip_next);
          }
        }
      }
#elif VARIANT != 1
      if (tx == kt) {
#pragma unroll
        for (int q = 0; q < T; q++) {
          if (q == kb) {
#pragma unroll
            for (int ia = 0; ia < T; ia++) {
              buf[ty + 16 * ia] = a[ia][q];
              bufU[ty + 16 * ia] = a[ia][q];
            }
            if (ty == kt) { bufU[k] = 1.0; buf[384] = a[q][q]; buf[385] = ip_next; }
          }
        }
      }
#endif
      if (rec && k == 20) tl[1] = clock64();
#if VARIANT != 2
      __syncthreads();
#endif
#if VARIANT == 5
      const double2 pp = *reinterpret_cast<const double2*>(buf + 384);
      const double piv = pp.x, ip = pp.y;
      if (!(piv > 0.0)) break;
      if (rec && k == 20) tl[2] = clock64();
      if (tid == 0) pivs[k] = piv;
      double ck_lo[T], ck_up[T], fr[T];
      {
        const double2* c2 = reinterpret_cast<const double2*>(buf + ty * T);
        const double2* f2 = reinterpret_cast<const double2*>(buf + tx * T);
#pragma unroll
        for (int ia = 0; ia < T; ia += 2) {
          const double2 v = c2[ia >> 1], w = f2[ia >> 1];
          ck_lo[ia] = v.x; ck_lo[ia + 1] = v.y; fr[ia] = w.x; fr[ia + 1] = w.y;
        }
#pragma unroll
        for (int ia = 0; ia < T; ia++) {
          const int i = ty + 16 * ia;
          ck_up[ia] = (i < k) ? ck_lo[ia] : ((i == k) ? 1.0 : 0.0);
        }
      }
#else
      const double piv = buf[384];
#if VARIANT != 1
      if (!(piv > 0.0)) break;
#endif
      const double ip = buf[385];
      if (tid == 0) pivs[k] = piv;
      double ck_lo[T], ck_up[T];
#pragma unroll
      for (int ia = 0; ia < T; ia++) {
        const int i = ty + 16 * ia;
        ck_lo[ia] = buf[i];
        ck_up[ia] = (i <= k) ? bufU[i] : 0.0;
      }
#endif
      const int jb0 = (k + 1) >> 4;
#if VARIANT != 3
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        if (jb >= jb0) {
          const int j = tx + 16 * jb;
#if VARIANT == 5
          const double f = (j > k && j < n) ? fr[jb] * ip : 0.0;
#else
          const double f = (j > k && j < n) ? buf[j] * ip : 0.0;
#endif
#pragma unroll
          for (int ia = 0; ia < T; ia++) {
            const double ck = (ia > jb) ? ck_lo[ia] : ((ia < jb) ? ck_up[ia] : (diag_lower ? ck_lo[ia] : ck_up[ia]));
            a[ia][jb] -= ck * f;
          }
        }
      }
#else
      acc += ck_lo[0] + ck_up[1] + ip * jb0;
#endif
      if (rec && k == 20) { tl[3] = clock64() + (long long)(a[0][0] * 0.0 + a[1][1] * 0.0 + a[2][2] * 0.0 + a[3][3] * 0.0 + a[3][0] * 0.0 + a[0][3] * 0.0); }
#if VARIANT != 4
      if (tx == ((k + 1) & 15) && ty == tx) {
        const int kn = (k + 1) >> 4;
#pragma unroll
        for (int q = 0; q < T; q++)
          if (q == kn) ip_next = fast_rcp(a[q][q]);
      }
#endif
    }
    total += clock64() - t0;
    if (rep == reps - 1 && (tid & 31) == 0) for (int q = 0; q < 6; q++) cyc[8 + (tid >> 5) * 6 + q] = tl[q];
#pragma unroll
    for (int ia = 0; ia < T; ia++)
#pragma unroll
      for (int jb = 0; jb < T; jb++) acc += a[ia][jb];
    __syncthreads();
  }
  out[tid] = acc + pivs[tid & 63];
  if (tid == 0) cyc[0] = total;
}
int main() {
  const int n = 60;
  double* hK = new double[64 * 64];
  for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) hK[i * 64 + j] = (i == j) ? 2.0 + 1e-3 * i : 0.5 / (1.0 + abs(i - j));
  for (int j = 0; j < 64; j++) hK[n * 64 + j] = 0.1 * j;
  double *K, *out; long long* cyc; long long h;
  cudaMalloc(&K, 64 * 64 * 8); cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8 * 64);
  cudaMemcpy(K, hK, 64 * 64 * 8, cudaMemcpyHostToDevice);
  const int reps = 200;
  for (int r = 0; r < 2; r++) { k_elim<<<1, 256>>>(K, out, cyc, n, reps); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  long long hh[64]; cudaMemcpy(hh, cyc, 8 * 64, cudaMemcpyDeviceToHost);
  long long base = hh[8];
  for (int w = 0; w < 8; w++) { printf("warp %d:", w); for (int q = 0; q < 6; q++) printf(" %6lld", hh[8 + w * 6 + q] - base); printf("\n"); }
  printf("VARIANT %d: %.1f cycles per elimination (n=%d) = %.1f cycles/step  [%s]\n", VARIANT, (double)h / reps, n,
         (double)h / reps / n, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
