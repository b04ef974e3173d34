// Third-generation elimination step: the pivot-column block index KB is a template
// parameter (outer loop unrolled at compile time), so tile indexing, most masks and the
// inactive-column skipping are resolved statically and the step has no uniform branches.
#include <cstdio>
#include <cuda_runtime.h>
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
struct ElimCtx { long long tl[6]; bool recrep;
  double* colbuf; double* pivs; int n; int tx, ty; bool diag_lower, own_diag; double ip_next; bool bad;
};
template <int KB>
__device__ __forceinline__ void elim_block(double (&a)[T][T], ElimCtx& c) {
  const int n = c.n, tx = c.tx, ty = c.ty;
  const int kend = (n - 16 * KB) < 16 ? (n - 16 * KB) : 16;
  for (int kt = 0; kt < kend; kt++) {
    const int k = 16 * KB + kt;
    double* buf = c.colbuf + (kt & 1) * COLBUF;
    const bool rec = c.recrep && KB == 1;
    if (rec && kt == 4) c.tl[0] = clock64();
    if (rec && kt == 5) c.tl[5] = clock64();
    if (tx == kt) {
      double2* b2 = reinterpret_cast<double2*>(buf + ty * T);
#pragma unroll
      for (int ia = 0; ia < T; ia += 2) b2[ia >> 1] = make_double2(a[ia][KB], a[ia + 1][KB]);
      if (ty == kt) { *reinterpret_cast<double2*>(buf + 384) = make_double2(a[KB][KB], c.ip_next); c.pivs[k] = a[KB][KB]; }
    }
    const bool row_lt = ty < kt, row_eq = ty == kt, col_gt = tx > kt;
    if (rec && kt == 4) c.tl[1] = clock64();
    __syncthreads();
    const double2 pp = *reinterpret_cast<const double2*>(buf + 384);
    c.bad = c.bad || !(pp.x > 0.0);
    if (rec && kt == 4) c.tl[2] = clock64() + (long long)(pp.x * 0.0);
    const double ip = pp.y;
    double ck_lo[T], fr[T];
    {
      const double2* c2 = reinterpret_cast<const double2*>(buf + ty * T);
      const double2* f2 = reinterpret_cast<const double2*>(buf + tx * T);
#pragma unroll
      for (int q = 0; q < T; q += 2) {
        const double2 v = c2[q >> 1], w = f2[q >> 1];
        ck_lo[q] = v.x; ck_lo[q + 1] = v.y; fr[q] = w.x; fr[q + 1] = w.y;
      }
    }
    // inverse-part multiplier of row block ia: rows i < k fully active for ia < KB, none for ia > KB
    const double ck_up_kb = row_lt ? ck_lo[KB] : (row_eq ? 1.0 : 0.0);
#pragma unroll
    for (int jb = KB; jb < T; jb++) {
      double f = fr[jb] * ip;
      if (jb == KB) f = col_gt ? f : 0.0;
      if (jb == T - 1) f = (tx + 16 * jb < n) ? f : 0.0;
#pragma unroll
      for (int ia = 0; ia < T; ia++) {
        if (ia > jb) a[ia][jb] -= ck_lo[ia] * f;                        // lower part
        else if (ia < jb) {                                             // inverse part
          if (ia < KB) a[ia][jb] -= ck_lo[ia] * f;
          else if (ia == KB) a[ia][jb] -= ck_up_kb * f;
        } else {                                                        // diagonal tile
          const double cu = (ia < KB) ? ck_lo[ia] : ((ia == KB) ? ck_up_kb : 0.0);
          a[ia][jb] -= (c.diag_lower ? ck_lo[ia] : cu) * f;
        }
      }
    }
    if (rec && kt == 4) c.tl[3] = clock64() + (long long)(a[0][1] * 0.0 + a[1][1] * 0.0 + a[2][2] * 0.0 + a[3][3] * 0.0 + a[3][1] * 0.0 + a[0][3] * 0.0);
    if (c.own_diag) {
      if (kt < 15) { if (tx == kt + 1) c.ip_next = fast_rcp(a[KB][KB]); }
      else if (KB + 1 < T) { if (tx == 0) c.ip_next = fast_rcp(a[KB + 1 < T ? KB + 1 : KB][KB + 1 < T ? KB + 1 : KB]); }
    }
  }
}
template <int KB>
__device__ __forceinline__ void elim_all(double (&a)[T][T], ElimCtx& c) {
  if (16 * KB < c.n) elim_block<KB>(a, c);
  if constexpr (KB + 1 < T) elim_all<KB + 1>(a, c);
}
__global__ void __launch_bounds__(256, 1) k_elim(const double* K, double* out, long long* cyc, int n, int reps) {
  __shared__ __align__(16) double colbuf[2 * COLBUF];
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
    ElimCtx c;
    c.colbuf = colbuf; c.pivs = pivs; c.n = n; c.tx = tx; c.ty = ty; c.diag_lower = ty >= tx; c.own_diag = ty == tx;
    c.ip_next = 0.0; c.bad = false; c.recrep = (rep == reps - 1); for (int q = 0; q < 6; q++) c.tl[q] = 0;
    if (tid == 0) c.ip_next = fast_rcp(a[0][0]);
    elim_all<0>(a, c);
    total += clock64() - t0;
    if (rep == reps - 1 && (tid & 31) == 0) for (int q = 0; q < 6; q++) cyc[8 + (tid >> 5) * 6 + q] = c.tl[q];
    if (c.bad) acc += 1.0;
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
  double ho[256]; cudaMemcpy(ho, out, 256 * 8, cudaMemcpyDeviceToHost);
  long long hh[64]; cudaMemcpy(hh, cyc, 8 * 64, cudaMemcpyDeviceToHost);
  long long base = hh[8];
  for (int w = 0; w < 8; w++) { printf("warp %d:", w); for (int q = 0; q < 6; q++) if (q != 4) printf(" %6lld", hh[8 + w * 6 + q] - base); printf("\n"); }
  printf("elim3: %.1f cycles per elimination (n=%d) = %.1f cycles/step  [%s] chk=%.6f\n", (double)h / reps, n,
         (double)h / reps / n, cudaGetErrorString(cudaGetLastError()), ho[5]);
  return 0;
}
