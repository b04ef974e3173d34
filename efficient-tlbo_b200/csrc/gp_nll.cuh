// Device-side GP negative log marginal likelihood machinery shared by the batched MAP fit
// (gp_fit.cu) and the MCMC hyper-parameter sampler (gp_mcmc.cu): shared-memory workspace,
// kernel-matrix build, square-root-free panel elimination (shared-memory and register-resident
// variants), NLL + gradient evaluation.  See gp_fit.cu for the algorithm notes.
#pragma once
#include "tlbo_common.cuh"

#define FIT_THREADS 256
// Optional phase timing (compile with -DTLBO_PHASE_TIMING): CTA 0 accumulates clock64() deltas
// per phase of the NLL evaluation / optimiser step; read back with tlbo_debug_phase_cycles().
#ifdef TLBO_PHASE_TIMING
__device__ long long g_phase_cycles[16];
#define PHASE_T0() long long _pt = clock64()
#define PHASE_MARK(idx)                                                        \
  do {                                                                         \
    const long long _now = clock64();                                          \
    if (threadIdx.x == 0 && blockIdx.x == 0) g_phase_cycles[idx] += _now - _pt; \
    _pt = _now;                                                                \
  } while (0)
#define PHASE_LOCALS() long long _pl[4] = {0, 0, 0, 0}
#define PHASE_LMARK(q) do { const long long _now = clock64(); _pl[q] += _now - _pt; _pt = _now; } while (0)
#define PHASE_LFLUSH(base)                                                     \
  do {                                                                         \
    if (threadIdx.x == 0 && blockIdx.x == 0)                                   \
      for (int _q = 0; _q < 4; _q++) g_phase_cycles[base + _q] += _pl[_q];     \
  } while (0)
#else
#define PHASE_T0() do {} while (0)
#define PHASE_MARK(idx) do {} while (0)
#define PHASE_LOCALS() do {} while (0)
#define PHASE_LMARK(q) do {} while (0)
#define PHASE_LFLUSH(base) do {} while (0)
#endif
#define FIT_COLBUF 392   // raw / inverse-masked / lower-masked copies of the pivot column + pivot, 1/pivot

struct FitWs {
  double* A;       // [(n+1) * lda]  elimination panel
  double* Xr;      // [n * d]  raw inputs, permuted columns
  double* Xs;      // [n * d]  cont columns / l, cat columns raw
  double* y;       // [n]
  double* piv;     // [n]
  double* dinv;    // [n]  1/sqrt(piv)
  double* alpha;   // [n]
  double* red;     // [8 * 32] reduction scratch
  double* hyp;     // [2 * TLBO_HMAX] per permuted column: l (cont) or 1/(2 l^2) (cat); then 1/l^3 (cat)
  double* colbuf;  // [2 * FIT_COLBUF] double-buffered pivot-column broadcast of the register-panel elimination
  int* pairs;      // [n (n-1) / 2] packed (i << 16 | j), i > j: built once per kernel (register-panel path)
  double* Kc;      // [npairs] kernel value per pair        } written by the kernel-matrix build, reused by
  double* Cg;      // [npairs] c e^{-t-hs} 5/3 (1 + t)      } the gradient contraction (NULL: recompute)
  int* acol;       // [TLBO_HMAX] permuted column of compact column q (register-panel path: columns that vary)
  int* rmask;      // [n] activity pattern per row (inactive_pattern; read only when the space has conditions)
  int n, lda, d, npairs;
  int nac, nak;    // varying continuous / categorical columns (find_active_columns)
};

__host__ __device__ inline int fit_lda(int n) { return (n & 1) ? n : n + 1; }

// pair tables of the register-panel path: index table always (tile > 0), value caches when
// they fit beside the panel (decided by the host plan)
__host__ __device__ inline size_t fit_pair_bytes(int ncap, int tile, int cache) {
  if (tile <= 0) return 0;
  const size_t np = (size_t)ncap * (ncap - 1) / 2 + 2;
  return ((np * 4 + 15) / 16) * 16 + (cache ? 2 * np * sizeof(double) : 0);
}
__host__ inline size_t fit_small_bytes(int ncap, int d, int tile, int cache) {
  // Xr, Xs, y, piv, dinv, alpha, red, hyp, colbuf, pair tables
  return sizeof(double) * ((size_t)2 * ncap * d + 4 * (size_t)ncap + 8 * 32 + 2 * TLBO_HMAX + 2 * FIT_COLBUF + 8 + 16 +
                           (size_t)ncap / 2 + 2) +
         fit_pair_bytes(ncap, tile, cache);
}
__host__ inline size_t fit_panel_bytes(int ncap) { return sizeof(double) * (size_t)(ncap + 1) * fit_lda(ncap); }

__device__ __forceinline__ void carve_ws(FitWs& S, unsigned char* smem, double* panel_global, int n, int ncap,
                                         int d, int tile, int cache) {
  double* p = reinterpret_cast<double*>(smem);
  S.n = n; S.d = d; S.lda = fit_lda(n);
  S.Xr = p; p += (size_t)ncap * d;
  S.Xs = p; p += (size_t)ncap * d;
  S.y = p; p += ncap;
  S.piv = p; p += ncap;
  S.dinv = p; p += ncap;
  S.alpha = p; p += ncap;
  S.red = p; p += 8 * 32;
  S.hyp = p; p += 2 * TLBO_HMAX;
  S.colbuf = p; p += 2 * FIT_COLBUF;
  S.acol = reinterpret_cast<int*>(p); p += 16;
  S.rmask = reinterpret_cast<int*>(p); p += ncap / 2 + 2;
  S.nac = 0; S.nak = 0;
  S.npairs = n * (n - 1) / 2;
  S.pairs = nullptr; S.Kc = nullptr; S.Cg = nullptr;
  if (tile > 0) {
    const size_t np = (size_t)ncap * (ncap - 1) / 2 + 2;
    if (cache) { S.Kc = p; p += np; S.Cg = p; p += np; }
    S.pairs = reinterpret_cast<int*>(p);
    p += ((np * 4 + 15) / 16) * 2;
  }
  S.A = panel_global ? panel_global : p;
}

__device__ __forceinline__ void load_problem(const FitWs& S, const SpaceDesc& sp, const double* X,
                                             const double* y, int ncap) {
  const int n = S.n, d = S.d;
  for (int idx = threadIdx.x; idx < n * d; idx += blockDim.x) {
    const int i = idx / d, p = idx - i * d;
    S.Xr[idx] = X[(size_t)i * d + sp.perm[p]];
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    S.y[i] = y[i];
    unsigned m = 0u;
    if (sp.has_cond)
      for (int p = 0; p < d; p++) m |= (X[(size_t)i * d + sp.perm[p]] <= -1.0) ? (1u << p) : 0u;
    S.rmask[i] = (int)m;
  }
  if (S.pairs) {
    for (int q = threadIdx.x; q < S.npairs; q += blockDim.x) {
      int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
      while (i * (i - 1) / 2 > q) i--;
      while ((i + 1) * i / 2 <= q) i++;
      S.pairs[q] = (i << 16) | (q - i * (i - 1) / 2);
    }
  }
  (void)ncap;
}

// Columns whose value is the same in every training row (the constant hyper-parameters of a search space,
// e.g. 4 of the 9 columns of the random-forest space) contribute an exact zero to every squared distance /
// Hamming count and to their own length-scale gradient: the register-panel path works on the compacted
// list of VARYING columns (bit-identical results, 5 instead of 9 columns per pair for that space).
// Call after load_problem; contains barriers.
__device__ __forceinline__ void find_active_columns(FitWs& S, const SpaceDesc& sp) {
  const int n = S.n, d = S.d, dc = sp.d_cont, tid = threadIdx.x;
  __syncthreads();
  if (tid < d) {
    const double v0 = S.Xr[tid];
    bool varies = false;
    for (int i = 1; i < n; i++) varies = varies || (S.Xr[(size_t)i * d + tid] != v0);
    S.red[tid] = varies ? 1.0 : 0.0;
  }
  __syncthreads();
  int nac = 0, nak = 0;
  for (int p = 0; p < dc; p++)
    if (S.red[p] != 0.0) { if (tid == 0) S.acol[nac] = p; nac++; }
  for (int p = dc; p < d; p++)
    if (S.red[p] != 0.0) { if (tid == 0) S.acol[nac + nak] = p; nak++; }
  S.nac = nac; S.nak = nak;
  __syncthreads();
}

// Steps 1-2.  Returns false (uniformly) if a pivot is not positive (Cholesky would fail).
// On success: A[i][k] i>k = L_ik*sqrt(piv_k) (unscaled), piv[], dinv[], and -- scaled --
// A[i][k] i<k = (L^-1)_ki,  A[n][k] = (L^-1 y)_k.   If scale_lower, A[i][k] i>=k = L_ik.
static __device__ bool build_and_eliminate(const FitWs& S, const SpaceDesc& sp, const double* theta,
                                    bool scale_lower) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont;
  const int tid = threadIdx.x, T = blockDim.x;
  double* A = S.A;
  __syncthreads();
  if (tid < d) {
    const double l = exp(theta[1 + tid]);
    if (tid < dc) { S.hyp[tid] = l; }
    else { S.hyp[tid] = 1.0 / (2.0 * l * l); S.hyp[TLBO_HMAX + tid] = 1.0 / (l * l * l); }
  }
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  __syncthreads();
  for (int idx = tid; idx < n * d; idx += T) {
    const int p = idx % d;
    double v = S.Xr[idx];
    if (p < dc) v = v / S.hyp[p];
    S.Xs[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < (n + 1) * n; idx += T) {
    const int i = idx / n, j = idx - i * n;
    double v;
    if (i == n) v = S.y[j];
    else if (i < j) v = 0.0;
    else if (i == j) v = c + noise;
    else {
      const double* xi = S.Xs + (size_t)i * d;
      const double* xj = S.Xs + (size_t)j * d;
      double s = 0.0, hs = 0.0;
      for (int p = 0; p < dc; p++) { const double df = xi[p] - xj[p]; s += df * df; }
      for (int p = dc; p < d; p++) if (xi[p] != xj[p]) hs += S.hyp[p];
      v = matern_hamming(c, s, hs);
      if (sp.has_cond && S.rmask[i] != S.rmask[j]) v = 0.0;                       // activity mask, gp_kernels.py:21-29
    }
    A[(size_t)i * lda + j] = v;
  }
  bool ok = true;
  const int tx = tid & 15, ty = tid >> 4, TY = T >> 4;
  for (int k = 0; k < n; k++) {
    __syncthreads();
    const double piv = A[(size_t)k * lda + k];
    if (!(piv > 0.0)) { ok = false; break; }
    const double ip = 1.0 / piv;
    if (tid == 0) S.piv[k] = piv;
    for (int j = k + 1 + tx; j < n; j += 16) {
      const double fj = A[(size_t)j * lda + k] * ip;
      for (int i = ty; i <= n; i += TY) {
        double ck;
        if (i >= j || i < k) ck = A[(size_t)i * lda + k];
        else if (i == k) ck = 1.0;
        else continue;
        A[(size_t)i * lda + j] -= ck * fj;
      }
    }
  }
  __syncthreads();
  if (!ok) return false;
  for (int k = tid; k < n; k += T) S.dinv[k] = 1.0 / sqrt(S.piv[k]);
  __syncthreads();
  for (int idx = tid; idx < (n + 1) * n; idx += T) {
    const int i = idx / n, k = idx - i * n;
    if (i < k || i == n) A[(size_t)i * lda + k] *= S.dinv[k];
    else if (scale_lower) {
      if (i == k) A[(size_t)i * lda + k] = sqrt(S.piv[k]);
      else A[(size_t)i * lda + k] *= S.dinv[k];
    }
  }
  __syncthreads();
  // alpha_i = sum_{k>=i} (L^-1)_ki z_k
  for (int i = tid; i < n; i += T) {
    const double* ri = A + (size_t)i * lda;
    const double* z = A + (size_t)n * lda;
    double a = S.dinv[i] * z[i];
    for (int k = i + 1; k < n; k++) a += ri[k] * z[k];
    S.alpha[i] = a;
  }
  __syncthreads();
  return true;
}

// 1/x from the hardware reciprocal seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps:
// ~52 cycles of dependent latency against ~112 for the IEEE division (tools/ubench/lat.cu),
// result within 1 ulp -- it sits on the critical path of every elimination step.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// One block of 16 elimination steps with the pivot-column block index KB a compile-time
// constant (the outer loop over KB is unrolled by template recursion): tile indexing, most of
// the activity masks and the skipping of already-eliminated column blocks resolve statically,
// so a step is: 2-3 128-bit stores by the 16 owners of the pivot column (one warp), ONE
// __syncthreads, five 128-bit loads, a handful of selects and the tile FMAs -- no uniform
// branches and no data-dependent branch (a non-positive pivot only raises a flag).
// Developed against tools/ubench/elim*.cu: 820 -> 503 cycles per step on B200.
__device__ __forceinline__ void sts2_if(double* p, double x, double y, bool pred) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1, %2}; }" ::"r"(addr), "d"(x), "d"(y),
               "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ void sts1_if(double* p, double x, bool pred) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(addr), "d"(x), "r"((int)pred)
               : "memory");
}

struct ElimCtx {
  double* colbuf; double* pivs; int n; int tx, ty; bool diag_lower, own_diag; double ip_next; bool bad;
};

// LOWER_ONLY: only the lower part (Cholesky factor and the y row) is eliminated -- what the
// gradient-free log-likelihood of the MCMC sampler needs; the inverse-part tiles are skipped (the
// upper-part threads of diagonal tiles compute values nobody reads).
template <int T, int KB, bool LOWER_ONLY = false>
__device__ __forceinline__ void elim_block(double (&a)[T][T], ElimCtx& c) {
  constexpr int TP = (T + 1) & ~1;             // packed per-lane stride (128-bit chunks)
  const int n = c.n, tx = c.tx, ty = c.ty;
  const int kend = (n - 16 * KB) < 16 ? (n - 16 * KB) : 16;
  for (int kt = 0; kt < kend; kt++) {
    const int k = 16 * KB + kt;
    double* buf = c.colbuf + (kt & 1) * FIT_COLBUF;
    {
      // branch-free publish (divergent branches on this path cost ~100 cycles per step):
      // predicated 128-bit stores by the 16 owners of the pivot column, pivot + 1/pivot by its owner
      const bool pub = (tx == kt), own = pub && (ty == kt);
#pragma unroll
      for (int q = 0; q < TP; q += 2)
        sts2_if(buf + ty * TP + q, a[q][KB], (q + 1 < T) ? a[(q + 1 < T) ? q + 1 : q][KB] : 0.0, pub);
      sts2_if(buf + 384, a[KB][KB], c.ip_next, own);
      sts1_if(c.pivs + k, a[KB][KB], own);
    }
    const bool row_lt = ty < kt, row_eq = ty == kt, col_gt = tx > kt;
    __syncthreads();
    const double2 pp = *reinterpret_cast<const double2*>(buf + 384);
    c.bad = c.bad || !(pp.x > 0.0);
    const double ip = pp.y;
    double ck_lo[TP], fr[TP];
    {
      const double2* c2 = reinterpret_cast<const double2*>(buf + ty * TP);
      const double2* f2 = reinterpret_cast<const double2*>(buf + tx * TP);
#pragma unroll
      for (int q = 0; q < TP; q += 2) {
        const double2 v = c2[q >> 1], w = f2[q >> 1];
        ck_lo[q] = v.x; ck_lo[q + 1] = v.y; fr[q] = w.x; fr[q + 1] = w.y;
      }
    }
    // inverse-part multiplier of row block KB (rows i < k are active, the pivot row enters with 1);
    // row blocks below KB are fully active, those above not yet
    const double ck_up_kb = row_lt ? ck_lo[KB] : (row_eq ? 1.0 : 0.0);
#pragma unroll
    for (int jb = KB; jb < T; jb++) {
      double f = (tx + 16 * jb < n) ? fr[jb] * ip : 0.0;
      if (jb == KB) f = col_gt ? f : 0.0;
#pragma unroll
      for (int ia = 0; ia < T; ia++) {
        if (ia > jb) a[ia][jb] -= ck_lo[ia] * f;                        // lower part (and the y row)
        else if (ia < jb) {                                             // inverse part
          if constexpr (!LOWER_ONLY) {
            if (ia < KB) a[ia][jb] -= ck_lo[ia] * f;
            else if (ia == KB) a[ia][jb] -= ck_up_kb * f;
          }
        } else if constexpr (LOWER_ONLY) {                              // diagonal tile, lower threads matter
          a[ia][jb] -= ck_lo[ia] * f;
        } else {                                                        // diagonal tile
          const double cu = (ia < KB) ? ck_lo[ia] : ((ia == KB) ? ck_up_kb : 0.0);
          a[ia][jb] -= (c.diag_lower ? ck_lo[ia] : cu) * f;
        }
      }
    }
    {
      // speculative and branch-free: every thread inverts its own candidate for the next pivot
      // (only the owner's value is ever published)
      constexpr int KN = (KB + 1 < T) ? KB + 1 : KB;
      const double pv = (kt < 15) ? a[KB][KB] : a[KN][KN];
      c.ip_next = fast_rcp(pv);
    }
  }
}

template <int T, int KB, bool LOWER_ONLY = false>
__device__ __forceinline__ void elim_all(double (&a)[T][T], ElimCtx& c) {
  if (16 * KB < c.n) elim_block<T, KB, LOWER_ONLY>(a, c);
  if constexpr (KB + 1 < T) elim_all<T, KB + 1, LOWER_ONLY>(a, c);
}

// Symmetric sweep of the augmented matrix [[K, y], [y^T, 0]] over the n pivots of K (Goodnight's SWP
// operator = Gauss-Jordan without pivoting on an SPD matrix):  a_kk <- -1/p,  a_ik <- a_ik / p,
// a_kj <- a_kj / p,  a_ij <- a_ij - a_ik a_kj / p.  After the n steps the registers hold -K^-1, row n holds
// alpha = K^-1 y, element (n, n) holds -y^T K^-1 y, and the pivots are those of the Cholesky
// factorisation (p_k = L_kk^2: same log-determinant, same positive-definiteness test).  Against the
// L / L^-1 elimination above this needs no write-back, no back substitution for alpha and no
// K^-1 = L^-T L^-1 contraction afterwards (12 k of 71 k cycles per evaluation at n = 60), for the same
// per-step cost; measured against an extended-precision solve its alpha and gradient errors are no larger
// than the Cholesky route's (cond 1e5..5e6: 1e-13..2e-11 either way).  Thread (ty, tx) keeps the tiles ia >= jb of
// its T x T register tile (the lower triangle of the tile grid, diagonal tiles in full: 20 instead of 32 FMAs per pair
// step); the entries of a pivot column that sit in dropped tiles are published by the ROW owners (ty == kt) by symmetry.
// Launch time 1.80 -> 1.75 ms (n = 60), 1.77 -> 1.62 ms (n = 45) against full tiles.
//
// PAIRS of pivots and select-free updates (round 2).  ncu and the phase clocks show the sweep bound by the
// instruction stream of a step (~100 instructions per thread and pivot, two warps per scheduler), not by its
// latency chain, so a step now does more arithmetic per instruction of overhead:
//  * the pivots k = kt + 16 KB and k' = k + 16 (tile columns KB, KB + 1 of the SAME owner half-warp tx == kt; the
//    2 x 2 pivot block P lives in ONE thread's tile) are swept together -- a block sweep equals the two single
//    sweeps (any pivot order is a valid sweep of an SPD matrix: same -K^-1, same product of pivots, positive
//    pivots iff K is positive definite): one publish / barrier / consume and one pass over the tile
//    (a_ij -= c_i u_j + c'_i u'_j with (u, u') = P^-1 (c, c')) per TWO pivots;
//  * the special cases of the sweep operator (pivot rows <- P^-1 rows, pivot columns likewise, pivot block <- -P^-1)
//    fall out of the SAME rank-2 formula when the published pivot columns carry P - I in the pivot rows: then
//    rows k, k' come out as (u_j, u'_j), columns as (u_i, u'_i), and the pivot block as 2 I - P^-1 -- the owner
//    subtracts 2 from its two diagonal elements.  No per-element selects (the single-pivot step had 9 double selects
//    and 4 extra multiplies per thread);
//  * P^-1 (determinant, one reciprocal, three products) is formed by the ONE owner thread behind a branch instead of
//    a speculative reciprocal in all 256 threads.
// Pivots without a partner (n - 16 KB < 16 leaves tile column KB + 1 short, odd T) take the single-pivot step in the
// same select-free form (published diagonal entry p - 1, owner fix -2).  piv[] receives det P and 1 for a pair
// (only the sum of logarithms and the signs are used).
template <int T, int KB>
__device__ __forceinline__ void sweep_pair_steps(double (&a)[T][T], ElimCtx& c, int kend, int& par) {
  constexpr int TP = (T + 1) & ~1;
  constexpr int K2 = KB + 1;
  static_assert(K2 < T, "pair needs two tile columns");
  const int tx = c.tx, ty = c.ty;
#pragma unroll 1
  for (int kt = 0; kt < kend; kt++) {
    double* buf = c.colbuf + par * FIT_COLBUF;
    par ^= 1;
    const bool pub = (tx == kt), own = pub && (ty == kt);
    // only tiles ia >= jb are kept (the matrix is symmetric): rows of tile-row blocks >= the pivot's come from the
    // column owners (tx == kt), the others from the row owners (ty == kt) by symmetry
    const bool rpub = (ty == kt);
#pragma unroll
    for (int q = 0; q < T; q++) {
      if (q >= KB) sts1_if(buf + ty * TP + q, a[q][KB], pub);
      else sts1_if(buf + tx * TP + q, a[KB][q], rpub);
      if (q >= K2) sts1_if(buf + 128 + ty * TP + q, a[q][K2], pub);
      else sts1_if(buf + 128 + tx * TP + q, a[K2][q], rpub);
    }
    if (own) {                                                 // one thread of one warp
      const double p11 = a[KB][KB], p12 = a[K2][KB], p22 = a[K2][K2];
      // p11 p22 - p12^2 with the rounding error of the square recovered (the conditional variance of the second
      // pivot is a small difference of O(1) products for strongly correlated rows)
      const double sq = p12 * p12, sqe = fma(p12, p12, -sq);
      const double det = fma(p11, p22, -sq) - sqe;
      const double idet = fast_rcp(det);
      *reinterpret_cast<double2*>(buf + 384) = make_double2(p22 * idet, -(p12 * idet));
      *reinterpret_cast<double2*>(buf + 386) = make_double2(p11 * idet, (p11 > 0.0 && det > 0.0) ? 1.0 : -1.0);
      buf[ty * TP + KB] = p11 - 1.0;                           // P - I in the pivot rows
      buf[128 + ty * TP + K2] = p22 - 1.0;
      c.pivs[16 * KB + kt] = det;
      c.pivs[16 * K2 + kt] = 1.0;
    }
    __syncthreads();
    const double2 qa = *reinterpret_cast<const double2*>(buf + 384);
    const double2 qb = *reinterpret_cast<const double2*>(buf + 386);
    c.bad = c.bad || !(qb.y > 0.0);
    const double q11 = qa.x, q12 = qa.y, q22 = qb.x;
    double ck[TP], dk[TP], u[TP], v[TP];
    {
      const double2* c2 = reinterpret_cast<const double2*>(buf + ty * TP);
      const double2* d2 = reinterpret_cast<const double2*>(buf + 128 + ty * TP);
      const double2* e2 = reinterpret_cast<const double2*>(buf + tx * TP);
      const double2* g2 = reinterpret_cast<const double2*>(buf + 128 + tx * TP);
#pragma unroll
      for (int q = 0; q < TP; q += 2) {
        const double2 cc = c2[q >> 1], dd = d2[q >> 1], ee = e2[q >> 1], gg = g2[q >> 1];
        ck[q] = cc.x; ck[q + 1] = cc.y; dk[q] = dd.x; dk[q + 1] = dd.y;
        u[q] = fma(q11, ee.x, q12 * gg.x); u[q + 1] = fma(q11, ee.y, q12 * gg.y);
        v[q] = fma(q12, ee.x, q22 * gg.x); v[q + 1] = fma(q12, ee.y, q22 * gg.y);
      }
    }
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
#pragma unroll
      for (int ia = jb; ia < T; ia++) a[ia][jb] = fma(-ck[ia], u[jb], fma(-dk[ia], v[jb], a[ia][jb]));
    }
    const double two = own ? 2.0 : 0.0;                        // pivot block: 2 I - P^-1  ->  -P^-1
    a[KB][KB] -= two;
    a[K2][K2] -= two;
  }
}

template <int T, int KB>
__device__ __forceinline__ void sweep_single_steps(double (&a)[T][T], ElimCtx& c, int kt0, int kt1, int& par) {
  constexpr int TP = (T + 1) & ~1;
  const int tx = c.tx, ty = c.ty;
#pragma unroll 1
  for (int kt = kt0; kt < kt1; kt++) {
    double* buf = c.colbuf + par * FIT_COLBUF;
    par ^= 1;
    const bool pub = (tx == kt), own = pub && (ty == kt);
    const bool rpub = (ty == kt);
#pragma unroll
    for (int q = 0; q < T; q++) {
      if (q >= KB) sts1_if(buf + ty * TP + q, a[q][KB], pub);
      else sts1_if(buf + tx * TP + q, a[KB][q], rpub);
    }
    if (own) {
      const double pv = a[KB][KB];
      *reinterpret_cast<double2*>(buf + 384) = make_double2(pv, fast_rcp(pv));
      buf[ty * TP + KB] = pv - 1.0;
      c.pivs[16 * KB + kt] = pv;
    }
    __syncthreads();
    const double2 pp = *reinterpret_cast<const double2*>(buf + 384);
    c.bad = c.bad || !(pp.x > 0.0);
    const double ip = pp.y;
    double ck[TP], f[TP];
    {
      const double2* c2 = reinterpret_cast<const double2*>(buf + ty * TP);
      const double2* f2 = reinterpret_cast<const double2*>(buf + tx * TP);
#pragma unroll
      for (int q = 0; q < TP; q += 2) {
        const double2 vv = c2[q >> 1], w = f2[q >> 1];
        ck[q] = vv.x; ck[q + 1] = vv.y; f[q] = w.x * ip; f[q + 1] = w.y * ip;
      }
    }
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
#pragma unroll
      for (int ia = jb; ia < T; ia++) a[ia][jb] = fma(-ck[ia], f[jb], a[ia][jb]);
    }
    a[KB][KB] -= own ? 2.0 : 0.0;                              // 2 - 1/p  ->  -1/p
  }
}

template <int T, int KB>
__device__ __forceinline__ void sweep_blocks(double (&a)[T][T], ElimCtx& c, int& par) {
  const int n = c.n;
  const int nk = (n - 16 * KB) < 0 ? 0 : ((n - 16 * KB) < 16 ? (n - 16 * KB) : 16);        // pivots in tile column KB
  if constexpr (KB + 1 < T) {
    const int np = (n - 16 * (KB + 1)) < 0 ? 0 : ((n - 16 * (KB + 1)) < 16 ? (n - 16 * (KB + 1)) : 16);
    sweep_pair_steps<T, KB>(a, c, np, par);
    sweep_single_steps<T, KB>(a, c, np, nk, par);
    if constexpr (KB + 2 < T) sweep_blocks<T, KB + 2>(a, c, par);
  } else {
    sweep_single_steps<T, KB>(a, c, 0, nk, par);
  }
}

template <int T>
__device__ __forceinline__ void sweep_all(double (&a)[T][T], ElimCtx& c) {
  int par = 0;
  sweep_blocks<T, 0>(a, c, par);
}

// ---------------------------------------------------------------------------------------
// Register-resident variant of steps 1-2 for n + 1 <= 16 * T.  The (n+1) x n panel lives in
// registers, 2-D cyclically distributed over the 16 x 16 thread grid (thread (ty, tx) owns
// rows ty + 16a, columns tx + 16b); per elimination step only the pivot column crosses
// shared memory (double-buffered, ONE __syncthreads per column).  Element-wise the
// arithmetic is identical to build_and_eliminate.  ncu on the shared-memory version showed
// ~1.8k cycles per column of exposed LDS latency in the per-thread row loop.
// On success the shared panel holds, scaled: A[i][k] (i < k) = (L^-1)_ki, A[n][k] = (L^-1 y)_k
// and -- scale_lower: A[i][k] (i >= k) = L_ik;  else: A[k][k] = 1/L_kk, A[i][k] (i > k) = 0
// (so that row i of the panel is column i of L^-1, zero-padded: what nll_eval_reg contracts).
// Steps 0-1 of the register-panel paths: column scaling and the kernel matrix (lower triangle, diagonal,
// y row) into the shared-memory panel, pair-value caches for the gradient.  Ends with a barrier.
__device__ __forceinline__ void kernel_panel_reg(const FitWs& S, const SpaceDesc& sp, const double* theta) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont;
  const int tid = threadIdx.x, NT = blockDim.x;
  double* A = S.A;
  PHASE_T0();
  __syncthreads();
  // every thread needs the scale of "its" column (p = tid & 31 < d <= 22): one exp per thread,
  // no shared-memory round trip before the scaling pass
  const int pcol = tid & 31;
  const int nac = S.nac, da = S.nac + S.nak;       // compact (varying) columns: continuous first
  double myscale = 1.0;
  if (pcol < da) {
    const int pp = S.acol[pcol];
    const double l = exp(theta[1 + pp]);
    myscale = (pcol < nac) ? 1.0 / l : 1.0;
    if (tid < da && tid >= nac) { S.hyp[tid] = 1.0 / (2.0 * l * l); S.hyp[TLBO_HMAX + tid] = 1.0 / (l * l * l); }
    for (int i = tid >> 5; i < n; i += NT >> 5) S.Xs[(size_t)i * da + pcol] = S.Xr[(size_t)i * d + pp] * myscale;
  }
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  __syncthreads();
  PHASE_MARK(0);
  // kernel matrix (lower triangle, diagonal, y row) through shared memory: ONE copy of the
  // exp / sqrt code over the pair table, then every thread picks up its register tile.
  // FOUR pairs in flight per thread: the sqrt / exp chains are ~60 dependent FP64 operations and only two
  // warps share a scheduler, so the phase is latency-bound (n = 60: 7 pairs per thread in two passes)
  constexpr int PF = 8;
  for (int qb = tid; qb < S.npairs; qb += PF * NT) {
    int qq[PF];
    bool has[PF];
    const double* xi[PF];
    const double* xj[PF];
    double ss[PF], hh[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int q = qb + u * NT;
      has[u] = q < S.npairs;
      qq[u] = has[u] ? q : qb;
      const int ij = S.pairs[qq[u]];
      xi[u] = S.Xs + (size_t)(ij >> 16) * da;
      xj[u] = S.Xs + (size_t)(ij & 0xffff) * da;
      ss[u] = 0.0; hh[u] = 0.0;
    }
    for (int p = 0; p < nac; p++) {
#pragma unroll
      for (int u = 0; u < PF; u++) { const double e = xi[u][p] - xj[u][p]; ss[u] += e * e; }
    }
    for (int p = nac; p < da; p++) {
      const double hp = S.hyp[p];
#pragma unroll
      for (int u = 0; u < PF; u++) hh[u] += (xi[u][p] != xj[u][p]) ? hp : 0.0;
    }
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const double t = sqrt_nonneg(ss[u]) * SQRT5;
      double ce = c * exp_nonpos(-t - hh[u]);
      const int ij = S.pairs[qq[u]];
      if (sp.has_cond && S.rmask[ij >> 16] != S.rmask[ij & 0xffff]) ce = 0.0;     // activity mask, gp_kernels.py:21-29
      // t^2 / 3 as a multiplication by the rounded 1/3 (<= 1 ulp from the IEEE quotient; the division is a
      // ~25-instruction subroutine on the critical path of every pair)
      const double v = ce * (1.0 + t + t * t * (1.0 / 3.0));
      if (has[u]) {
        A[(size_t)(ij >> 16) * lda + (ij & 0xffff)] = v;
        if (S.Kc) { S.Kc[qq[u]] = v; S.Cg[qq[u]] = ce * (5.0 / 3.0) * (t + 1.0); }
      }
    }
  }
  for (int j = tid; j < n; j += NT) {
    A[(size_t)j * lda + j] = c + noise;
    A[(size_t)n * lda + j] = S.y[j];
  }
  __syncthreads();
  PHASE_MARK(1);
}

template <int T>
__device__ bool build_and_eliminate_reg(const FitWs& S, const SpaceDesc& sp, const double* theta,
                                        bool scale_lower) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont;
  const int tid = threadIdx.x, NT = blockDim.x;
  // ty = fast index: the 16 owners of a pivot column (fixed tx) sit in ONE warp, so the other
  // seven warps skip the publish code with a uniform branch
  const int tx = tid >> 4, ty = tid & 15;
  const int nac = S.nac, da = S.nac + S.nak;       // compact (varying) columns, see find_active_columns
  double* A = S.A;
  kernel_panel_reg(S, sp, theta);
  PHASE_T0();
  double a[T][T];
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    const int i = ty + 16 * ia;
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
      const int j = tx + 16 * jb;
      a[ia][jb] = (i <= n && j < n && i >= j) ? A[(size_t)i * lda + j] : 0.0;
    }
  }
  PHASE_MARK(1);
  // Tile (ia, jb) holds lower-part elements (i >= j, incl. the y row) when ia > jb, upper-part
  // (inverse) elements when ia < jb; on diagonal tiles the class is fixed per thread.  The 16
  // owners of pivot column k publish it three ways -- raw (multiplier of lower-part rows),
  // masked to the active inverse rows (i < k, 1 at i == k), masked to the rows j > k that
  // still get eliminated -- and the owner of the pivot publishes 1/pivot, so the consumers'
  // tile update is T*T unpredicated FMAs with no per-step division, select or compare.
  ElimCtx ec;
  ec.colbuf = S.colbuf; ec.pivs = S.piv; ec.n = n; ec.tx = tx; ec.ty = ty;
  ec.diag_lower = ty >= tx; ec.own_diag = ty == tx; ec.ip_next = 0.0; ec.bad = false;
  ec.ip_next = fast_rcp(a[0][0]);
  elim_all<T, 0>(a, ec);
  const bool ok = !__syncthreads_or(ec.bad ? 1 : 0);
  PHASE_MARK(2);
  if (!ok) return false;
  for (int k = tid; k < n; k += NT) S.dinv[k] = 1.0 / sqrt(S.piv[k]);
  __syncthreads();
#pragma unroll
  for (int jb = 0; jb < T; jb++) {
    const int k = tx + 16 * jb;
    if (k < n) {
      const double dk = S.dinv[k];
#pragma unroll
      for (int ia = 0; ia < T; ia++) {
        const int i = ty + 16 * ia;
        if (i <= n) {
          double v;
          if (i < k || i == n) v = a[ia][jb] * dk;
          else if (scale_lower) v = (i == k) ? sqrt(S.piv[k]) : a[ia][jb] * dk;
          else v = (i == k) ? dk : 0.0;
          A[(size_t)i * lda + k] = v;
        }
      }
    }
  }
  __syncthreads();
  // alpha_i = sum_{k>=i} (L^-1)_ki z_k
  for (int i = tid; i < n; i += NT) {
    const double* ri = A + (size_t)i * lda;
    const double* z = A + (size_t)n * lda;
    double al = S.dinv[i] * z[i];
    for (int k = i + 1; k < n; k++) al += ri[k] * z[k];
    S.alpha[i] = al;
  }
  __syncthreads();
  PHASE_MARK(3);
  return true;
}

// Gradient-free variant for the MCMC sampler (gp_mcmc.cu): kernel matrix + LOWER_ONLY elimination,
// then  sum_k [ z_k^2 / 2 + log(piv_k) / 2 ]  (z = L^-1 y) straight from the register tiles -- no
// panel write-back, no alpha, no pair-value caches.  Returns false (uniformly) on a non-positive
// pivot; *nll_data is uniform on return.
template <int T>
__device__ bool lml_eliminate_reg(const FitWs& S, const SpaceDesc& sp, const double* theta, double* nll_data) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int tx = tid >> 4, ty = tid & 15;
  double* A = S.A;
  __syncthreads();
  const int pcol = tid & 31;
  double myscale = 1.0;
  if (pcol < d) {
    const double l = exp(theta[1 + pcol]);
    myscale = (pcol < dc) ? 1.0 / l : 1.0;
    if (tid < d && tid >= dc) S.hyp[tid] = 1.0 / (2.0 * l * l);
  }
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  if (pcol < d)
    for (int i = tid >> 5; i < n; i += NT >> 5) S.Xs[(size_t)i * d + pcol] = S.Xr[(size_t)i * d + pcol] * myscale;
  __syncthreads();
  for (int q0 = tid; q0 < S.npairs; q0 += 2 * NT) {
    const int q1 = q0 + NT;
    const bool has1 = q1 < S.npairs;
    const int ij0 = S.pairs[q0], ij1 = S.pairs[has1 ? q1 : q0];
    const int i0 = ij0 >> 16, j0 = ij0 & 0xffff, i1 = ij1 >> 16, j1 = ij1 & 0xffff;
    const double* xi0 = S.Xs + (size_t)i0 * d;
    const double* xj0 = S.Xs + (size_t)j0 * d;
    const double* xi1 = S.Xs + (size_t)i1 * d;
    const double* xj1 = S.Xs + (size_t)j1 * d;
    double s0 = 0.0, s1 = 0.0, h0 = 0.0, h1 = 0.0;
    for (int p = 0; p < dc; p++) {
      const double e0 = xi0[p] - xj0[p], e1 = xi1[p] - xj1[p];
      s0 += e0 * e0; s1 += e1 * e1;
    }
    for (int p = dc; p < d; p++) {
      const double hp = S.hyp[p];
      h0 += (xi0[p] != xj0[p]) ? hp : 0.0;
      h1 += (xi1[p] != xj1[p]) ? hp : 0.0;
    }
    const double t0 = sqrt_nonneg(s0) * SQRT5, t1 = sqrt_nonneg(s1) * SQRT5;
    double v0 = c * exp_nonpos(-t0 - h0) * (1.0 + t0 + t0 * t0 / 3.0);
    double v1 = c * exp_nonpos(-t1 - h1) * (1.0 + t1 + t1 * t1 / 3.0);
    if (sp.has_cond) {
      if (S.rmask[i0] != S.rmask[j0]) v0 = 0.0;
      if (S.rmask[i1] != S.rmask[j1]) v1 = 0.0;
    }
    A[(size_t)i0 * lda + j0] = v0;
    if (has1) A[(size_t)i1 * lda + j1] = v1;
  }
  for (int j = tid; j < n; j += NT) {
    A[(size_t)j * lda + j] = c + noise;
    A[(size_t)n * lda + j] = S.y[j];
  }
  __syncthreads();
  double a[T][T];
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    const int i = ty + 16 * ia;
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
      const int j = tx + 16 * jb;
      a[ia][jb] = (i <= n && j < n && i >= j) ? A[(size_t)i * lda + j] : 0.0;
    }
  }
  ElimCtx ec;
  ec.colbuf = S.colbuf; ec.pivs = S.piv; ec.n = n; ec.tx = tx; ec.ty = ty;
  ec.diag_lower = ty >= tx; ec.own_diag = ty == tx; ec.ip_next = 0.0; ec.bad = false;
  ec.ip_next = fast_rcp(a[0][0]);
  elim_all<T, 0, true>(a, ec);
  const bool ok = !__syncthreads_or(ec.bad ? 1 : 0);      // also orders the pivots in S.piv
  if (!ok) return false;
  // the owner of (row n, column k) holds z_k sqrt(piv_k); pivots come from shared memory
  double part = 0.0;
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    if (ty + 16 * ia == n) {
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int k = tx + 16 * jb;
        if (k < n) { const double v = a[ia][jb]; part += 0.5 * (v * v) / S.piv[k]; }
      }
    }
  }
  for (int k = tid; k < n; k += NT) part += 0.5 * log(S.piv[k]);
  *nll_data = block_sum(part, S.red);
  return true;
}

// log-priors of GaussianProcess._nll and the final sign / finiteness rules (thread 0).
__device__ __forceinline__ void nll_finish(const double* theta, int H, int n, double nll_data, double* out) {
  double lml = -nll_data - 0.5 * (double)n * LOG_2PI;
  // lognormal prior on c (gp_base_prior.py:389-449), sigma = 1, mean = 0
  const double v0 = exp(theta[0]);
  double lp0, gp0;
  if (v0 <= 0.0) { lp0 = -1e25; gp0 = 0.0; }
  else {
    const double lv = log(v0);
    lp0 = -(lv * lv) / 2.0 - log(2.5066282746310002 * v0);
    gp0 = -(1.0 + lv) / v0 * v0;
  }
  // horseshoe prior on s2 (gp_base_prior.py:289-355), scale = 0.1
  const double vn = exp(theta[H - 1]);
  const double s2 = 0.1 * 0.1;
  double lpn, gpn;
  if (vn == 0.0) { lpn = INFINITY; gpn = INFINITY; }
  else {
    const double a = log(1.0 + 3.0 * (s2 / (vn * vn)));
    lpn = log(a + VERY_SMALL_NUMBER);
    double b = 3.0 * s2 + vn * vn;
    b *= log(3.0 * s2 * (1.0 / (vn * vn)) + 1.0);
    b = fmax(b, 1e-14);
    gpn = -(6.0 * s2) / b;
  }
  lml += lp0 + lpn;
  out[1] += gp0;
  out[H] += gpn;
  bool fin = isfinite(lml);
  for (int h = 0; h < H; h++) fin = fin && isfinite(out[1 + h]);
  if (!fin) {
    out[0] = 1e25;
    for (int h = 0; h < H; h++) out[1 + h] = 0.0;
  } else {
    out[0] = -lml;
    for (int h = 0; h < H; h++) out[1 + h] = -out[1 + h];
  }
}

// K^-1_ij = sum_k U[i][k] U[j][k] over the zero-padded rows of the panel (row i of U is zero
// left of column i), for the lower-triangle tiles only.  With the column block KB static, row
// blocks above KB are skipped at compile time and so are the tiles right of the diagonal.
template <int T, int KB>
__device__ __forceinline__ void kin_all(double (&kin)[T][T], const double* A, int lda, int n, int ty, int tx) {
  const int k1 = (16 * KB + 16 < n) ? 16 * KB + 16 : n;
  for (int k = 16 * KB; k < k1; k++) {
    double ui[KB + 1], uj[KB + 1];
#pragma unroll
    for (int q = 0; q <= KB; q++) {
      const int i = ty + 16 * q, j = tx + 16 * q;
      ui[q] = (i < n) ? A[(size_t)i * lda + k] : 0.0;
      uj[q] = (j < n) ? A[(size_t)j * lda + k] : 0.0;
    }
#pragma unroll
    for (int ia = 0; ia <= KB; ia++)
#pragma unroll
      for (int jb = 0; jb <= ia; jb++) kin[ia][jb] += ui[ia] * uj[jb];
  }
  if constexpr (KB + 1 < T) kin_all<T, KB + 1>(kin, A, lda, n, ty, tx);
}

// nll_eval on the register-panel path: kernel matrix, symmetric sweep (sweep_block: -K^-1, alpha, pivots in
// ONE pass over the register tiles), W = alpha alpha^T - K^-1 published once and contracted with dK/dtheta
// in one linear pair loop.
template <int T, int DMAX>
__device__ bool nll_eval_reg(const FitWs& S, const SpaceDesc& sp, const double* theta, double* out) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont, H = d + 2;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int tx = tid >> 4, ty = tid & 15;
  const int nac = S.nac, da = S.nac + S.nak;       // compact (varying) columns, see find_active_columns
  double* A = S.A;
  kernel_panel_reg(S, sp, theta);
  PHASE_T0();
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  // full symmetric tile of [[K, y], [y^T, 0]]: thread (ty, tx) owns rows ty + 16 ia, columns tx + 16 jb
  double a[T][T];
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    const int i = ty + 16 * ia;
#pragma unroll
    for (int jb = 0; jb < T; jb++) {
      const int j = tx + 16 * jb;
      const int hi = i > j ? i : j, lo = i > j ? j : i;
      a[ia][jb] = (ia >= jb && hi <= n && lo < n) ? A[(size_t)hi * lda + lo] : 0.0;      // tiles ia < jb are never used
    }
  }
  PHASE_MARK(1);
  ElimCtx ec;
  ec.colbuf = S.colbuf; ec.pivs = S.piv; ec.n = n; ec.tx = tx; ec.ty = ty;
  ec.diag_lower = ty >= tx; ec.own_diag = ty == tx; ec.bad = false;
  ec.ip_next = 0.0;
  sweep_all<T>(a, ec);
  // row n of the swept matrix is alpha, element (n, n) is -y^T K^-1 y: published before the barrier
#pragma unroll
  for (int ia = 0; ia < T; ia++) {
    if (ty + 16 * ia == n) {
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int j = tx + 16 * jb;
        if (j < n) S.alpha[j] = a[ia][jb];
        else if (j == n) S.red[32 + 27] = a[ia][jb];
      }
    }
  }
  const bool ok = !__syncthreads_or(ec.bad ? 1 : 0);
  PHASE_MARK(2);
  if (!ok) {
    if (tid == 0) out[0] = 1e25;
    if (tid < H) out[1 + tid] = 0.0;
    __syncthreads();
    return false;
  }
  double part = 0.0;
  for (int k = tid; k < n; k += NT) part += 0.5 * log(S.piv[k]);
  // While the first warps take the logarithms of the pivots, four lanes of the LAST warp (no pivots: n + 1 <= 16 T
  // < NT - 32) evaluate the log-priors of GaussianProcess._nll -- the reference's formulas (nll_finish states them for
  // one thread) with two lock-step logarithms instead of seven sequential transcendental calls; range-specialised log
  // and reciprocals (<= 2 ulp).  S.red[32 + 24 .. 26] is not touched until the finish.
  if ((tid >> 5) == (NT >> 5) - 1 && (tid & 31) < 4) {
    const int pl = tid & 31;
    const double s2 = 0.1 * 0.1, vn2 = noise * noise;
    const double ivn2 = fast_rcp(vn2);
    const double arg = (pl == 0) ? 1.0 + 3.0 * (s2 * ivn2)
                     : (pl == 1) ? 3.0 * s2 * ivn2 + 1.0
                     : (pl == 2) ? c : 2.5066282746310002 * c;
    const double lg = log_pos(arg);
    const double lg_b = __shfl_sync(0xfu, lg, 1), lv = __shfl_sync(0xfu, lg, 2), l25 = __shfl_sync(0xfu, lg, 3);
    if (pl == 0) {
      // lognormal prior on c (gp_base_prior.py:389-449), sigma = 1, mean = 0
      double lp0, gp0;
      if (c <= 0.0) { lp0 = -1e25; gp0 = 0.0; }
      else { lp0 = -(lv * lv) / 2.0 - l25; gp0 = -(1.0 + lv) * fast_rcp(c) * c; }
      // horseshoe prior on s2 (gp_base_prior.py:289-355), scale = 0.1
      double lpn, gpn;
      if (noise == 0.0) { lpn = INFINITY; gpn = INFINITY; }
      else {
        lpn = log_pos(lg + VERY_SMALL_NUMBER);
        double bq = (3.0 * s2 + vn2) * lg_b;
        bq = fmax(bq, 1e-14);
        gpn = -(6.0 * s2) * fast_rcp(bq);
      }
      S.red[32 + 24] = lp0 + lpn;
      S.red[32 + 25] = gp0;
      S.red[32 + 26] = gpn;
    }
  }
  if (tid == 0) part -= 0.5 * S.red[32 + 27];
  PHASE_MARK(3);
  const double nll_data = block_sum(part, S.red);
  PHASE_MARK(4);
  // publish W = alpha alpha^T - K^-1 (strict lower part into the panel, diagonal into piv[]: both are dead
  // by now) and contract it with dK/dtheta in ONE linear pair loop
  {
#pragma unroll
    for (int ia = 0; ia < T; ia++) {
      const int i = ty + 16 * ia;
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int j = tx + 16 * jb;
        if (i < n && j < i) A[(size_t)i * lda + j] = S.alpha[i] * S.alpha[j] + a[ia][jb];
        else if (i < n && j == i) S.piv[i] = S.alpha[i] * S.alpha[i] + a[ia][jb];
      }
    }
  }
  __syncthreads();
  PHASE_MARK(6);
  double g0 = 0.0, gn = 0.0;
  double gacc[DMAX];
#pragma unroll
  for (int p = 0; p < DMAX; p++) gacc[p] = 0.0;
  constexpr int PF = 4;                       // pairs in flight per thread (as in kernel_panel_reg)
  for (int qb = tid; qb < S.npairs; qb += PF * NT) {
    const double* xi[PF];
    const double* xj[PF];
    double WK[PF], cf[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int q = qb + u * NT;
      const bool has = q < S.npairs;
      const int qq = has ? q : qb;
      const int ij = S.pairs[qq];
      const int i = ij >> 16, j = ij & 0xffff;
      const double W = has ? A[(size_t)i * lda + j] : 0.0;          // a masked pair adds zeros
      xi[u] = S.Xs + (size_t)i * da;
      xj[u] = S.Xs + (size_t)j * da;
      if (S.Kc) {
        WK[u] = W * S.Kc[qq];
        cf[u] = W * S.Cg[qq];
      } else {
        double s = 0.0, h = 0.0;
        for (int p = 0; p < nac; p++) { const double e = xi[u][p] - xj[u][p]; s += e * e; }
        for (int p = nac; p < da; p++) h += (xi[u][p] != xj[u][p]) ? S.hyp[p] : 0.0;
        const double t = sqrt_nonneg(s) * SQRT5;
        double ce = c * exp_nonpos(-t - h);
        if (sp.has_cond && S.rmask[i] != S.rmask[j]) ce = 0.0;
        WK[u] = W * (ce * (1.0 + t + t * t * (1.0 / 3.0)));
        cf[u] = W * (ce * (5.0 / 3.0) * (t + 1.0));
      }
    }
#pragma unroll
    for (int u = 0; u < PF; u++) g0 += WK[u];
#pragma unroll
    for (int p = 0; p < DMAX; p++) {
      if (p < nac) {
#pragma unroll
        for (int u = 0; u < PF; u++) { const double e = xi[u][p] - xj[u][p]; gacc[p] += cf[u] * (e * e); }
      } else if (p < da) {
        const double h3 = S.hyp[TLBO_HMAX + p];
#pragma unroll
        for (int u = 0; u < PF; u++) gacc[p] += (xi[u][p] != xj[u][p]) ? WK[u] * h3 : 0.0;
      }
    }
  }
  for (int i = tid; i < n; i += NT) {
    const double W = S.piv[i];
    g0 += 0.5 * W * c;
    gn += 0.5 * W * noise;
  }
  PHASE_MARK(7);
  const int lane = tid & 31, wid = tid >> 5, nw = NT >> 5;
  {
    // g0, gn and the da varying columns' sums in COMPACT order through transposed warp reductions (16 values in 16
    // exchanges; the per-value butterflies were 5 exchanges each behind a branch per column): lanes r < 16 hold the
    // warp total of value r.  S.red was last read before the barrier in front of the W publish: no barrier here
    double vals[16];
    vals[0] = g0; vals[1] = gn;
#pragma unroll
    for (int p = 0; p < 14; p++) vals[2 + p] = (p < DMAX) ? gacc[p < DMAX ? p : 0] : 0.0;
    const double tot = warp_transpose_sum16(vals);
    if (lane < 16) S.red[wid * 32 + lane] = tot;
    if constexpr (DMAX > 14) {
#pragma unroll
      for (int p = 0; p < 16; p++) vals[p] = (14 + p < DMAX) ? gacc[14 + p < DMAX ? 14 + p : 0] : 0.0;
      const double tot2 = warp_transpose_sum16(vals);
      if (lane < 8) S.red[wid * 32 + 16 + lane] = tot2;       // da <= d <= 22: compact slots 16 .. 23
    }
  }
  __syncthreads();
  // warp 0 sums the per-warp partials and applies the final rules
  double gsum = 0.0;
  if (wid == 0) {
    if (lane < H) {
      // theta_h  ->  compact slot: amplitude 0, noise 1, varying column q at 2 + q; constant columns: zero gradient
      int r = (lane == 0) ? 0 : ((lane == H - 1) ? 1 : -1);
      for (int q = 0; q < da; q++)
        if (S.acol[q] == lane - 1 && lane < H - 1) r = 2 + q;
      if (r >= 0)
        for (int w = 0; w < nw; w++) gsum += S.red[w * 32 + r];       // d LML / d theta_h (without priors)
    }
    // final sign / finiteness rules of GaussianProcess._nll (gp.py:183-197); the log-priors were computed by warp 1
    // at the start of the evaluation
    const double lml = -nll_data - 0.5 * (double)n * LOG_2PI + S.red[32 + 24];
    if (lane == 0) gsum += S.red[32 + 25];
    if (lane == H - 1) gsum += S.red[32 + 26];
    const bool fin = __all_sync(0xffffffffu, isfinite(lml) && (lane >= H || isfinite(gsum)));
    if (lane == 0) out[0] = fin ? -lml : 1e25;
    if (lane < H) out[1 + lane] = fin ? -gsum : 0.0;
  }
  __syncthreads();
  PHASE_MARK(8);
  return true;
}

// One evaluation of GaussianProcess._nll (gp.py:164-197).  Result: out[0] = f, out[1..H] = g
// (shared memory), visible to all threads on return.  Returns Cholesky success.
template <int DMAX>
__device__ bool nll_eval(const FitWs& S, const SpaceDesc& sp, const double* theta, double* out) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont, H = d + 2;
  const int tid = threadIdx.x, T = blockDim.x;
  const bool ok = build_and_eliminate(S, sp, theta, false);
  if (!ok) {
    if (tid == 0) out[0] = 1e25;
    if (tid < H) out[1 + tid] = 0.0;
    __syncthreads();
    return false;
  }
  const double* A = S.A;
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  // data fit and log-determinant
  double part = 0.0;
  for (int k = tid; k < n; k += T) {
    const double z = A[(size_t)n * lda + k];
    part += 0.5 * z * z + 0.5 * log(S.piv[k]);
  }
  const double nll_data = block_sum(part, S.red);

  // gradient accumulators: g0 (log c), gacc[p] (length scales, permuted order), gn (log s2)
  double g0 = 0.0, gn = 0.0;
  double gacc[DMAX];
#pragma unroll
  for (int p = 0; p < DMAX; p++) gacc[p] = 0.0;

  const int npairs = n * (n - 1) / 2;
  for (int q = tid; q < npairs; q += T) {
    int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)q)) * 0.5);
    while (i * (i - 1) / 2 > q) i--;
    while ((i + 1) * i / 2 <= q) i++;
    const int j = q - i * (i - 1) / 2;   // j < i
    const double* ri = A + (size_t)i * lda;
    const double* rj = A + (size_t)j * lda;
    double kin = S.dinv[i] * rj[i];
    for (int k = i + 1; k < n; k++) kin += ri[k] * rj[k];
    const double W = S.alpha[i] * S.alpha[j] - kin;
    const double* xi = S.Xs + (size_t)i * d;
    const double* xj = S.Xs + (size_t)j * d;
    double s = 0.0, hs = 0.0;
    for (int p = 0; p < dc; p++) { const double df = xi[p] - xj[p]; s += df * df; }
    for (int p = dc; p < d; p++) if (xi[p] != xj[p]) hs += S.hyp[p];
    const double t = sqrt(s) * SQRT5;
    const double e = exp(-t);
    const double eh = (sp.has_cond && S.rmask[i] != S.rmask[j]) ? 0.0 : exp(-hs);   // activity mask (K and dK)
    const double Kc = c * ((1.0 + t + t * t / 3.0) * e) * eh;
    const double WK = W * Kc;
    g0 += WK;
    const double coef = W * c * eh * (5.0 / 3.0) * (t + 1.0) * e;
#pragma unroll
    for (int p = 0; p < DMAX; p++) {
      if (p < dc) { const double df = xi[p] - xj[p]; gacc[p] += coef * (df * df); }
      else if (p < d) { if (xi[p] != xj[p]) gacc[p] += WK * S.hyp[TLBO_HMAX + p]; }
    }
  }
  for (int i = tid; i < n; i += T) {
    const double* ri = A + (size_t)i * lda;
    double kin = S.dinv[i] * S.dinv[i];
    for (int k = i + 1; k < n; k++) kin += ri[k] * ri[k];
    const double W = S.alpha[i] * S.alpha[i] - kin;
    g0 += 0.5 * W * c;
    gn += 0.5 * W * noise;
  }
  // reduce H values across the block
  const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  __syncthreads();
  {
    double v = warp_sum(g0);
    if (lane == 0) S.red[wid * 32 + 0] = v;
    v = warp_sum(gn);
    if (lane == 0) S.red[wid * 32 + (d + 1)] = v;
#pragma unroll
    for (int p = 0; p < DMAX; p++) {
      if (p < d) {
        v = warp_sum(gacc[p]);
        if (lane == 0) S.red[wid * 32 + 1 + p] = v;
      }
    }
  }
  __syncthreads();
  if (tid < H) {
    double v = 0.0;
    for (int w = 0; w < nw; w++) v += S.red[w * 32 + tid];
    out[1 + tid] = v;       // d LML / d theta_h (without priors)
  }
  __syncthreads();
  if (tid == 0) nll_finish(theta, H, n, nll_data, out);
  __syncthreads();
  return true;
}

// T = 0: shared-memory panel (any n that fits); T > 0: register panel, n + 1 <= 16 T.
template <int T>
__device__ __forceinline__ bool build_and_eliminate_any(const FitWs& S, const SpaceDesc& sp, const double* theta,
                                                        bool scale_lower) {
  if constexpr (T == 0) return build_and_eliminate(S, sp, theta, scale_lower);
  else return build_and_eliminate_reg<T>(S, sp, theta, scale_lower);
}
template <int T, int DMAX>
__device__ __forceinline__ bool nll_eval_any(const FitWs& S, const SpaceDesc& sp, const double* theta, double* out) {
  if constexpr (T == 0) return nll_eval<DMAX>(S, sp, theta, out);
  else return nll_eval_reg<T, DMAX>(S, sp, theta, out);
}
// Register-panel tile for a batch of row capacity ncap: the panel holds n + 1 rows (the y row rides along),
// n + 1 <= 16 T.  The tile follows the CAPACITY (T = ceil(ncap / 16)), so the caller's capacity must leave
// room for the y row when it is a multiple of 16 (n <= ncap - 1; tlbo_b200/ops.py:batch_cap; the kernels
// reject a violating surrogate through its status instead of computing garbage).
__host__ inline int panel_tile_for(int ncap) {
  if (ncap <= 64) return 4;
  if (ncap <= 80) return 5;
  if (ncap <= 96) return 6;
  if (ncap <= 112) return 7;
  if (ncap <= 128) return 8;
  return 0;
}
