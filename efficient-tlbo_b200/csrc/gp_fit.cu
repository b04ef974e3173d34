// Batched GP fit on device: one CTA per (surrogate, restart) runs the complete
// bound-constrained L-BFGS-B MAP optimisation of the GP hyper-parameters with zero host
// round trips; a second kernel picks the best restart and factorises.
//
// Replaces reference tlbo/model/gp.py:98-248 (GaussianProcess._train/_nll/_optimize) and
// the sklearn-0.21 / skopt GaussianProcessRegressor.fit + log_marginal_likelihood under it
// (kernel stack: tlbo/model/gp_kernels.py, priors: tlbo/model/gp_base_prior.py).
//
// Per NLL evaluation, inside one CTA with the whole problem in shared memory:
//   1. K (lower triangle) from theta: Matern-5/2 x Hamming x c + s2*I.
//   2. ONE square-root-free elimination sweep over the (n+1) x n panel [K ; I ; y^T]
//      (one __syncthreads per column) that leaves, in a single n x n array, the pivots of
//      L, the columns of L below the diagonal, L^-T above it and L^-1 y in the extra row.
//   3. alpha = L^-T (L^-1 y);  LML = -1/2 |L^-1 y|^2 - 1/2 sum log piv - n/2 log 2pi.
//   4. gradient: 1/2 tr((alpha alpha^T - K^-1) dK/dtheta_h) with K^-1_ij formed on the fly
//      from rows of L^-T and dK/dtheta recomputed per pair -- never materialised (n^2 H).
//   5. + log-priors (lognormal on c, horseshoe on s2) and their gradients.
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
#include "lbfgsb_dense.cuh"
#include "lbfgsb_warp.cuh"
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
  int n, lda, d, npairs;
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
  return sizeof(double) * ((size_t)2 * ncap * d + 4 * (size_t)ncap + 8 * 32 + 2 * TLBO_HMAX + 2 * FIT_COLBUF + 8) +
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
  for (int i = threadIdx.x; i < n; i += blockDim.x) S.y[i] = y[i];
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

// Steps 1-2.  Returns false (uniformly) if a pivot is not positive (Cholesky would fail).
// On success: A[i][k] i>k = L_ik*sqrt(piv_k) (unscaled), piv[], dinv[], and -- scaled --
// A[i][k] i<k = (L^-1)_ki,  A[n][k] = (L^-1 y)_k.   If scale_lower, A[i][k] i>=k = L_ik.
__device__ bool build_and_eliminate(const FitWs& S, const SpaceDesc& sp, const double* theta,
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

template <int T, int KB>
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
          if (ia < KB) a[ia][jb] -= ck_lo[ia] * f;
          else if (ia == KB) a[ia][jb] -= ck_up_kb * f;
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

template <int T, int KB>
__device__ __forceinline__ void elim_all(double (&a)[T][T], ElimCtx& c) {
  if (16 * KB < c.n) elim_block<T, KB>(a, c);
  if constexpr (KB + 1 < T) elim_all<T, KB + 1>(a, c);
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
template <int T>
__device__ bool build_and_eliminate_reg(const FitWs& S, const SpaceDesc& sp, const double* theta,
                                        bool scale_lower) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont;
  const int tid = threadIdx.x, NT = blockDim.x;
  // ty = fast index: the 16 owners of a pivot column (fixed tx) sit in ONE warp, so the other
  // seven warps skip the publish code with a uniform branch
  const int tx = tid >> 4, ty = tid & 15;
  double* A = S.A;
  PHASE_T0();
  __syncthreads();
  // every thread needs the scale of "its" column (p = tid & 31 < d <= 22): one exp per thread,
  // no shared-memory round trip before the scaling pass
  const int pcol = tid & 31;
  double myscale = 1.0;
  if (pcol < d) {
    const double l = exp(theta[1 + pcol]);
    myscale = (pcol < dc) ? 1.0 / l : 1.0;
    if (tid < d) {
      if (tid < dc) { S.hyp[tid] = l; }
      else { S.hyp[tid] = 1.0 / (2.0 * l * l); S.hyp[TLBO_HMAX + tid] = 1.0 / (l * l * l); }
    }
  }
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  if (pcol < d)
    for (int i = tid >> 5; i < n; i += NT >> 5) S.Xs[(size_t)i * d + pcol] = S.Xr[(size_t)i * d + pcol] * myscale;
  __syncthreads();
  PHASE_MARK(0);
  // kernel matrix (lower triangle, diagonal, y row) through shared memory: ONE copy of the
  // exp / sqrt code over the pair table, then every thread picks up its register tile
  // two pairs in flight per thread (independent sqrt / exp chains)
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
    const double ce0 = c * exp_nonpos(-t0 - h0), ce1 = c * exp_nonpos(-t1 - h1);
    const double v0 = ce0 * (1.0 + t0 + t0 * t0 / 3.0), v1 = ce1 * (1.0 + t1 + t1 * t1 / 3.0);
    A[(size_t)i0 * lda + j0] = v0;
    if (S.Kc) { S.Kc[q0] = v0; S.Cg[q0] = ce0 * (5.0 / 3.0) * (t0 + 1.0); }
    if (has1) {
      A[(size_t)i1 * lda + j1] = v1;
      if (S.Kc) { S.Kc[q1] = v1; S.Cg[q1] = ce1 * (5.0 / 3.0) * (t1 + 1.0); }
    }
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

// nll_eval on the register-panel path: the gradient contraction uses the same 2-D cyclic
// tiling -- K^-1_ij = sum_k U[i][k] U[j][k] over the zero-padded rows of the panel, T x T
// outer products per k (2T shared loads per T^2 FMAs instead of 2 per FMA).
template <int T, int DMAX>
__device__ bool nll_eval_reg(const FitWs& S, const SpaceDesc& sp, const double* theta, double* out) {
  const int n = S.n, lda = S.lda, d = S.d, dc = sp.d_cont, H = d + 2;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int tx = tid >> 4, ty = tid & 15;
  const bool ok = build_and_eliminate_reg<T>(S, sp, theta, false);
  if (!ok) {
    if (tid == 0) out[0] = 1e25;
    if (tid < H) out[1 + tid] = 0.0;
    __syncthreads();
    return false;
  }
  const double* A = S.A;
  const double c = exp(theta[0]);
  const double noise = exp(theta[d + 1]);
  double part = 0.0;
  for (int k = tid; k < n; k += NT) {
    const double z = A[(size_t)n * lda + k];
    part += 0.5 * z * z + 0.5 * log(S.piv[k]);
  }
  PHASE_T0();
  const double nll_data = block_sum(part, S.red);
  PHASE_MARK(4);

  double kin[T][T];
#pragma unroll
  for (int ia = 0; ia < T; ia++)
#pragma unroll
    for (int jb = 0; jb < T; jb++) kin[ia][jb] = 0.0;
  kin_all<T, 0>(kin, A, lda, n, ty, tx);
  // publish W = alpha alpha^T - K^-1 (strict lower part into the panel, diagonal into piv[]:
  // both are dead by now) and contract it with dK/dtheta in ONE linear pair loop
  __syncthreads();
  PHASE_MARK(5);
  {
    double* Aw = S.A;
#pragma unroll
    for (int ia = 0; ia < T; ia++) {
      const int i = ty + 16 * ia;
#pragma unroll
      for (int jb = 0; jb < T; jb++) {
        const int j = tx + 16 * jb;
        if (i < n && j < i) Aw[(size_t)i * lda + j] = S.alpha[i] * S.alpha[j] - kin[ia][jb];
        else if (i < n && j == i) S.piv[i] = S.alpha[i] * S.alpha[i] - kin[ia][jb];
      }
    }
  }
  __syncthreads();
  PHASE_MARK(6);
  double g0 = 0.0, gn = 0.0;
  double gacc[DMAX];
#pragma unroll
  for (int p = 0; p < DMAX; p++) gacc[p] = 0.0;
  for (int q0 = tid; q0 < S.npairs; q0 += 2 * NT) {
    const int q1 = q0 + NT;
    const bool has1 = q1 < S.npairs;
    const int ij0 = S.pairs[q0], ij1 = S.pairs[has1 ? q1 : q0];
    const int i0 = ij0 >> 16, j0 = ij0 & 0xffff, i1 = ij1 >> 16, j1 = ij1 & 0xffff;
    const double W0 = A[(size_t)i0 * lda + j0];
    const double W1 = has1 ? A[(size_t)i1 * lda + j1] : 0.0;        // a masked second pair adds zeros
    const double* xi0 = S.Xs + (size_t)i0 * d;
    const double* xj0 = S.Xs + (size_t)j0 * d;
    const double* xi1 = S.Xs + (size_t)i1 * d;
    const double* xj1 = S.Xs + (size_t)j1 * d;
    double WK0, WK1, cf0, cf1;
    if (S.Kc) {
      WK0 = W0 * S.Kc[q0]; cf0 = W0 * S.Cg[q0];
      WK1 = W1 * S.Kc[has1 ? q1 : q0]; cf1 = W1 * S.Cg[has1 ? q1 : q0];
    } else {
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
      const double ce0 = c * exp_nonpos(-t0 - h0), ce1 = c * exp_nonpos(-t1 - h1);
      WK0 = W0 * (ce0 * (1.0 + t0 + t0 * t0 / 3.0)); cf0 = W0 * (ce0 * (5.0 / 3.0) * (t0 + 1.0));
      WK1 = W1 * (ce1 * (1.0 + t1 + t1 * t1 / 3.0)); cf1 = W1 * (ce1 * (5.0 / 3.0) * (t1 + 1.0));
    }
    g0 += WK0;
    g0 += WK1;
#pragma unroll
    for (int p = 0; p < DMAX; p++) {
      if (p < dc) {
        const double e0 = xi0[p] - xj0[p], e1 = xi1[p] - xj1[p];
        gacc[p] += cf0 * (e0 * e0);
        gacc[p] += cf1 * (e1 * e1);
      } else if (p < d) {
        const double h3 = S.hyp[TLBO_HMAX + p];
        gacc[p] += (xi0[p] != xj0[p]) ? WK0 * h3 : 0.0;
        gacc[p] += (xi1[p] != xj1[p]) ? WK1 * h3 : 0.0;
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
  __syncthreads();
  {
    // one interleaved butterfly over all H slots (independent shuffles pipeline)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g0 += __shfl_xor_sync(0xffffffffu, g0, o);
      gn += __shfl_xor_sync(0xffffffffu, gn, o);
#pragma unroll
      for (int p = 0; p < DMAX; p++) gacc[p] += __shfl_xor_sync(0xffffffffu, gacc[p], o);
    }
    if (lane == 0) {
      S.red[wid * 32 + 0] = g0;
      S.red[wid * 32 + (d + 1)] = gn;
#pragma unroll
      for (int p = 0; p < DMAX; p++)
        if (p < d) S.red[wid * 32 + 1 + p] = gacc[p];
    }
  }
  __syncthreads();
  if (tid < H) {
    double v = 0.0;
    for (int w = 0; w < nw; w++) v += S.red[w * 32 + tid];
    out[1 + tid] = v;
  }
  __syncthreads();
  if (tid == 0) nll_finish(theta, H, n, nll_data, out);
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
    const double e = exp(-t), eh = exp(-hs);
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
__host__ inline int panel_tile_for(int ncap) {
  if (ncap + 1 <= 64) return 4;
  if (ncap + 1 <= 80) return 5;
  if (ncap + 1 <= 96) return 6;
  if (ncap + 1 <= 112) return 7;
  if (ncap + 1 <= 128) return 8;
  return 0;
}

// ---------------------------------------------------------------------------------------
struct FitArgs {
  const double* X; const double* y; const int32_t* n;
  int B, R, ncap, d;
  SpaceDesc sp;
  const double* p0; const double* lo; const double* hi;
  int do_optimize;
  double* restart_theta; double* restart_f; int32_t* restart_info;
  double* panels;        // global panels (NULL -> shared memory)
  size_t panel_stride;   // doubles
  int cache;             // pair-value caches present in the shared-memory plan
};

// Optimiser state size in shared memory (either variant).
__host__ __device__ inline size_t fit_state_bytes() {
  const size_t a = sizeof(LbfgsbState), b = sizeof(LbwState);
  return (((a > b ? a : b) + 15) / 16) * 16;
}

// WARP_OPT = true: warp-cooperative optimiser (lbfgsb_warp.cuh, default);
//            false: single-thread reference statement of the same optimiser (lbfgsb_dense.cuh).
template <int DMAX, bool WARP_OPT, int T>
__global__ void __launch_bounds__(FIT_THREADS, 1) gp_fit_kernel(FitArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x / a.R, r = blockIdx.x - b * a.R;
  const int n = a.n[b], d = a.d, H = d + 2;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ double th[TLBO_HMAX];
  __shared__ double fg[1 + TLBO_HMAX];
  __shared__ int act;
  LbfgsbState* st = reinterpret_cast<LbfgsbState*>(smem);
  LbwState* sw = reinterpret_cast<LbwState*>(smem);
  unsigned char* rest = smem + fit_state_bytes();
  FitWs S;
  carve_ws(S, rest, a.panels ? a.panels + (size_t)blockIdx.x * a.panel_stride : nullptr, n, a.ncap, d, T, a.cache);
  load_problem(S, a.sp, a.X + (size_t)b * a.ncap * d, a.y + (size_t)b * a.ncap, a.ncap);
  if (tid < H) th[tid] = a.p0[(size_t)r * H + tid];
  __syncthreads();
  int bumps = 0;
  if (r == 0) {
    // gp.py:127-140: the initial GaussianProcessRegressor.fit at the default theta retries
    // with noise += 1 on LinAlgError; the bumped theta then becomes start point 0.
    for (int tries = 0; tries < 10; tries++) {
      if (build_and_eliminate_any<T>(S, a.sp, th, false)) break;
      if (tid == 0) th[H - 1] = log(exp(th[H - 1]) + 1.0);
      bumps++;
      __syncthreads();
    }
  }
  int nfev = 0;
  if (!a.do_optimize) {
    nll_eval_any<T, DMAX>(S, a.sp, th, fg);
    nfev = 1;
    if (tid < H) a.restart_theta[((size_t)b * a.R + r) * H + tid] = th[tid];
    if (tid == 0) {
      a.restart_f[(size_t)b * a.R + r] = fg[0];
      int32_t* info = a.restart_info + ((size_t)b * a.R + r) * 4;
      info[0] = nfev; info[1] = 0; info[2] = -1; info[3] = bumps;
    }
    return;
  }
  if (WARP_OPT) { if (wid == 0) lbw_init(*sw, H, th, a.lo, a.hi, lane); }
  else { if (tid == 0) lbfgsb_init(*st, H, th, a.lo, a.hi); }
  __syncthreads();
  for (;;) {
    if (tid < H) th[tid] = WARP_OPT ? sw->x[tid] : st->x[tid];
    __syncthreads();
    nll_eval_any<T, DMAX>(S, a.sp, th, fg);
    nfev++;
    PHASE_T0();
    if (WARP_OPT) {
      if (wid == 0) {
        const int rc = lbw_feed(*sw, fg[0], fg + 1, lane);
        if (lane == 0) act = rc;
      }
    } else {
      if (tid == 0) act = lbfgsb_feed(*st, fg[0], fg + 1);
    }
    __syncthreads();
    PHASE_MARK(9);
#ifdef TLBO_PHASE_TIMING
    if (tid == 0 && blockIdx.x == 0) g_phase_cycles[15] += 1;
#endif
    if (act == LB_DONE) break;
  }
  if (tid < H) a.restart_theta[((size_t)b * a.R + r) * H + tid] = WARP_OPT ? sw->x[tid] : st->x[tid];
  if (tid == 0) {
    a.restart_f[(size_t)b * a.R + r] = WARP_OPT ? sw->f_last : st->f_last;
    int32_t* info = a.restart_info + ((size_t)b * a.R + r) * 4;
    info[0] = nfev; info[1] = WARP_OPT ? sw->iter : st->iter; info[2] = WARP_OPT ? sw->status : st->status;
    info[3] = bumps;
  }
}

// ---------------------------------------------------------------------------------------
struct FacArgs {
  const double* X; const double* y; const int32_t* n;
  int B, R, ncap, d;
  SpaceDesc sp;
  const double* theta_in;     // [B, R, H] candidates (R = 1 -> given theta)
  const double* f_in;         // [B, R] or NULL
  const double* ymean; const double* ystd;   // device copies [B]
  tlbo_pack pack; int slot0;
  double* L; double* Kinv; int32_t* status;
  double* panels; size_t panel_stride;
};

__global__ void __launch_bounds__(FIT_THREADS, 1) gp_factorize_kernel(FacArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x;
  const int n = a.n[b], d = a.d, H = d + 2, dc = a.sp.d_cont, ncap = a.ncap;
  const int tid = threadIdx.x, T = blockDim.x;
  __shared__ double th[TLBO_HMAX];
  __shared__ int best;
  FitWs S;
  carve_ws(S, smem, a.panels ? a.panels + (size_t)b * a.panel_stride : nullptr, n, ncap, d, 0, 0);
  load_problem(S, a.sp, a.X + (size_t)b * ncap * d, a.y + (size_t)b * ncap, ncap);
  if (tid == 0) {
    // gp.py:241-248: first strict improvement wins
    int br = -1;
    if (a.f_in) {
      double fstar = INFINITY;
      for (int r = 0; r < a.R; r++) {
        const double f = a.f_in[(size_t)b * a.R + r];
        if (f < fstar) { fstar = f; br = r; }
      }
    } else br = 0;
    best = br;
  }
  __syncthreads();
  const int slot = a.slot0 + b;
  const int pncap = a.pack.ncap;
  if (best < 0) {
    if (tid == 0) { a.status[b] = 2; a.pack.d_n[slot] = 0; }
    return;
  }
  if (tid < H) th[tid] = a.theta_in[((size_t)b * a.R + best) * H + tid];
  __syncthreads();
  const bool ok = build_and_eliminate(S, a.sp, th, a.L != nullptr);
  if (tid == 0) { a.status[b] = ok ? 0 : 1; a.pack.d_n[slot] = ok ? n : 0; }
  if (tid < H) a.pack.d_theta[(size_t)slot * H + tid] = th[tid];
  if (!ok) return;
  const int lda = S.lda;
  const double* A = S.A;
  // hyp row
  double* hyp = a.pack.d_hyp + (size_t)slot * TLBO_HYP_STRIDE;
  if (tid == 0) {
    const double c = exp(th[0]), noise = exp(th[H - 1]);
    hyp[0] = c; hyp[1] = noise; hyp[2] = a.ymean[b]; hyp[3] = a.ystd[b]; hyp[4] = c + noise;
    hyp[5] = 0.0; hyp[6] = 0.0; hyp[7] = 0.0;
  }
  if (tid < d) hyp[8 + tid] = (tid < dc) ? 1.0 / S.hyp[tid] : S.hyp[tid];
  // Xs (multiplied by 1/l: the predict kernels scale candidates the same way)
  double* Xs = a.pack.d_Xs + (size_t)slot * pncap * d;
  for (int idx = tid; idx < pncap * d; idx += T) {
    const int i = idx / d, p = idx - i * d;
    double v = 0.0;
    if (i < n) { v = S.Xr[(size_t)i * d + p]; if (p < dc) v = v * (1.0 / S.hyp[p]); }
    Xs[idx] = v;
  }
  double* al = a.pack.d_alpha + (size_t)slot * pncap;
  for (int i = tid; i < pncap; i += T) al[i] = (i < n) ? S.alpha[i] : 0.0;
  const int ldl = TLBO_LDL(pncap);
  double* Li = a.pack.d_Linv + (size_t)slot * pncap * ldl;
  for (int idx = tid; idx < pncap * ldl; idx += T) {
    const int k = idx / ldl, i = idx - k * ldl;    // Linv[k][i]
    double v = 0.0;
    if (k < n && i < n) {
      if (i < k) v = A[(size_t)i * lda + k];
      else if (i == k) v = S.dinv[k];
    }
    Li[idx] = v;
  }
  if (a.L) {
    double* Lo = a.L + (size_t)b * ncap * ncap;
    for (int idx = tid; idx < ncap * ncap; idx += T) {
      const int i = idx / ncap, k = idx - i * ncap;
      Lo[idx] = (i < n && k <= i) ? A[(size_t)i * lda + k] : 0.0;
    }
  }
  if (a.Kinv) {
    double* Ko = a.Kinv + (size_t)b * ncap * ncap;
    for (int idx = tid; idx < ncap * ncap; idx += T) {
      const int i = idx / ncap, j = idx - i * ncap;
      double v = 0.0;
      if (i < n && j < n) {
        const int hi = i > j ? i : j, lo = i > j ? j : i;
        const double* rh = A + (size_t)hi * lda;
        const double* rl = A + (size_t)lo * lda;
        v = S.dinv[hi] * (hi == lo ? S.dinv[hi] : rl[hi]);
        for (int k = hi + 1; k < n; k++) v += rh[k] * rl[k];
      }
      Ko[idx] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
struct NllArgs {
  const double* X; const double* y; const int32_t* n;
  int B, ncap, d;
  SpaceDesc sp;
  const double* theta; double* nll; double* grad; int32_t* status;
  double* panels; size_t panel_stride;
  int cache;
};

template <int DMAX, int T>
__global__ void __launch_bounds__(FIT_THREADS, 1) gp_nll_grad_kernel(NllArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x;
  const int n = a.n[b], d = a.d, H = d + 2;
  const int tid = threadIdx.x;
  __shared__ double th[TLBO_HMAX];
  __shared__ double fg[1 + TLBO_HMAX];
  FitWs S;
  carve_ws(S, smem, a.panels ? a.panels + (size_t)b * a.panel_stride : nullptr, n, a.ncap, d, T, a.cache);
  load_problem(S, a.sp, a.X + (size_t)b * a.ncap * d, a.y + (size_t)b * a.ncap, a.ncap);
  if (tid < H) th[tid] = a.theta[(size_t)b * H + tid];
  __syncthreads();
  const bool ok = nll_eval_any<T, DMAX>(S, a.sp, th, fg);
  if (tid == 0) { a.nll[b] = fg[0]; a.status[b] = ok ? 0 : 1; }
  if (tid < H) a.grad[(size_t)b * H + tid] = fg[1 + tid];
}

// ---------------------------------------------------------------------------------------
// host side
static const size_t kSmemLimit = 227 * 1024;
static int g_fit_variant = 1;   // 1 = warp-cooperative optimiser + register panel, 0 = single-thread statement

#ifdef TLBO_PHASE_TIMING
extern "C" int tlbo_debug_phase_cycles(long long* h_out, int reset) {
  cudaMemcpyFromSymbol(h_out, g_phase_cycles, sizeof(long long) * 16);
  cudaMemcpyFromSymbol(h_out + 16, g_feed_cycles, sizeof(long long) * 8);
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
    cudaMemcpyToSymbol(g_feed_cycles, z, sizeof(long long) * 8);
  }
  return 0;
}
#endif

extern "C" int tlbo_set_fit_variant(int variant) {
  const int old = g_fit_variant;
  g_fit_variant = variant ? 1 : 0;
  return old;
}

struct WorkLayout {
  int tile;              // register-panel tile (0 = shared-memory panel path)
  int cache;             // pair-value caches (Kc, Cg) fit in shared memory
  bool global_panels;
  size_t smem_fit, smem_fac;
  size_t panel_stride;   // doubles
};

static WorkLayout plan(int ncap, int d) {
  WorkLayout w;
  const size_t panel = fit_panel_bytes(ncap);
  const size_t st = fit_state_bytes();
  int tile = g_fit_variant != 0 ? panel_tile_for(ncap) : 0;
  int cache = tile > 0 ? 1 : 0;
  size_t small = fit_small_bytes(ncap, d, tile, cache);
  if (tile > 0 && st + small + panel + 1024 > kSmemLimit) { cache = 0; small = fit_small_bytes(ncap, d, tile, 0); }
  if (tile > 0 && st + small + panel + 1024 > kSmemLimit) { tile = 0; small = fit_small_bytes(ncap, d, 0, 0); }
  w.tile = tile;
  w.cache = cache;
  w.global_panels = (st + small + panel + 1024) > kSmemLimit;
  w.smem_fit = st + small + (w.global_panels ? 0 : panel);
  w.smem_fac = small + (w.global_panels ? 0 : panel);
  w.panel_stride = panel / sizeof(double);
  return w;
}

extern "C" int64_t tlbo_gp_fit_workspace_bytes(int B, int R, int ncap, int d) {
  // restart buffers + (possibly) global panels + device copies of ymean/ystd
  const int H = d + 2;
  size_t bytes = 0;
  bytes += (size_t)B * R * H * 8 + (size_t)B * R * 8 + (size_t)B * R * 16 + 2 * (size_t)B * 8 + 256;
  WorkLayout w = plan(ncap, d);
  if (w.global_panels) bytes += (size_t)B * R * w.panel_stride * 8;
  return (int64_t)bytes;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { tlbo_set_error(cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}

static int launch_factorize(const double* d_X, const double* d_y, const int32_t* d_n, int B, int R, int ncap,
                            int d, const SpaceDesc& sp, const double* theta_in, const double* f_in,
                            const double* d_ymean, const double* d_ystd, const tlbo_pack* pack, int slot0,
                            double* d_L, double* d_Kinv, int32_t* d_status, double* panels,
                            const WorkLayout& w, cudaStream_t s) {
  FacArgs fa;
  fa.X = d_X; fa.y = d_y; fa.n = d_n; fa.B = B; fa.R = R; fa.ncap = ncap; fa.d = d; fa.sp = sp;
  fa.theta_in = theta_in; fa.f_in = f_in; fa.ymean = d_ymean; fa.ystd = d_ystd;
  fa.pack = *pack; fa.slot0 = slot0; fa.L = d_L; fa.Kinv = d_Kinv; fa.status = d_status;
  fa.panels = w.global_panels ? panels : nullptr; fa.panel_stride = w.panel_stride;
  int rc = set_smem(gp_factorize_kernel, w.smem_fac);
  if (rc) return rc;
  gp_factorize_kernel<<<B, FIT_THREADS, w.smem_fac, s>>>(fa);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int check_common(int B, int ncap, int d, const tlbo_pack* pack, int slot0) {
  if (B <= 0 || ncap <= 0 || d <= 0 || d > TLBO_MAX_D) { tlbo_set_error("bad B/ncap/d"); return TLBO_E_BADARG; }
  if (pack) {
    if (pack->d != d || pack->ncap < ncap || slot0 < 0 || slot0 + B > pack->M) {
      tlbo_set_error("pack does not match the batch (d / ncap / slots)");
      return TLBO_E_BADARG;
    }
  }
  return 0;
}

extern "C" int tlbo_gp_fit_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap,
                                   int d, const int32_t* h_types, const double* d_p0, int R,
                                   const double* d_lo, const double* d_hi, const double* h_ymean,
                                   const double* h_ystd, int do_optimize, const tlbo_pack* pack, int slot0,
                                   double* d_restart_theta, double* d_restart_f, int32_t* d_restart_info,
                                   int32_t* d_status, void* d_work, void* stream) {
  int rc = check_common(B, ncap, d, pack, slot0);
  if (rc) return rc;
  if (R <= 0 || !pack || !d_status || !d_work) { tlbo_set_error("bad R / NULL pack, status or work"); return TLBO_E_BADARG; }
  SpaceDesc sp;
  rc = tlbo_make_space(d, h_types, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int H = d + 2;
  if (!do_optimize) R = 1;
  WorkLayout w = plan(ncap, d);
  // carve the workspace
  unsigned char* wp = (unsigned char*)d_work;
  double* w_theta = (double*)wp; wp += (size_t)B * R * H * 8;
  double* w_f = (double*)wp; wp += (size_t)B * R * 8;
  int32_t* w_info = (int32_t*)wp; wp += (size_t)B * R * 16;
  double* w_ymean = (double*)wp; wp += (size_t)B * 8;
  double* w_ystd = (double*)wp; wp += (size_t)B * 8;
  wp = (unsigned char*)(((uintptr_t)wp + 255) & ~(uintptr_t)255);
  double* panels = (double*)wp;
  TLBO_CUDA_CHECK(cudaMemcpyAsync(w_ymean, h_ymean, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  TLBO_CUDA_CHECK(cudaMemcpyAsync(w_ystd, h_ystd, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  double* rt = d_restart_theta ? d_restart_theta : w_theta;
  double* rf = d_restart_f ? d_restart_f : w_f;
  int32_t* ri = d_restart_info ? d_restart_info : w_info;

  FitArgs a;
  a.X = d_X; a.y = d_y; a.n = d_n; a.B = B; a.R = R; a.ncap = ncap; a.d = d; a.sp = sp;
  a.p0 = d_p0; a.lo = d_lo; a.hi = d_hi; a.do_optimize = do_optimize;
  a.restart_theta = rt; a.restart_f = rf; a.restart_info = ri;
  a.panels = w.global_panels ? panels : nullptr; a.panel_stride = w.panel_stride;
  a.cache = w.cache;
#define TLBO_LAUNCH_FIT(DM, WO, TT)                                          \
  do {                                                                       \
    rc = set_smem(gp_fit_kernel<DM, WO, TT>, w.smem_fit);                    \
    if (rc) return rc;                                                       \
    gp_fit_kernel<DM, WO, TT><<<B * R, FIT_THREADS, w.smem_fit, s>>>(a);     \
  } while (0)
#define TLBO_LAUNCH_FIT_T(DM)                                                \
  do {                                                                       \
    switch (tile) {                                                          \
      case 4: TLBO_LAUNCH_FIT(DM, true, 4); break;                           \
      case 5: TLBO_LAUNCH_FIT(DM, true, 5); break;                           \
      case 6: TLBO_LAUNCH_FIT(DM, true, 6); break;                           \
      case 7: TLBO_LAUNCH_FIT(DM, true, 7); break;                           \
      case 8: TLBO_LAUNCH_FIT(DM, true, 8); break;                           \
      default: TLBO_LAUNCH_FIT(DM, true, 0); break;                          \
    }                                                                        \
  } while (0)
  const int tile = w.tile;
  if (g_fit_variant != 0) {
    if (d <= 12) TLBO_LAUNCH_FIT_T(12); else TLBO_LAUNCH_FIT_T(24);
  } else {
    if (d <= 12) TLBO_LAUNCH_FIT(12, false, 0); else TLBO_LAUNCH_FIT(24, false, 0);
  }
#undef TLBO_LAUNCH_FIT_T
#undef TLBO_LAUNCH_FIT
  TLBO_CUDA_CHECK(cudaGetLastError());
  return launch_factorize(d_X, d_y, d_n, B, R, ncap, d, sp, rt, rf, w_ymean, w_ystd, pack, slot0, nullptr,
                          nullptr, d_status, panels, w, s);
}

extern "C" int tlbo_gp_factorize_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B,
                                         int ncap, int d, const int32_t* h_types, const double* d_theta,
                                         const double* h_ymean, const double* h_ystd, const tlbo_pack* pack,
                                         int slot0, double* d_L, double* d_Kinv, int32_t* d_status,
                                         void* d_work, void* stream) {
  int rc = check_common(B, ncap, d, pack, slot0);
  if (rc) return rc;
  if (!pack || !d_status || !d_work) { tlbo_set_error("NULL pack, status or work"); return TLBO_E_BADARG; }
  SpaceDesc sp;
  rc = tlbo_make_space(d, h_types, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  WorkLayout w = plan(ncap, d);
  unsigned char* wp = (unsigned char*)d_work;
  double* w_ymean = (double*)wp; wp += (size_t)B * 8;
  double* w_ystd = (double*)wp; wp += (size_t)B * 8;
  wp = (unsigned char*)(((uintptr_t)wp + 255) & ~(uintptr_t)255);
  double* panels = (double*)wp;
  TLBO_CUDA_CHECK(cudaMemcpyAsync(w_ymean, h_ymean, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  TLBO_CUDA_CHECK(cudaMemcpyAsync(w_ystd, h_ystd, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  return launch_factorize(d_X, d_y, d_n, B, 1, ncap, d, sp, d_theta, nullptr, w_ymean, w_ystd, pack, slot0,
                          d_L, d_Kinv, d_status, panels, w, s);
}

extern "C" int tlbo_gp_nll_grad_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B,
                                        int ncap, int d, const int32_t* h_types, const double* d_theta,
                                        double* d_nll, double* d_grad, int32_t* d_status, void* d_work,
                                        void* stream) {
  int rc = check_common(B, ncap, d, nullptr, 0);
  if (rc) return rc;
  SpaceDesc sp;
  rc = tlbo_make_space(d, h_types, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  WorkLayout w = plan(ncap, d);
  NllArgs a;
  a.X = d_X; a.y = d_y; a.n = d_n; a.B = B; a.ncap = ncap; a.d = d; a.sp = sp;
  a.theta = d_theta; a.nll = d_nll; a.grad = d_grad; a.status = d_status;
  a.panels = w.global_panels ? (double*)d_work : nullptr; a.panel_stride = w.panel_stride;
  a.cache = w.cache;
  if (w.global_panels && !d_work) { tlbo_set_error("workspace required for this ncap"); return TLBO_E_BADARG; }
#define TLBO_LAUNCH_NLL(DM, TT)                                              \
  do {                                                                       \
    rc = set_smem(gp_nll_grad_kernel<DM, TT>, w.smem_fac);                   \
    if (rc) return rc;                                                       \
    gp_nll_grad_kernel<DM, TT><<<B, FIT_THREADS, w.smem_fac, s>>>(a);        \
  } while (0)
#define TLBO_LAUNCH_NLL_T(DM)                                                \
  do {                                                                       \
    switch (tile) {                                                          \
      case 4: TLBO_LAUNCH_NLL(DM, 4); break;                                 \
      case 5: TLBO_LAUNCH_NLL(DM, 5); break;                                 \
      case 6: TLBO_LAUNCH_NLL(DM, 6); break;                                 \
      case 7: TLBO_LAUNCH_NLL(DM, 7); break;                                 \
      case 8: TLBO_LAUNCH_NLL(DM, 8); break;                                 \
      default: TLBO_LAUNCH_NLL(DM, 0); break;                                \
    }                                                                        \
  } while (0)
  const int tile = w.tile;
  if (d <= 12) TLBO_LAUNCH_NLL_T(12); else TLBO_LAUNCH_NLL_T(24);
#undef TLBO_LAUNCH_NLL_T
#undef TLBO_LAUNCH_NLL
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}
