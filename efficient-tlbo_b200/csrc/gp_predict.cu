// Posterior mean / variance kernels and the fused ensemble predict + EI + arg-max.
//
// Replaces reference tlbo/model/gp.py:250-305 (GaussianProcess._predict over skopt's
// GaussianProcessRegressor.predict(return_std=True)), tlbo/facade/rgpe.py:141-153 /
// base_facade.py:145-183 (ensemble combination), base_facade.py:110-143 (variance floor),
// tlbo/acquisition_function/acquisition.py:130-177 (EI) and
// tlbo/optimizer/ei_optimization.py:106-132 + tlbo/framework/smbo_offline.py:261-263
// (ordering and first-unevaluated selection).
//
// Variance form: the reference evaluates diag - k^T K_inv k with K_inv = L^-T L^-1; here the
// same quantity is evaluated as diag - |L^-1 k|^2 (triangular form, n^2 + 2n flops per
// surrogate and candidate by SURVEY 8d's convention), which is the sum of squares the
// reference's einsum expands to and cannot go negative by cancellation inside the sum.
//
// Fused kernel structure (per CTA: 8 warps, each owning NW tiles of 8 candidates):
//   for each surrogate m (accumulation order):
//     stage Xs_m, alpha_m, hyp_m in shared memory
//     K* phase : lane (g = lane>>2, t = lane&3) computes k(x_cand[g], x_train[4*js + t]) for
//                every j-step js -- exactly the B fragment (k = t, n = g) of an FP64
//                mma.m8n8k4 -- parks it in a lane-private shared-memory slot, and
//                accumulates k * alpha for the mean
//     V phase  : L^-1 streamed through shared memory in row chunks; per pair of 8-row blocks
//                acc[8 x 8] += Linv[8 x 4] * Kstar[4 x 8] on the FP64 tensor pipe (DMMA),
//                lower-triangular blocks only; |v|^2 accumulated from the C fragments
//     combine  : per candidate un-standardise, clip, weight, accumulate mu / var
//   EI, NaN rule, exclusion of evaluated rows, lexicographic (ei, key, idx) arg-max.
#include "tlbo_common.cuh"
#include <float.h>
#include <string.h>

#define PRED_THREADS 256
#define PRED_WARPS 8

// ---------------------------------------------------------------------------------------
// Small-query predict: one warp per (query row, surrogate).
struct PredSmallArgs {
  tlbo_pack pack;
  SpaceDesc sp;
  const double* Xq; int Nq;
  double* mu; double* var;
};

// ---- TMA (bulk async copy) helpers: candidate tiles of the fused kernel, L^-1 of the small-query kernel ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            unsigned long long* bar) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(gmem_src), "r"(bytes), "r"(mb)
               : "memory");
}
// the copy alone: the caller has announced the byte count of several copies with one expect_tx
__device__ __forceinline__ void tma_copy_1d(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            unsigned long long* bar) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned mb = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(gmem_src), "r"(bytes), "r"(mb)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 26); spin++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (done) return;
  }
  asm volatile("trap;");       // never hang the device on a lost transaction
}

#define SMALL_WARP_DOUBLES(ncap) ((ncap) + 2 * TLBO_HMAX)

// Posterior of surrogate m at one query row by ONE warp (kbuf: SMALL_WARP_DOUBLES(ncap) doubles of warp-private
// shared memory; xv: lane p < d holds the row's permuted column p, unscaled).  All lanes return the un-standardised
// mean / variance.  Every global load sits in a batch of independent loads (the kernel is a chain of L2 latencies,
// not of flops): scalars + scales, then the training rows eight columns at a time, then L^-1 sixteen rows at a time
// with the lanes across the columns (coalesced) and one transposed reduction per row block.
__device__ __forceinline__ void predict_small_item(const PredSmallArgs& a, int m, double xv, double* kbuf,
                                                   double& mu_out, double& var_out) {
  const int lane = threadIdx.x & 31;
  const int ncap = a.pack.ncap, d = a.pack.d, dc = a.sp.d_cont, ldl = TLBO_LDL(ncap);
  const double* Xs = a.pack.d_Xs + (size_t)m * ncap * d;
  const double* al = a.pack.d_alpha + (size_t)m * ncap;
  const double* L = a.pack.d_Linv + (size_t)m * ncap * ldl;
  double* xq = kbuf + ncap;            // [TLBO_HMAX] query row, permuted, continuous columns scaled
  double* hw = xq + TLBO_HMAX;         // [TLBO_HMAX] 1 / l of the continuous, Hamming weights of the categorical columns
  const int n = a.pack.d_n[m];
  const double* hyp = a.pack.d_hyp + (size_t)m * TLBO_HYP_STRIDE;
  const double c = hyp[0], ymean = hyp[2], ystd = hyp[3], diag = hyp[4];
  __syncwarp();
  unsigned qmask = 0u;
  if (lane < d) {
    const double sc = hyp[8 + lane];
    if (a.sp.has_cond && xv <= -1.0) qmask = 1u << lane;
    xq[lane] = (lane < dc) ? xv * sc : xv;
    hw[lane] = sc;
  }
  if (a.sp.has_cond) qmask = __reduce_or_sync(0xffffffffu, qmask);
  __syncwarp();
  double mpart = 0.0;
  for (int j = lane; j < n; j += 32) {
    const double* xj = Xs + (size_t)j * d;
    const double aj = al[j];
    double s = 0.0, hs = 0.0;
    unsigned jm = 0u;
    for (int p0 = 0; p0 < d; p0 += 8) {
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; u++) x[u] = (p0 + u < d) ? xj[p0 + u] : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int p = p0 + u;
        if (p < dc) { const double df = xq[p] - x[u]; s += df * df; }
        else if (p < d) { if (xq[p] != x[u]) hs += hw[p]; }
        // activity mask (gp_kernels.py:21-29); a training entry is inactive iff its raw value was <= -1, i.e. its
        // stored value (continuous columns are pre-multiplied by 1 / l > 0) is <= -(1 / l)
        if (a.sp.has_cond && p < d) jm |= (x[u] <= ((p < dc) ? -hw[p] : -1.0)) ? (1u << p) : 0u;
      }
    }
    double k = matern_hamming(c, s, hs);
    if (a.sp.has_cond && jm != qmask) k = 0.0;
    kbuf[j] = k;
    mpart += k * aj;
  }
  __syncwarp();
  double qpart = 0.0;
  for (int rb = 0; rb < n; rb += 16) {
    double v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = 0.0;
    const int jend = (rb + 16 < n) ? rb + 16 : n;        // columns 0 .. jend - 1 reach this row block (lower triangle)
    for (int j0 = 0; j0 < jend; j0 += 32) {
      const int j = j0 + lane;
      const double kj = (j < n) ? kbuf[j] : 0.0;
      const double* col = L + (size_t)rb * ldl + j;
      double l[16];
#pragma unroll
      for (int r = 0; r < 16; r++) l[r] = (rb + r < n && j <= rb + r) ? col[(size_t)r * ldl] : 0.0;   // one batch
#pragma unroll
      for (int r = 0; r < 16; r++) v[r] += l[r] * kj;
    }
    const double vi = warp_transpose_sum16(v);           // lanes r, r + 16: (L^-1 k)[rb + r]
    if (lane < 16) qpart += vi * vi;
  }
  const double mean = warp_sum(mpart);
  const double qq = warp_sum(qpart);
  double yv = diag - qq;
  if (yv < 0.0) yv = 0.0;
  const double sd = sqrt(yv);
  double var = sd * sd;
  if (!(var >= VERY_SMALL_NUMBER)) var = (var != var) ? var : VERY_SMALL_NUMBER;   // np.clip keeps NaN
  mu_out = mean * ystd + ymean;
  var_out = var * (ystd * ystd);
}

__global__ void __launch_bounds__(256) gp_predict_small_kernel(PredSmallArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* kbuf = reinterpret_cast<double*>(smem) + (size_t)wid * SMALL_WARP_DOUBLES(a.pack.ncap);
  const long long total = (long long)a.Nq * a.pack.M;
  for (long long item = (long long)blockIdx.x * warps + wid; item < total; item += (long long)gridDim.x * warps) {
    const int m = (int)(item / a.Nq), q = (int)(item - (long long)m * a.Nq);
    double mu, var;
    const double xv = (lane < a.pack.d) ? a.Xq[(size_t)q * a.pack.d + a.sp.perm[lane]] : 0.0;
    predict_small_item(a, m, xv, kbuf, mu, var);
    if (lane == 0) {
      a.mu[(size_t)m * a.Nq + q] = mu;
      a.var[(size_t)m * a.Nq + q] = var;
    }
  }
}

extern "C" int tlbo_gp_predict(const tlbo_pack* pack, const int32_t* h_types, const double* d_Xq, int Nq,
                               double* d_mu, double* d_var, void* stream) {
  if (!pack || Nq <= 0) { tlbo_set_error("bad pack / Nq"); return TLBO_E_BADARG; }
  PredSmallArgs a;
  a.pack = *pack;
  int rc = tlbo_make_space(pack->d, h_types, &a.sp);
  if (rc) return rc;
  a.Xq = d_Xq; a.Nq = Nq; a.mu = d_mu; a.var = d_var;
  const int warps = 8;
  const size_t smem = (size_t)warps * SMALL_WARP_DOUBLES(pack->ncap) * sizeof(double);
  if (smem > 200 * 1024) { tlbo_set_error("ncap too large for predict_small"); return TLBO_E_TOOLARGE; }
  if (smem > 48 * 1024)
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(gp_predict_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long items = (long long)Nq * pack->M;
  int grid = (int)((items + warps - 1) / warps);
  if (grid > 148 * 8) grid = 148 * 8;
  gp_predict_small_kernel<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(a);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Ensemble combination and EI of one query row from its per-surrogate posteriors, shared by the fused pool kernel
// and the small-query kernel below (same operations in the same order: the two must agree bit for bit).
struct EnsAcc { double mu, var, wsum; };

__device__ __forceinline__ void ens_combine(int mode, bool first, double wm, double mu_m, double var_m, EnsAcc& e) {
  if (mode == 0) {
    // rgpe.py:150-152 / base_facade.py:170-171: `mu += w * mu_t`, `var += w * w * var_t` -- numpy rounds the
    // product and the sum separately (no FMA contraction: nvcc's -fmad would fuse them)
    e.mu = __dadd_rn(e.mu, __dmul_rn(wm, mu_m));
    e.var = __dadd_rn(e.var, __dmul_rn(__dmul_rn(wm, wm), var_m));
  } else if (mode == 2) {
    const double iv = 1.0 / var_m;
    e.var = __dadd_rn(e.var, __dmul_rn(iv, wm));
    e.mu = __dadd_rn(e.mu, __dmul_rn(__dmul_rn(iv, mu_m), wm));
  } else {
    if (first) { e.mu = mu_m; e.var = var_m; }
    else { e.mu = __dadd_rn(e.mu, __dmul_rn(wm, mu_m)); e.wsum = __dadd_rn(e.wsum, wm); }
  }
}

// Final mean / variance of the ensemble (mode rules, facade variance floor) and EI
// (acquisition.py:130-177); NaN EI is returned as NaN (the callers apply their own NaN rule).
__device__ __forceinline__ double ens_ei(int mode, const EnsAcc& e, double eta, double par, double var_floor,
                                         double& mu_out, double& var_out) {
  double mu = e.mu, var = e.var;
  if (mode == 1) mu = mu / (0.75 + e.wsum);
  if (mode == 2) {
    // POGPE (pogpe.py:57-60): precision-weighted product of experts
    double tmp = e.var;
    if (tmp == 0.0) tmp = 1e-5;
    var = 1.0 / tmp;
    mu = __dmul_rn(e.mu, var);
  }
  if (var < var_floor) var = var_floor;
  if (var != var) var = var_floor;
  const double s = sqrt(var);
  double f;
  if (s == 0.0) f = 0.0;
  else {
    const double dm = eta - mu - par;
    const double z = dm / s;
    const double cdf = 0.5 * erfc(-z * 0.7071067811865476);
    const double pdf = exp(-0.5 * z * z) * 0.3989422804014327;
    // two products and one sum, each rounded (numpy semantics, no FMA contraction)
    f = __dadd_rn(__dmul_rn(dm, cdf), __dmul_rn(s, pdf));
    // z < -37.5: Phi(z) is subnormal and the two terms (equal to ~1e-3 relative) carry
    // no relative precision, so the sign of their difference is rounding luck of the
    // libm in use (glibc happens to stay >= 0); EI there is < 1e-305 * s.
    if (f < 0.0 && cdf < 2.2250738585072014e-308) f = 0.0;
  }
  mu_out = mu; var_out = var;
  return f;
}

// ---------------------------------------------------------------------------------------
// Small-query ensemble EI in ONE launch (the online scenario's hill-climb step: ~30 neighbours x K + 1 surrogates):
// one CTA per (query row, surrogate) -- ~800 small CTAs, five to a SM, so that the chains of dependent latencies of
// different posteriors overlap.  The row's posteriors meet in d_mu / d_var; the LAST CTA of a row (ticket counter)
// loads them side by side, thread 0 combines them in accumulation order, scores EI and stores it.  Xq and ei may be
// mapped pinned HOST memory (UVA): results are written over the bus, no copy engine involved, and a host that
// pre-filled ei[] with TLBO_EI_PENDING can spin on the rows instead of synchronising the stream.
#define SMALL_EI_INLINE_DOUBLES 1024

struct SmallEiArgs {
  PredSmallArgs p;       // p.mu / p.var: per-surrogate posteriors [M, Nq] (scratch the combination reads)
  const double* w; const int32_t* skip; const int32_t* order; int mode;
  double eta, par, var_floor;
  double* ei;            // [Nq]
  unsigned* counter;     // [Nq], zero on entry; left zero
  int nst;               // > 0: surrogates with n <= nst rows have Xs / alpha / L^-1 bulk-copied to shared memory
  int inlined;           // the query rows travel in `xin` (kernel arguments, constant bank) instead of p.Xq
  double xin[SMALL_EI_INLINE_DOUBLES];
};

#define SMALL_EI_WARPS 4
#define SMALL_EI_THREADS (SMALL_EI_WARPS * 32)
// doubles of the staging area: L^-1 rows at stride nst (row count rounded up to 16) | Xs | alpha  (nst even)
#define SMALL_EI_LROWS(nst) (((nst) + 15) & ~15)          /* the matrix-vector product reads whole blocks of 16 rows */
#define SMALL_EI_STAGE_DOUBLES(nst, d) ((size_t)SMALL_EI_LROWS(nst) * (nst) + (size_t)(nst) * (d) + (nst) + 2)

// CTA (q, m): the posterior of surrogate m at query row q by four warps.
//   stage   : (n <= nst) the surrogate's training rows, alpha and the first n columns of the first n rows of L^-1 by
//             bulk copies (TMA; one per row of L^-1, issued by as many threads) -- in flight while the scalars load
//   K*      : one training row per thread, k_j to shared memory, k_j alpha_j accumulated
//   L^-1 k  : one block of 16 rows per warp and round, lanes across the columns, no predicates (the pack holds exact
//             zeros above the diagonal), one transposed reduction per block
//   finish  : mean / variance to d_mu / d_var; the LAST CTA of a row (ticket counter) loads the row's M posteriors
//             side by side, thread 0 combines them in accumulation order, scores EI and stores it.
__global__ void __launch_bounds__(SMALL_EI_THREADS, 6) ensemble_ei_small_kernel(const __grid_constant__ SmallEiArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int M = a.p.pack.M, Nq = a.p.Nq, q = blockIdx.x, m = blockIdx.y;
  const int ncap = a.p.pack.ncap, d = a.p.pack.d, dc = a.p.sp.d_cont, nst = a.nst;
  double* kbuf = reinterpret_cast<double*>(smem);          // [ncap] K*
  double* xq = kbuf + ncap;                                // [TLBO_HMAX] query row, permuted, continuous columns scaled
  double* hw = xq + TLBO_HMAX;                             // [TLBO_HMAX] 1 / l | Hamming weights
  double* sL = hw + TLBO_HMAX;                             // staging area
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ double s_red[2 * SMALL_EI_WARPS];
  __shared__ unsigned s_ticket;
  const int n = a.p.pack.d_n[m];
  const double* hyp = a.p.pack.d_hyp + (size_t)m * TLBO_HYP_STRIDE;
  const double* Xs = a.p.pack.d_Xs + (size_t)m * ncap * d;
  const double* al = a.p.pack.d_alpha + (size_t)m * ncap;
  const double* L = a.p.pack.d_Linv + (size_t)m * ncap * TLBO_LDL(ncap);
  int ldl = TLBO_LDL(ncap);
  const bool staged = n > 0 && n <= nst;                   // uniform
  if (n > 0) {
    if (staged) {
      double* sX = sL + (size_t)SMALL_EI_LROWS(nst) * nst;
      double* sA = sX + (((size_t)nst * d + 1) & ~(size_t)1);
      const int ne = (n + 1) & ~1;
      const unsigned bx = (unsigned)((((size_t)n * d + 1) & ~(size_t)1) * sizeof(double));
      const unsigned ba = (unsigned)(ne * sizeof(double)), br = (unsigned)(ne * sizeof(double));
      if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned mb0 = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
        const unsigned mb1 = (unsigned)__cvta_generic_to_shared(&s_bar[1]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb0), "r"(bx + ba) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb1), "r"(br * (unsigned)n) : "memory");
        tma_copy_1d(sX, Xs, bx, &s_bar[0]);
        tma_copy_1d(sA, al, ba, &s_bar[0]);
      }
      __syncthreads();                                     // barriers initialised and armed before any row copy
      for (int i = tid; i < n; i += SMALL_EI_THREADS) tma_copy_1d(sL + (size_t)i * nst, L + (size_t)i * ldl, br, &s_bar[1]);
      Xs = sX; al = sA; L = sL; ldl = nst;
    }
    const double c = hyp[0];
    unsigned qm = 0u;
    if (tid < d) {
      const double xv = a.inlined ? a.xin[q * d + a.p.sp.perm[tid]] : a.p.Xq[(size_t)q * d + a.p.sp.perm[tid]];
      const double sc = hyp[8 + tid];
      if (a.p.sp.has_cond && xv <= -1.0) qm = 1u << tid;
      xq[tid] = (tid < dc) ? xv * sc : xv;
      hw[tid] = sc;
    }
    unsigned qmask = 0u;
    if (a.p.sp.has_cond) {                                 // d <= TLBO_HMAX < 32: the pattern bits live in warp 0
      qm = __reduce_or_sync(0xffffffffu, qm);
      if (tid == 0) s_ticket = qm;
    }
    __syncthreads();
    if (a.p.sp.has_cond) qmask = s_ticket;
    if (staged) mbar_wait(&s_bar[0], 0);
    double mpart = 0.0;
    for (int j = tid; j < n; j += SMALL_EI_THREADS) {
      const double* xj = Xs + (size_t)j * d;
      double s = 0.0, hs = 0.0;
      unsigned jm = 0u;
      for (int p0 = 0; p0 < d; p0 += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = (p0 + u < d) ? xj[p0 + u] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int p = p0 + u;
          if (p < dc) { const double df = xq[p] - x[u]; s += df * df; }
          else if (p < d) { if (xq[p] != x[u]) hs += hw[p]; }
          // activity mask (gp_kernels.py:21-29), as in predict_small_item
          if (a.p.sp.has_cond && p < d) jm |= (x[u] <= ((p < dc) ? -hw[p] : -1.0)) ? (1u << p) : 0u;
        }
      }
      double k = matern_hamming(c, s, hs);
      if (a.p.sp.has_cond && jm != qmask) k = 0.0;
      kbuf[j] = k;
      mpart += k * al[j];
    }
    __syncthreads();
    if (staged) mbar_wait(&s_bar[1], 0);
    double qpart = 0.0;
    for (int rb = 16 * wid; rb < n; rb += 16 * SMALL_EI_WARPS) {
      double v[16];
#pragma unroll
      for (int r = 0; r < 16; r++) v[r] = 0.0;
      const int jend = (rb + 16 < n) ? rb + 16 : n;        // columns 0 .. jend - 1 reach this row block
      for (int j0 = 0; j0 < jend; j0 += 32) {
        const int j = j0 + lane;
        const bool jok = j < n;
        const double kj = jok ? kbuf[j] : 0.0;
        const double* col = L + (size_t)rb * ldl + (jok ? j : 0);
        // rows >= n of the staging area are never written: their sums are dropped below, never mixed in
#pragma unroll
        for (int r0 = 0; r0 < 16; r0 += 8) {               // eight loads in flight (six CTAs per SM: 80 registers)
          double l[8];
#pragma unroll
          for (int r = 0; r < 8; r++)
            l[r] = (staged || rb + r0 + r < n) ? col[(size_t)(r0 + r) * ldl] : 0.0;
#pragma unroll
          for (int r = 0; r < 8; r++) v[r0 + r] += l[r] * kj;
        }
      }
      const double vi = warp_transpose_sum16(v);           // lanes r, r + 16: (L^-1 k)[rb + r]
      if (lane < 16 && rb + lane < n) qpart += vi * vi;
    }
    mpart = warp_sum(mpart);
    qpart = warp_sum(qpart);
    if (lane == 0) { s_red[wid] = mpart; s_red[SMALL_EI_WARPS + wid] = qpart; }
  }
  __syncthreads();
  if (tid == 0) {
    double mu = 0.0, var = 0.0;
    if (n > 0) {
      double mean = 0.0, qq = 0.0;
#pragma unroll
      for (int w = 0; w < SMALL_EI_WARPS; w++) { mean += s_red[w]; qq += s_red[SMALL_EI_WARPS + w]; }
      const double ymean = hyp[2], ystd = hyp[3], diag = hyp[4];
      double yv = diag - qq;
      if (yv < 0.0) yv = 0.0;
      const double sd = sqrt(yv);
      var = sd * sd;
      if (!(var >= VERY_SMALL_NUMBER)) var = (var != var) ? var : VERY_SMALL_NUMBER;   // np.clip keeps NaN
      mu = mean * ystd + ymean;
      var = var * (ystd * ystd);
    }
    a.p.mu[(size_t)m * Nq + q] = mu;
    a.p.var[(size_t)m * Nq + q] = var;
    __threadfence();
    s_ticket = atomicAdd(a.counter + q, 1u);
  }
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  // ---- last CTA of row q: every surrogate's posterior is in d_mu / d_var
  __threadfence();
  double* s_mu = reinterpret_cast<double*>(smem);           // reuse of the buffers above: [M] x 3 + int[M]
  double* s_var = s_mu + M;
  double* s_w = s_var + M;
  int* s_m = reinterpret_cast<int*>(s_w + M);               // slot of accumulation position mi, -1: dropped
  for (int mi = tid; mi < M; mi += blockDim.x) {
    const int mm = a.order ? a.order[mi] : mi;
    const bool keep = !(a.skip && a.skip[mm]) && a.p.pack.d_n[mm] > 0;
    s_m[mi] = keep ? mm : -1;
    s_w[mi] = a.w[mm];
    s_mu[mi] = __ldcg(a.p.mu + (size_t)mm * Nq + q);
    s_var[mi] = __ldcg(a.p.var + (size_t)mm * Nq + q);
  }
  __syncthreads();
  if (tid == 0) {
    EnsAcc e; e.mu = 0.0; e.var = 0.0; e.wsum = 0.0;
    for (int mi = 0; mi < M; mi++)
      if (s_m[mi] >= 0) ens_combine(a.mode, mi == 0, s_w[mi], s_mu[mi], s_var[mi], e);
    double mu, var;
    // one 8-byte store per row, no fence: a host that pre-filled ei[] with TLBO_EI_PENDING sees each row complete
    // the moment its value lands (system-wide fences cost ~3 us each on this path)
    a.ei[q] = ens_ei(a.mode, e, a.eta, a.par, a.var_floor, mu, var);
    a.counter[q] = 0u;
  }
}

extern "C" int tlbo_ensemble_ei_small(const tlbo_pack* pack, const int32_t* h_types, const double* d_w,
                                      const int32_t* d_skip, const int32_t* d_order, int mode, const double* Xq,
                                      const double* h_Xq, int Nq, double eta, double par, double var_floor,
                                      double* d_mu, double* d_var, double* ei, uint32_t* d_counter, void* stream) {
  if (!pack || !d_w || (!Xq && !h_Xq) || Nq <= 0 || !d_mu || !d_var || !ei || !d_counter) {
    tlbo_set_error("tlbo_ensemble_ei_small: bad arguments");
    return TLBO_E_BADARG;
  }
  static thread_local SmallEiArgs a;       // 8 KB of inline rows: not on the stack; one per calling thread (trial threads)
  a.p.pack = *pack;
  int rc = tlbo_make_space(pack->d, h_types, &a.p.sp);
  if (rc) return rc;
  a.p.Xq = Xq; a.p.Nq = Nq; a.p.mu = d_mu; a.p.var = d_var;
  a.w = d_w; a.skip = d_skip; a.order = d_order; a.mode = mode;
  a.eta = eta; a.par = par; a.var_floor = var_floor;
  a.ei = ei; a.counter = d_counter;
  const int M = pack->M, d = pack->d;
  a.inlined = (h_Xq && (long long)Nq * d <= SMALL_EI_INLINE_DOUBLES) ? 1 : 0;
  if (a.inlined) memcpy(a.xin, h_Xq, sizeof(double) * (size_t)Nq * d);
  else if (!Xq) { tlbo_set_error("tlbo_ensemble_ei_small: query set too large to inline and no device-readable Xq"); return TLBO_E_BADARG; }
  // K* / query buffers + (when four CTAs still fit on an SM) the staging area for surrogates of up to nst rows --
  // nst from the pack's nmax hint (0 = unknown: the capacity)
  const size_t small = sizeof(double) * (size_t)SMALL_WARP_DOUBLES(pack->ncap);
  int nst = (pack->nmax > 0 && pack->nmax < pack->ncap) ? pack->nmax : pack->ncap;
  nst = (nst + 1) & ~1;
  const size_t stage = sizeof(double) * SMALL_EI_STAGE_DOUBLES(nst, d);
  a.nst = (small + stage <= 52 * 1024) ? nst : 0;
  size_t smem = small + (a.nst ? stage : 0);
  const size_t tail = sizeof(double) * 3 * (size_t)M + sizeof(int) * (size_t)M;
  if (tail > smem) smem = tail;
  if (smem > 200 * 1024 || M > 65535) { tlbo_set_error("ncap / M too large for ensemble_ei_small"); return TLBO_E_TOOLARGE; }
  static size_t smem_set = 0;
  static bool carveout_set = false;
  if (!carveout_set) {
    // ~800 CTAs of ~30 KB each must all be resident (six to a SM): ask for the largest shared-memory carve-out
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(ensemble_ei_small_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  if (smem > 48 * 1024 && smem > smem_set) {
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(ensemble_ei_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)Nq, (unsigned)M);
  ensemble_ei_small_kernel<<<grid, SMALL_EI_THREADS, smem, (cudaStream_t)stream>>>(a);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Fused predict + EI + arg-max.
struct Best { double ei; double key; long long idx; };

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  // lexsort((key, ei))[::-1] order: larger ei first, then larger key, then larger index
  if (a.ei != b.ei) return a.ei > b.ei;
  if (a.key != b.key) return a.key > b.key;
  return a.idx > b.idx;
}

__device__ __forceinline__ Best warp_best(Best v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best w;
    w.ei = __shfl_xor_sync(0xffffffffu, v.ei, o);
    w.key = __shfl_xor_sync(0xffffffffu, v.key, o);
    w.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
    if (better(w, v)) v = w;
  }
  return v;
}

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct FusedArgs {
  tlbo_pack pack;
  SpaceDesc sp;
  const double* w; const int32_t* skip; const int32_t* order; int mode;
  const double* cand; long long N;
  double eta, par, var_floor;
  const double* keys; const long long* excluded; int n_excl; long long idx_offset;
  double* mu; double* var; double* ei;
  double* partial;       // [grid, 4]
  const double* cmu; const double* cvar; int ncached;   // resident per-surrogate posteriors [ncached, N]
  double* pm_mu; double* pm_var;                        // optional per-surrogate outputs [M, N]
  int rc_rows;           // Linv rows per shared-memory chunk (multiple of 16)
  int njmax;             // j-steps (of 4 training rows) the shared-memory plan provides for
};

// shared-memory plan (doubles): cand [TILE * DP] | Xs [(4 njmax + 4) * DP] | alpha [4 njmax + 4] |
//   hyp [32 + 32] | Linv chunk [rc_rows * ldl] | spill [8 warps][njmax][NW][32] | comb [8 warps][NW*8][2]
// DCP / DKP > 0: specialised layout -- continuous / categorical column counts padded to the
// compile-time DCP / DKP (pad columns are zero on both sides, so they contribute nothing), rows
// padded to DP = DCP + DKP doubles and read with 128-bit loads, two training rows in flight per
// lane.  DCP == 0: generic layout for any d (predicated loops).  ncu on the generic kernel showed
// 52% of all samples in the fully unrolled, predicated 22-column distance loop.
template <int NW, int DCP, int DKP, int MINB>
__global__ void __launch_bounds__(PRED_THREADS, MINB) predict_ei_fused_kernel(FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ncap = a.pack.ncap, d = a.pack.d, dc = a.sp.d_cont, ldl = TLBO_LDL(ncap);
  constexpr int WC = NW * 8;                 // candidates per warp
  constexpr int TILE = PRED_WARPS * WC;      // candidates per CTA iteration
  constexpr bool FAST = DCP > 0;
  const int DP = FAST ? (DCP + DKP) : d;     // row length of the staged candidate / training rows
  const int nrows = 4 * a.njmax + 4;
  double* s_cand = reinterpret_cast<double*>(smem);
  double* s_Xs = s_cand + (size_t)TILE * DP;
  double* s_alpha = s_Xs + (size_t)nrows * DP;
  double* s_hyp = s_alpha + nrows;
  double* s_hypp = s_hyp + TLBO_HYP_STRIDE;   // padded per-column scales: [DCP] 1/l, [DKP] 1/(2 l^2)
  double* s_Linv = s_hypp + TLBO_HYP_STRIDE;
  double* s_spill = s_Linv + (size_t)a.rc_rows * ldl;
  double* s_comb = s_spill + (size_t)PRED_WARPS * a.njmax * NW * 32;
  double* my_spill = s_spill + (size_t)wid * a.njmax * NW * 32;
  double* my_comb = s_comb + (size_t)wid * WC * 2;
  double* s_raw = s_comb + (size_t)PRED_WARPS * WC * 2;     // TMA landing buffer: TILE * d raw candidate rows
  // active-surrogate list (accumulation order, skipped / empty slots dropped): weight, slot, n, flags
  double* s_lw = s_raw + (size_t)TILE * d;
  int* s_lm = reinterpret_cast<int*>(s_lw + a.pack.M);
  int* s_ln = s_lm + a.pack.M;
  int* s_lf = s_ln + a.pack.M;            // bit 0: first entry of the order (mode 1 initialiser); bits 8..: cached run length
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ int s_L, s_neval, s_evali;
  if (tid == 0) mbar_init(&s_mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (wid == 0) {
    // order-preserving compaction by warp 0: the per-tile loops below then touch no global metadata
    int count = 0, neval = 0, evali = -1;
    for (int base = 0; base < a.pack.M; base += 32) {
      const int mi = base + lane;
      int m = 0, n = 0; bool keep = false;
      if (mi < a.pack.M) {
        m = a.order ? a.order[mi] : mi;
        n = a.pack.d_n[m];
        keep = !(a.skip && a.skip[m]) && n > 0;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      const int pos = count + __popc(mask & ((1u << lane) - 1u));
      const bool cached = keep && a.cmu && m < a.ncached;
      if (keep) { s_lw[pos] = a.w[m]; s_lm[pos] = m; s_ln[pos] = n; s_lf[pos] = (mi == 0) ? 1 : 0; }
      const unsigned emask = __ballot_sync(0xffffffffu, keep && !cached);
      if (emask) { neval += __popc(emask); evali = count + __popc(mask & ((1u << (31 - __clz(emask))) - 1u)); }
      count += __popc(mask);
    }
    __syncwarp();
    // run lengths of consecutive resident (cached) entries, from the back
    if (lane == 0) {
      int run = 0;
      for (int i = count - 1; i >= 0; i--) {
        const bool cached = a.cmu && s_lm[i] < a.ncached;
        run = cached ? run + 1 : 0;
        s_lf[i] |= run << 8;
      }
      s_L = count; s_neval = neval; s_evali = evali;
    }
  }
  __syncthreads();
  const int L = s_L;
  unsigned tma_phase = 0;

  Best best; best.ei = -INFINITY; best.key = -INFINITY; best.idx = -1;
  double negcount = 0.0;
  const long long ntiles = (a.N + TILE - 1) / TILE;
  // A tile's rows are contiguous in the pool: one 1-D bulk copy (TMA) lands them in shared memory
  // while the previous tile is still being scored; tiles whose byte count or address is not a
  // multiple of 16 (an odd tail) take the plain coalesced-load path.
  auto tma_bytes = [&](long long t) -> unsigned {
    const long long r0 = t * TILE;
    const long long rows = (a.N - r0 < TILE) ? a.N - r0 : TILE;
    const unsigned long long bytes = (unsigned long long)rows * d * sizeof(double);
    const unsigned long long addr = (unsigned long long)(a.cand + (size_t)r0 * d);
    return ((bytes & 15ull) == 0 && (addr & 15ull) == 0) ? (unsigned)bytes : 0u;
  };
  if (blockIdx.x < ntiles && tid == 0) {
    const unsigned b0 = tma_bytes(blockIdx.x);
    if (b0) tma_load_1d(s_raw, a.cand + (size_t)blockIdx.x * TILE * d, b0, &s_mbar);
  }
  // ---- staging helpers (all threads; the callers place the barriers)
  auto stage_model = [&](int m, int n) {
    const double* gX = a.pack.d_Xs + (size_t)m * ncap * d;
    const double* gA = a.pack.d_alpha + (size_t)m * ncap;
    const double* gH = a.pack.d_hyp + (size_t)m * TLBO_HYP_STRIDE;
    if (FAST) {
      for (int idx = tid; idx < nrows * DP; idx += PRED_THREADS) {
        const int j = idx / DP, pp = idx - j * DP;
        const int p = (pp < DCP) ? pp : dc + (pp - DCP);
        const bool ok = (j < n) && ((pp < DCP) ? (pp < dc) : (p < d));
        s_Xs[idx] = ok ? gX[(size_t)j * d + p] : 0.0;
      }
      for (int idx = tid; idx < nrows; idx += PRED_THREADS) s_alpha[idx] = (idx < n) ? gA[idx] : 0.0;
      if (tid < TLBO_HYP_STRIDE) {
        s_hyp[tid] = gH[tid];
        const int p = (tid < DCP) ? tid : dc + (tid - DCP);
        const bool ok = (tid < DCP) ? (tid < dc) : (tid < DCP + DKP && p < d);
        s_hypp[tid] = ok ? gH[8 + p] : 0.0;
      }
    } else {
      for (int idx = tid; idx < n * d; idx += PRED_THREADS) s_Xs[idx] = gX[idx];
      for (int idx = tid; idx < nrows; idx += PRED_THREADS) s_alpha[idx] = (idx < n) ? gA[idx] : 0.0;
      if (tid < TLBO_HYP_STRIDE) s_hyp[tid] = gH[tid];
    }
  };
  auto stage_linv = [&](int m, int r0, int rows) {
    // only columns < r0 + rows are ever read (lower triangular): copy that many, not ldl
    const double* src = a.pack.d_Linv + (size_t)m * ncap * ldl + (size_t)r0 * ldl;
    const int ncol = (r0 + rows + 3) & ~3;
    const int c4 = ncol >> 2;                       // 4-column groups per row
    for (int idx = tid; idx < rows * c4; idx += PRED_THREADS) {
      const int rr = idx / c4, cc = (idx - rr * c4) << 2;
      const double2 v0 = *reinterpret_cast<const double2*>(src + (size_t)rr * ldl + cc);
      const double2 v1 = *reinterpret_cast<const double2*>(src + (size_t)rr * ldl + cc + 2);
      *reinterpret_cast<double2*>(s_Linv + (size_t)rr * ldl + cc) = v0;
      *reinterpret_cast<double2*>(s_Linv + (size_t)rr * ldl + cc + 2) = v1;
    }
  };
  // Same copy with 16-byte asynchronous copies (LDGSTS): issued before the K* phase, awaited after it
  auto stage_linv_async = [&](int m, int r0, int rows) {
    const double* src = a.pack.d_Linv + (size_t)m * ncap * ldl + (size_t)r0 * ldl;
    const int ncol = (r0 + rows + 3) & ~3;
    const int c2 = ncol >> 1;                       // 2-column (16-byte) groups per row
    for (int idx = tid; idx < rows * c2; idx += PRED_THREADS) {
      const int rr = idx / c2, cc = (idx - rr * c2) << 1;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(s_Linv + (size_t)rr * ldl + cc);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)rr * ldl + cc) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Exactly one surrogate is evaluated per tile (the usual per-iteration launch: the target, all
  // sources resident) and its L^-1 fits one chunk: stage it ONCE per CTA, not once per tile.
  bool persist = false;
  if (s_neval == 1) {
    const int n1 = s_ln[s_evali];
    if (((n1 + 7) >> 3) * 8 <= a.rc_rows) {
      persist = true;
      stage_model(s_lm[s_evali], n1);
      stage_linv(s_lm[s_evali], 0, ((n1 + 7) >> 3) * 8);
    }
  }
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TILE;
    if (a.cmu && lane < WC) {
      // resident posteriors of this tile: pull them towards L2 now, while the candidate tile is
      // staged -- the combination loop below reads them in accumulation order
      const long long prow = row0 + wid * WC + lane;
      if (prow < a.N) {
        for (int i = 0; i < L; i++) {
          const int m = s_lm[i];
          if (m >= a.ncached) continue;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.cmu + (size_t)m * a.N + prow));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.cvar + (size_t)m * a.N + prow));
        }
      }
    }
    __syncthreads();                                        // previous tile's readers of s_cand are done
    // stage the candidate tile (permuted [and padded] columns) in ONE pass
    {
      const long long lim = (a.N - row0 < TILE ? a.N - row0 : TILE) * d;
      const bool via_tma = tma_bytes(tile) != 0;            // uniform
      const double* src = via_tma ? s_raw : a.cand + (size_t)row0 * d;
      if (via_tma) { mbar_wait(&s_mbar, tma_phase); tma_phase ^= 1; }
      if (FAST) {
        for (int idx = tid; idx < TILE * DP; idx += PRED_THREADS) {
          const int r = idx / DP, pp = idx - r * DP;
          const int p = (pp < DCP) ? ((pp < dc) ? pp : -1) : ((dc + pp - DCP < d) ? dc + pp - DCP : -1);
          double v = 0.0;
          if (p >= 0) {
            const long long sidx = (long long)r * d + a.sp.perm[p];
            if (sidx < lim) v = src[sidx];
          }
          s_cand[idx] = v;
        }
      } else {
        for (int idx = tid; idx < TILE * d; idx += PRED_THREADS) {
          const int r = idx / d, p = idx - r * d;
          const long long sidx = (long long)r * d + a.sp.perm[p];
          s_cand[idx] = (sidx < lim) ? src[sidx] : 0.0;
        }
      }
      __syncthreads();                                      // s_cand complete, s_raw free again
      const long long nxt = tile + gridDim.x;
      if (nxt < ntiles && tid == 0) {
        const unsigned bn = tma_bytes(nxt);
        if (bn) tma_load_1d(s_raw, a.cand + (size_t)nxt * TILE * d, bn, &s_mbar);
      }
    }
    // per-lane ensemble accumulators: lane cl < WC owns candidate cl of this warp
    EnsAcc acc; acc.mu = 0.0; acc.var = 0.0; acc.wsum = 0.0;
    const long long myrow = row0 + wid * WC + lane;
    const bool myok = lane < WC && myrow < a.N;
    auto combine = [&](bool first, double wm, double mu_m, double var_m) {
      ens_combine(a.mode, first, wm, mu_m, var_m, acc);
    };
    int li = 0;
    while (li < L) {
      const int run = s_lf[li] >> 8;
      if (run > 0) {
        // resident posteriors (they do not depend on the target observations): two coalesced 8-byte
        // reads per candidate and surrogate.  Up to CB surrogates' loads are issued back to back,
        // then accumulated in the fixed order -- one memory latency per batch, not per surrogate.
        constexpr int CB = 8;
        const int r = run < CB ? run : CB;
        double cm[CB], cv[CB];
#pragma unroll
        for (int b = 0; b < CB; b++) {
          cm[b] = 0.0; cv[b] = 0.0;
          if (b < r && myok) {
            const size_t off = (size_t)s_lm[li + b] * a.N + myrow;
            cm[b] = __ldg(a.cmu + off);
            cv[b] = __ldg(a.cvar + off);
          }
        }
        if (lane < WC) {
#pragma unroll
          for (int b = 0; b < CB; b++)
            if (b < r) combine((s_lf[li + b] & 1) != 0, s_lw[li + b], cm[b], cv[b]);
        }
        li += r;
        continue;
      }
      const int m = s_lm[li], n = s_ln[li];
      const double wm = s_lw[li];
      const bool first = (s_lf[li] & 1) != 0;
      li++;
      const int nib = (n + 7) >> 3;           // 8-row blocks
      if (!persist) {
        __syncthreads();                      // previous surrogate's readers of s_Xs / s_Linv are done
        stage_model(m, n);
        // first L^-1 chunk: in flight while the K* phase below runs
        stage_linv_async(m, 0, (nib * 8 < a.rc_rows) ? nib * 8 : a.rc_rows);
        __syncthreads();
      }
      const double c = s_hyp[0], ymean = s_hyp[2], ystd = s_hyp[3], diag = s_hyp[4];
      const int nj = (n + 3) >> 2;            // j-steps of 4 training rows
      // ---- K* phase
      double mpart[NW];
      if constexpr (FAST) {
        constexpr int DPC = DCP + DKP;
#pragma unroll
        for (int w = 0; w < NW; w++) {
          mpart[w] = 0.0;
          const double* xc = s_cand + (size_t)(wid * WC + w * 8 + g) * DPC;
          double xs[DPC];
#pragma unroll
          for (int p = 0; p < DPC; p += 2) {
            const double2 v = *reinterpret_cast<const double2*>(xc + p);
            xs[p] = (p < DCP) ? v.x * s_hypp[p] : v.x;
            xs[p + 1] = (p + 1 < DCP) ? v.y * s_hypp[p + 1] : v.y;
          }
          double hk[DKP > 0 ? DKP : 1];
#pragma unroll
          for (int q = 0; q < DKP; q++) hk[q] = s_hypp[DCP + q];
          // two training rows in flight per lane (independent sqrt / exp chains)
          for (int js = 0; js < nj; js += 2) {
            const int j0 = 4 * js + t, j1 = j0 + 4;
            const double* x0 = s_Xs + (size_t)j0 * DPC;
            const double* x1 = x0 + 4 * DPC;
            double s0 = 0.0, s1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
            for (int p = 0; p < DPC; p += 2) {
              const double2 u0 = *reinterpret_cast<const double2*>(x0 + p);
              const double2 u1 = *reinterpret_cast<const double2*>(x1 + p);
              if (p < DCP) {
                const double e0 = xs[p] - u0.x, e1 = xs[p] - u1.x;
                s0 += e0 * e0; s1 += e1 * e1;
              } else {
                h0 += (xs[p] != u0.x) ? hk[(p - DCP) < DKP ? (p - DCP) : 0] : 0.0;
                h1 += (xs[p] != u1.x) ? hk[(p - DCP) < DKP ? (p - DCP) : 0] : 0.0;
              }
              if (p + 1 < DCP) {
                const double e0 = xs[p + 1] - u0.y, e1 = xs[p + 1] - u1.y;
                s0 += e0 * e0; s1 += e1 * e1;
              } else {
                h0 += (xs[p + 1] != u0.y) ? hk[(p + 1 - DCP) < DKP ? (p + 1 - DCP) : 0] : 0.0;
                h1 += (xs[p + 1] != u1.y) ? hk[(p + 1 - DCP) < DKP ? (p + 1 - DCP) : 0] : 0.0;
              }
            }
            double k0 = matern_hamming_fast(c, s0, h0), k1 = matern_hamming_fast(c, s1, h1);
            k0 = (j0 < n) ? k0 : 0.0;
            k1 = (j1 < n) ? k1 : 0.0;
            mpart[w] += k0 * s_alpha[j0];
            my_spill[((size_t)js * NW + w) * 32 + lane] = k0;
            if (js + 1 < nj) {
              mpart[w] += k1 * s_alpha[j1];
              my_spill[((size_t)(js + 1) * NW + w) * 32 + lane] = k1;
            }
          }
        }
      } else {
#pragma unroll
        for (int w = 0; w < NW; w++) {
          mpart[w] = 0.0;
          const double* xc = s_cand + (size_t)(wid * WC + w * 8 + g) * d;
          double xs[TLBO_MAX_D];
#pragma unroll
          for (int p = 0; p < TLBO_MAX_D; p++)
            if (p < d) xs[p] = (p < dc) ? xc[p] * s_hyp[8 + p] : xc[p];
          for (int js = 0; js < nj; js++) {
            const int j = 4 * js + t;
            double k = 0.0;
            if (j < n) {
              const double* xj = s_Xs + (size_t)j * d;
              double s = 0.0, hs = 0.0;
#pragma unroll
              for (int p = 0; p < TLBO_MAX_D; p++) {
                if (p < dc) { const double df = xs[p] - xj[p]; s += df * df; }
                else if (p < d) { if (xs[p] != xj[p]) hs += s_hyp[8 + p]; }
              }
              k = matern_hamming(c, s, hs);
              mpart[w] += k * s_alpha[j];
            }
            my_spill[((size_t)js * NW + w) * 32 + lane] = k;
          }
        }
      }
      // ---- V phase: q = |Linv k|^2 via DMMA, Linv streamed in row chunks (resident when `persist`)
      double qpart[NW][2];
#pragma unroll
      for (int w = 0; w < NW; w++) { qpart[w][0] = 0.0; qpart[w][1] = 0.0; }
      for (int r0 = 0; r0 < nib * 8; r0 += a.rc_rows) {
        const int rows = (nib * 8 - r0 < a.rc_rows) ? nib * 8 - r0 : a.rc_rows;
        if (!persist) {
          if (r0 == 0) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
          } else {
            __syncthreads();
            stage_linv(m, r0, rows);
            __syncthreads();
          }
        }
        for (int ib = 0; ib * 8 < rows; ib += 2) {
          const bool two = (ib + 1) * 8 < rows;
          const int ig = (r0 >> 3) + ib;      // global 8-row block index of the first block
          const int jend = two ? 2 * (ig + 2) : 2 * (ig + 1);
          double acc0[NW][2], acc1[NW][2];
#pragma unroll
          for (int w = 0; w < NW; w++) { acc0[w][0] = acc0[w][1] = 0.0; acc1[w][0] = acc1[w][1] = 0.0; }
          const double* rowA0 = s_Linv + (size_t)(ib * 8 + g) * ldl + t;
          const double* rowA1 = rowA0 + (size_t)8 * ldl;
          const int jlim = jend < nj ? jend : nj;
          for (int js = 0; js < jlim; js++) {
            const double a0 = rowA0[4 * js];
            const double a1 = two ? rowA1[4 * js] : 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
              const double b = my_spill[((size_t)js * NW + w) * 32 + lane];
              dmma8x8x4(acc0[w][0], acc0[w][1], a0, b);
              if (two) dmma8x8x4(acc1[w][0], acc1[w][1], a1, b);
            }
          }
#pragma unroll
          for (int w = 0; w < NW; w++) {
            qpart[w][0] += acc0[w][0] * acc0[w][0] + acc1[w][0] * acc1[w][0];
            qpart[w][1] += acc0[w][1] * acc0[w][1] + acc1[w][1] * acc1[w][1];
          }
        }
      }
      // ---- reduce: mean over the 4 lanes of a group, q over the 8 groups
      __syncwarp();
#pragma unroll
      for (int w = 0; w < NW; w++) {
        double mv = mpart[w];
        mv += __shfl_xor_sync(0xffffffffu, mv, 1);
        mv += __shfl_xor_sync(0xffffffffu, mv, 2);
        double q0 = qpart[w][0], q1 = qpart[w][1];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          q0 += __shfl_xor_sync(0xffffffffu, q0, o);
          q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        }
        if (t == 0) my_comb[(w * 8 + g) * 2 + 0] = mv;                 // mean of candidate g
        if (g == 0) { my_comb[(w * 8 + 2 * t) * 2 + 1] = q0; my_comb[(w * 8 + 2 * t + 1) * 2 + 1] = q1; }
      }
      __syncwarp();
      if (lane < WC) {
        const double mean = my_comb[lane * 2 + 0], qq = my_comb[lane * 2 + 1];
        double yv = diag - qq;
        if (yv < 0.0) yv = 0.0;
        const double sd = sqrt(yv);
        double var = sd * sd;
        if (!(var >= VERY_SMALL_NUMBER)) var = (var != var) ? var : VERY_SMALL_NUMBER;
        const double mu_m = mean * ystd + ymean;
        const double var_m = var * (ystd * ystd);
        if (a.pm_mu && myok) { a.pm_mu[(size_t)m * a.N + myrow] = mu_m; a.pm_var[(size_t)m * a.N + myrow] = var_m; }
        combine(first, wm, mu_m, var_m);
      }
      __syncwarp();
    }
    // ---- EI + selection
    if (lane < WC) {
      const long long row = row0 + wid * WC + lane;
      if (row < a.N) {
        double mu, var;
        double f = ens_ei(a.mode, acc, a.eta, a.par, a.var_floor, mu, var);
        if (f < 0.0) negcount += 1.0;
        if (f != f) f = -DBL_MAX;
        if (a.mu) a.mu[row] = mu;
        if (a.var) a.var[row] = var;
        if (a.ei) a.ei[row] = f;
        const long long gidx = a.idx_offset + row;
        bool excl = false;
        int lo = 0, hi = a.n_excl - 1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const long long v = a.excluded[mid];
          if (v == gidx) { excl = true; break; }
          if (v < gidx) lo = mid + 1; else hi = mid - 1;
        }
        if (!excl) {
          Best cnd; cnd.ei = f; cnd.key = a.keys ? a.keys[row] : 0.0; cnd.idx = gidx;
          if (best.idx < 0 || better(cnd, best)) best = cnd;
        }
      }
    }
  }
  // ---- block arg-max
  __shared__ Best s_best[PRED_WARPS];
  __shared__ double s_neg[PRED_WARPS];
  Best wb = best;
  if (wb.idx < 0) { wb.ei = -INFINITY; wb.key = -INFINITY; }
  wb = warp_best(wb);
  const double wn = warp_sum(negcount);
  __syncthreads();
  if (lane == 0) { s_best[wid] = wb; s_neg[wid] = wn; }
  __syncthreads();
  if (tid == 0) {
    Best bb = s_best[0];
    double nn = s_neg[0];
    for (int w = 1; w < PRED_WARPS; w++) { if (better(s_best[w], bb)) bb = s_best[w]; nn += s_neg[w]; }
    double* o = a.partial + (size_t)blockIdx.x * 4;
    o[0] = bb.ei; o[1] = bb.key; o[2] = (double)bb.idx; o[3] = nn;
  }
}

#define ARGMAX_FINAL_THREADS 320
__global__ void __launch_bounds__(ARGMAX_FINAL_THREADS) argmax_final_kernel(const double* partial, int nparts, double* out) {
  // one partial per thread (<= 296 CTAs), warp arg-max, then warp 0 merges the warp results
  __shared__ Best s_b[ARGMAX_FINAL_THREADS / 32];
  __shared__ double s_n[ARGMAX_FINAL_THREADS / 32];
  Best b; b.ei = -INFINITY; b.key = -INFINITY; b.idx = -1;
  double nn = 0.0;
  for (int i = threadIdx.x; i < nparts; i += ARGMAX_FINAL_THREADS) {
    const double2 p0 = *reinterpret_cast<const double2*>(partial + (size_t)i * 4);
    const double2 p1 = *reinterpret_cast<const double2*>(partial + (size_t)i * 4 + 2);
    Best c; c.ei = p0.x; c.key = p0.y; c.idx = (long long)p1.x;
    nn += p1.y;
    if (c.idx >= 0 && (b.idx < 0 || better(c, b))) b = c;
  }
  if (b.idx < 0) { b.ei = -INFINITY; b.key = -INFINITY; }
  b = warp_best(b);
  nn = warp_sum(nn);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_b[wid] = b; s_n[wid] = nn; }
  __syncthreads();
  if (wid == 0) {
    Best c; c.ei = -INFINITY; c.key = -INFINITY; c.idx = -1;
    double n2 = 0.0;
    if (lane < ARGMAX_FINAL_THREADS / 32) { c = s_b[lane]; n2 = s_n[lane]; }
    if (c.idx < 0) { c.ei = -INFINITY; c.key = -INFINITY; }
    c = warp_best(c);
    n2 = warp_sum(n2);
    if (lane == 0) { out[0] = c.ei; out[1] = c.key; out[2] = (double)c.idx; out[3] = n2; }
  }
}

struct FusedPlan { int nw; int rc_rows; size_t smem; int grid; int njmax; int dcp, dkp; int minb; };

// Specialised layouts: (continuous, categorical) column counts are padded up to one of these.
static void fused_pick_layout(int dc, int dk, int* dcp, int* dkp) {
  if (dc >= 1 && dc <= 8 && dk == 0) { *dcp = 8; *dkp = 0; }
  else if (dc >= 1 && dc <= 8 && dk <= 2) { *dcp = 8; *dkp = 2; }
  else if (dc > 8 && dc <= 20 && dk == 0) { *dcp = 20; *dkp = 0; }
  else { *dcp = 0; *dkp = 0; }
}

static size_t fused_fixed_bytes(int nw, int njmax, int dp, int d, int M) {
  const size_t tile = (size_t)PRED_WARPS * nw * 8;
  const size_t nrows = 4 * (size_t)njmax + 4;
  return sizeof(double) * (tile * dp + nrows * dp + nrows + 2 * TLBO_HYP_STRIDE +
                           (size_t)PRED_WARPS * njmax * nw * 32 + (size_t)PRED_WARPS * nw * 8 * 2 +
                           tile * d /* TMA landing buffer */ + 2) +
         (size_t)M * (sizeof(double) + 3 * sizeof(int)) /* active-surrogate list */ + 16;
}

static bool fused_plan(int ncap, int nmax, int d, int dc, int dk, long long N, int M, FusedPlan* pl) {
  const int ldl = TLBO_LDL(ncap);
  if (nmax <= 0 || nmax > ncap) nmax = ncap;
  const int njmax = (nmax + 3) / 4;
  const int nb16 = ((nmax + 15) / 16) * 16;
  fused_pick_layout(dc, dk, &pl->dcp, &pl->dkp);
  const int dp = pl->dcp > 0 ? pl->dcp + pl->dkp : d;
  pl->njmax = njmax;
  // one CTA: 227 KB minus slack; two CTAs per SM: 2 x (dynamic + static (~300 B) + 1 KB reserved) <= 228 KB
  const size_t limit1 = 227 * 1024 - 2048, limit2 = (228 * 1024) / 2 - 1024 - 512;
  // preferred: two CTAs per SM (16 warps hide the sqrt / exp latency chains) with NW = 2
  if (pl->dcp > 0) {
    const size_t fixed = fused_fixed_bytes(2, njmax, dp, d, M);
    if (fixed + sizeof(double) * 16 * ldl <= limit2) {
      size_t rows = ((limit2 - fixed) / (sizeof(double) * ldl) / 16) * 16;
      if (rows > (size_t)nb16) rows = nb16;
      pl->nw = 2; pl->minb = 2; pl->rc_rows = (int)rows;
      pl->smem = fixed + sizeof(double) * rows * ldl;
      const long long ntiles = (N + 127) / 128;
      pl->grid = (int)(ntiles < 296 ? (ntiles < 1 ? 1 : ntiles) : 296);
      return true;
    }
  }
  for (int nw = 4; nw >= 1; nw >>= 1) {
    if (pl->dcp > 0 && nw == 1) break;
    const size_t tile = (size_t)PRED_WARPS * nw * 8;
    const size_t fixed = fused_fixed_bytes(nw, njmax, dp, d, M);
    if (fixed + sizeof(double) * 16 * ldl > limit1) continue;
    size_t rows = ((limit1 - fixed) / (sizeof(double) * ldl) / 16) * 16;
    if (rows > (size_t)nb16) rows = nb16;
    if (rows < 16) continue;
    pl->nw = nw; pl->minb = 1;
    pl->rc_rows = (int)rows;
    pl->smem = fixed + sizeof(double) * rows * ldl;
    const long long ntiles = (N + (long long)tile - 1) / (long long)tile;
    long long grid = ntiles < 148 ? ntiles : 148;
    pl->grid = (int)(grid < 1 ? 1 : grid);
    return true;
  }
  if (pl->dcp > 0) {           // specialised layout does not fit: fall back to the generic one
    pl->dcp = pl->dkp = 0;
    return fused_plan(ncap, nmax, d, -1, -1, N, M, pl);
  }
  return false;
}

extern "C" int64_t tlbo_predict_ei_workspace_bytes(int64_t N) {
  (void)N;
  return (int64_t)((296 * 4 + 8 + 256 + 8) * sizeof(double) + 256);
}

template <int NW, int DCP, int DKP, int MINB>
static int fused_launch(const FusedArgs& a, const FusedPlan& pl, cudaStream_t s) {
  TLBO_CUDA_CHECK(cudaFuncSetAttribute(predict_ei_fused_kernel<NW, DCP, DKP, MINB>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  predict_ei_fused_kernel<NW, DCP, DKP, MINB><<<pl.grid, PRED_THREADS, pl.smem, s>>>(a);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static int fused_entry(const tlbo_pack* pack, const int32_t* h_types, const double* d_w, const int32_t* d_skip,
                       const int32_t* d_order, int mode, const double* d_cand, int64_t N, double eta, double par,
                       double var_floor, const double* d_keys, const int64_t* d_excluded, int n_excl,
                       int64_t idx_offset, double* d_best, double* d_mu, double* d_var, double* d_ei,
                       const double* d_cache_mu, const double* d_cache_var, int n_cached, double* d_model_mu,
                       double* d_model_var, void* d_work, void* stream) {
  if (!pack || N <= 0 || !d_w || !d_cand || !d_best || !d_work) { tlbo_set_error("bad arguments"); return TLBO_E_BADARG; }
  if (pack->ncap % 8 != 0) { tlbo_set_error("pack ncap must be a multiple of 8"); return TLBO_E_BADARG; }
  FusedArgs a;
  a.pack = *pack;
  int rc = tlbo_make_space(pack->d, h_types, &a.sp);
  if (rc) return rc;
  FusedPlan pl;
  if (pack->M <= 0 || pack->M > 4096) { tlbo_set_error("pack M out of range for the fused kernel"); return TLBO_E_TOOLARGE; }
  if (!fused_plan(pack->ncap, pack->nmax, pack->d, a.sp.d_cont, a.sp.d_cat, N, pack->M, &pl)) {
    tlbo_set_error("surrogates too large for the fused kernel");
    return TLBO_E_TOOLARGE;
  }
  a.w = d_w; a.skip = d_skip; a.order = d_order; a.mode = mode; a.cand = d_cand; a.N = N;
  a.eta = eta; a.par = par; a.var_floor = var_floor;
  a.keys = d_keys; a.excluded = (const long long*)d_excluded; a.n_excl = d_excluded ? n_excl : 0;
  a.idx_offset = idx_offset; a.mu = d_mu; a.var = d_var; a.ei = d_ei;
  a.cmu = (d_cache_mu && d_cache_var && n_cached > 0) ? d_cache_mu : nullptr;
  a.cvar = d_cache_var; a.ncached = n_cached;
  a.pm_mu = (d_model_mu && d_model_var) ? d_model_mu : nullptr; a.pm_var = d_model_var;
  a.partial = (double*)d_work; a.rc_rows = pl.rc_rows; a.njmax = pl.njmax;
  cudaStream_t s = (cudaStream_t)stream;
  if (pl.dcp == 8 && pl.dkp == 2) rc = (pl.nw == 2) ? fused_launch<2, 8, 2, 2>(a, pl, s) : fused_launch<4, 8, 2, 1>(a, pl, s);
  else if (pl.dcp == 8 && pl.dkp == 0) rc = (pl.nw == 2) ? fused_launch<2, 8, 0, 2>(a, pl, s) : fused_launch<4, 8, 0, 1>(a, pl, s);
  else if (pl.dcp == 20) rc = (pl.nw == 2) ? fused_launch<2, 20, 0, 2>(a, pl, s) : fused_launch<4, 20, 0, 1>(a, pl, s);
  else if (pl.nw == 4) rc = fused_launch<4, 0, 0, 1>(a, pl, s);
  else if (pl.nw == 2) rc = fused_launch<2, 0, 0, 1>(a, pl, s);
  else rc = fused_launch<1, 0, 0, 1>(a, pl, s);
  if (rc) return rc;
  argmax_final_kernel<<<1, ARGMAX_FINAL_THREADS, 0, s>>>(a.partial, pl.grid, d_best);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Merge of `nparts` (ei, key, row, #negative) quadruples with the rule of the fused kernel's own final stage:
// the exchange step of the row-sharded candidate pool (the quadruples all-gathered from the ranks).
extern "C" int tlbo_merge_best(const double* d_quads, int nparts, double* d_best, void* stream) {
  if (!d_quads || !d_best || nparts <= 0) { tlbo_set_error("bad arguments"); return TLBO_E_BADARG; }
  argmax_final_kernel<<<1, ARGMAX_FINAL_THREADS, 0, (cudaStream_t)stream>>>(d_quads, nparts, d_best);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int tlbo_predict_ei_fused(const tlbo_pack* pack, const int32_t* h_types, const double* d_w,
                                     const int32_t* d_skip, const int32_t* d_order, int mode,
                                     const double* d_cand, int64_t N,
                                     double eta, double par, double var_floor, const double* d_keys,
                                     const int64_t* d_excluded, int n_excl, int64_t idx_offset, double* d_best,
                                     double* d_mu, double* d_var, double* d_ei, void* d_work, void* stream) {
  return fused_entry(pack, h_types, d_w, d_skip, d_order, mode, d_cand, N, eta, par, var_floor, d_keys, d_excluded,
                     n_excl, idx_offset, d_best, d_mu, d_var, d_ei, nullptr, nullptr, 0, nullptr, nullptr, d_work,
                     stream);
}

extern "C" int tlbo_predict_ei_fused_cached(const tlbo_pack* pack, const int32_t* h_types, const double* d_w,
                                            const int32_t* d_skip, const int32_t* d_order, int mode,
                                            const double* d_cand, int64_t N, double eta, double par,
                                            double var_floor, const double* d_keys, const int64_t* d_excluded,
                                            int n_excl, int64_t idx_offset, double* d_best, double* d_mu,
                                            double* d_var, double* d_ei, const double* d_cache_mu,
                                            const double* d_cache_var, int n_cached, void* d_work, void* stream) {
  return fused_entry(pack, h_types, d_w, d_skip, d_order, mode, d_cand, N, eta, par, var_floor, d_keys, d_excluded,
                     n_excl, idx_offset, d_best, d_mu, d_var, d_ei, d_cache_mu, d_cache_var, n_cached, nullptr,
                     nullptr, d_work, stream);
}

extern "C" int tlbo_predict_pool_posteriors(const tlbo_pack* pack, const int32_t* h_types, const double* d_cand,
                                            int64_t N, double* d_model_mu, double* d_model_var, void* d_work,
                                            void* stream) {
  if (!pack || !d_model_mu || !d_model_var || !d_work) { tlbo_set_error("bad arguments"); return TLBO_E_BADARG; }
  // ensemble outputs are not wanted: unit weights, no exclusion; the arg-max result lands in the
  // tail of the workspace and is ignored
  double* scratch = (double*)d_work;
  double* d_w = scratch + 296 * 4 + 8;
  double* d_best = d_w + 256;
  if (pack->M > 256) { tlbo_set_error("too many surrogates for tlbo_predict_pool_posteriors"); return TLBO_E_TOOLARGE; }
  TLBO_CUDA_CHECK(cudaMemsetAsync(d_w, 0, sizeof(double) * 256, (cudaStream_t)stream));
  return fused_entry(pack, h_types, d_w, nullptr, nullptr, 0, d_cand, N, 0.0, 0.0, 0.0, nullptr, nullptr, 0, 0, d_best,
                     nullptr, nullptr, nullptr, nullptr, nullptr, 0, d_model_mu, d_model_var, d_work, stream);
}
