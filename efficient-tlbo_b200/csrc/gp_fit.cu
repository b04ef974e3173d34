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
#include "gp_nll.cuh"
#include "lbfgsb_dense.cuh"
#include "lbfgsb_warp.cuh"


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
  if (T > 0 && n + 1 > 16 * T) {
    // capacity without room for the y row (see panel_tile_for): no restart result -> the factorisation
    // reports status 2 for this surrogate
    if (tid == 0) {
      a.restart_f[(size_t)b * a.R + r] = INFINITY;
      int32_t* info = a.restart_info + ((size_t)b * a.R + r) * 4;
      info[0] = 0; info[1] = 0; info[2] = LB_ERROR; info[3] = 0;
    }
    return;
  }
  LbfgsbState* st = reinterpret_cast<LbfgsbState*>(smem);
  LbwState* sw = reinterpret_cast<LbwState*>(smem);
  unsigned char* rest = smem + fit_state_bytes();
  FitWs S;
  carve_ws(S, rest, a.panels ? a.panels + (size_t)blockIdx.x * a.panel_stride : nullptr, n, a.ncap, d, T, a.cache);
  load_problem(S, a.sp, a.X + (size_t)b * a.ncap * d, a.y + (size_t)b * a.ncap, a.ncap);
  if (tid < H) th[tid] = a.p0[(size_t)r * H + tid];
  if constexpr (T > 0) find_active_columns(S, a.sp);
  __syncthreads();
  // ONE evaluation call site (the unrolled evaluation is ~100 KB of code; a second copy costs
  // instruction-cache misses on the critical path): the evaluation at the start point doubles as
  // the positive-definiteness probe of gp.py:127-140 -- the initial GaussianProcessRegressor.fit at
  // the default theta retries with noise += 1 on LinAlgError (restart 0 only; bumped evaluations do
  // not count as function evaluations), and the bumped theta then becomes start point 0.
  int bumps = 0, nfev = 0;
  if (a.do_optimize) {
    if (WARP_OPT) { if (wid == 0) lbw_init(*sw, H, th, a.lo, a.hi, lane); }
    else { if (tid == 0) lbfgsb_init(*st, H, th, a.lo, a.hi); }
    __syncthreads();
  }
  for (;;) {
    if (a.do_optimize && tid < H) th[tid] = WARP_OPT ? sw->x[tid] : st->x[tid];
    __syncthreads();
    const bool ok = nll_eval_any<T, DMAX>(S, a.sp, th, fg);
    if (r == 0 && !ok && nfev == 0 && bumps < 10) {
      if (tid == 0) th[H - 1] = log(exp(th[H - 1]) + 1.0);
      bumps++;
      __syncthreads();
      if (a.do_optimize) {
        if (WARP_OPT) { if (wid == 0) lbw_init(*sw, H, th, a.lo, a.hi, lane); }
        else { if (tid == 0) lbfgsb_init(*st, H, th, a.lo, a.hi); }
        __syncthreads();
      }
      continue;
    }
    nfev++;
    if (!a.do_optimize) {
      if (tid < H) a.restart_theta[((size_t)b * a.R + r) * H + tid] = th[tid];
      if (tid == 0) {
        a.restart_f[(size_t)b * a.R + r] = fg[0];
        int32_t* info = a.restart_info + ((size_t)b * a.R + r) * 4;
        info[0] = nfev; info[1] = 0; info[2] = -1; info[3] = bumps;
      }
      return;
    }
    PHASE_T0();
    if (WARP_OPT) {
      if (wid == 0) {
        const int rc = lbw_feed<DMAX>(*sw, fg[0], fg + 1, lane);
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
  if (T > 0 && n + 1 > 16 * T) {      // capacity without room for the y row (see panel_tile_for)
    if (tid == 0) { a.nll[b] = NAN; a.status[b] = 3; }
    return;
  }
  FitWs S;
  carve_ws(S, smem, a.panels ? a.panels + (size_t)b * a.panel_stride : nullptr, n, a.ncap, d, T, a.cache);
  load_problem(S, a.sp, a.X + (size_t)b * a.ncap * d, a.y + (size_t)b * a.ncap, a.ncap);
  if (tid < H) th[tid] = a.theta[(size_t)b * H + tid];
  if constexpr (T > 0) find_active_columns(S, a.sp);
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
#ifdef TLBO_FIT_FEW_TILES   /* development builds: one register-panel instantiation (fast SASS inspection) */
#define TLBO_LAUNCH_FIT_T(DM)                                                \
  do {                                                                       \
    if (tile == 4) TLBO_LAUNCH_FIT(DM, true, 4);                             \
    else { tlbo_set_error("development build: tile 4 only"); return TLBO_E_BADARG; } \
  } while (0)
#else
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
#endif
  const int tile = w.tile;
  // DMAX also sizes the unrolled loops of the warp optimiser: H = d + 2 <= DMAX
  if (g_fit_variant != 0) {
    if (d <= 10) TLBO_LAUNCH_FIT_T(12); else TLBO_LAUNCH_FIT_T(24);
  } else {
    if (d <= 10) TLBO_LAUNCH_FIT(12, false, 0); else TLBO_LAUNCH_FIT(24, false, 0);
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
#ifdef TLBO_FIT_FEW_TILES
#define TLBO_LAUNCH_NLL_T(DM)                                                \
  do {                                                                       \
    if (tile == 4) TLBO_LAUNCH_NLL(DM, 4);                                   \
    else { tlbo_set_error("development build: tile 4 only"); return TLBO_E_BADARG; } \
  } while (0)
#else
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
#endif
  const int tile = w.tile;
  if (d <= 12) TLBO_LAUNCH_NLL_T(12); else TLBO_LAUNCH_NLL_T(24);
#undef TLBO_LAUNCH_NLL_T
#undef TLBO_LAUNCH_NLL
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}
