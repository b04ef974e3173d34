// MCMC marginalisation of the GP hyper-parameters on device (BASELINE config 5).
//
// Replaces reference tlbo/model/gp_mcmc.py:116-282 (GaussianProcessMCMC._train: an
// emcee.EnsembleSampler over the kernel hyper-parameters, 250 burn-in + 250 chain steps),
// :294-332 (_ll: sklearn GaussianProcessRegressor.log_marginal_likelihood(theta) + priors incl.
// the Tophat bound priors of base_gp.py:92-132 / gp_base_prior.py:155-240) and :384-440
// (_predict: mean of the walker means, var(means) + mean(vars)).
//
// emcee is not vendored by the reference (requirements.txt:16, unpinned); its published
// algorithm is the affine-invariant stretch move of Goodman & Weare (2010) as implemented by
// emcee 3.x (`EnsembleSampler.sample` -> `RedBlueMove.propose` -> `StretchMove.get_proposal`):
// per step the walkers are split into two random halves; each walker s of the active half
// proposes  q = c - (c - s) * zz  against a random walker c of the other half with
// zz = ((a - 1) u + 1)^2 / a, a = 2, and is accepted iff
// log u' < (ndim - 1) log zz + lp(q) - lp(s).  EVERY random number of the chain (split
// permutation, u, partner index, u') is independent of the chain state, so the host draws the
// whole table up front from the legacy MT19937 stream in emcee's order and the chain runs here
// with zero host round trips.
//
// One persistent CTA per (surrogate, active-half walker): a half-step is one log-probability
// evaluation (kernel matrix + register-panel elimination of gp_nll.cuh, no gradient) followed by
// a barrier among the W/2 CTAs of that surrogate (global atomic counter; the kernel is launched
// cooperatively so that all CTAs are co-resident).  Latency-bound like the MAP fit: the figure of
// merit is microseconds per half-step, not a roofline fraction.
#undef TLBO_PHASE_TIMING
#include "gp_nll.cuh"

struct McmcArgs {
  const double* X; const double* y; const int32_t* n;
  int B, W, Ns, ncap, d, nhalf;
  SpaceDesc sp;
  double* pos; double* lp;                       // [B, W, H], [B, W]
  const int32_t* sidx; const int32_t* cidx;      // [NT, nhalf, Ns]
  const double* zz; const double* fac; const double* logu;
  const int32_t* table_of;                       // [B] or NULL
  const double* lo_raw; const double* hi_raw;    // [H] Tophat bounds, original scale
  const double* theta0;                          // [H] or NULL
  unsigned int* barrier;                         // [B]
  int32_t* naccept; int32_t* status;
  double* panels; size_t panel_stride; int cache;
};

__device__ __forceinline__ void mcmc_barrier(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(ctr) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// GaussianProcessMCMC._ll (gp_mcmc.py:294-332) at th[0..H) (shared memory; clamped in place to
// [-50, 50] as the reference clamps the proposal it is handed).  Uniform return value.
template <int T>
__device__ double mcmc_logprob(const FitWs& S, const SpaceDesc& sp, double* th, int H, const double* lo_raw,
                               const double* hi_raw) {
  const int tid = threadIdx.x, n = S.n;
  __syncthreads();
  if (tid < H) {
    double v = th[tid];
    if (v < -50.0) v = -50.0;
    if (v > 50.0) v = 50.0;
    th[tid] = v;
  }
  __syncthreads();
  // Tophat bound priors (gp_base_prior.py:196-211 via Prior.lnprob :35-51): -inf outside.  A
  // proposal outside the box has log-probability -inf whatever the likelihood is, so the
  // factorisation is skipped for it (same result: finite + -inf, -inf + -inf, nan -> -inf).
  bool inside = true;
  if (tid < H) {
    const double t = th[tid], v = exp(t);
    inside = !(v < lo_raw[tid] || v > hi_raw[tid]) && (t == t);
  }
  if (!__syncthreads_and(inside ? 1 : 0)) return -INFINITY;
  double nll_data;
  if constexpr (T > 0) {
    if (!lml_eliminate_reg<T>(S, sp, th, &nll_data)) return -INFINITY;      // sklearn: LinAlgError -> -inf
  } else {
    if (!build_and_eliminate(S, sp, th, false)) return -INFINITY;
    double part = 0.0;
    for (int k = tid; k < n; k += blockDim.x) {
      const double z = S.A[(size_t)n * S.lda + k];
      part += 0.5 * z * z + 0.5 * log(S.piv[k]);
    }
    nll_data = block_sum(part, S.red);
  }
  double lml = -nll_data - 0.5 * (double)n * LOG_2PI;
  // priors in the reference's order: [lognormal, tophat] on c, [tophat] per length scale,
  // [horseshoe(0.1), tophat] on the noise (base_gp.py:92-132); the tophats contribute 0 here
  {
    const double v0 = exp(th[0]);
    double lp0;
    if (v0 <= 0.0) lp0 = -1e25;
    else { const double lv = log(v0); lp0 = -(lv * lv) / 2.0 - log(2.5066282746310002 * v0); }
    lml += lp0;
    const double vn = exp(th[H - 1]);
    const double s2 = 0.1 * 0.1;
    double lpn;
    if (vn == 0.0) lpn = INFINITY;
    else lpn = log(log(1.0 + 3.0 * (s2 / (vn * vn))) + VERY_SMALL_NUMBER);
    lml += lpn;
  }
  return isfinite(lml) ? lml : -INFINITY;
}

template <int T>
__global__ void __launch_bounds__(FIT_THREADS, 1) gp_mcmc_kernel(McmcArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x / a.Ns, k = blockIdx.x - b * a.Ns;
  const int n = a.n[b], d = a.d, H = d + 2, W = a.W;
  const int tid = threadIdx.x;
  __shared__ double th[TLBO_HMAX];
  __shared__ double s_flag[2];
  FitWs S;
  carve_ws(S, smem, a.panels ? a.panels + (size_t)blockIdx.x * a.panel_stride : nullptr, n, a.ncap, d, T, a.cache);
  load_problem(S, a.sp, a.X + (size_t)b * a.ncap * d, a.y + (size_t)b * a.ncap, a.ncap);
  if constexpr (T > 0) find_active_columns(S, a.sp);     // build_and_eliminate_reg works on the varying columns
  __syncthreads();
  double* pos = a.pos + (size_t)b * W * H;
  double* lp = a.lp + (size_t)b * W;
  unsigned int* bar = a.barrier + b;
  unsigned int phase = 0;
  if (k == 0 && a.theta0) {
    // gp_mcmc.py:141: self.gp.fit(X, y) at the kernel's default theta; a LinAlgError there is not
    // caught by the reference -> reported through status
    if (tid < H) th[tid] = a.theta0[tid];
    __syncthreads();
    const bool ok = build_and_eliminate_any<T>(S, a.sp, th, false);
    if (tid == 0) a.status[b] = ok ? 0 : 1;
  }
  // emcee run_mcmc: log-probability of the initial ensemble
  for (int w = k; w < W; w += a.Ns) {
    __syncthreads();
    if (tid < H) th[tid] = pos[(size_t)w * H + tid];
    const double l = mcmc_logprob<T>(S, a.sp, th, H, a.lo_raw, a.hi_raw);
    if (tid == 0) lp[w] = l;
  }
  mcmc_barrier(bar, (++phase) * (unsigned)a.Ns);
  const size_t tab = (size_t)(a.table_of ? a.table_of[b] : 0) * a.nhalf * a.Ns;
  // the random tables do not depend on the chain state: the entries of half-step h + 1 are fetched
  // before the barrier that ends half-step h, off the critical path
  int s_nx = 0, c_nx = 0;
  double zz_nx = 0.0, fac_nx = 0.0, logu_nx = 0.0;
  if (a.nhalf > 0) {
    const size_t off = tab + k;
    s_nx = a.sidx[off]; c_nx = a.cidx[off]; zz_nx = a.zz[off]; fac_nx = a.fac[off]; logu_nx = a.logu[off];
  }
  for (int h = 0; h < a.nhalf; h++) {
    const int s = s_nx, c = c_nx;
    const double zz = zz_nx, fac = fac_nx, logu = logu_nx;
    if (tid < H) {
      // StretchMove.get_proposal: c - (c - s) * zz, three separately rounded operations
      const double cs = __ldcg(pos + (size_t)c * H + tid), ss = __ldcg(pos + (size_t)s * H + tid);
      th[tid] = __dsub_rn(cs, __dmul_rn(__dsub_rn(cs, ss), zz));
    }
    if (h + 1 < a.nhalf) {
      const size_t off = tab + (size_t)(h + 1) * a.Ns + k;
      s_nx = a.sidx[off]; c_nx = a.cidx[off]; zz_nx = a.zz[off]; fac_nx = a.fac[off]; logu_nx = a.logu[off];
    }
    const double lnew = mcmc_logprob<T>(S, a.sp, th, H, a.lo_raw, a.hi_raw);
    if (tid == 0) {
      // RedBlueMove.propose: lnpdiff = factors + new_log_probs - old_log_probs; accept iff
      // log(u) < lnpdiff  (nan compares false)
      const double lold = __ldcg(lp + s);
      const double diff = __dsub_rn(__dadd_rn(fac, lnew), lold);
      s_flag[1] = (logu < diff) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (s_flag[1] != 0.0) {
      if (tid < H) pos[(size_t)s * H + tid] = th[tid];
      if (tid == 0) { lp[s] = lnew; a.naccept[(size_t)b * W + s] += 1; }
    }
    mcmc_barrier(bar, (++phase) * (unsigned)a.Ns);
  }
}

// ---------------------------------------------------------------------------------------
// gp_mcmc.py:418-440: m = mean_w mu_w, v = var_w(mu_w) + mean_w(var_w), clipped at machine eps.
// Sums run over the walkers in order with separately rounded operations, as numpy's
// add.reduce over axis 0 does.
__global__ void __launch_bounds__(256) mcmc_marginalize_kernel(const double* __restrict__ mu,
                                                               const double* __restrict__ var, int G, int W,
                                                               long long N, double* __restrict__ gmu,
                                                               double* __restrict__ gvar) {
  // two passes over the walkers' means (the second one hits L2): measured faster than keeping the W
  // values in registers (138 registers, one CTA per SM)
  const long long total = (long long)G * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long g = idx / N, j = idx - g * N;
    const double* pm = mu + (size_t)g * W * N + j;
    const double* pv = var + (size_t)g * W * N + j;
    double sm = pm[0], sv = pv[0];
    for (int w = 1; w < W; w++) {
      sm = __dadd_rn(sm, pm[(size_t)w * N]);
      sv = __dadd_rn(sv, pv[(size_t)w * N]);
    }
    const double m = sm / (double)W;
    const double d0 = __dsub_rn(pm[0], m);
    double ss = __dmul_rn(d0, d0);
    for (int w = 1; w < W; w++) {
      const double dw = __dsub_rn(pm[(size_t)w * N], m);
      ss = __dadd_rn(ss, __dmul_rn(dw, dw));
    }
    double v = __dadd_rn(ss / (double)W, sv / (double)W);
    const double eps = 2.220446049250313e-16;
    if (v < eps) v = eps;            // np.clip(v, eps, inf); nan stays nan
    gmu[idx] = m;
    gvar[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------
static const size_t kSmemLimitMcmc = 227 * 1024;

struct McmcLayout { int tile, cache; bool global_panels; size_t smem, panel_stride; };

static McmcLayout plan_mcmc(int ncap, int d) {
  McmcLayout w;
  const size_t panel = fit_panel_bytes(ncap);
  int tile = panel_tile_for(ncap);
  int cache = 0;                       // no gradient: the pair-value caches are not needed
  size_t small = fit_small_bytes(ncap, d, tile, cache);
  if (tile > 0 && small + panel + 1024 > kSmemLimitMcmc) { tile = 0; small = fit_small_bytes(ncap, d, 0, 0); }
  w.tile = tile; w.cache = cache;
  w.global_panels = (small + panel + 1024) > kSmemLimitMcmc;
  w.smem = small + (w.global_panels ? 0 : panel);
  w.panel_stride = panel / sizeof(double);
  return w;
}

extern "C" int64_t tlbo_gp_mcmc_workspace_bytes(int B, int W, int ncap, int d) {
  McmcLayout w = plan_mcmc(ncap, d);
  size_t bytes = (size_t)B * sizeof(unsigned int) + 256;
  if (w.global_panels) bytes += (size_t)B * (W / 2) * w.panel_stride * 8;
  return (int64_t)bytes;
}

template <int T>
static int launch_mcmc(McmcArgs& a, const McmcLayout& w, int b0, int nb, cudaStream_t s) {
  // one cooperative launch per group of surrogates whose CTAs are co-resident
  McmcArgs g = a;
  g.X = a.X + (size_t)b0 * a.ncap * a.d; g.y = a.y + (size_t)b0 * a.ncap; g.n = a.n + b0; g.B = nb;
  const int H = a.d + 2;
  g.pos = a.pos + (size_t)b0 * a.W * H; g.lp = a.lp + (size_t)b0 * a.W;
  g.table_of = a.table_of ? a.table_of + b0 : nullptr;
  g.barrier = a.barrier + b0; g.naccept = a.naccept + (size_t)b0 * a.W; g.status = a.status + b0;
  g.panels = a.panels ? a.panels + (size_t)b0 * a.Ns * w.panel_stride : nullptr;
  void* params[] = {&g};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)gp_mcmc_kernel<T>, dim3(nb * a.Ns), dim3(FIT_THREADS),
                                              params, w.smem, s);
  if (e != cudaSuccess) { tlbo_set_error(cudaGetErrorString(e)); return (int)e; }
  return 0;
}

template <int T>
static int run_mcmc(McmcArgs& a, const McmcLayout& w, cudaStream_t s) {
  if (w.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gp_mcmc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem);
    if (e != cudaSuccess) { tlbo_set_error(cudaGetErrorString(e)); return (int)e; }
  }
  int dev = 0, sms = 0, per_sm = 0;
  TLBO_CUDA_CHECK(cudaGetDevice(&dev));
  TLBO_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TLBO_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gp_mcmc_kernel<T>, FIT_THREADS, w.smem));
  const int resident = sms * per_sm;
  if (resident < a.Ns) { tlbo_set_error("walker half does not fit on the device"); return TLBO_E_TOOLARGE; }
  const int per_launch = resident / a.Ns;
  for (int b0 = 0; b0 < a.B; b0 += per_launch) {
    const int nb = (a.B - b0 < per_launch) ? a.B - b0 : per_launch;
    const int rc = launch_mcmc<T>(a, w, b0, nb, s);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int tlbo_gp_mcmc_chain(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap, int d,
                                  const int32_t* h_types, int W, int nsteps, double* d_pos, double* d_lp,
                                  const int32_t* d_sidx, const int32_t* d_cidx, const double* d_zz,
                                  const double* d_factors, const double* d_logu, const int32_t* d_table_of,
                                  const double* d_prior_lo, const double* d_prior_hi, const double* d_theta0,
                                  int32_t* d_naccept, int32_t* d_status, void* d_work, void* stream) {
  if (B <= 0 || ncap <= 0 || d <= 0 || d > TLBO_MAX_D || W < 2 || (W & 1) || nsteps < 0) {
    tlbo_set_error("bad B / ncap / d / W (even, >= 2) / nsteps");
    return TLBO_E_BADARG;
  }
  if (!d_X || !d_y || !d_n || !d_pos || !d_lp || !d_prior_lo || !d_prior_hi || !d_naccept || !d_status || !d_work ||
      (nsteps > 0 && (!d_sidx || !d_cidx || !d_zz || !d_factors || !d_logu))) {
    tlbo_set_error("NULL argument");
    return TLBO_E_BADARG;
  }
  SpaceDesc sp;
  int rc = tlbo_make_space(d, h_types, &sp);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  McmcLayout w = plan_mcmc(ncap, d);
  McmcArgs a;
  a.X = d_X; a.y = d_y; a.n = d_n; a.B = B; a.W = W; a.Ns = W / 2; a.ncap = ncap; a.d = d; a.nhalf = 2 * nsteps;
  a.sp = sp; a.pos = d_pos; a.lp = d_lp; a.sidx = d_sidx; a.cidx = d_cidx; a.zz = d_zz; a.fac = d_factors;
  a.logu = d_logu; a.table_of = d_table_of; a.lo_raw = d_prior_lo; a.hi_raw = d_prior_hi; a.theta0 = d_theta0;
  a.naccept = d_naccept; a.status = d_status; a.cache = w.cache; a.panel_stride = w.panel_stride;
  unsigned char* wp = (unsigned char*)d_work;
  a.barrier = (unsigned int*)wp;
  wp += (((size_t)B * sizeof(unsigned int)) + 255) & ~(size_t)255;
  a.panels = w.global_panels ? (double*)wp : nullptr;
  TLBO_CUDA_CHECK(cudaMemsetAsync(a.barrier, 0, (size_t)B * sizeof(unsigned int), s));
  TLBO_CUDA_CHECK(cudaMemsetAsync(d_status, 0, (size_t)B * sizeof(int32_t), s));
  switch (w.tile) {
    case 4: return run_mcmc<4>(a, w, s);
    case 5: return run_mcmc<5>(a, w, s);
    case 6: return run_mcmc<6>(a, w, s);
    case 7: return run_mcmc<7>(a, w, s);
    case 8: return run_mcmc<8>(a, w, s);
    default: return run_mcmc<0>(a, w, s);
  }
}

extern "C" int tlbo_mcmc_marginalize(const double* d_mu, const double* d_var, int G, int W, int64_t N, double* d_gmu,
                                     double* d_gvar, void* stream) {
  if (G <= 0 || W <= 0 || N < 0 || !d_mu || !d_var || !d_gmu || !d_gvar) {
    tlbo_set_error("bad G / W / N or NULL argument");
    return TLBO_E_BADARG;
  }
  if (N == 0) return 0;
  const long long total = (long long)G * N;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mcmc_marginalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_mu, d_var, G, W, (long long)N, d_gmu,
                                                                               d_gvar);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}
