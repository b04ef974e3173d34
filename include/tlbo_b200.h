/* tlbo_b200 -- C ABI of the B200-native surrogate-and-acquisition hot path.
 *
 * The reference (thomas-young-2013/efficient-tlbo) is pure Python and has no
 * FFI; its boundary for this path is the duck-typed Python surrogate / facade /
 * acquisition interface.  This header declares the C entry
 * points that the Python mirror of that interface (efficient-tlbo_b200/tlbo_b200)
 * binds with ctypes, each citing the reference call site(s) it replaces.
 * INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every array is dense row-major float64 unless
 *    stated; pointers named d_* are DEVICE pointers (current CUDA device), pointers
 *    named h_* are HOST pointers.
 *  - `stream` is a cudaStream_t passed as void*; every entry point only enqueues
 *    work on that stream (no implicit synchronisation) unless stated.
 *  - return value: 0 on success, a positive cudaError_t, or a negative TLBO_E_*.
 *    tlbo_last_error() returns a static message for the calling thread.
 *  - input columns follow the reference's encoding
 *    (tlbo/config_space/util.py:11-30): `h_types[d]` is the vector produced by
 *    get_types (tlbo/model/util_funcs.py:14-55): 0 = continuous / integer /
 *    constant column, >0 = categorical with that many choices.
 *  - theta layout (log space) = the reference kernel tree's
 *    (tlbo/model/model_builder.py:28-72): [log c, log l_cont.., log l_cat.., log s2],
 *    H = d + 2.
 */
#ifndef TLBO_B200_H
#define TLBO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLBO_MAX_D 22          /* H = d + 2 <= 24 */
#define TLBO_E_BADARG (-1)
#define TLBO_E_TOOLARGE (-2)
/* h_types[d] everywhere below: 0 = numerical column, k > 0 = categorical with k choices (tlbo/model/util_funcs.py:14-55
 * get_types).  Bit 30 of h_types[0] set: the configuration space has conditions (tlbo/model/base_gp.py:134-146
 * _set_has_conditions), and every kernel value k(x, x') is multiplied by the activity mask of
 * tlbo/model/gp_kernels.py:21-29 -- 1 iff x and x' have the same set of columns <= -1 (inactive). */
#define TLBO_TYPES_HAS_CONDITIONS 0x40000000

const char* tlbo_last_error(void);
int tlbo_version(void);
/* cudaStreamSynchronize(stream): lets a host binding wait for the stream it launched on without going through
 * its framework's stream objects (the online local search waits a few hundred times per BO iteration). */
int tlbo_stream_synchronize(void* stream);

/* A batch ("pack") of fitted GP surrogates laid out for the predict kernels.
 * All arrays are device memory owned by the caller; per-model arrays are strided by
 * `ncap` rows (ncap >= max n, multiple of 8).  Columns of Xs are PERMUTED: the
 * continuous columns first (scaled by 1/l), then the categorical ones (raw codes).
 *   hyp row (TLBO_HYP_STRIDE doubles): [c, s2, y_mean, y_std, c+s2, 0,0,0,
 *        then per permuted column p: cont -> 1/l_p, cat -> 1/(2 l_p^2)]
 *   Linv rows are strided by TLBO_LDL(ncap) = ncap + 4 doubles (bank-conflict-free
 *   fragment loads for the FP64 mma in the fused predict kernel).                  */
#define TLBO_HYP_STRIDE 32
#define TLBO_LDL(ncap) ((ncap) + 4)
typedef struct tlbo_pack {
  int32_t M;            /* number of surrogates                                   */
  int32_t ncap;         /* row capacity / stride                                  */
  int32_t d;            /* input columns                                          */
  int32_t nmax;         /* hint: largest n among the slots in use (0 = unknown -> ncap);
                           sizes the shared-memory plan of the fused predict kernel        */
  int32_t* d_n;         /* [M]            observations per surrogate              */
  double* d_Xs;         /* [M, ncap, d]   permuted + scaled training inputs       */
  double* d_hyp;        /* [M, TLBO_HYP_STRIDE]                                   */
  double* d_alpha;      /* [M, ncap]      K^-1 y                                  */
  double* d_Linv;       /* [M, ncap, TLBO_LDL(ncap)] row-major lower-triangular L^-1 */
  double* d_theta;      /* [M, d+2]       fitted log hyper-parameters             */
} tlbo_pack;

/* Batched MAP fit of B independent GP surrogates -- replaces
 * GaussianProcess._train/_nll/_optimize (tlbo/model/gp.py:98-248) and the skopt /
 * sklearn GaussianProcessRegressor.fit + log_marginal_likelihood they call, for every
 * surrogate of build_source_surrogates / build_single_surrogate
 * (tlbo/facade/base_facade.py:58-108) at once.
 *   d_X   [B, ncap, d]  raw inputs, reference column order (non-finite already -> -1)
 *   d_y   [B, ncap]     targets AFTER the reference's y normalisation
 *   d_n   [B]           rows in use
 *   d_p0  [R, H]        L-BFGS-B start points (row 0 = kernel default theta; rows 1..
 *                       = the sampled restarts of gp.py:211-239), shared by all B
 *   d_lo, d_hi [H]      log-space bounds
 *   h_ymean, h_ystd [B] BaseGP._normalize_y statistics (base_gp.py:48-64), stored in the pack
 *   do_optimize         0 -> keep p0 row 0 (gp.py do_optimize=False)
 * Outputs: pack slots slot0 .. slot0+B-1 are filled (n, Xs, hyp, alpha, Linv, theta);
 *   d_restart_theta [B, R, H], d_restart_f [B, R], d_restart_info [B, R, 4] int32
 *   (nfev, nit, status, noise bumps) and d_status [B] int32 (0 ok, 1 final Cholesky
 *   failed -> the reference would raise LinAlgError, 2 no finite optimum) may be NULL
 *   except d_status.  d_work: scratch of tlbo_gp_fit_workspace_bytes() bytes.       */
int tlbo_gp_fit_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap, int d,
                        const int32_t* h_types, const double* d_p0, int R, const double* d_lo,
                        const double* d_hi, const double* h_ymean, const double* h_ystd, int do_optimize,
                        const tlbo_pack* pack, int slot0, double* d_restart_theta, double* d_restart_f,
                        int32_t* d_restart_info, int32_t* d_status, void* d_work, void* stream);
int64_t tlbo_gp_fit_workspace_bytes(int B, int R, int ncap, int d);
/* Selects the L-BFGS-B implementation used by tlbo_gp_fit_batched: 1 (default) = warp-cooperative
 * (csrc/lbfgsb_warp.cuh), 0 = the single-thread statement of the same algorithm
 * (csrc/lbfgsb_dense.cuh, the one cross-checked against scipy on the host).  Returns the
 * previous value.  Both follow scipy.optimize.fmin_l_bfgs_b (reference gp.py:241-248).     */
int tlbo_set_fit_variant(int variant);

/* Factorise B surrogates at GIVEN theta (skopt GaussianProcessRegressor.fit with
 * optimizer=None: L = chol(K), alpha_ = K^-1 y, K_inv_ = L^-T L^-1) and fill pack slots.
 * Optional dense outputs for parity checks: d_L [B, ncap, ncap], d_Kinv [B, ncap, ncap]. */
int tlbo_gp_factorize_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap,
                              int d, const int32_t* h_types, const double* d_theta, const double* h_ymean,
                              const double* h_ystd, const tlbo_pack* pack, int slot0, double* d_L,
                              double* d_Kinv, int32_t* d_status, void* d_work, void* stream);

/* NLL (+ log-priors) and gradient at given theta for B surrogates -- one evaluation of
 * GaussianProcess._nll (tlbo/model/gp.py:164-197).  d_nll [B], d_grad [B, H],
 * d_status [B] (1 = Cholesky failed -> nll 1e25, grad 0 as the reference returns).    */
int tlbo_gp_nll_grad_batched(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap, int d,
                             const int32_t* h_types, const double* d_theta, double* d_nll, double* d_grad,
                             int32_t* d_status, void* d_work, void* stream);

/* Per-surrogate posterior mean / variance at a SMALL query set -- replaces
 * GaussianProcess._predict (tlbo/model/gp.py:250-305) + skopt predict(return_std=True)
 * for the K source predicts at the n target points (tlbo/facade/rgpe.py:27-30,
 * topo_variant1.py:27-35) and the CV-fold predicts (rgpe.py:56-70).
 * d_Xq [Nq, d] reference column order; outputs d_mu, d_var [M, Nq] un-standardised,
 * variance clipped at 1e-10 before un-standardising as gp.py:297-300 does.            */
int tlbo_gp_predict(const tlbo_pack* pack, const int32_t* h_types, const double* d_Xq, int Nq, double* d_mu,
                    double* d_var, void* stream);

/* Fused ensemble predict + EI + arg-max over a candidate pool -- replaces, in one pass,
 * (K+1) x GaussianProcess.predict, RGPE.predict / combine_predictions
 * (tlbo/facade/rgpe.py:141-153, base_facade.py:145-183), the facade variance floor
 * (base_facade.py:140-141), EI._compute (tlbo/acquisition_function/acquisition.py:130-177),
 * the NaN rule of __call__ (:73-76), lexsort((rand, ei))[::-1]
 * (tlbo/optimizer/ei_optimization.py:126-132) and the first-unevaluated scan
 * (tlbo/framework/smbo_offline.py:261-263).
 *   d_w [M]           ensemble weight of every pack slot
 *   d_skip [M] int32  1 = leave the surrogate out (RGPE ignored_flag), may be NULL
 *   d_order [M] int32 pack slots in ACCUMULATION order (floating-point sums follow the
 *                     reference's loop order: RGPE adds the target first, rgpe.py:142-152;
 *                     combine_predictions adds it last, base_facade.py:152-160); NULL = 0..M-1
 *   mode              0: mu = sum w mu_m, var = sum w^2 var_m   (RGPE / 'idp_lc')
 *                     1: TST (tst.py:64-73): first slot in order = target t,
 *                        mu = (mu_t + sum_{m != t} w_m mu_m) / (0.75 + sum_{m != t} w_m), var = var_t
 *                     2: POGPE / "gogpe" (pogpe.py:36-60): product of experts,
 *                        var = 1 / sum_m (w_m / var_m) (a zero sum -> 1e-5), mu = var * sum_m (w_m mu_m / var_m)
 *   d_cand [N, d]     candidate rows, reference column order
 *   eta, par          EI incumbent and xi
 *   var_floor         facade / model variance threshold (1e-10 or 1e-5)
 *   d_keys [N]        tie-break keys rng.rand(N); NULL -> ties resolved by larger index
 *   d_excluded [n_excl] int64 SORTED global row ids already evaluated (may be NULL)
 *   idx_offset        global row id of d_cand row 0 (candidate-pool sharding)
 * Outputs: d_best [4] doubles = (ei, key, global idx as double, #EI<0) of the best
 *   non-excluded row; optional per-row d_mu, d_var, d_ei [N] (after floor / NaN rule).
 *   d_work: scratch of tlbo_predict_ei_workspace_bytes(N) bytes.                       */
int tlbo_predict_ei_fused(const tlbo_pack* pack, const int32_t* h_types, const double* d_w,
                          const int32_t* d_skip, const int32_t* d_order, int mode, const double* d_cand,
                          int64_t N, double eta,
                          double par, double var_floor, const double* d_keys, const int64_t* d_excluded,
                          int n_excl, int64_t idx_offset, double* d_best, double* d_mu, double* d_var,
                          double* d_ei, void* d_work, void* stream);
/* Exchange step of the row-sharded candidate pool (the reference has no counterpart: its
 * _sort_configs_by_acq_value, tlbo/optimizer/ei_optimization.py:106-132, orders ONE table in one process): merges
 * `nparts` (ei, key, global row, #negative EI) quadruples -- the per-rank results of tlbo_predict_ei_fused*,
 * all-gathered into d_quads[nparts][4] -- with the same lexicographic rule (EI desc, key desc, row desc among
 * row >= 0) into d_best[4], on the device. */
int tlbo_merge_best(const double* d_quads, int nparts, double* d_best, void* stream);

int64_t tlbo_predict_ei_workspace_bytes(int64_t N);

/* Resident per-surrogate posteriors.  The source surrogates of a facade never change after
 * build_source_surrogates (tlbo/facade/base_facade.py:58-91) and OfflineSearch scores the SAME
 * candidate table every iteration (tlbo/optimizer/ei_offline_optimizer.py:20-29), so their
 * posterior mean / variance over the pool -- which the reference recomputes in every
 * RGPE.predict / combine_predictions call -- is computed ONCE and kept in HBM:
 *   tlbo_predict_pool_posteriors fills d_model_mu / d_model_var [M, N] (un-standardised, GP-level
 *   1e-10 clip applied) for every slot of `pack`;
 *   tlbo_predict_ei_fused_cached is tlbo_predict_ei_fused with slots m < n_cached read from
 *   d_cache_mu / d_cache_var [n_cached, N] instead of being evaluated.  Results are bit-identical
 *   to the uncached call (same values, same accumulation order).                               */
int tlbo_predict_pool_posteriors(const tlbo_pack* pack, const int32_t* h_types, const double* d_cand, int64_t N,
                                 double* d_model_mu, double* d_model_var, void* d_work, void* stream);
int tlbo_predict_ei_fused_cached(const tlbo_pack* pack, const int32_t* h_types, const double* d_w,
                                 const int32_t* d_skip, const int32_t* d_order, int mode, const double* d_cand,
                                 int64_t N, double eta, double par, double var_floor, const double* d_keys,
                                 const int64_t* d_excluded, int n_excl, int64_t idx_offset, double* d_best,
                                 double* d_mu, double* d_var, double* d_ei, const double* d_cache_mu,
                                 const double* d_cache_var, int n_cached, void* d_work, void* stream);

/* Integer mis-ranked-pair counts -- replaces the pure-Python loops of RGPE.train /
 * TransBO_RGPE.train (tlbo/facade/rgpe.py:74-104, topo.py:72-104) and TST.train
 * (tlbo/facade/tst.py:37-43).
 *   d_y [n]; d_ys [S, Mr, n] sampled (or predicted) values;
 *   d_row_lo, d_row_hi [Mr] int32 row range of i; j runs over 0..n-1, or over i+1..n-1
 *   when upper_only != 0 (TST).  d_counts [S, Mr] int64 =
 *   #{(i, j): (y_i < y_j) xor (ys_i < ys_j)}.  Bit-exact.                              */
int tlbo_rank_loss_counts(const double* d_y, const double* d_ys, int n, int S, int Mr, const int32_t* d_row_lo,
                          const int32_t* d_row_hi, int upper_only, int64_t* d_counts, void* stream);

/* RGPE weights and weight-dilution flags from the pair-count table, on the device -- replaces the host tail of
 * RGPE.train (tlbo/facade/rgpe.py:106-133: np.argmin per bootstrap sample, the histogram / num_sample, and
 * `median(source losses) > sorted(target losses)[int(0.95 S)]`), so that the fused predict + EI launch can follow the
 * count kernel on the stream without a host round trip.
 *   d_counts [S, K + n_folds] int64 (tlbo_rank_loss_counts[_normal]); target loss = sum of its fold columns, or
 *   n_sq when n_folds == 0 (rgpe.py:84-87); d_w [K + 1] = histogram / S; d_skip [K + 1] int32 (1 = ignored source,
 *   0 for the target).  Bit-exact against numpy (integer arithmetic, one division per weight).  S <= 32, K <= 127. */
int tlbo_rgpe_weights(const int64_t* d_counts, int S, int K, int n_folds, int64_t n_sq, double* d_w, int32_t* d_skip,
                      void* stream);

/* Same counts with the bootstrap draws formed on the device: d_z [S, Mr, n] are STANDARD normal
 * draws (host, legacy np.random stream, reference draw order); the sampled values of
 * rgpe.py:77,97  np.random.normal(mu, var)  are loc + scale * z with loc = d_loc[d_src[m]],
 * scale = d_scale[d_src[m]] (rows of the [*, n] outputs of tlbo_gp_predict), evaluated as a
 * rounded product plus a rounded sum exactly like numpy's legacy normal().  Saves the
 * device -> host -> device round trip of the predictions inside RGPE.train.                 */
int tlbo_rank_loss_counts_normal(const double* d_y, const double* d_z, const double* d_loc, const double* d_scale,
                                 const int32_t* d_src, int n, int S, int Mr, const int32_t* d_row_lo,
                                 const int32_t* d_row_hi, int64_t* d_counts, void* stream);

/* Pairwise logistic ranking loss and gradient -- replaces Loss_func / Loss_der with
 * func_id = 3 (tlbo/utils/scipy_solver.py:6-78) inside the SLSQP loop of scipy_solve.
 *   d_A [n, m] predictions, d_y [n], d_w [m];  d_out [2 + m] = (loss, #pairs, grad[m]). */
int tlbo_pairwise_logistic_loss_grad(const double* d_A, const double* d_y, const double* d_w, int n, int m,
                                     double* d_out, void* stream);
/* Same, with the weight vector taken from HOST memory (m <= 64) and passed as a kernel argument:
 * one SLSQP callback = one launch, no host -> device copy.  d_out may be mapped pinned host
 * memory (unified addressing), so the callback only has to synchronise the stream.           */
int tlbo_pairwise_logistic_loss_grad_hw(const double* d_A, const double* d_y, const double* h_w, int n, int m,
                                        double* d_out, void* stream);
/* Same, and WAITS for the result: h_out must be mapped pinned host memory; it is pre-filled with the
 * TLBO_EI_PENDING payload and the call returns when the kernel's 2 + m eight-byte stores have all landed (a spin on
 * the mapped words instead of a stream synchronisation: one SLSQP callback = one C call).                          */
int tlbo_pairwise_logistic_loss_grad_hw_wait(const double* d_A, const double* d_y, const double* h_w, int n, int m,
                                             double* h_out, void* stream);

/* MCMC marginalisation of the GP hyper-parameters (BASELINE config 5) -- replaces
 * GaussianProcessMCMC._train / _ll (tlbo/model/gp_mcmc.py:116-282,294-332): the
 * emcee.EnsembleSampler chain over the kernel hyper-parameters (emcee 3.x stretch move, a = 2,
 * two randomly split halves per step; un-vendored dependency, requirements.txt:16) with
 * log-probability = sklearn GaussianProcessRegressor.log_marginal_likelihood(theta) + lognormal
 * prior on c + horseshoe(0.1) prior on s2 + Tophat bound priors on every dimension
 * (tlbo/model/base_gp.py:92-132, gp_base_prior.py:155-240), -inf on a failed Cholesky or a
 * non-finite value.  B surrogates x W walkers (W even) advance `nsteps` steps in ONE persistent
 * cooperative launch per group of co-resident surrogates; no host round trips.
 *   d_X, d_y, d_n        as tlbo_gp_fit_batched (y after BaseGP._normalize_y)
 *   d_pos [B, W, H]      in: initial ensemble (prior samples, gp_mcmc.py:153-171 or the previous
 *                        chain's p0); out: final positions (sampler.get_chain()[-1])
 *   d_lp  [B, W]         out: log-probability of the final positions (nsteps = 0: of d_pos, i.e. one
 *                        batched evaluation of _ll)
 *   random tables, host-drawn from the sampler's legacy MT19937 stream in emcee's order, shape
 *   [n_tables, 2 * nsteps, W/2], half-step h = 2 * step + split:
 *     d_sidx   int32  walker ids of the active half (ascending)
 *     d_cidx   int32  partner walker id (complementary half) of each active walker
 *     d_zz            stretch factors ((a - 1) u + 1)^2 / a
 *     d_factors       (H - 1) * log(zz)
 *     d_logu          log of the acceptance uniforms
 *   d_table_of [B] int32 table used by surrogate b (NULL: table 0 for all)
 *   d_prior_lo, d_prior_hi [H]  Tophat bounds in the ORIGINAL scale
 *   d_theta0 [H]         kernel default theta for the initial GaussianProcessRegressor.fit of
 *                        gp_mcmc.py:141 (d_status[b] = 1 if its Cholesky fails); may be NULL
 *   d_naccept [B, W] int32  accepted proposals per walker (accumulated, caller zeroes)
 *   d_work               scratch of tlbo_gp_mcmc_workspace_bytes() bytes                         */
int tlbo_gp_mcmc_chain(const double* d_X, const double* d_y, const int32_t* d_n, int B, int ncap, int d,
                       const int32_t* h_types, int W, int nsteps, double* d_pos, double* d_lp,
                       const int32_t* d_sidx, const int32_t* d_cidx, const double* d_zz, const double* d_factors,
                       const double* d_logu, const int32_t* d_table_of, const double* d_prior_lo,
                       const double* d_prior_hi, const double* d_theta0, int32_t* d_naccept, int32_t* d_status,
                       void* d_work, void* stream);
int64_t tlbo_gp_mcmc_workspace_bytes(int B, int W, int ncap, int d);

/* Hyper-parameter-marginalised posterior -- replaces GaussianProcessMCMC._predict
 * (tlbo/model/gp_mcmc.py:418-440): for each of G surrogates with W walker GPs each,
 *   m = mean_w mu_w,  v = var_w(mu_w) + mean_w(var_w),  v clipped at DBL_EPSILON.
 * d_mu, d_var [G * W, N] per-walker posteriors (outputs of tlbo_gp_predict /
 * tlbo_predict_pool_posteriors); d_gmu, d_gvar [G, N].                                            */
int tlbo_mcmc_marginalize(const double* d_mu, const double* d_var, int G, int W, int64_t N, double* d_gmu,
                          double* d_gvar, void* stream);

/* Host-side helper of the online scenario's local search (no device work): the complete
 * one-exchange neighbourhood of a configuration in the order ConfigSpace's lazy generator
 * get_one_exchange_neighbourhood(configuration, seed) yields it -- the call the reference makes once
 * per hill-climb step (tlbo/optimizer/ei_optimization.py:266-267, tlbo/config_space/__init__.py:10).
 * Reproduces the generator's private numpy RandomState(seed) stream (legacy MT19937, masked-rejection
 * randint / shuffle, polar-method normal).
 *   h_array [d]        the configuration (reference encoding; non-finite = inactive)
 *   h_types [d]        get_types vector; h_const_mask [d] 1 = Constant hyper-parameter (may be NULL)
 *   num_neighbors, stdev   the reference binds 8 and 0.05 (tlbo/config_space/__init__.py:10); ConfigSpace's own
 *                          defaults are 4 and 0.2
 *   h_out [max_rows, d], *h_n_rows   neighbours, one changed column per row                      */
int tlbo_one_exchange_neighbourhood(const double* h_array, int d, const int32_t* h_types,
                                    const uint8_t* h_const_mask, uint32_t seed, int num_neighbors, double stdev,
                                    double* h_out, int max_rows, int32_t* h_n_rows);

/* OBTL's sampled ranking loss (reference tlbo/facade/obtl.py:95-147, calculate_weight_by_sampling +
 * calculate_ranking_loss / penalty_func): loss[s, m] = sum_{i<j} log(1 + exp(-10 z_ij)) / n^2 with
 * y_pred = loc[m] + scale[m] * z[s, m] (scale = the predictive variance, as the reference writes it) and
 * z_ij = (y_i < y_j) ? y_pred_j - y_pred_i : y_pred_i - y_pred_j.
 *   d_y [n]; d_z [S, Mr, n] standard-normal draws (host legacy np.random, reference draw order);
 *   d_loc, d_scale [Mr, n]; d_loss [S, Mr]                                                               */
int tlbo_pair_penalty_loss_normal(const double* d_y, const double* d_z, const double* d_loc, const double* d_scale,
                                  int n, int S, int Mr, double* d_loss, void* stream);

/* Ensemble EI of a SMALL query set in one launch -- the acquisition call of one hill-climb step
 * (tlbo/optimizer/ei_optimization.py:270-281 evaluates the neighbours one by one and stops at the first improvement;
 * acquisition.py:130-177 over RGPE.predict, rgpe.py:141-153): one CTA per (query row, surrogate) computes the
 * posterior (as tlbo_gp_predict; training rows, alpha and L^-1 staged in shared memory by bulk copies); the last CTA
 * of a row combines its surrogates in accumulation order (d_w / d_skip / d_order / mode / var_floor as
 * tlbo_predict_ei_fused -- same operations) and scores EI.
 *   Xq [Nq, d] and ei [Nq] may be device memory or MAPPED pinned host memory (UVA);
 *   h_Xq: optional HOST-readable copy of the rows (may equal Xq when that is host memory): when Nq * d <= 1024 the
 *   rows then travel inside the kernel arguments (constant bank) and Xq is not read (may be NULL);
 *   ei: EI per row, NaN kept (the caller applies acquisition.py:72-76).  Each row is ONE 8-byte store with no
 *   fence behind it: a host that pre-fills ei[] with TLBO_EI_PENDING (a NaN payload no computation produces) may
 *   spin until no row holds it instead of synchronising the stream;
 *   d_mu, d_var: [M, Nq] per-surrogate posteriors (scratch the combination reads; left there);
 *   d_counter: Nq device uint32, zero on entry, zero on completion.  Thread-safe (own buffers).                     */
#define TLBO_EI_PENDING_BITS 0x7ff8dead0000beefULL
int tlbo_ensemble_ei_small(const tlbo_pack* pack, const int32_t* h_types, const double* d_w, const int32_t* d_skip,
                           const int32_t* d_order, int mode, const double* Xq, const double* h_Xq, int Nq, double eta,
                           double par, double var_floor, double* d_mu, double* d_var, double* ei, uint32_t* d_counter,
                           void* stream);

/* One hill-climb of the online scenario's local search (tlbo/optimizer/ei_optimization.py:239-294,
 * LocalSearch._one_iter), native driver: per step tlbo_one_exchange_neighbourhood(h_config, seed = the next of the
 * pre-drawn rng.randint(0, 10000) values), ONE launch scoring the whole neighbourhood
 * (tlbo_ensemble_ei_small), a spin on the EI rows, accept = the first neighbour with EI > incumbent.
 *   view: the K + 1 surrogates; d_w / d_skip / d_order / mode / eta / par / var_floor: as tlbo_predict_ei_fused;
 *   h_config [d], h_acq: incumbent and its EI, updated in place;
 *   h_X [cap, d] host staging for the neighbourhood; h_ei [cap]: MAPPED pinned host memory (EI written by the
 *   kernel over the bus, polled row by row);
 *   d_mu, d_var [M * cap] device scratch; d_counter: cap zeroed device uint32 (left zero);
 *   h_counters int32[5]: += seeds consumed, neighbours the lazy reference loop inspects, acquisition evaluations,
 *   rows with EI < 0 (the caller raises as EI._compute does), steps of this climb (caller zeroes [4] per climb).
 * Returns 0: climb ended; 1: seeds ran out (call again with more, state is up to date).  Thread-safe with per-thread buffers. */
int tlbo_local_search_climb(const tlbo_pack* view, const int32_t* h_types, const uint8_t* h_const_mask,
                            const double* d_w, const int32_t* d_skip, const int32_t* d_order, int mode, double eta,
                            double par, double var_floor, int num_neighbors, double stdev, const uint32_t* h_seeds,
                            int n_seeds, int max_steps, double* h_config, double* h_acq, double* h_X, double* d_mu,
                            double* d_var, double* h_ei, uint32_t* d_counter, int cap, int32_t* h_counters,
                            void* stream);

/* numpy's LEGACY RandomState.random_sample on the device, bit for bit, with the MT19937 state resident in HBM:
 * the `rng.rand(N)` tie-break keys that `_sort_configs_by_acq_value` (reference
 * tlbo/optimizer/ei_optimization.py:106-132) draws over the whole candidate table on every maximisation.
 *   d_state  uint32[625]: key[624] then pos, i.e. RandomState.get_state()[1] and [2]; advanced in place
 *   d_out    double[N]                                                                                   */
int tlbo_mt19937_random_sample(uint32_t* d_state, int64_t N, double* d_out, void* stream);

/* The same stream from C thread blocks at once (the same reference call site: rng.rand(N) over the WHOLE table,
 * 20 ms of one CTA at N = 1e7).  MT19937 is linear over GF(2): the window of 624 state words that starts J words
 * further down the sequence is g_J(F) W with g_J(t) = t^J mod phi(t).
 *   tlbo_mt19937_jump: d_states[c] (c = 0 .. C-1, uint32[625] each) = the generator state d_base advanced by the
 *     jump whose polynomial's set-bit positions are d_bitpos[c * stride .. + d_nbits[c]) (slot 0 is a copy of
 *     d_base; the jumps are multiples of 624 words, so the in-block position carries over);
 *   tlbo_mt19937_random_sample_sliced: slice c delivers doubles [c S, min((c + 1) S, N)) from d_states[c];
 *     S a multiple of 312 (one state block), (C - 1) S < N <= C S; d_state_out (uint32[625],
 *     a separate buffer) receives the generator state after all N doubles -- the base of the
 *     next call's jump and what RandomState.set_state takes.                                            */
int tlbo_mt19937_jump(const uint32_t* d_base, uint32_t* d_states, int C, const int32_t* d_bitpos,
                      const int32_t* d_nbits, int stride, void* stream);
int tlbo_mt19937_random_sample_sliced(const uint32_t* d_states, int C, int64_t N, int64_t S, double* d_out,
                                      uint32_t* d_state_out, void* stream);

/* Host-side helper (no device work): the host half of build_single_surrogate + GaussianProcess._train for
 * the 1 + 5 training sets a facade builds per BO iteration from the same (X, y) -- all rows, and all rows
 * except one contiguous validation range (reference tlbo/facade/base_facade.py:93-108, rgpe.py:55-70,
 * topo_variant1.py CV folds, tlbo/model/gp.py:115-122, base_gp.py:48-64,148-151, utils/normalization.py).
 * Applies the all-equal guards (h_y_inout is mutated exactly where the reference mutates the caller's y),
 * the facade-level normalisation (0 none, 1 standardize, 2 zero-one), the non-finite -> -1 imputation and
 * the GP-level standardisation, bit-identical to the numpy statement (numpy's pairwise summation), and
 * writes the zero-padded upload buffers of tlbo_gp_fit_batched.
 *   h_X [n, d], h_y_inout [n]; job b keeps every row except [h_skip_lo[b], h_skip_hi[b])
 *   h_outer_guard [n_jobs] (may be NULL): 1 = apply rgpe.py:62-63's guard on the original y first
 *   h_Xout [n_jobs, ncap, d], h_yout [n_jobs, ncap], h_nout / h_ymean / h_ystd [n_jobs]                    */
int tlbo_host_prepare_jobs(const double* h_X, double* h_y_inout, int n, int d, int n_jobs,
                           const int32_t* h_skip_lo, const int32_t* h_skip_hi, const int32_t* h_outer_guard,
                           int normalize, int normalize_y, int ncap, double* h_Xout, double* h_yout,
                           int32_t* h_nout, double* h_ymean, double* h_ystd);

/* Device-rate probes used by bench.py for the roofline denominators (FP64 FMA pipe and
 * FP64 mma.sync rate are not in MEASURED_PEAKS.json).  Returns achieved TFLOP/s in
 * h_out[0] (DFMA) and h_out[1] (DMMA m8n8k4); synchronous.                             */
int tlbo_fp64_peak_probe(double* h_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
