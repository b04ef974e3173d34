// Ensemble-weighting kernels.
//
//  * rank_loss_counts: bit-exact integer mis-ranked-pair counts -- replaces the pure-Python
//    O(S * (K+1) * n^2) loops of RGPE.train / TransBO_RGPE.train
//    (reference tlbo/facade/rgpe.py:74-104, tlbo/facade/topo.py:72-104) and the discordant
//    pair count of TST.train (tlbo/facade/tst.py:37-43).  The normal draws themselves are
//    produced on the host by legacy np.random in the reference's draw order and uploaded;
//    the kernel only compares and counts.
//  * pairwise_logistic_loss_grad: Loss_func / Loss_der with func_id = 3
//    (reference tlbo/utils/scipy_solver.py:6-78) evaluated once per SLSQP callback.
#include "tlbo_common.cuh"

#define RL_THREADS 128

__global__ void __launch_bounds__(RL_THREADS) rank_loss_kernel(const double* __restrict__ y,
                                                               const double* __restrict__ ys, int n, int Mr,
                                                               const int32_t* __restrict__ row_lo,
                                                               const int32_t* __restrict__ row_hi, int upper_only,
                                                               const double* __restrict__ loc,
                                                               const double* __restrict__ scale,
                                                               const int32_t* __restrict__ src,
                                                               long long* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sy = reinterpret_cast<double*>(smem);
  double* ss = sy + n;
  __shared__ unsigned long long total;
  const int item = blockIdx.x;           // s * Mr + m
  const int m = item % Mr;
  const double* zs = ys + (size_t)item * n;
  if (loc) {
    // ys holds standard-normal draws z: sample = loc + scale * z with a rounded product and a
    // rounded sum -- the arithmetic of numpy's legacy normal(loc, scale) (no FMA contraction)
    const double* lo_ = loc + (size_t)src[m] * n;
    const double* sc_ = scale + (size_t)src[m] * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      sy[i] = y[i];
      ss[i] = __dadd_rn(lo_[i], __dmul_rn(sc_[i], zs[i]));
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sy[i] = y[i]; ss[i] = zs[i]; }
  }
  if (threadIdx.x == 0) total = 0ull;
  __syncthreads();
  const int lo = row_lo[m], hi = row_hi[m];
  unsigned int cnt = 0;
  // work item = (i, j-chunk of 32): lanes take consecutive j, warps take rows
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = lo + wid; i < hi; i += nw) {
    const double yi = sy[i], si = ss[i];
    const int j0 = upper_only ? i + 1 : 0;
    for (int j = j0 + lane; j < n; j += 32) {
      const bool a = yi < sy[j];
      const bool b = si < ss[j];
      cnt += (a != b) ? 1u : 0u;
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) atomicAdd(&total, (unsigned long long)cnt);
  __syncthreads();
  if (threadIdx.x == 0) counts[item] = (long long)total;
}

static int rank_loss_launch(const double* d_y, const double* d_ys, int n, int S, int Mr, const int32_t* d_row_lo,
                            const int32_t* d_row_hi, int upper_only, const double* d_loc, const double* d_scale,
                            const int32_t* d_src, int64_t* d_counts, void* stream) {
  if (n <= 0 || S <= 0 || Mr <= 0) { tlbo_set_error("bad n/S/Mr"); return TLBO_E_BADARG; }
  const size_t smem = 2 * (size_t)n * sizeof(double);
  if (smem > 200 * 1024) { tlbo_set_error("n too large for rank_loss_counts"); return TLBO_E_TOOLARGE; }
  if (smem > 48 * 1024)
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(rank_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rank_loss_kernel<<<S * Mr, RL_THREADS, smem, (cudaStream_t)stream>>>(d_y, d_ys, n, Mr, d_row_lo, d_row_hi,
                                                                        upper_only, d_loc, d_scale, d_src,
                                                                        (long long*)d_counts);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int tlbo_rank_loss_counts(const double* d_y, const double* d_ys, int n, int S, int Mr,
                                     const int32_t* d_row_lo, const int32_t* d_row_hi, int upper_only,
                                     int64_t* d_counts, void* stream) {
  return rank_loss_launch(d_y, d_ys, n, S, Mr, d_row_lo, d_row_hi, upper_only, nullptr, nullptr, nullptr, d_counts,
                          stream);
}

extern "C" int tlbo_rank_loss_counts_normal(const double* d_y, const double* d_z, const double* d_loc,
                                            const double* d_scale, const int32_t* d_src, int n, int S, int Mr,
                                            const int32_t* d_row_lo, const int32_t* d_row_hi, int64_t* d_counts,
                                            void* stream) {
  if (!d_loc || !d_scale || !d_src) { tlbo_set_error("NULL loc/scale/src"); return TLBO_E_BADARG; }
  return rank_loss_launch(d_y, d_z, n, S, Mr, d_row_lo, d_row_hi, 0, d_loc, d_scale, d_src, d_counts, stream);
}

// ---------------------------------------------------------------------------------------
// RGPE weights from the pair-count table, on the device (reference tlbo/facade/rgpe.py:106-133): losses[s][k] =
// counts[s][k] for the K sources and the sum of the fold counts for the target (n^2 when there are no folds);
// w_k = #{s : argmin_k' losses[s][k'] == k} / S (np.argmin: first minimum); a source is ignored when the median of
// its losses exceeds the 95th-percentile loss of the target (sorted(column)[int(S * 0.5)] > sorted(target)[int(S *
// 0.95)]).  Integer arithmetic and one division per weight: the same doubles as numpy.  One CTA; it lets the fused
// predict + EI launch follow the count kernel on the stream without a host round trip (the host recomputes the same
// table later for its own bookkeeping).
#define RW_MAXS 32
#define RW_MAXM 127

__global__ void __launch_bounds__(128) rgpe_weights_kernel(const int64_t* __restrict__ counts, int S, int K, int n_folds,
                                                           long long n_sq, double* __restrict__ w,
                                                           int32_t* __restrict__ skip) {
  __shared__ long long L[RW_MAXS][RW_MAXM + 1];
  __shared__ int hist[RW_MAXM + 1];
  __shared__ long long thr;
  const int tid = threadIdx.x, M = K + 1, Mr = K + n_folds;
  for (int idx = tid; idx < S * M; idx += blockDim.x) {
    const int s = idx / M, k = idx - s * M;
    long long v;
    if (k < K) v = counts[(size_t)s * Mr + k];
    else if (n_folds == 0) v = n_sq;
    else { v = 0; for (int f = 0; f < n_folds; f++) v += counts[(size_t)s * Mr + K + f]; }
    L[s][k] = v;
  }
  for (int k = tid; k < M; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  for (int s = tid; s < S; s += blockDim.x) {
    int best = 0;
    for (int k = 1; k < M; k++) if (L[s][k] < L[s][best]) best = k;
    atomicAdd(&hist[best], 1);
  }
  // order statistic r of column k by ONE WARP (S <= 32: lane i ranks sample i against the column): the value x with
  // #less <= r < #less + #equal
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  auto kth = [&](int k, int r) -> long long {
    const long long x = (lane < S) ? L[lane][k] : 0;
    int less = 0, eq = 0;
    for (int j = 0; j < S; j++) { const long long v = L[j][k]; less += v < x; eq += v == x; }
    const unsigned hit = __ballot_sync(0xffffffffu, lane < S && less <= r && r < less + eq);
    return __shfl_sync(0xffffffffu, x, __ffs(hit) - 1);
  };
  if (wid == 0) { const long long t = kth(K, (int)(S * 0.95)); if (lane == 0) thr = t; }
  __syncthreads();
  for (int k = wid; k < M; k += nw) {
    const long long med = (k < K) ? kth(k, (int)(S * 0.5)) : 0;
    if (lane == 0) {
      w[k] = (double)hist[k] / (double)S;
      skip[k] = (k < K && med > thr) ? 1 : 0;
    }
  }
}

extern "C" int tlbo_rgpe_weights(const int64_t* d_counts, int S, int K, int n_folds, int64_t n_sq, double* d_w,
                                 int32_t* d_skip, void* stream) {
  if (!d_counts || !d_w || !d_skip || S <= 0 || S > RW_MAXS || K < 0 || K + 1 > RW_MAXM + 1 || n_folds < 0) {
    tlbo_set_error("tlbo_rgpe_weights: bad arguments");
    return TLBO_E_BADARG;
  }
  rgpe_weights_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(d_counts, S, K, n_folds, (long long)n_sq, d_w, d_skip);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
#define PL_THREADS 256
#define PL_MAXM 64

struct PlWeights { double w[PL_MAXM]; };

// One CTA.  Phase 1: p = A w.  Phase 2: all n^2 ordered pairs spread over the 256 threads
// (thread (r, c): row r, partners q = c, c + 4, ...), each pair with y_q > y_r contributes
// log(1 + e^{p_r - p_q}) to the loss and +-sigma(p_r - p_q) to the per-row coefficients.
// Phase 3: grad = A^T coef / #pairs.  hw != NULL: weights arrive as a kernel argument (no copy).
__global__ void __launch_bounds__(PL_THREADS) pair_logistic_kernel(const double* __restrict__ A,
                                                                   const double* __restrict__ y,
                                                                   const double* __restrict__ w, PlWeights hw,
                                                                   int use_hw, int n, int m,
                                                                   double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sp = reinterpret_cast<double*>(smem);   // predictions [n]
  double* sy = sp + n;                            // [n]
  double* sc = sy + n;                            // per-row gradient coefficient [n]
  double* red = sc + n;                           // [32]
  double* sw = red + 32;                          // [m]
  const int tid = threadIdx.x, T = blockDim.x;
  for (int k = tid; k < m; k += T) sw[k] = use_hw ? hw.w[k] : w[k];
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    double p = 0.0;
    for (int k = 0; k < m; k++) p += A[(size_t)i * m + k] * sw[k];
    sp[i] = p;
    sy[i] = y[i];
    sc[i] = 0.0;
  }
  __syncthreads();
  // Row r owns the pairs in which it is the LOWER-y member (reference pair (i, j) with
  // y_i > y_j has j = r): loss term log(1 + exp(p_r - p_i)), coefficient +s on row r, -s on row i.
  double loss = 0.0, npairs = 0.0;
  const int lanes_per_row = 4;
  const int c = tid & (lanes_per_row - 1);
  const int rows_per_pass = T / lanes_per_row;
  for (int r0 = 0; r0 < n; r0 += rows_per_pass) {          // uniform trip count: shuffles below
    const int r = r0 + tid / lanes_per_row;
    const bool valid = r < n;
    const double yr = valid ? sy[r] : 0.0, pr = valid ? sp[r] : 0.0;
    double coef = 0.0;
    if (valid) {
      for (int q = c; q < n; q += lanes_per_row) {
        const double yq = sy[q];
        if (yq > yr) {
          const double ez = exp(pr - sp[q]);
          loss += log(1.0 + ez);
          npairs += 1.0;
          coef += ez / (1.0 + ez);
        } else if (yq < yr) {
          const double ez = exp(sp[q] - pr);
          coef -= ez / (1.0 + ez);
        }
      }
    }
    coef += __shfl_xor_sync(0xffffffffu, coef, 1);
    coef += __shfl_xor_sync(0xffffffffu, coef, 2);
    if (valid && c == 0) sc[r] = coef;
  }
  const double tl = block_sum(loss, red);
  const double tp = block_sum(npairs, red);
  __syncthreads();
  if (tid == 0) { out[0] = (tp > 0.0) ? tl / tp : 0.0; out[1] = tp; }
  for (int k = tid; k < m; k += T) {
    double gk = 0.0;
    for (int r = 0; r < n; r++) gk += sc[r] * A[(size_t)r * m + k];
    out[2 + k] = (tp > 0.0) ? gk / tp : 0.0;
  }
}

static int pair_logistic_launch(const double* d_A, const double* d_y, const double* d_w, const double* h_w, int n,
                                int m, double* d_out, void* stream) {
  if (n <= 0 || m <= 0) { tlbo_set_error("bad n/m"); return TLBO_E_BADARG; }
  if (h_w && m > PL_MAXM) { tlbo_set_error("m too large for host-side weights"); return TLBO_E_TOOLARGE; }
  const size_t smem = (3 * (size_t)n + 32 + m) * sizeof(double);
  if (smem > 200 * 1024) { tlbo_set_error("n too large for pairwise_logistic"); return TLBO_E_TOOLARGE; }
  if (smem > 48 * 1024)
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(pair_logistic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PlWeights hw;
  if (h_w) for (int k = 0; k < m; k++) hw.w[k] = h_w[k];
  pair_logistic_kernel<<<1, PL_THREADS, smem, (cudaStream_t)stream>>>(d_A, d_y, d_w, hw, h_w ? 1 : 0, n, m, d_out);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int tlbo_pairwise_logistic_loss_grad(const double* d_A, const double* d_y, const double* d_w, int n,
                                                int m, double* d_out, void* stream) {
  return pair_logistic_launch(d_A, d_y, d_w, nullptr, n, m, d_out, stream);
}

extern "C" int tlbo_pairwise_logistic_loss_grad_hw(const double* d_A, const double* d_y, const double* h_w, int n,
                                                   int m, double* d_out, void* stream) {
  if (!h_w) { tlbo_set_error("NULL h_w"); return TLBO_E_BADARG; }
  return pair_logistic_launch(d_A, d_y, nullptr, h_w, n, m, d_out, stream);
}

// One SLSQP callback in one call: h_out (MAPPED pinned host memory, 2 + m doubles) is pre-filled with the
// TLBO_EI_PENDING payload, the kernel overwrites every entry with one 8-byte store, and the host spins until no entry
// holds the payload -- no cudaStreamSynchronize (a driver round trip of ~10 us, ~30 times per TOPO iteration).  The
// stream is queried now and then so that a failed launch surfaces as its error instead of a hang.
extern "C" int tlbo_pairwise_logistic_loss_grad_hw_wait(const double* d_A, const double* d_y, const double* h_w, int n,
                                                        int m, double* h_out, void* stream) {
  if (!h_w || !h_out) { tlbo_set_error("NULL h_w / h_out"); return TLBO_E_BADARG; }
  double pending;
  const uint64_t bits = TLBO_EI_PENDING_BITS;
  memcpy(&pending, &bits, sizeof(pending));
  const int rows = 2 + m;
  for (int j = 0; j < rows; j++) h_out[j] = pending;
  int rc = pair_logistic_launch(d_A, d_y, nullptr, h_w, n, m, h_out, stream);
  if (rc) return rc;
  const volatile uint64_t* q = reinterpret_cast<const volatile uint64_t*>(h_out);
  int j = 0;
  for (unsigned spin = 1;; spin++) {
    while (j < rows && q[j] != bits) j++;
    if (j == rows) return 0;
    if ((spin & 0x3ff) == 0) {
      cudaError_t e = cudaStreamQuery((cudaStream_t)stream);
      if (e == cudaSuccess) {
        while (j < rows && q[j] != bits) j++;
        if (j == rows) return 0;
        tlbo_set_error("pairwise_logistic launch finished without reporting every output");
        return TLBO_E_BADARG;
      }
      if (e != cudaErrorNotReady) { tlbo_set_error(cudaGetErrorString(e)); return (int)e; }
    }
  }
}

// ---------------------------------------------------------------------------------------
// OBTL's sampled ranking loss (reference tlbo/facade/obtl.py:95-147): for every bootstrap sample s and
// surrogate m, y_pred = normal(mu_m, var_m) (scale = the predictive VARIANCE, as written there) and
//   loss = sum_{i<j} log(1 + exp(-10 z_ij)) / n^2,   z_ij = (y_i < y_j) ? y_pred_j - y_pred_i : y_pred_i - y_pred_j
// -- the pure-Python double loop of calculate_ranking_loss, 50 x (K + 1) x n^2 / 2 penalty evaluations per BO
// iteration.  One CTA per (sample, surrogate); the standard-normal draws come from the host's GLOBAL
// legacy np.random stream in the reference's order and the affine map runs here with a rounded product
// and a rounded sum (numpy's legacy normal).  A float loss: the sum order differs from the reference's
// sequential loop (relative 1e-15); what the facade consumes is the arg-min over surrogates per sample.
__global__ void __launch_bounds__(RL_THREADS) pair_penalty_kernel(const double* __restrict__ y,
                                                                  const double* __restrict__ z, int n, int Mr,
                                                                  const double* __restrict__ loc,
                                                                  const double* __restrict__ scale,
                                                                  double* __restrict__ loss) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* sy = reinterpret_cast<double*>(smem);
  double* ss = sy + n;
  __shared__ double red[32];
  const int item = blockIdx.x;           // s * Mr + m
  const int m = item % Mr;
  const double* zs = z + (size_t)item * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sy[i] = y[i];
    ss[i] = __dadd_rn(loc[(size_t)m * n + i], __dmul_rn(scale[(size_t)m * n + i], zs[i]));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double acc = 0.0;
  for (int i = wid; i < n; i += nw) {
    const double yi = sy[i], si = ss[i];
    for (int j = i + 1 + lane; j < n; j += 32) {
      double zz = (yi < sy[j]) ? (ss[j] - si) : (si - ss[j]);
      zz *= 10.0;
      acc += log(1.0 + exp(-zz));
    }
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) loss[item] = tot / ((double)n * (double)n);
}

extern "C" int tlbo_pair_penalty_loss_normal(const double* d_y, const double* d_z, const double* d_loc,
                                             const double* d_scale, int n, int S, int Mr, double* d_loss,
                                             void* stream) {
  if (!d_y || !d_z || !d_loc || !d_scale || !d_loss || n <= 0 || S <= 0 || Mr <= 0) {
    tlbo_set_error("tlbo_pair_penalty_loss_normal: bad arguments");
    return TLBO_E_BADARG;
  }
  const size_t smem = 2 * (size_t)n * sizeof(double);
  if (smem > 200 * 1024) { tlbo_set_error("n too large for pair_penalty_loss"); return TLBO_E_TOOLARGE; }
  if (smem > 48 * 1024)
    TLBO_CUDA_CHECK(cudaFuncSetAttribute(pair_penalty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pair_penalty_kernel<<<S * Mr, RL_THREADS, smem, (cudaStream_t)stream>>>(d_y, d_z, n, Mr, d_loc, d_scale, d_loss);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// Tie-break keys on the device: numpy's LEGACY `RandomState.random_sample` (MT19937, two 32-bit outputs
// per double: (a >> 5) * 2^26 + (b >> 6)) / 2^53), bit for bit, with the generator state resident in HBM.
//
// `_sort_configs_by_acq_value` (reference tlbo/optimizer/ei_optimization.py:106-132) draws `rng.rand(N)`
// over the WHOLE candidate table on every maximisation; at N = 1e6 that is ~5 ms of sequential host work and
// an 8 MB upload per BO iteration -- more than everything else in the iteration.  MT19937 regenerates its
// 624-word state in blocks; inside a block, word i depends on old words i, i + 1 and on word i + 397 (old for
// i < 227, else the NEW word i - 227), so a block splits into three data-parallel phases (227 + 227 + 169
// words) plus the last word.  One CTA walks the blocks (about 0.3 us each), tempers and converts 312 doubles
// per block with coalesced stores; it runs on a side stream under the fit kernel, which leaves 82 SMs idle.
// d_state: uint32[625] = key[624], pos -- numpy's `get_state()[1]`, `[2]`.
#define MT_N 624
#define MT_M 397
#define MT_THREADS 576

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((uint32_t)(-(int32_t)(y & 1u)) & 0x9908b0dfu);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
__device__ __forceinline__ double mt_double(uint32_t w1, uint32_t w2) {
  const double a = (double)(w1 >> 5), b = (double)(w2 >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;      // every step exact in binary64
}

// Warps 0-7 twist block b + 1 (from `cur` into `nxt`, three phases, three barriers) while warps 8-17 temper,
// convert and store the doubles of block b straight from `cur`: a block costs the three shared-memory hops of
// the twist and nothing else.  The look-ahead twist is speculative -- if the stream ends inside block b, block
// b + 1 is simply not committed, exactly as numpy leaves its key array untouched until the next draw.
#define MT_TWIST_THREADS 256
#define MT_WRITE_THREADS 320
// Grid of C CTAs: CTA c starts from its own window states_in[c] (the generator state c * S doubles further down
// the stream, see tlbo_mt19937_jump) and delivers doubles [c * S, min((c + 1) * S, N)); the LAST CTA's final
// state is the generator state after all N doubles and goes to state_out.  C = 1: the plain sequential walk.
__global__ void __launch_bounds__(MT_THREADS) mt19937_random_sample_kernel(const uint32_t* __restrict__ states_in,
                                                                           uint32_t* __restrict__ state_out,
                                                                           long long Ntot, long long S,
                                                                           double* __restrict__ out_all) {
  __shared__ uint32_t keybuf[2][MT_N];
  const int tid = threadIdx.x;
  const int wtid = tid - MT_TWIST_THREADS;          // >= 0: writer thread
  const uint32_t* state = states_in + (size_t)blockIdx.x * (MT_N + 1);
  const long long lo = (long long)blockIdx.x * S;
  const long long N = (Ntot - lo < S) ? Ntot - lo : S;
  double* out = out_all + lo;
  int cb = 0;
  for (int i = tid; i < MT_N; i += MT_THREADS) keybuf[0][i] = state[i];
  int pos = (int)state[MT_N];
  __syncthreads();
  long long produced = 0;
  bool carry_valid = false;
  uint32_t carry = 0;
  while (produced < N) {                      // every control variable is uniform over the CTA
    const uint32_t* cur = keybuf[cb];
    uint32_t* nxt = keybuf[cb ^ 1];
    // ---- bookkeeping of what block `cur` still has to deliver (positions >= pos)
    int start = pos;
    const long long base = produced;
    const bool had_carry = carry_valid && start < MT_N;
    if (had_carry) { produced++; start++; }
    const int first = start;                  // first word of the first whole pair
    long long pairs = (MT_N - start) / 2;
    if (pairs < 0) pairs = 0;
    const long long remain = N - produced;
    if (pairs > remain) pairs = remain;
    produced += pairs;
    start += 2 * (int)pairs;
    bool new_carry = false;
    if (produced < N && start < MT_N) { new_carry = true; start++; }      // odd word: first half of the next double
    // ---- writers: block `cur`;  twisters: phase 1 of the next block
    if (wtid >= 0) {
      if (had_carry && wtid == MT_WRITE_THREADS - 1) out[base] = mt_double(carry, mt_temper(cur[first - 1]));
      const long long ob = base + (had_carry ? 1 : 0);
      for (int k = wtid; k < (int)pairs; k += MT_WRITE_THREADS)
        out[ob + k] = mt_double(mt_temper(cur[first + 2 * k]), mt_temper(cur[first + 2 * k + 1]));
    } else if (tid < 227) {
      nxt[tid] = mt_twist(cur[tid], cur[tid + 1], cur[tid + MT_M]);
    }
    const uint32_t next_carry = new_carry ? mt_temper(cur[start - 1]) : 0u;
    __syncthreads();
    if (tid < 227) nxt[227 + tid] = mt_twist(cur[227 + tid], cur[228 + tid], nxt[tid]);
    __syncthreads();
    if (tid < 169) nxt[454 + tid] = mt_twist(cur[454 + tid], cur[455 + tid], nxt[227 + tid]);
    else if (tid == 169) nxt[623] = mt_twist(cur[623], nxt[0], nxt[396]);
    __syncthreads();
    carry_valid = new_carry;
    carry = next_carry;
    if (start >= MT_N && produced < N) { cb ^= 1; pos = 0; }     // commit the look-ahead block
    else pos = start;
  }
  if (blockIdx.x == gridDim.x - 1) {
    for (int i = tid; i < MT_N; i += MT_THREADS) state_out[i] = keybuf[cb][i];
    if (tid == 0) state_out[MT_N] = (uint32_t)pos;
  }
}

extern "C" int tlbo_mt19937_random_sample(uint32_t* d_state, int64_t N, double* d_out, void* stream) {
  if (!d_state || !d_out || N < 0) { tlbo_set_error("tlbo_mt19937_random_sample: bad arguments"); return TLBO_E_BADARG; }
  if (N == 0) return 0;
  mt19937_random_sample_kernel<<<1, MT_THREADS, 0, (cudaStream_t)stream>>>(d_state, d_state, (long long)N,
                                                                            (long long)N, d_out);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---- multi-CTA generation: jump-ahead ---------------------------------------------------------------------
// MT19937 is linear over GF(2): the window of 624 words that starts J words further down the word sequence is
// g_J(F) W with g_J(t) = t^J mod phi(t) (phi: the degree-19937 characteristic polynomial, F: the one-word
// shift), i.e. word j of the jumped window is the XOR over the set bits i of g_J of x_{i + j}.  The host
// supplies the bit positions (tlbo_b200/utils/mt_jump.py; they depend on the slice length only and are computed
// once per maximiser); one CTA per slice extends the base window by 32 blocks in shared memory (the same
// three-phase twist as the generator) and XORs ~10^4 shifted copies, one word per thread.  The jump is by a
// multiple of 624 words, so the in-block position carries over unchanged.
#define MT_JUMP_BLOCKS 33
#define MT_JUMP_THREADS 640
__global__ void __launch_bounds__(MT_JUMP_THREADS) mt19937_jump_kernel(const uint32_t* __restrict__ base,
                                                                       uint32_t* __restrict__ states,
                                                                       const int32_t* __restrict__ bitpos,
                                                                       const int32_t* __restrict__ nbits, int stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* seq = reinterpret_cast<uint32_t*>(smem);
  const int tid = threadIdx.x, c = blockIdx.x;
  uint32_t* dst = states + (size_t)c * (MT_N + 1);
  if (c == 0) {                                      // slice 0 starts at the base state itself
    for (int i = tid; i <= MT_N; i += MT_JUMP_THREADS) dst[i] = base[i];
    return;
  }
  for (int i = tid; i < MT_N; i += MT_JUMP_THREADS) seq[i] = base[i];
  __syncthreads();
  for (int b = 1; b < MT_JUMP_BLOCKS; b++) {
    const uint32_t* cur = seq + (size_t)(b - 1) * MT_N;
    uint32_t* nxt = seq + (size_t)b * MT_N;
    if (tid < 227) nxt[tid] = mt_twist(cur[tid], cur[tid + 1], cur[tid + MT_M]);
    __syncthreads();
    if (tid < 227) nxt[227 + tid] = mt_twist(cur[227 + tid], cur[228 + tid], nxt[tid]);
    __syncthreads();
    if (tid < 169) nxt[454 + tid] = mt_twist(cur[454 + tid], cur[455 + tid], nxt[227 + tid]);
    else if (tid == 169) nxt[623] = mt_twist(cur[623], nxt[0], nxt[396]);
    __syncthreads();
  }
  const int32_t* bp = bitpos + (size_t)c * stride;
  const int nb = nbits[c];
  if (tid < MT_N) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int q = 0;
    for (; q + 4 <= nb; q += 4) {
      r0 ^= seq[bp[q] + tid]; r1 ^= seq[bp[q + 1] + tid]; r2 ^= seq[bp[q + 2] + tid]; r3 ^= seq[bp[q + 3] + tid];
    }
    for (; q < nb; q++) r0 ^= seq[bp[q] + tid];
    dst[tid] = (r0 ^ r1) ^ (r2 ^ r3);
  }
  if (tid == 0) dst[MT_N] = base[MT_N];
}

extern "C" int tlbo_mt19937_jump(const uint32_t* d_base, uint32_t* d_states, int C, const int32_t* d_bitpos,
                                 const int32_t* d_nbits, int stride, void* stream) {
  if (!d_base || !d_states || !d_bitpos || !d_nbits || C <= 0 || stride <= 0) {
    tlbo_set_error("tlbo_mt19937_jump: bad arguments");
    return TLBO_E_BADARG;
  }
  const size_t smem = (size_t)MT_JUMP_BLOCKS * MT_N * sizeof(uint32_t);
  TLBO_CUDA_CHECK(cudaFuncSetAttribute(mt19937_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mt19937_jump_kernel<<<C, MT_JUMP_THREADS, smem, (cudaStream_t)stream>>>(d_base, d_states, d_bitpos, d_nbits, stride);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int tlbo_mt19937_random_sample_sliced(const uint32_t* d_states, int C, int64_t N, int64_t S, double* d_out,
                                                 uint32_t* d_state_out, void* stream) {
  if (!d_states || !d_out || !d_state_out || N < 0 || C <= 0 || S <= 0 || (S % 312) != 0 ||
      (int64_t)(C - 1) * S >= N + (N == 0) || (int64_t)C * S < N) {
    tlbo_set_error("tlbo_mt19937_random_sample_sliced: bad arguments (S must be a multiple of 312, (C-1) S < N <= C S)");
    return TLBO_E_BADARG;
  }
  if (N == 0) return 0;
  mt19937_random_sample_kernel<<<C, MT_THREADS, 0, (cudaStream_t)stream>>>(d_states, d_state_out, (long long)N,
                                                                            (long long)S, d_out);
  TLBO_CUDA_CHECK(cudaGetLastError());
  return 0;
}
