// Host-side native helper of the online scenario's local search: the complete one-exchange
// neighbourhood of a configuration, in the order ConfigSpace's lazy generator
// `util.get_one_exchange_neighbourhood(configuration, seed)` yields it (the reference calls it once
// per hill-climb step: tlbo/optimizer/ei_optimization.py:266-267 via tlbo/config_space/__init__.py:10).
//
// The generator owns a private numpy `RandomState(seed)`; reproducing its stream needs numpy's LEGACY
// generator bit for bit: MT19937 seeded by `init_genrand` (integer seeds), `randint` / `shuffle` by
// masked rejection on 32-bit outputs, `normal` by the polar Box-Muller method with a cached second
// variate on 53-bit doubles built from two outputs.  In Python this loop costs ~0.25 ms per
// hill-climb step (about 80 scalar draws through the interpreter) -- more than the device work of the
// step; here it is a few microseconds.  tests/test_online_search_cpu.py checks it draw for draw
// against the Python statement of the same algorithm (tlbo_b200/config_space).
#include <float.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#include "tlbo_common.cuh"

namespace {

struct LegacyRandom {
  uint32_t key[624];
  int pos;
  bool has_gauss;
  double gauss;

  explicit LegacyRandom(uint32_t seed) : pos(624), has_gauss(false), gauss(0.0) {
    for (int i = 0; i < 624; i++) {
      key[i] = seed;
      seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
  }
  void refill() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & A);
    }
    for (; i < 623; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & A);
    }
    y = (key[623] & UPPER) | (key[0] & LOWER);
    key[623] = key[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & A);
    pos = 0;
  }
  uint32_t next32() {
    if (pos == 624) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  double next_double() {
    const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  // masked rejection on [0, max] (random_interval / bounded_masked_uint32)
  uint32_t interval(uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (next32() & mask)) > max) { }
    return v;
  }
  double legacy_gauss() {
    if (has_gauss) {
      const double t = gauss;
      has_gauss = false;
      gauss = 0.0;
      return t;
    }
    double f, x1, x2, r2;
    do {
      x1 = 2.0 * next_double() - 1.0;
      x2 = 2.0 * next_double() - 1.0;
      r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    gauss = f * x1;
    has_gauss = true;
    return f * x2;
  }
};

}  // namespace

extern "C" int tlbo_one_exchange_neighbourhood(const double* h_array, int d, const int32_t* h_types,
                                               const uint8_t* h_const_mask, uint32_t seed, int num_neighbors,
                                               double stdev, double* h_out, int max_rows, int32_t* h_n_rows) {
  if (!h_array || !h_types || !h_out || !h_n_rows || d <= 0 || d > TLBO_MAX_D || num_neighbors < 0) {
    tlbo_set_error("bad arguments");
    return TLBO_E_BADARG;
  }
  LegacyRandom rnd(seed);
  int remaining[TLBO_MAX_D];
  int stack_len[TLBO_MAX_D];
  bool stack_made[TLBO_MAX_D];
  double stacks[TLBO_MAX_D][64];
  int usable = 0, used = 0;
  for (int j = 0; j < d; j++) {
    if (h_types[j] > 0) remaining[j] = h_types[j] - 1;
    else remaining[j] = (h_const_mask && h_const_mask[j]) ? 0 : num_neighbors;
    if (h_types[j] > 64) { tlbo_set_error("categorical column with more than 64 choices"); return TLBO_E_TOOLARGE; }
    stack_made[j] = false;
    stack_len[j] = 0;
    const bool fin = isfinite(h_array[j]);
    if (fin) usable++;
    if (fin && remaining[j] == 0) used++;
  }
  int rows = 0;
  while (used < usable) {
    const int index = (int)rnd.interval((uint32_t)(d - 1));
    if (remaining[index] == 0 || !isfinite(h_array[index])) continue;
    const double value = h_array[index];
    double neighbour = 0.0;
    bool have = false;
    if (h_types[index] > 0) {
      if (!stack_made[index]) {
        int n = 0;
        for (int c = 0; c < h_types[index]; c++) if (c != (int)value) stacks[index][n++] = (double)c;
        for (int i = n - 1; i >= 1; i--) {            // RandomState.shuffle on a list
          const int j = (int)rnd.interval((uint32_t)i);
          const double t = stacks[index][i]; stacks[index][i] = stacks[index][j]; stacks[index][j] = t;
        }
        stack_len[index] = n;
        stack_made[index] = true;
      }
      neighbour = stacks[index][--stack_len[index]];  // list.pop()
      have = true;
    } else {
      for (int it = 0; it < 101; it++) {
        const double cand = value + stdev * rnd.legacy_gauss();
        if (cand >= 0.0 && cand <= 1.0) { neighbour = cand; have = true; break; }
      }
      if (!have) {
        remaining[index] = 0;
        used++;
        continue;
      }
    }
    if (rows >= max_rows) { tlbo_set_error("neighbourhood larger than the output buffer"); return TLBO_E_TOOLARGE; }
    double* row = h_out + (size_t)rows * d;
    for (int j = 0; j < d; j++) row[j] = h_array[j];
    row[index] = neighbour;
    rows++;
    remaining[index]--;
    if (remaining[index] == 0) used++;
  }
  *h_n_rows = rows;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Host half of `build_single_surrogate` + `GaussianProcess._train` for a batch of cross-validation jobs
// (reference tlbo/facade/base_facade.py:93-108, tlbo/facade/rgpe.py:55-70, tlbo/facade/topo_variant1.py,
// tlbo/model/gp.py:115-122, tlbo/model/base_gp.py:48-64,148-151, tlbo/utils/normalization.py:4-28).
//
// Per BO iteration the facades build 1 + 5 training sets from the same (X, y): all rows, and all rows
// except one contiguous validation range.  Each set gets the all-equal guard(s), the facade-level
// normalisation, the inactive-value imputation and the GP-level standardisation -- a few dozen numpy
// calls on 50-element arrays per set, ~0.09 ms each in the interpreter, on the critical path in front of
// the device fit.  This helper writes all of them straight into the padded upload buffer.
//
// Every floating-point result is bit-identical to numpy's: `np.mean` / `np.std` reduce with numpy's
// pairwise summation (8 interleaved partial sums up to 128 elements, recursive halving above), the
// variance is mean(|x - mean|^2) with the mean divided first, and each elementwise operation is a single
// IEEE operation (this file is compiled for the host without FMA contraction).
// tests/test_host_logic.py checks it against the numpy statement on thousands of random cases.
namespace {

double np_pairwise_sum(const double* a, long n) {
  if (n < 8) {
    double res = 0.0;
    for (long i = 0; i < n; i++) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    long i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  }
  long n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}
inline double np_sum(const double* a, long n) { return 0.0 + np_pairwise_sum(a, n); }   // add.reduce starts from the identity
inline double np_mean(const double* a, long n) { return np_sum(a, n) / (double)n; }
double np_std(const double* a, long n, double* tmp) {
  const double m = np_sum(a, n) / (double)n;
  for (long i = 0; i < n; i++) { const double x = a[i] - m; tmp[i] = x * x; }
  return sqrt(np_sum(tmp, n) / (double)n);
}

}  // namespace

extern "C" int tlbo_host_prepare_jobs(const double* h_X, double* h_y_inout, int n, int d, int n_jobs,
                                      const int32_t* h_skip_lo, const int32_t* h_skip_hi,
                                      const int32_t* h_outer_guard, int normalize, int normalize_y, int ncap,
                                      double* h_Xout, double* h_yout, int32_t* h_nout, double* h_ymean,
                                      double* h_ystd) {
  if (!h_X || !h_y_inout || !h_skip_lo || !h_skip_hi || !h_Xout || !h_yout || !h_nout || !h_ymean || !h_ystd ||
      n <= 0 || d <= 0 || n_jobs <= 0 || ncap < n || normalize < 0 || normalize > 2) {
    tlbo_set_error("tlbo_host_prepare_jobs: bad arguments");
    return TLBO_E_BADARG;
  }
  double* tmp = new double[2 * (size_t)n];
  double* yc = tmp + n;
  for (int b = 0; b < n_jobs; b++) {
    const int lo = h_skip_lo[b], hi = h_skip_hi[b];
    if (lo < 0 || hi < lo || hi > n || (hi - lo) >= n) {
      delete[] tmp;
      tlbo_set_error("tlbo_host_prepare_jobs: bad validation range");
      return TLBO_E_BADARG;
    }
    const int m = n - (hi - lo);
    auto src = [&](int i) { return i < lo ? i : i + (hi - lo); };      // kept row i -> row of (X, y)
    // caller-side guard on the ORIGINAL y (rgpe.py:62-63): y[rows] all equal -> y[rows[0]] += 1e-4
    if (h_outer_guard && h_outer_guard[b]) {
      bool alleq = true;
      const double y0 = h_y_inout[src(0)];
      for (int i = 1; i < m && alleq; i++) alleq = (h_y_inout[src(i)] == y0);
      if (alleq) h_y_inout[src(0)] += 1e-4;
    }
    const bool whole = (hi == lo);             // the job's y IS the caller's array (no fancy-index copy)
    for (int i = 0; i < m; i++) yc[i] = h_y_inout[src(i)];
    // build_single_surrogate (base_facade.py:96-104)
    if (normalize == 1 || normalize == 2) {
      bool alleq = true;
      for (int i = 1; i < m && alleq; i++) alleq = (yc[i] == yc[0]);
      if (alleq) {
        yc[0] += 1e-4;
        if (whole) h_y_inout[0] += 1e-4;       // in-place on the caller's array
      }
      if (normalize == 1) {
        const double mean = np_mean(yc, m), sd = np_std(yc, m, tmp);
        for (int i = 0; i < m; i++) yc[i] = (yc[i] - mean) / sd;
      } else {
        double mn = yc[0], mx = yc[0];
        for (int i = 1; i < m; i++) { mn = yc[i] < mn ? yc[i] : mn; mx = yc[i] > mx ? yc[i] : mx; }
        for (int i = 0; i < m; i++) yc[i] = (yc[i] - mn) / (mx - mn);
      }
    }
    // GaussianProcess._train: standardise again (base_gp.py:48-64); std == 0 -> 1
    double ym = 0.0, ys = 1.0;
    if (normalize_y) {
      ym = np_mean(yc, m);
      ys = np_std(yc, m, tmp);
      if (ys == 0.0) ys = 1.0;
      for (int i = 0; i < m; i++) yc[i] = (yc[i] - ym) / ys;
    }
    h_ymean[b] = ym; h_ystd[b] = ys; h_nout[b] = m;
    double* yo = h_yout + (size_t)b * ncap;
    double* Xo = h_Xout + (size_t)b * ncap * d;
    for (int i = 0; i < m; i++) {
      yo[i] = yc[i];
      const double* xr = h_X + (size_t)src(i) * d;
      for (int p = 0; p < d; p++) { const double v = xr[p]; Xo[(size_t)i * d + p] = isfinite(v) ? v : -1.0; }
    }
    for (int i = m; i < ncap; i++) {
      yo[i] = 0.0;
      for (int p = 0; p < d; p++) Xo[(size_t)i * d + p] = 0.0;
    }
  }
  delete[] tmp;
  return 0;
}

// Completion of a small-query launch: the kernel overwrites every EI row (pre-filled with TLBO_EI_PENDING) with one
// 8-byte store into mapped pinned memory.  Spinning on the rows costs ~1 us after the last store lands;
// cudaStreamSynchronize adds a driver round trip to every hill-climb step.  The stream is queried now and then so
// that a failed launch surfaces as its CUDA error instead of a hang.
static inline bool ei_pending(const volatile double* p) {
  uint64_t bits;
  const double v = *p;
  memcpy(&bits, &v, sizeof(bits));
  return bits == TLBO_EI_PENDING_BITS;
}

static int wait_for_rows(const double* h_ei, int rows, cudaStream_t s) {
  int j = 0;
  for (unsigned spin = 1;; spin++) {
    while (j < rows && !ei_pending(h_ei + j)) j++;
    if (j == rows) return 0;
    if ((spin & 0x3ff) == 0) {
      cudaError_t e = cudaStreamQuery(s);
      if (e == cudaSuccess) {
        while (j < rows && !ei_pending(h_ei + j)) j++;
        if (j == rows) return 0;
        tlbo_set_error("small-query launch finished without reporting every row");
        return TLBO_E_BADARG;
      }
      if (e != cudaErrorNotReady) { tlbo_set_error(cudaGetErrorString(e)); return (int)e; }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// One hill-climb of the online scenario's local search (reference tlbo/optimizer/ei_optimization.py:239-294,
// LocalSearch._one_iter) driven from native code: per step the one-exchange neighbourhood of the incumbent
// (generator seed = the next pre-drawn rng.randint(0, 10000)), ONE launch scoring the whole neighbourhood
// (tlbo_ensemble_ei_small: every (row, surrogate) posterior, ensemble, EI), a spin on the EI rows as they land in
// mapped memory, and the reference's accept rule (first neighbour with EI > incumbent, NaN -> -DBL_MAX as
// AbstractAcquisitionFunction.__call__ does).  A climb is ~25 such steps; in the interpreter a step costs ~50 us
// of Python on top of the launches.  Buffers (all caller-owned): h_X [cap, d] host staging for the neighbourhood (it
// travels inside the kernel arguments), h_ei [cap] MAPPED pinned host memory (the kernel stores EI over the bus),
// d_mu / d_var [M * cap] device scratch, d_counter cap zeroed device uint32 (left zero).  Re-entrant across threads with their own buffers.
// Returns 0 when the climb ended (no improving neighbour or max_steps), 1 when the pre-drawn seeds ran out
// (call again with more seeds: the incumbent / counters are up to date), < 0 / CUDA codes on error.
extern "C" int tlbo_local_search_climb(const tlbo_pack* view, const int32_t* h_types, const uint8_t* h_const_mask,
                                       const double* d_w, const int32_t* d_skip, const int32_t* d_order, int mode,
                                       double eta, double par, double var_floor, int num_neighbors, double stdev,
                                       const uint32_t* h_seeds, int n_seeds, int max_steps, double* h_config,
                                       double* h_acq, double* h_X, double* d_mu, double* d_var, double* h_ei,
                                       uint32_t* d_counter, int cap, int32_t* h_counters, void* stream) {
  if (!view || !h_types || !h_seeds || !h_config || !h_acq || !h_X || !d_mu || !d_var || !h_ei ||
      !d_counter || !h_counters || cap <= 0) {
    tlbo_set_error("tlbo_local_search_climb: bad arguments");
    return TLBO_E_BADARG;
  }
  const int d = view->d;
  cudaStream_t s = (cudaStream_t)stream;
  // h_counters: [0] seeds consumed (= hill-climb steps), [1] neighbours the lazy reference loop inspects,
  //             [2] acquisition evaluations, [3] rows with EI < 0 (the caller raises), [4] steps of THIS climb so far
  int used = 0;
  for (;;) {
    if (used >= n_seeds) { h_counters[0] += used; return 1; }
    int rows = 0;
    int rc = tlbo_one_exchange_neighbourhood(h_config, d, h_types, h_const_mask, h_seeds[used], num_neighbors, stdev,
                                             h_X, cap, &rows);
    if (rc) return rc;
    used++;
    h_counters[4] += 1;
    bool changed = false;
    if (rows > 0) {
      for (long i = 0; i < (long)rows * d; i++)
        if (!isfinite(h_X[i])) h_X[i] = -1.0;               // EI._to_device / GaussianProcess._impute_inactive
      // one launch: every (neighbour, surrogate) posterior, the ensemble and EI; the rows go up inside the kernel
      // arguments, EI comes back through mapped pinned host memory (no copy-engine hop, no synchronisation call)
      {
        double pending;
        const uint64_t bits = TLBO_EI_PENDING_BITS;
        memcpy(&pending, &bits, sizeof(pending));
        for (int j = 0; j < rows; j++) h_ei[j] = pending;
      }
      rc = tlbo_ensemble_ei_small(view, h_types, d_w, d_skip, d_order, mode, h_X, h_X, rows, eta, par, var_floor, d_mu,
                                  d_var, h_ei, d_counter, stream);
      if (rc) return rc;
      rc = wait_for_rows(h_ei, rows, s);
      if (rc) return rc;
      h_counters[2] += 1;
      int first = -1, negative = 0;
      for (int j = 0; j < rows; j++) negative += (h_ei[j] < 0.0) ? 1 : 0;             // EI._compute raises on these
      if (negative) { h_counters[3] += negative; h_counters[0] += used; return 0; }
      for (int j = 0; j < rows; j++) {
        double a = h_ei[j];
        if (a != a) a = -DBL_MAX;                                                      // acquisition.py:72-76
        if (a > *h_acq) { first = j; break; }
      }
      h_counters[1] += (first >= 0) ? first + 1 : rows;
      if (first >= 0) {
        for (int p = 0; p < d; p++) h_config[p] = h_X[(size_t)first * d + p];
        *h_acq = h_ei[first];
        changed = true;
      }
    }
    if (!changed || (max_steps >= 0 && h_counters[4] == max_steps)) { h_counters[0] += used; return 0; }
  }
}
