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
