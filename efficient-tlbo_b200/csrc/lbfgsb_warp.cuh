// Warp-resident L-BFGS-B (device only): the same algorithm, state machine and termination tests as
// lbfgsb_dense.cuh (which stays the host-testable statement of the optimiser, checked against scipy in
// tests/test_lbfgsb_host.py), organised for LATENCY: one step of the optimiser sits between two
// dependent NLL evaluations of the fit kernel, with the other seven warps of the CTA parked at a barrier.
//
// Round-1 version: 24 k cycles per step (rebuild_B 5.8 k, Cauchy 2.6 k, subspace step 8.3 k, line-search
// bookkeeping 5.3 k; profiles/README.md).  What this version does differently:
//   * the problem size is a template parameter (NMAX = 12 or 24): every loop over variables is unrolled,
//     every vector a lane needs in full (s_k, y_k, B s, d, the pivot column ...) is read with 128-bit
//     broadcast loads from shared memory into registers -- a "gather" is one store per lane, one
//     __syncwarp and NMAX/2 loads, and cross-lane sums are a gather plus a local, fixed-order sum
//     (identical in every lane) instead of five dependent shuffle levels;
//   * lane i keeps row i of the dense limited-memory matrix B in registers from the rebuild through the
//     Cauchy search to the subspace step (B is symmetric; it never touches shared memory);
//   * scalar logic (dcsrch, break-point bookkeeping, stpmx, convergence tests) is executed redundantly by
//     all lanes on identical data: no lane-0 sections, no flag broadcasts, no divergence;
//   * the reduced Cholesky of the subspace step runs on the FULL matrix with identity rows / columns at
//     the fixed variables (bit-identical arithmetic for the free block: the fixed rows contribute exact
//     zeros), right-looking and register-resident with the forward substitution folded in;
//   * the (s, y) pairs live in a ring buffer (no shifting), 1 / y's is stored with the pair, and the
//     reciprocals on the rebuild's critical chain use the rcp.approx seed + 2 Newton steps (<= 1 ulp).
//
// Replaces scipy.optimize.fmin_l_bfgs_b of reference tlbo/model/gp.py:241-248.
#pragma once
#include "lbfgsb_dense.cuh"

#define LBW_SCR 4

struct __align__(16) LbwState {
  // lane-indexed vectors, zero beyond n (all 16-byte aligned: read back with 128-bit loads)
  double x[LB_NMAX], l[LB_NMAX], u[LB_NMAX], g[LB_NMAX];
  double t[LB_NMAX], r[LB_NMAX], d[LB_NMAX], z[LB_NMAX];
  double S[LB_M][LB_NMAX], Y[LB_M][LB_NMAX];     // ring buffer: pair k (oldest first) at (head + k) % LB_M
  double scr[LBW_SCR][LB_NMAX];                   // rotating gather scratch
  double ysinv[LB_M + 2];                         // 1 / y_k's_k
  double L[LB_NMAX][LB_NMAX + 1];                 // reduced Cholesky factor (rows), for the back substitution
  double theta, f, fold, f_last;
  double stp, stpmx, gd, gdold, dnorm, dtd, sbgnrm;
  double ls_ginit, ls_gtest, ls_gx, ls_gy, ls_finit, ls_fx, ls_fy, ls_stx, ls_sty, ls_stmin, ls_stmax, ls_width,
      ls_width1;
  int n, col, head, ifun, iback, ls_brackt, ls_stage, phase, iter, nfgv, nskip, status;
};

#ifdef TLBO_PHASE_TIMING
#define LBW_T0() long long _lt = clock64()
#define LBW_MARK(idx) do { const long long _n = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) g_feed_cycles[idx] += _n - _lt; _lt = _n; } while (0)
__device__ long long g_feed_cycles[8];
#else
#define LBW_T0() do {} while (0)
#define LBW_MARK(idx) do {} while (0)
#endif

namespace lbw {

// Scalars carried in registers during one call (identical in every lane); lane 0 writes them back.
struct Regs {
  double theta, f, fold, stp, stpmx, gd, gdold, dnorm, dtd, sbgnrm;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
  int col, head, ifun, iback, brackt, stage, phase, iter, nfgv, nskip, status;
};

__device__ __forceinline__ void load_regs(const LbwState& s, Regs& R) {
  R.theta = s.theta; R.f = s.f; R.fold = s.fold; R.stp = s.stp; R.stpmx = s.stpmx; R.gd = s.gd; R.gdold = s.gdold;
  R.dnorm = s.dnorm; R.dtd = s.dtd; R.sbgnrm = s.sbgnrm;
  R.ginit = s.ls_ginit; R.gtest = s.ls_gtest; R.gx = s.ls_gx; R.gy = s.ls_gy; R.finit = s.ls_finit; R.fx = s.ls_fx;
  R.fy = s.ls_fy; R.stx = s.ls_stx; R.sty = s.ls_sty; R.stmin = s.ls_stmin; R.stmax = s.ls_stmax;
  R.width = s.ls_width; R.width1 = s.ls_width1;
  R.col = s.col; R.head = s.head; R.ifun = s.ifun; R.iback = s.iback; R.brackt = s.ls_brackt; R.stage = s.ls_stage;
  R.phase = s.phase; R.iter = s.iter; R.nfgv = s.nfgv; R.nskip = s.nskip; R.status = s.status;
}
__device__ __forceinline__ void store_regs(LbwState& s, const Regs& R, int lane) {
  if (lane == 0) {
    s.theta = R.theta; s.f = R.f; s.fold = R.fold; s.stp = R.stp; s.stpmx = R.stpmx; s.gd = R.gd; s.gdold = R.gdold;
    s.dnorm = R.dnorm; s.dtd = R.dtd; s.sbgnrm = R.sbgnrm;
    s.ls_ginit = R.ginit; s.ls_gtest = R.gtest; s.ls_gx = R.gx; s.ls_gy = R.gy; s.ls_finit = R.finit; s.ls_fx = R.fx;
    s.ls_fy = R.fy; s.ls_stx = R.stx; s.ls_sty = R.sty; s.ls_stmin = R.stmin; s.ls_stmax = R.stmax;
    s.ls_width = R.width; s.ls_width1 = R.width1;
    s.col = R.col; s.head = R.head; s.ifun = R.ifun; s.iback = R.iback; s.ls_brackt = R.brackt; s.ls_stage = R.stage;
    s.phase = R.phase; s.iter = R.iter; s.nfgv = R.nfgv; s.nskip = R.nskip; s.status = R.status;
  }
  __syncwarp();
}

// 1/x from the hardware seed and two Newton steps (<= 1 ulp; ~56 cycles against ~112 for the IEEE division)
__device__ __forceinline__ double rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// 1/sqrt(x), x > 0: hardware seed and two coupled Newton steps
__device__ __forceinline__ double rsqrt_pos(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double h = 0.5 * r, g = x * r;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  h = fma(h, e, h);
  return 2.0 * h;
}

template <int NMAX>
__device__ __forceinline__ void ldvec(const double* p, double (&v)[NMAX]) {
#pragma unroll
  for (int j = 0; j < NMAX; j += 2) {
    const double2 q = *reinterpret_cast<const double2*>(p + j);
    v[j] = q.x; v[j + 1] = q.y;
  }
}
// every lane's `v` (zero for lanes >= n) into registers of all lanes
template <int NMAX>
__device__ __forceinline__ void gather(LbwState& s, int& slot, int lane, double v, double (&out)[NMAX]) {
  double* b = s.scr[slot];
  slot = (slot + 1) & (LBW_SCR - 1);
  if (lane < NMAX) b[lane] = v;
  __syncwarp();
  ldvec<NMAX>(b, out);
}
template <int NMAX>
__device__ __forceinline__ double vsum(const double (&v)[NMAX]) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int j = 0; j < NMAX; j += 4) { a0 += v[j]; a1 += v[j + 1]; a2 += v[j + 2]; a3 += v[j + 3]; }
  return (a0 + a1) + (a2 + a3);
}
template <int NMAX>
__device__ __forceinline__ double vdot(const double (&a)[NMAX], const double (&b)[NMAX]) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int j = 0; j < NMAX; j += 4) {
    a0 = fma(a[j], b[j], a0); a1 = fma(a[j + 1], b[j + 1], a1);
    a2 = fma(a[j + 2], b[j + 2], a2); a3 = fma(a[j + 3], b[j + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
template <int NMAX>
__device__ __forceinline__ double allsum(LbwState& s, int& slot, int lane, double v) {
  double t[NMAX];
  gather<NMAX>(s, slot, lane, v, t);
  return vsum<NMAX>(t);
}

struct StepIO { double stx, fx, dx, sty, fy, dy, stp; int brackt; };
__device__ __noinline__ StepIO dcstep_ni(StepIO io, double fp, double dp, double stpmin, double stpmax) {
  lb::dcstep(io.stx, io.fx, io.dx, io.sty, io.fy, io.dy, io.stp, fp, dp, io.brackt, stpmin, stpmax);
  return io;
}

// dcsrch over the register copy of the line-search state (every lane, identical data).
__device__ __forceinline__ int dcsrch(Regs& R, double f, double g, int start) {
  const double ftol = 1.0e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = R.stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (start) {
    if (R.stp < stpmin) return 4;
    if (R.stp > stpmax) return 4;
    if (g >= 0.0) return 4;
    R.brackt = 0;
    R.stage = 1;
    R.finit = f;
    R.ginit = g;
    R.gtest = ftol * g;
    R.width = stpmax - stpmin;
    R.width1 = R.width * 2.0;                // (width / 0.5, exactly)
    R.stx = 0.0; R.fx = f; R.gx = g;
    R.sty = 0.0; R.fy = f; R.gy = g;
    R.stmin = 0.0;
    R.stmax = R.stp + xtrapu * R.stp;
    return 1;
  }
  const double ftest = R.finit + R.stp * R.gtest;
  if (R.stage == 1 && f <= ftest && g >= 0.0) R.stage = 2;
  int ret = 1;
  if (R.brackt && (R.stp <= R.stmin || R.stp >= R.stmax)) ret = 3;
  if (R.brackt && R.stmax - R.stmin <= xtol * R.stmax) ret = 3;
  if (R.stp == stpmax && f <= ftest && g <= R.gtest) ret = 3;
  if (R.stp == stpmin && (f > ftest || g >= R.gtest)) ret = 3;
  if (f <= ftest && fabs(g) <= gtol * (-R.ginit)) ret = 2;
  if (ret != 1) return ret;
  {
    // ONE out-of-line dcstep (a fifth of the optimiser's code; only line-search continuations get here):
    // the modified-function variant of stage 1 differs in its arguments only
    const bool modified = (R.stage == 1 && f <= R.fx && f > ftest);
    StepIO io;
    io.stx = R.stx; io.sty = R.sty; io.stp = R.stp; io.brackt = R.brackt;
    io.fx = R.fx; io.dx = R.gx; io.fy = R.fy; io.dy = R.gy;
    double fp = f, dp = g;
    if (modified) {
      fp = f - R.stp * R.gtest;
      io.fx = R.fx - R.stx * R.gtest;
      io.fy = R.fy - R.sty * R.gtest;
      dp = g - R.gtest;
      io.dx = R.gx - R.gtest;
      io.dy = R.gy - R.gtest;
    }
    io = dcstep_ni(io, fp, dp, R.stmin, R.stmax);
    R.stx = io.stx; R.sty = io.sty; R.stp = io.stp; R.brackt = io.brackt;
    if (modified) {
      R.fx = io.fx + R.stx * R.gtest;
      R.fy = io.fy + R.sty * R.gtest;
      R.gx = io.dx + R.gtest;
      R.gy = io.dy + R.gtest;
    } else {
      R.fx = io.fx; R.fy = io.fy; R.gx = io.dx; R.gy = io.dy;
    }
  }
  if (R.brackt) {
    if (fabs(R.sty - R.stx) >= 0.66 * R.width1) R.stp = R.stx + 0.5 * (R.sty - R.stx);
    R.width1 = R.width;
    R.width = fabs(R.sty - R.stx);
  }
  if (R.brackt) {
    R.stmin = lb::dmin(R.stx, R.sty);
    R.stmax = lb::dmax(R.stx, R.sty);
  } else {
    R.stmin = R.stp + xtrapl * (R.stp - R.stx);
    R.stmax = R.stp + xtrapu * (R.stp - R.stx);
  }
  R.stp = lb::dmax(R.stp, stpmin);
  R.stp = lb::dmin(R.stp, stpmax);
  if ((R.brackt && (R.stp <= R.stmin || R.stp >= R.stmax)) || (R.brackt && R.stmax - R.stmin <= xtol * R.stmax))
    R.stp = R.stx;
  return 1;
}

// infinity norm of the projected gradient (lane-owned x, g, l, u)
template <int NMAX>
__device__ __forceinline__ double projgr(LbwState& s, int& slot, int lane, bool act, double xl, double gl, double ll,
                                         double ul) {
  double v = 0.0;
  if (act) {
    double gi = gl;
    if (gi < 0.0) gi = lb::dmax(xl - ul, gi);
    else gi = lb::dmin(xl - ll, gi);
    v = fabs(gi);
  }
  double t[NMAX];
  gather<NMAX>(s, slot, lane, v, t);
  double m0 = 0.0, m1 = 0.0;
#pragma unroll
  for (int j = 0; j < NMAX; j += 2) { m0 = lb::dmax(m0, t[j]); m1 = lb::dmax(m1, t[j + 1]); }
  return lb::dmax(m0, m1);
}

// Row `lane` of the dense limited-memory matrix: theta*I followed by the stored rank-2 updates, oldest first.
template <int NMAX>
__device__ __forceinline__ void rebuild_B(LbwState& s, const Regs& R, int& slot, int lane, bool act,
                                          double (&Bc)[NMAX]) {
#pragma unroll
  for (int j = 0; j < NMAX; j++) Bc[j] = (act && j == lane) ? R.theta : 0.0;
#pragma unroll 1
  for (int k = 0; k < R.col; k++) {
    int p = R.head + k;
    if (p >= LB_M) p -= LB_M;
    double sk[NMAX], yk[NMAX], bs[NMAX];
    ldvec<NMAX>(s.S[p], sk);
    ldvec<NMAX>(s.Y[p], yk);
    const double ykl = (lane < NMAX) ? s.Y[p][lane] : 0.0;
    const double c1 = ykl * s.ysinv[p];
    const double a = vdot<NMAX>(Bc, sk);                 // (B s_k)_lane
    gather<NMAX>(s, slot, lane, a, bs);
    const double sBs = vdot<NMAX>(sk, bs);
    const double c2 = a * rcp(sBs);
#pragma unroll
    for (int j = 0; j < NMAX; j++) Bc[j] = fma(-bs[j], c2, fma(yk[j], c1, Bc[j]));
  }
}

// Back-tracking branch of the v3.0 projection step (subsm): the sequential rule of the published code over
// the free variables in index order, on shared-memory copies; identical in every lane.  Returns alpha.
__device__ __noinline__ double project_backtrack(const LbwState& s, const double* dv, const double* zv, unsigned fm,
                                                 int n, int* ibd_out) {
  double alpha = 1.0, temp1 = 1.0;
  int ibd = -1;
  for (int a = 0; a < n; a++) {
    if ((fm >> a) & 1u) {
      const double dk = dv[a];
      if (dk < 0.0) {
        const double temp2 = s.l[a] - zv[a];
        if (temp2 >= 0.0) temp1 = 0.0;
        else if (dk * alpha < temp2) temp1 = temp2 / dk;
      } else if (dk > 0.0) {
        const double temp2 = s.u[a] - zv[a];
        if (temp2 <= 0.0) temp1 = 0.0;
        else if (dk * alpha > temp2) temp1 = temp2 / dk;
      }
      if (temp1 < alpha) { alpha = temp1; ibd = a; }
    }
  }
  *ibd_out = ibd;
  return alpha;
}

// One iteration start: rebuild B, generalised Cauchy point, subspace minimisation (v3.0 projection step),
// first entry of the line search.  xl / gl: the lane's current iterate and gradient (s.g holds g).
template <int NMAX>
__device__ __forceinline__ int new_iteration(LbwState& s, Regs& R, int& slot, int lane, double xl, double gl) {
  const int n = s.n;
  const bool act = lane < n;
  const double ll = act ? s.l[lane] : 0.0, ul = act ? s.u[lane] : 0.0;
  LBW_T0();
#pragma unroll 1
  for (int guard = 0; guard < 4; guard++) {
    double Bc[NMAX];
    rebuild_B<NMAX>(s, R, slot, lane, act, Bc);
    LBW_MARK(0);
    // ---- generalised Cauchy point
    double zl = xl;
    int iw = 0;
    if (R.sbgnrm > 0.0) {
      double dd = 0.0, tb = 0.0, zfix = 0.0;
      bool hasb = false;
      if (act) {
        const double neggi = -gl;
        const double tl = xl - ll, tu = ul - xl;
        const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
        if (xlower) { if (neggi <= 0.0) iw = 1; }
        else if (xupper) { if (neggi >= 0.0) iw = 2; }
        else { if (fabs(neggi) <= 0.0) iw = -3; }
        if (iw == 0) {
          dd = neggi;
          if (neggi < 0.0) { tb = tl / (-neggi); hasb = true; }
          else if (neggi > 0.0) { tb = tu / neggi; hasb = true; }
        }
      }
      const int nbreak0 = __popc(__ballot_sync(0xffffffffu, hasb));
      if (__ballot_sync(0xffffffffu, dd != 0.0) != 0u) {
        double ddv[NMAX], tmp[NMAX];
        gather<NMAX>(s, slot, lane, dd, ddv);
        double f1 = -vdot<NMAX>(ddv, ddv);
        double Bd = vdot<NMAX>(Bc, ddv);
        double f2 = allsum<NMAX>(s, slot, lane, dd * Bd);
        const double f2_org = -R.theta * f1;
        double dtm = -f1 / f2, tsum = 0.0, tj = 0.0;
        int nleft = nbreak0;
        bool vertex = false;
        while (nleft > 0) {
          // smallest remaining break point, lowest index on ties
          gather<NMAX>(s, slot, lane, hasb ? tb : INFINITY, tmp);
          double bt = INFINITY;
          int bi = -1;
#pragma unroll
          for (int j = 0; j < NMAX; j++)
            if (tmp[j] < bt) { bt = tmp[j]; bi = j; }
          if (bi < 0) break;                               // cannot happen (nleft counts the finite entries)
          const double tj0 = tj;
          tj = bt;
          const double dt = tj - tj0;
          if (dtm < dt) break;
          tsum += dt;
          nleft--;
          if (lane == bi) {
            hasb = false;
            const double dibp = dd;
            dd = 0.0;
            if (dibp > 0.0) { zfix = ul - xl; zl = ul; iw = 2; }
            else { zfix = ll - xl; zl = ll; iw = 1; }
          }
          if (nleft == 0 && nbreak0 == n) { dtm = dt; vertex = true; break; }
          const double zc = (dd != 0.0) ? tsum * dd : zfix;
          gather<NMAX>(s, slot, lane, dd, ddv);
          Bd = vdot<NMAX>(Bc, ddv);
          f1 = allsum<NMAX>(s, slot, lane, gl * dd + zc * Bd);
          f2 = allsum<NMAX>(s, slot, lane, dd * Bd);
          f2 = lb::dmax(LB_EPSMCH * f2_org, f2);
          if (nleft > 0) dtm = -f1 / f2;
          else dtm = 0.0;
        }
        if (!vertex) {
          if (dtm <= 0.0) dtm = 0.0;
          tsum += dtm;
          if (act) zl += tsum * dd;
        }
      }
    }
    LBW_MARK(1);
    // ---- subspace minimisation over the free variables
    const bool freev = act && iw <= 0;
    const unsigned fm = __ballot_sync(0xffffffffu, freev);
    bool ok = true;
    if (fm != 0u && R.col > 0) {
      double tmp[NMAX];
      gather<NMAX>(s, slot, lane, act ? zl - xl : 0.0, tmp);
      const double rr = freev ? -(gl + vdot<NMAX>(Bc, tmp)) : 0.0;
      // Cholesky of the full matrix with identity rows / columns at the fixed variables, right-looking,
      // row `lane` in registers; the forward substitution rides along in `y`
      double w[NMAX];
#pragma unroll
      for (int j = 0; j < NMAX; j++) w[j] = (freev && ((fm >> j) & 1u)) ? Bc[j] : ((j == lane) ? 1.0 : 0.0);
      double y = rr, my_rp = 1.0;
#pragma unroll
      for (int k = 0; k < NMAX; k++) {
        if (k < n && ok) {
          // the pivot travels by shuffle so that its reciprocal square root starts while the rest of the
          // column is gathered
          const double pk = __shfl_sync(0xffffffffu, w[k], k);
          double colk[NMAX];
          gather<NMAX>(s, slot, lane, w[k], colk);         // A[j][k] from lane j (valid for j >= k)
          if (!(pk > 0.0)) ok = false;
          else {
            const double rp = rsqrt_pos(pk);
            const double lk = w[k] * rp;                   // L[lane][k] (lane >= k); lane k: sqrt(pk)
            if (lane == k) my_rp = rp;
            if (lane >= k && lane < NMAX) s.L[lane][k] = lk;
            const double yk = __shfl_sync(0xffffffffu, y, k) * rp;
            if (lane == k) y = yk;
            else if (lane > k) y = fma(-lk, yk, y);
#pragma unroll
            for (int j = k + 1; j < NMAX; j++) w[j] = fma(-lk, colk[j] * rp, w[j]);
          }
        }
      }
      if (ok) {
        __syncwarp();
        double lt[NMAX];                                    // column `lane` of L: L[q][lane], q > lane
#pragma unroll
        for (int q = 0; q < NMAX; q++) lt[q] = (lane < NMAX && q < n) ? s.L[q][lane < NMAX ? lane : 0] : 0.0;
#pragma unroll
        for (int q = NMAX - 1; q >= 0; q--) {
          if (q < n) {
            const double xq = __shfl_sync(0xffffffffu, y * my_rp, q);
            if (lane == q) y = xq;
            else if (lane < q) y = fma(-lt[q], xq, y);
          }
        }
        const double dsub = freev ? y : 0.0;
        // projection step: every free lane clamps its own variable; the rare back-tracking branch (a
        // clamp was active AND the projected direction is not a descent direction) applies the
        // sequential rule of the published code on gathered copies, identically in every lane
        double znew = zl;
        bool hit = false;
        if (freev) {
          double xk = lb::dmax(ll, zl + dsub);
          xk = lb::dmin(ul, xk);
          znew = xk;
          hit = (xk == ll || xk == ul);
        }
        if (__ballot_sync(0xffffffffu, hit) == 0u) zl = znew;
        else {
          const double dd_p = allsum<NMAX>(s, slot, lane, act ? (znew - xl) * gl : 0.0);
          if (!(dd_p > 0.0)) zl = znew;
          else {
            // (out of line: rare, and a fifth of this function's code when unrolled in place)
            double* dvs = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
            double* zvs = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
            if (lane < NMAX) { dvs[lane] = dsub; zvs[lane] = act ? zl : 0.0; }
            __syncwarp();
            int ibd = -1;
            const double alpha = project_backtrack(s, dvs, zvs, fm, n, &ibd);
            double dmine = dsub;
            if (alpha < 1.0 && lane == ibd) {
              if (dsub > 0.0) { zl = ul; dmine = 0.0; }
              else if (dsub < 0.0) { zl = ll; dmine = 0.0; }
            }
            if (freev) zl += alpha * dmine;
          }
        }
      }
    }
    LBW_MARK(2);
    if (!ok) {                                              // non-PD reduced matrix: refresh memory, retry
      R.col = 0; R.head = 0; R.theta = 1.0;
      continue;
    }
    // ---- lnsrlb, first entry
    const double dl = act ? zl - xl : 0.0;
    // every lane forms its bound ratio for stpmx (the divisions run in parallel); ONE exchange for d, the
    // bound distances and the ratios
    double a2 = 0.0, ratio = 0.0;
    if (act && R.iter != 0) {
      a2 = (dl < 0.0) ? ll - xl : ul - xl;
      ratio = (dl != 0.0) ? a2 / dl : 0.0;
    }
    double dv[NMAX], gv[NMAX], a2v[NMAX], rv[NMAX];
    {
      double* b0 = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
      double* b1 = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
      double* b2 = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
      if (lane < NMAX) { b0[lane] = dl; b1[lane] = a2; b2[lane] = ratio; }
      __syncwarp();
      ldvec<NMAX>(b0, dv);
      ldvec<NMAX>(b1, a2v);
      ldvec<NMAX>(b2, rv);
    }
    ldvec<NMAX>(s.g, gv);
    const double dtd = vdot<NMAX>(dv, dv);
    const double gd = vdot<NMAX>(gv, dv);
    double stpmx = 1.0e10;
    if (R.iter == 0) stpmx = 1.0;
    else {
      // the running rule of lnsrlb in index order on the gathered values
      double m = 1.0e10;
#pragma unroll
      for (int i = 0; i < NMAX; i++) {
        const double a1 = dv[i], b2 = a2v[i];
        if (a1 < 0.0) {
          if (b2 >= 0.0) m = 0.0;
          else if (a1 * m < b2) m = rv[i];
        } else if (a1 > 0.0) {
          if (b2 <= 0.0) m = 0.0;
          else if (a1 * m > b2) m = rv[i];
        }
      }
      stpmx = m;
    }
    R.dtd = dtd;
    R.dnorm = sqrt(dtd);
    R.stpmx = stpmx;
    R.stp = 1.0;
    R.fold = R.f;
    R.ifun = 0;
    R.iback = 0;
    R.gd = gd;
    R.gdold = gd;
    if (!(gd >= 0.0)) (void)dcsrch(R, R.f, gd, 1);
    if (gd >= 0.0) {
      if (R.col == 0) {
        R.status = LB_ABNORMAL;
        LBW_MARK(3);
        return LB_DONE;
      }
      R.col = 0; R.head = 0; R.theta = 1.0;
      continue;
    }
    R.ifun = 1;
    R.nfgv += 1;
    R.iback = 0;
    R.phase = 1;
    if (act) {
      s.d[lane] = dl; s.t[lane] = xl; s.r[lane] = gl; s.z[lane] = zl;
      s.x[lane] = (R.stp == 1.0) ? zl : R.stp * dl + xl;
    }
    LBW_MARK(3);
    return LB_EVAL;
  }
  R.status = LB_ERROR;
  return LB_DONE;
}

}  // namespace lbw

// All 32 lanes of one warp call these with identical arguments.
__device__ inline void lbw_init(LbwState& s, int n, const double* x0, const double* l, const double* u, int lane) {
  {
    double* p = reinterpret_cast<double*>(&s);
    const int nd = (int)(offsetof(LbwState, theta) / sizeof(double));
    for (int i = lane; i < nd; i += 32) p[i] = 0.0;
  }
  __syncwarp();
  if (lane == 0) {
    s.n = n;
    s.col = 0; s.head = 0; s.theta = 1.0;
    s.iter = 0; s.nfgv = 0; s.nskip = 0; s.phase = 0; s.status = LB_ERROR;
    s.ifun = 0; s.iback = 0;
    s.ls_brackt = 0; s.ls_stage = 1;
    s.ls_ginit = s.ls_gtest = s.ls_gx = s.ls_gy = s.ls_finit = s.ls_fx = s.ls_fy = 0.0;
    s.ls_stx = s.ls_sty = s.ls_stmin = s.ls_stmax = s.ls_width = s.ls_width1 = 0.0;
    s.stp = 1.0; s.stpmx = 1.0; s.gd = s.gdold = 0.0; s.f = s.fold = s.f_last = 0.0;
    s.dnorm = s.dtd = 0.0;
    s.sbgnrm = 0.0;
  }
  if (lane < n) {
    s.l[lane] = l[lane]; s.u[lane] = u[lane];
    double v = x0[lane];
    if (v < l[lane]) v = l[lane];
    if (v > u[lane]) v = u[lane];
    s.x[lane] = v;
  }
  __syncwarp();
}

template <int NMAX>
__device__ __noinline__ int lbw_feed(LbwState& s, double f, const double* g, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  const double factr = 1.0e7, pgtol = 1.0e-5;
  const int maxiter = 15000, maxfun = 15000, maxls = 20;
  LBW_T0();
  lbw::Regs R;
  lbw::load_regs(s, R);
  int slot = 0;
  double gl = act ? g[lane] : 0.0;
  double xl = act ? s.x[lane] : 0.0;
  const double ll = act ? s.l[lane] : 0.0, ul = act ? s.u[lane] : 0.0;
  if (act) s.g[lane] = gl;
  if (lane == 0) s.f_last = f;
  R.f = f;
  __syncwarp();
  int rc = LB_DONE;
  bool iterate = false;                     // start a new iteration (ONE call site of new_iteration)
  if (R.phase == 0) {
    R.nfgv = 1;
    R.sbgnrm = lbw::projgr<NMAX>(s, slot, lane, act, xl, gl, ll, ul);
    if (R.sbgnrm <= pgtol) R.status = LB_CONV_PGTOL;
    else iterate = true;
  } else {
    // ---- inside the line search
    double dvv[NMAX], gvv[NMAX];
    lbw::ldvec<NMAX>(s.d, dvv);
    lbw::ldvec<NMAX>(s.g, gvv);
    const double gd = lbw::vdot<NMAX>(gvv, dvv);
    R.gd = gd;
    const int lsret = lbw::dcsrch(R, f, gd, 0);
    if (lsret != 2 && lsret != 3) {
      R.ifun += 1;
      R.nfgv += 1;
      R.iback = R.ifun - 1;
      if (R.iback >= maxls) {
        // line search failed: restore the previous iterate
        xl = act ? s.t[lane] : 0.0;
        gl = act ? s.r[lane] : 0.0;
        if (act) { s.x[lane] = xl; s.g[lane] = gl; }
        __syncwarp();
        R.f = R.fold;
        R.nfgv -= 1;
        if (R.col == 0) { R.iter += 1; R.status = LB_ABNORMAL; }
        else { R.col = 0; R.head = 0; R.theta = 1.0; iterate = true; }
      } else {
        if (act) s.x[lane] = (R.stp == 1.0) ? s.z[lane] : R.stp * s.d[lane] + s.t[lane];
        rc = LB_EVAL;
      }
    } else {
      // ---- NEW_X
      R.iter += 1;
      R.sbgnrm = lbw::projgr<NMAX>(s, slot, lane, act, xl, gl, ll, ul);
      bool done = false;
      if (R.iter >= maxiter) { R.status = LB_STOP_MAXITER; done = true; }
      else if (R.nfgv > maxfun) { R.status = LB_STOP_MAXFUN; done = true; }
      else if (R.sbgnrm <= pgtol) { R.status = LB_CONV_PGTOL; done = true; }
      else {
        const double ddum0 = lb::dmax(fabs(R.fold), lb::dmax(fabs(R.f), 1.0));
        if ((R.fold - R.f) <= LB_EPSMCH * factr * ddum0) { R.status = LB_CONV_FACTR; done = true; }
      }
      if (!done) {
        // ---- BFGS pair
        const double rl = act ? gl - s.r[lane] : 0.0;          // y
        double sl = act ? s.d[lane] : 0.0;                      // s
        double dr, ddum;
        if (R.stp == 1.0) { dr = gd - R.gdold; ddum = -R.gdold; }
        else { dr = (gd - R.gdold) * R.stp; sl *= R.stp; ddum = -R.gdold * R.stp; }
        if (dr <= LB_EPSMCH * ddum) R.nskip += 1;
        else {
          double t1[NMAX], t2[NMAX];
          {
            double* b0 = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
            double* b1 = s.scr[slot]; slot = (slot + 1) & (LBW_SCR - 1);
            if (lane < NMAX) { b0[lane] = rl * rl; b1[lane] = rl * sl; }
            __syncwarp();
            lbw::ldvec<NMAX>(b0, t1);
            lbw::ldvec<NMAX>(b1, t2);
          }
          const double rr = lbw::vsum<NMAX>(t1), ys = lbw::vsum<NMAX>(t2);
          int p;
          if (R.col == LB_M) { p = R.head; R.head = (R.head + 1 == LB_M) ? 0 : R.head + 1; }
          else { p = R.head + R.col; if (p >= LB_M) p -= LB_M; R.col += 1; }
          if (act) { s.S[p][lane] = sl; s.Y[p][lane] = rl; }
          if (lane == 0) s.ysinv[p] = 1.0 / ys;
          R.theta = rr / dr;
          __syncwarp();
        }
        iterate = true;
      }
    }
  }
  LBW_MARK(4);
  if (iterate) rc = lbw::new_iteration<NMAX>(s, R, slot, lane, xl, gl);
  lbw::store_regs(s, R, lane);
  return rc;
}
