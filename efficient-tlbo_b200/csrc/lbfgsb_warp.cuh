// Warp-cooperative L-BFGS-B (device only): the same algorithm, state machine and
// termination tests as lbfgsb_dense.cuh (which stays the host-testable statement of the
// optimiser, checked against scipy in tests/test_lbfgsb_host.py), with the linear algebra
// spread over the 32 lanes of one warp:
//   * dense limited-memory matrix rebuild (m rank-2 updates)      lane = column of B
//   * B*v products in the Cauchy search / subspace minimisation   lane = row, symmetric access
//   * reduced Cholesky factorisation and the two triangular solves lane = row
//   * all dot products                                             xor-butterfly (identical in every lane)
// Scalar logic (line search dcsrch/dcstep, break-point bookkeeping, projection step) runs on
// lane 0 over the shared-memory state and is published with __syncwarp().
//
// ncu on the single-thread version showed 63% of the fit kernel's wall time inside this
// state machine with 255 threads parked at the barrier (profiles/r1a_*); this version is
// the remedy.  Replaces scipy.optimize.fmin_l_bfgs_b of reference tlbo/model/gp.py:241-248.
#pragma once
#include "lbfgsb_dense.cuh"

#define LW_LD (LB_NMAX + 1)

struct LbwState {
  int n;
  double x[LB_NMAX], l[LB_NMAX], u[LB_NMAX];
  double f, g[LB_NMAX];
  double t[LB_NMAX], r[LB_NMAX], fold;
  double d[LB_NMAX], z[LB_NMAX];
  double S[LB_M][LB_NMAX], Y[LB_M][LB_NMAX];
  double B[LB_NMAX][LB_NMAX];          // symmetric; lanes read column `lane` (= row) conflict-free
  double Wk[LB_NMAX][LW_LD];           // reduced Cholesky factor (lower)
  // per-variable scratch of the Cauchy search / subspace step
  double dd[LB_NMAX], tb[LB_NMAX], zfix[LB_NMAX], rr[LB_NMAX], dsub[LB_NMAX], xp[LB_NMAX];
  int hasb[LB_NMAX], ind[LB_NMAX], iwhere[LB_NMAX];
  double theta;
  int col;
  double stp, stpmx, gd, gdold, dnorm, dtd;
  int ifun, iback;
  int ls_brackt, ls_stage;
  double ls_ginit, ls_gtest, ls_gx, ls_gy, ls_finit, ls_fx, ls_fy, ls_stx, ls_sty, ls_stmin, ls_stmax, ls_width,
      ls_width1;
  int phase, iter, nfgv, nskip, status;
  double sbgnrm, factr, pgtol;
  int maxiter, maxfun, maxls;
  double f_last;
  int flag;                             // lane-0 -> warp broadcast slot
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

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int bcast_flag(LbwState& s, int lane, int v) {
  if (lane == 0) s.flag = v;
  __syncwarp();
  const int r = s.flag;
  __syncwarp();
  return r;
}

// (B v)_lane with v in shared memory; B symmetric so column `lane` is read (conflict-free)
__device__ __forceinline__ double matvec_row(const LbwState& s, const double* v, int lane) {
  double a = 0.0;
  if (lane < s.n)
    for (int j = 0; j < s.n; j++) a += s.B[j][lane] * v[j];
  return a;
}

__device__ __forceinline__ double projgr(const LbwState& s, int lane) {
  double v = 0.0;
  if (lane < s.n) {
    double gi = s.g[lane];
    if (gi < 0.0) gi = lb::dmax(s.x[lane] - s.u[lane], gi);
    else gi = lb::dmin(s.x[lane] - s.l[lane], gi);
    v = fabs(gi);
  }
  return wmax(v);
}

// Column `lane` of B stays in registers across the col rank-2 updates (no shared-memory
// round trip per pair); written back once.  NMAX (12 or 24) bounds the unrolled loops.
template <int NMAX>
__device__ void rebuild_B_t(LbwState& s, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  double Bc[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; j++) Bc[j] = (j == lane) ? s.theta : 0.0;
  for (int k = 0; k < s.col; k++) {
    const double* sk = s.S[k];
    const double* yk = s.Y[k];
    double skv[NMAX];
#pragma unroll
    for (int j = 0; j < NMAX; j++) skv[j] = (j < n) ? sk[j] : 0.0;
    // (B s_k)_lane, summed in the order of the single-thread statement
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < NMAX; j++) a += Bc[j] * skv[j];
    if (!act) a = 0.0;
    const double skl = act ? sk[lane] : 0.0, ykl = act ? yk[lane] : 0.0;
    double p1 = skl * a, p2 = ykl * skl;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    const double i1 = 1.0 / p1, i2 = 1.0 / p2;            // 1 / s'Bs, 1 / y's
    const double c1 = ykl * i2, c2 = a * i1;
    // entry (j, lane) += y_j y_lane / y's - (Bs)_j (Bs)_lane / s'Bs ; (Bs)_j from lane j
#pragma unroll
    for (int j = 0; j < NMAX; j++) {
      const double bsj = __shfl_sync(0xffffffffu, a, j);
      const double ykj = (j < n) ? yk[j] : 0.0;
      Bc[j] += ykj * ykl * i2 - bsj * a * i1;
    }
    (void)c1; (void)c2;
  }
  __syncwarp();
  if (act) {
#pragma unroll
    for (int j = 0; j < NMAX; j++)
      if (j < n) s.B[j][lane] = Bc[j];
  }
  __syncwarp();
}

__device__ __forceinline__ void rebuild_B(LbwState& s, int lane) {
  if (s.n <= 12) rebuild_B_t<12>(s, lane);
  else rebuild_B_t<LB_NMAX>(s, lane);
}

__device__ __forceinline__ void reset_memory(LbwState& s, int lane) {
  if (lane == 0) { s.col = 0; s.theta = 1.0; }
  __syncwarp();
  rebuild_B(s, lane);
}

// Generalised Cauchy point; result in s.z, s.iwhere updated.
__device__ void cauchy(LbwState& s, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  if (s.sbgnrm <= 0.0) {
    if (act) s.z[lane] = s.x[lane];
    __syncwarp();
    return;
  }
  double dd = 0.0, tb = 0.0, zfix = 0.0;
  bool hasb = false;
  if (act) {
    const double neggi = -s.g[lane];
    const double tl = s.x[lane] - s.l[lane], tu = s.u[lane] - s.x[lane];
    const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
    int iw = 0;
    if (xlower) { if (neggi <= 0.0) iw = 1; }
    else if (xupper) { if (neggi >= 0.0) iw = 2; }
    else { if (fabs(neggi) <= 0.0) iw = -3; }
    s.iwhere[lane] = iw;
    if (iw == 0) {
      dd = neggi;
      if (neggi < 0.0) { tb = tl / (-neggi); hasb = true; }
      else if (neggi > 0.0) { tb = tu / neggi; hasb = true; }
    }
    s.z[lane] = s.x[lane];
    s.dd[lane] = dd;
  }
  __syncwarp();
  const int nbreak0 = __popc(__ballot_sync(0xffffffffu, hasb));
  if (__ballot_sync(0xffffffffu, dd != 0.0) == 0u) return;

  double f1 = -wsum(dd * dd);
  double Bd = matvec_row(s, s.dd, lane);
  double f2 = wsum(dd * Bd);
  const double f2_org = -s.theta * f1;
  double dtm = -f1 / f2, tsum = 0.0, tj = 0.0;
  int nleft = nbreak0;
  bool vertex = false;
  const double gl = act ? s.g[lane] : 0.0;
  while (nleft > 0) {
    // smallest remaining break point, lowest index on ties
    double bt = hasb ? tb : INFINITY;
    int bi = hasb ? lane : 64;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ot < bt || (ot == bt && oi < bi)) { bt = ot; bi = oi; }
    }
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
      if (dibp > 0.0) { zfix = s.u[lane] - s.x[lane]; s.z[lane] = s.u[lane]; s.iwhere[lane] = 2; }
      else { zfix = s.l[lane] - s.x[lane]; s.z[lane] = s.l[lane]; s.iwhere[lane] = 1; }
      s.dd[lane] = 0.0;
    }
    __syncwarp();
    if (nleft == 0 && nbreak0 == n) { dtm = dt; vertex = true; break; }
    const double zc = (dd != 0.0) ? tsum * dd : zfix;
    Bd = matvec_row(s, s.dd, lane);
    f1 = wsum(gl * dd + zc * Bd);
    f2 = wsum(dd * Bd);
    f2 = lb::dmax(LB_EPSMCH * f2_org, f2);
    if (nleft > 0) dtm = -f1 / f2;
    else dtm = 0.0;
  }
  if (!vertex) {
    if (dtm <= 0.0) dtm = 0.0;
    tsum += dtm;
    if (act) s.z[lane] += tsum * dd;
  }
  __syncwarp();
}

// Cholesky solve of the free-variable block (lane = row).  Returns false if not PD.
__device__ bool solve_free(LbwState& s, int nsub, int lane) {
  const bool act = lane < nsub;
  if (act) {
    const int ka = s.ind[lane];
    for (int b = 0; b <= lane; b++) s.Wk[lane][b] = s.B[s.ind[b]][ka];
  }
  __syncwarp();
  bool ok = true;
  for (int k = 0; k < nsub; k++) {
    double v = 0.0;
    if (act && lane >= k) {
      v = s.Wk[lane][k];
      for (int q = 0; q < k; q++) v -= s.Wk[lane][q] * s.Wk[k][q];
    }
    const double pk = __shfl_sync(0xffffffffu, v, k);
    if (!(pk > 0.0)) { ok = false; break; }
    const double rp = rsqrt(pk);                 // 1 / L_kk
    if (act && lane == k) { s.Wk[k][k] = pk * rp; s.xp[k] = rp; }
    else if (act && lane > k) s.Wk[lane][k] = v * rp;
    __syncwarp();
  }
  if (!ok) return false;
  // forward substitution (column oriented; same subtraction order per row as the scalar code)
  double v = act ? s.rr[lane] : 0.0;
  for (int q = 0; q < nsub; q++) {
    double o = 0.0;
    if (lane == q) o = v * s.xp[q];
    o = __shfl_sync(0xffffffffu, o, q);
    if (lane == q) v = o;
    else if (act && lane > q) v -= s.Wk[lane][q] * o;
  }
  // back substitution with L^T
  for (int i = nsub - 1; i >= 0; i--) {
    double o = 0.0;
    if (lane == i) o = v * s.xp[i];
    o = __shfl_sync(0xffffffffu, o, i);
    if (lane == i) v = o;
    else if (act && lane < i) v -= s.Wk[i][lane] * o;
  }
  if (act) s.dsub[lane] = v;
  __syncwarp();
  return true;
}

// Subspace minimisation with the v3.0 projection step.  s.z: Cauchy point in, minimiser out.
__device__ bool subsm(LbwState& s, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  const bool freev = act && s.iwhere[lane] <= 0;
  const unsigned fm = __ballot_sync(0xffffffffu, freev);
  const int nsub = __popc(fm);
  if (nsub == 0 || s.col == 0) return true;
  const int pos = __popc(fm & ((1u << lane) - 1u));
  if (freev) s.ind[pos] = lane;
  if (act) s.xp[lane] = s.z[lane] - s.x[lane];       // zc
  __syncwarp();
  if (lane < nsub) {
    const int k = s.ind[lane];
    double v = 0.0;
    for (int j = 0; j < n; j++) v += s.B[j][k] * s.xp[j];
    s.rr[lane] = -(s.g[k] + v);
  }
  __syncwarp();
  if (!solve_free(s, nsub, lane)) return false;
  // projection step: every free lane clamps its own variable; the rare back-tracking branch
  // (a clamp was active AND the projected direction is not a descent direction) keeps the
  // exact sequential rule on lane 0
  const double zl = act ? s.z[lane] : 0.0;
  double znew = zl;
  bool hit = false;
  if (freev) {
    double xk = lb::dmax(s.l[lane], zl + s.dsub[pos]);
    xk = lb::dmin(s.u[lane], xk);
    znew = xk;
    hit = (xk == s.l[lane] || xk == s.u[lane]);
  }
  if (__ballot_sync(0xffffffffu, hit) == 0u) {
    if (freev) s.z[lane] = znew;
    __syncwarp();
    return true;
  }
  const double dd_p = wsum(act ? (znew - s.x[lane]) * s.g[lane] : 0.0);
  if (!(dd_p > 0.0)) {
    if (freev) s.z[lane] = znew;
    __syncwarp();
    return true;
  }
  if (lane == 0) {
    // z still holds the Cauchy point here
    double alpha = 1.0, temp1 = 1.0;
    int ibd = -1;
    for (int a = 0; a < nsub; a++) {
      const int k = s.ind[a];
      const double dk = s.dsub[a];
      if (dk < 0.0) {
        const double temp2 = s.l[k] - s.z[k];
        if (temp2 >= 0.0) temp1 = 0.0;
        else if (dk * alpha < temp2) temp1 = temp2 / dk;
      } else if (dk > 0.0) {
        const double temp2 = s.u[k] - s.z[k];
        if (temp2 <= 0.0) temp1 = 0.0;
        else if (dk * alpha > temp2) temp1 = temp2 / dk;
      }
      if (temp1 < alpha) { alpha = temp1; ibd = a; }
    }
    if (alpha < 1.0) {
      const double dk = s.dsub[ibd];
      const int k = s.ind[ibd];
      if (dk > 0.0) { s.z[k] = s.u[k]; s.dsub[ibd] = 0.0; }
      else if (dk < 0.0) { s.z[k] = s.l[k]; s.dsub[ibd] = 0.0; }
    }
    for (int a = 0; a < nsub; a++) s.z[s.ind[a]] += alpha * s.dsub[a];
  }
  __syncwarp();
  return true;
}

// dcsrch on lane 0 over the shared state (same code path as the scalar optimiser).
__device__ __forceinline__ int dcsrch_lane0(LbwState& s, double f, double g, int start) {
  const double ftol = 1.0e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = s.stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (start) {
    if (s.stp < stpmin) return 4;
    if (s.stp > stpmax) return 4;
    if (g >= 0.0) return 4;
    s.ls_brackt = 0;
    s.ls_stage = 1;
    s.ls_finit = f;
    s.ls_ginit = g;
    s.ls_gtest = ftol * g;
    s.ls_width = stpmax - stpmin;
    s.ls_width1 = s.ls_width / 0.5;
    s.ls_stx = 0.0; s.ls_fx = f; s.ls_gx = g;
    s.ls_sty = 0.0; s.ls_fy = f; s.ls_gy = g;
    s.ls_stmin = 0.0;
    s.ls_stmax = s.stp + xtrapu * s.stp;
    return 1;
  }
  const double ftest = s.ls_finit + s.stp * s.ls_gtest;
  if (s.ls_stage == 1 && f <= ftest && g >= 0.0) s.ls_stage = 2;
  int ret = 1;
  if (s.ls_brackt && (s.stp <= s.ls_stmin || s.stp >= s.ls_stmax)) ret = 3;
  if (s.ls_brackt && s.ls_stmax - s.ls_stmin <= xtol * s.ls_stmax) ret = 3;
  if (s.stp == stpmax && f <= ftest && g <= s.ls_gtest) ret = 3;
  if (s.stp == stpmin && (f > ftest || g >= s.ls_gtest)) ret = 3;
  if (f <= ftest && fabs(g) <= gtol * (-s.ls_ginit)) ret = 2;
  if (ret != 1) return ret;
  if (s.ls_stage == 1 && f <= s.ls_fx && f > ftest) {
    const double fm = f - s.stp * s.ls_gtest;
    double fxm = s.ls_fx - s.ls_stx * s.ls_gtest;
    double fym = s.ls_fy - s.ls_sty * s.ls_gtest;
    const double gm = g - s.ls_gtest;
    double gxm = s.ls_gx - s.ls_gtest;
    double gym = s.ls_gy - s.ls_gtest;
    lb::dcstep(s.ls_stx, fxm, gxm, s.ls_sty, fym, gym, s.stp, fm, gm, s.ls_brackt, s.ls_stmin, s.ls_stmax);
    s.ls_fx = fxm + s.ls_stx * s.ls_gtest;
    s.ls_fy = fym + s.ls_sty * s.ls_gtest;
    s.ls_gx = gxm + s.ls_gtest;
    s.ls_gy = gym + s.ls_gtest;
  } else {
    lb::dcstep(s.ls_stx, s.ls_fx, s.ls_gx, s.ls_sty, s.ls_fy, s.ls_gy, s.stp, f, g, s.ls_brackt, s.ls_stmin,
               s.ls_stmax);
  }
  if (s.ls_brackt) {
    if (fabs(s.ls_sty - s.ls_stx) >= 0.66 * s.ls_width1) s.stp = s.ls_stx + 0.5 * (s.ls_sty - s.ls_stx);
    s.ls_width1 = s.ls_width;
    s.ls_width = fabs(s.ls_sty - s.ls_stx);
  }
  if (s.ls_brackt) {
    s.ls_stmin = lb::dmin(s.ls_stx, s.ls_sty);
    s.ls_stmax = lb::dmax(s.ls_stx, s.ls_sty);
  } else {
    s.ls_stmin = s.stp + xtrapl * (s.stp - s.ls_stx);
    s.ls_stmax = s.stp + xtrapu * (s.stp - s.ls_stx);
  }
  s.stp = lb::dmax(s.stp, stpmin);
  s.stp = lb::dmin(s.stp, stpmax);
  if ((s.ls_brackt && (s.stp <= s.ls_stmin || s.stp >= s.ls_stmax)) ||
      (s.ls_brackt && s.ls_stmax - s.ls_stmin <= xtol * s.ls_stmax))
    s.stp = s.ls_stx;
  return 1;
}

// Start a new iteration (Cauchy point, subspace step, first entry of the line search).
__device__ __noinline__ int new_iteration(LbwState& s, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  for (int guard = 0; guard < 4; guard++) {
    LBW_T0();
    cauchy(s, lane);
    LBW_MARK(1);
    const int nfree = __popc(__ballot_sync(0xffffffffu, act && s.iwhere[lane] <= 0));
    bool ok = true;
    if (nfree > 0 && s.col > 0) ok = subsm(s, lane);
    LBW_MARK(2);
    if (!ok) { reset_memory(s, lane); continue; }
    const double dl = act ? s.z[lane] - s.x[lane] : 0.0;
    const double gl = act ? s.g[lane] : 0.0;
    const double dtd = wsum(dl * dl);
    const double gd = wsum(gl * dl);
    // stpmx: the sequential rule is a running minimum of the per-variable bound ratios
    double stpmx = 1.0e10;
    if (s.iter == 0) stpmx = 1.0;
    else {
      // every lane forms its bound ratio (the divisions run in parallel); lane 0 then applies the
      // exact sequential rule of lnsrlb (the comparison uses the running value)
      double a2 = 0.0, ratio = 0.0;
      if (act) {
        a2 = (dl < 0.0) ? s.l[lane] - s.x[lane] : s.u[lane] - s.x[lane];
        ratio = (dl != 0.0) ? a2 / dl : 0.0;
        s.d[lane] = dl; s.dd[lane] = a2; s.tb[lane] = ratio;
      }
      __syncwarp();
      if (lane == 0) {
        double m = 1.0e10;
        for (int i = 0; i < n; i++) {
          const double a1 = s.d[i], b2 = s.dd[i];
          if (a1 < 0.0) {
            if (b2 >= 0.0) m = 0.0;
            else if (a1 * m < b2) m = s.tb[i];
          } else if (a1 > 0.0) {
            if (b2 <= 0.0) m = 0.0;
            else if (a1 * m > b2) m = s.tb[i];
          }
        }
        s.stpmx = m;
      }
      __syncwarp();
      stpmx = s.stpmx;
    }
    if (act) { s.d[lane] = dl; s.t[lane] = s.x[lane]; s.r[lane] = s.g[lane]; }
    int code = 0;      // 0: evaluate, 1: abnormal done, 2: reset and retry
    if (lane == 0) {
      s.dtd = dtd;
      s.dnorm = sqrt(dtd);
      s.stpmx = stpmx;
      s.stp = 1.0;
      s.fold = s.f;
      s.ifun = 0;
      s.iback = 0;
      s.gd = gd;
      s.gdold = gd;
      if (!(gd >= 0.0)) (void)dcsrch_lane0(s, s.f, gd, 1);
      if (gd >= 0.0) {
        if (s.col == 0) { s.status = LB_ABNORMAL; code = 1; }
        else code = 2;
      } else {
        s.ifun = 1;
        s.nfgv += 1;
        s.iback = 0;
        s.phase = 1;
      }
    }
    code = bcast_flag(s, lane, code);
    LBW_MARK(3);
    if (code == 1) return LB_DONE;
    if (code == 2) { reset_memory(s, lane); continue; }
    if (act) {
      if (s.stp == 1.0) s.x[lane] = s.z[lane];
      else s.x[lane] = s.stp * s.d[lane] + s.t[lane];
    }
    __syncwarp();
    return LB_EVAL;
  }
  if (lane == 0) s.status = LB_ERROR;
  __syncwarp();
  return LB_DONE;
}

}  // namespace lbw

// All 32 lanes of one warp call these with identical arguments.
__device__ inline void lbw_init(LbwState& s, int n, const double* x0, const double* l, const double* u, int lane) {
  if (lane == 0) {
    s.n = n;
    s.factr = 1.0e7; s.pgtol = 1.0e-5; s.maxiter = 15000; s.maxfun = 15000; s.maxls = 20;
    s.col = 0; s.theta = 1.0;
    s.iter = 0; s.nfgv = 0; s.nskip = 0; s.phase = 0; s.status = LB_ERROR;
    s.ls_brackt = 0; s.ls_stage = 1;
    s.ls_ginit = s.ls_gtest = s.ls_gx = s.ls_gy = s.ls_finit = s.ls_fx = s.ls_fy = 0.0;
    s.ls_stx = s.ls_sty = s.ls_stmin = s.ls_stmax = s.ls_width = s.ls_width1 = 0.0;
    s.stp = 1.0; s.stpmx = 1.0; s.gd = s.gdold = 0.0; s.f = s.fold = s.f_last = 0.0;
    s.sbgnrm = 0.0; s.flag = 0;
  }
  if (lane < n) {
    s.l[lane] = l[lane]; s.u[lane] = u[lane];
    double v = x0[lane];
    if (v < l[lane]) v = l[lane];
    if (v > u[lane]) v = u[lane];
    s.x[lane] = v;
    s.iwhere[lane] = 0;
  }
  __syncwarp();
  lbw::rebuild_B(s, lane);
}

__device__ __noinline__ int lbw_feed(LbwState& s, double f, const double* g, int lane) {
  const int n = s.n;
  const bool act = lane < n;
  LBW_T0();
  if (lane == 0) { s.f = f; s.f_last = f; }
  if (act) s.g[lane] = g[lane];
  __syncwarp();
  if (s.phase == 0) {
    const double nrm = lbw::projgr(s, lane);
    int done = 0;
    if (lane == 0) {
      s.nfgv = 1;
      s.sbgnrm = nrm;
      if (nrm <= s.pgtol) { s.status = LB_CONV_PGTOL; done = 1; }
    }
    done = lbw::bcast_flag(s, lane, done);
    if (done) return LB_DONE;
    return lbw::new_iteration(s, lane);
  }
  // ---- inside the line search
  const double gd = lbw::wsum(act ? s.g[lane] * s.d[lane] : 0.0);
  int code = 0;   // 0: another trial point, 1: done (abnormal), 2: reset + new iteration, 3: NEW_X
  if (lane == 0) {
    s.gd = gd;
    const int lsret = lbw::dcsrch_lane0(s, f, gd, 0);
    if (lsret != 2 && lsret != 3) {
      s.ifun += 1;
      s.nfgv += 1;
      s.iback = s.ifun - 1;
      if (s.iback >= s.maxls) {
        for (int i = 0; i < n; i++) { s.x[i] = s.t[i]; s.g[i] = s.r[i]; }
        s.f = s.fold;
        s.nfgv -= 1;
        if (s.col == 0) { s.iter += 1; s.status = LB_ABNORMAL; code = 1; }
        else code = 2;
      }
    } else code = 3;
  }
  code = lbw::bcast_flag(s, lane, code);
  if (code == 1) return LB_DONE;
  if (code == 2) { lbw::reset_memory(s, lane); return lbw::new_iteration(s, lane); }
  if (code == 0) {
    if (act) {
      if (s.stp == 1.0) s.x[lane] = s.z[lane];
      else s.x[lane] = s.stp * s.d[lane] + s.t[lane];
    }
    __syncwarp();
    return LB_EVAL;
  }
  // ---- NEW_X
  const double nrm = lbw::projgr(s, lane);
  const double rl = act ? s.g[lane] - s.r[lane] : 0.0;
  const double rr = lbw::wsum(rl * rl);
  int done = 0, store = 0;
  if (lane == 0) {
    s.iter += 1;
    s.sbgnrm = nrm;
    if (s.iter >= s.maxiter) { s.status = LB_STOP_MAXITER; done = 1; }
    else if (s.nfgv > s.maxfun) { s.status = LB_STOP_MAXFUN; done = 1; }
    else if (nrm <= s.pgtol) { s.status = LB_CONV_PGTOL; done = 1; }
    else {
      const double ddum0 = lb::dmax(fabs(s.fold), lb::dmax(fabs(s.f), 1.0));
      if ((s.fold - s.f) <= LB_EPSMCH * s.factr * ddum0) { s.status = LB_CONV_FACTR; done = 1; }
    }
    if (!done) {
      double dr, ddum;
      if (s.stp == 1.0) { dr = s.gd - s.gdold; ddum = -s.gdold; }
      else { dr = (s.gd - s.gdold) * s.stp; ddum = -s.gdold * s.stp; }
      if (dr <= LB_EPSMCH * ddum) s.nskip += 1;
      else {
        store = 1;
        if (s.col == LB_M) store = 2;      // drop the oldest pair first
        s.theta = rr / dr;
      }
    }
  }
  done = lbw::bcast_flag(s, lane, done ? 1 : (store ? 2 + store : 0));
  if (done == 1) return LB_DONE;
  if (act) {
    s.r[lane] = rl;
    if (s.stp != 1.0) s.d[lane] *= s.stp;
  }
  if (done >= 3) {
    if (done == 4) {
      if (act)
        for (int k = 0; k + 1 < LB_M; k++) { s.S[k][lane] = s.S[k + 1][lane]; s.Y[k][lane] = s.Y[k + 1][lane]; }
      if (lane == 0) s.col = LB_M - 1;
      __syncwarp();
    }
    if (act) { s.S[s.col][lane] = s.d[lane]; s.Y[s.col][lane] = rl; }
    __syncwarp();
    if (lane == 0) s.col += 1;
    __syncwarp();
    LBW_MARK(4);
    lbw::rebuild_B(s, lane);
    LBW_MARK(0);
  } else {
    __syncwarp();
  }
  return lbw::new_iteration(s, lane);
}
