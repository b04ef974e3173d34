// Bound-constrained limited-memory BFGS (L-BFGS-B 3.0 semantics) for tiny problems.
//
// Replaces, on device, the 11 x scipy.optimize.fmin_l_bfgs_b runs of
// GaussianProcess._optimize (reference tlbo/model/gp.py:199-248; scipy defaults
// m=10, factr=1e7, pgtol=1e-5, maxls=20, maxfun=maxiter=15000).
//
// Design: the GP hyper-parameter vector has H = d+2 <= 24 entries, so instead of
// the compact (W, M) representation of Byrd/Lu/Nocedal/Zhu the limited-memory
// matrix  B = theta*I - W M W^T  is materialised as a dense H x H matrix (the m
// most recent BFGS updates applied to theta*I -- identical in exact arithmetic,
// Byrd/Nocedal/Schnabel 1994 Thm 2.3).  The generalised Cauchy point, the
// subspace minimisation with the v3.0 projection step (Morales/Nocedal 2011), the
// More'-Thuente line search (dcsrch/dcstep, ftol=1e-3, gtol=0.9, xtol=0.1) and
// every termination test follow the published algorithm step by step, so the
// iterates track scipy's to rounding.  One thread runs this state machine; the
// surrounding CTA evaluates f and g (reverse communication).
//
// The header is plain C++ (TLBO_HD expands to __host__ __device__ under nvcc) so
// the very same code is compiled for the host by tests/hostsim and checked
// against scipy without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TLBO_HD __host__ __device__ __forceinline__
#define TLBO_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TLBO_HD inline
#define TLBO_HD_NOINLINE inline
#endif

#define LB_NMAX 24
#define LB_M 10

enum { LB_EVAL = 0, LB_DONE = 1 };
enum {
  LB_CONV_PGTOL = 0,      // CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL
  LB_CONV_FACTR = 1,      // CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH
  LB_ABNORMAL = 2,        // ABNORMAL_TERMINATION_IN_LNSRCH
  LB_STOP_MAXITER = 3,    // STOP: TOTAL NO. of ITERATIONS REACHED LIMIT
  LB_STOP_MAXFUN = 4,     // STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT
  LB_ERROR = 5
};

struct LbfgsbState {
  int n;
  double x[LB_NMAX], l[LB_NMAX], u[LB_NMAX];
  double f, g[LB_NMAX];
  double t[LB_NMAX], r[LB_NMAX], fold;      // previous iterate, previous gradient
  double d[LB_NMAX], z[LB_NMAX];
  double S[LB_M][LB_NMAX], Y[LB_M][LB_NMAX];
  double B[LB_NMAX][LB_NMAX];
  double Wk[LB_NMAX][LB_NMAX];   // scratch for the reduced Cholesky (kept off the stack)
  double theta;
  int col;
  int iwhere[LB_NMAX];
  // line search
  double stp, stpmx, gd, gdold, dnorm, dtd;
  int ifun, iback;
  // dcsrch state
  int ls_brackt, ls_stage;
  double ls_ginit, ls_gtest, ls_gx, ls_gy, ls_finit, ls_fx, ls_fy, ls_stx, ls_sty, ls_stmin, ls_stmax,
      ls_width, ls_width1;
  // driver
  int phase;          // 0: waiting for f(x0); 1: in line search
  int iter, nfgv, nskip;
  int status;
  double sbgnrm;
  double factr, pgtol;
  int maxiter, maxfun, maxls;
  double f_last;      // last evaluated f (what scipy's python driver reports as `fun`)
};

namespace lb {

TLBO_HD double dmax(double a, double b) { return a > b ? a : b; }
TLBO_HD double dmin(double a, double b) { return a < b ? a : b; }
#define LB_EPSMCH 2.220446049250313e-16

// projgr: infinity norm of the projected gradient
TLBO_HD double projgr(const LbfgsbState& s) {
  double nrm = 0.0;
  for (int i = 0; i < s.n; i++) {
    double gi = s.g[i];
    if (gi < 0.0) gi = dmax(s.x[i] - s.u[i], gi);
    else gi = dmin(s.x[i] - s.l[i], gi);
    nrm = dmax(nrm, fabs(gi));
  }
  return nrm;
}

// Rebuild the dense limited-memory matrix from the stored pairs (oldest first).
TLBO_HD void rebuild_B(LbfgsbState& s) {
  const int n = s.n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) s.B[i][j] = (i == j) ? s.theta : 0.0;
  double Bs[LB_NMAX];
  for (int k = 0; k < s.col; k++) {
    const double* sk = s.S[k];
    const double* yk = s.Y[k];
    double sBs = 0.0, ys = 0.0;
    for (int i = 0; i < n; i++) {
      double a = 0.0;
      for (int j = 0; j < n; j++) a += s.B[i][j] * sk[j];
      Bs[i] = a;
    }
    for (int i = 0; i < n; i++) { sBs += sk[i] * Bs[i]; ys += yk[i] * sk[i]; }
    const double i1 = 1.0 / sBs, i2 = 1.0 / ys;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) s.B[i][j] += yk[i] * yk[j] * i2 - Bs[i] * Bs[j] * i1;
  }
}

TLBO_HD void reset_memory(LbfgsbState& s) {
  s.col = 0;
  s.theta = 1.0;
  rebuild_B(s);
}

// Generalised Cauchy point (subroutine cauchy).  Result in s.z; s.iwhere updated.
TLBO_HD void cauchy(LbfgsbState& s) {
  const int n = s.n;
  double* xcp = s.z;
  if (s.sbgnrm <= 0.0) {
    for (int i = 0; i < n; i++) xcp[i] = s.x[i];
    return;
  }
  double dd[LB_NMAX], tb[LB_NMAX];
  bool hasb[LB_NMAX];
  int nbreak = 0;
  for (int i = 0; i < n; i++) {
    const double neggi = -s.g[i];
    const double tl = s.x[i] - s.l[i], tu = s.u[i] - s.x[i];
    const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
    int iw = 0;
    if (xlower) { if (neggi <= 0.0) iw = 1; }
    else if (xupper) { if (neggi >= 0.0) iw = 2; }
    else { if (fabs(neggi) <= 0.0) iw = -3; }
    s.iwhere[i] = iw;
    hasb[i] = false;
    if (iw != 0) {
      dd[i] = 0.0;
    } else {
      dd[i] = neggi;
      if (neggi < 0.0) { tb[i] = tl / (-neggi); hasb[i] = true; nbreak++; }
      else if (neggi > 0.0) { tb[i] = tu / neggi; hasb[i] = true; nbreak++; }
    }
    xcp[i] = s.x[i];
  }
  bool anyd = false;
  for (int i = 0; i < n; i++) anyd = anyd || (dd[i] != 0.0);
  if (!anyd) return;

  // zc = xcp - x accumulated so far (fixed variables only; free ones move by tsum*d)
  double zfix[LB_NMAX];
  for (int i = 0; i < n; i++) zfix[i] = 0.0;
  double f1 = 0.0, f2 = 0.0;
  double Bd[LB_NMAX];
  for (int i = 0; i < n; i++) f1 -= dd[i] * dd[i];
  for (int i = 0; i < n; i++) {
    double a = 0.0;
    for (int j = 0; j < n; j++) a += s.B[i][j] * dd[j];
    Bd[i] = a;
  }
  for (int i = 0; i < n; i++) f2 += dd[i] * Bd[i];
  const double f2_org = -s.theta * f1;
  double dtm = -f1 / f2;
  double tsum = 0.0;
  const int nbreak0 = nbreak;
  int nleft = nbreak;
  double tj = 0.0;
  bool vertex = false;
  while (nleft > 0) {
    // smallest remaining breakpoint
    int ibp = -1;
    for (int i = 0; i < n; i++)
      if (hasb[i] && (ibp < 0 || tb[i] < tb[ibp])) ibp = i;
    const double tj0 = tj;
    tj = tb[ibp];
    const double dt = tj - tj0;
    if (dtm < dt) break;
    tsum += dt;
    nleft--;
    hasb[ibp] = false;
    const double dibp = dd[ibp];
    dd[ibp] = 0.0;
    if (dibp > 0.0) { zfix[ibp] = s.u[ibp] - s.x[ibp]; xcp[ibp] = s.u[ibp]; s.iwhere[ibp] = 2; }
    else { zfix[ibp] = s.l[ibp] - s.x[ibp]; xcp[ibp] = s.l[ibp]; s.iwhere[ibp] = 1; }
    if (nleft == 0 && nbreak0 == n) { dtm = dt; vertex = true; break; }
    // derivatives of the quadratic along the new piece: f1 = g'd + z'Bd, f2 = d'Bd
    double zc[LB_NMAX];
    for (int i = 0; i < n; i++) zc[i] = (dd[i] != 0.0) ? tsum * dd[i] : zfix[i];
    f1 = 0.0; f2 = 0.0;
    for (int i = 0; i < n; i++) {
      double a = 0.0;
      for (int j = 0; j < n; j++) a += s.B[i][j] * dd[j];
      Bd[i] = a;
    }
    for (int i = 0; i < n; i++) { f1 += (s.g[i] * dd[i] + zc[i] * Bd[i]); f2 += dd[i] * Bd[i]; }
    f2 = dmax(LB_EPSMCH * f2_org, f2);
    if (nleft > 0) dtm = -f1 / f2;
    else { dtm = 0.0; }   // no direction left (bnded) -- f1 = f2 = 0 in the original
  }
  if (vertex) return;
  if (dtm <= 0.0) dtm = 0.0;
  tsum += dtm;
  for (int i = 0; i < n; i++) xcp[i] += tsum * dd[i];
}

// Dense Cholesky solve of the free-variable block.  Returns false if not PD.
TLBO_HD bool solve_free(LbfgsbState& s, const int* ind, int nsub, const double* rhs, double* out) {
  double (*A)[LB_NMAX] = s.Wk;
  for (int a = 0; a < nsub; a++)
    for (int b = 0; b <= a; b++) A[a][b] = s.B[ind[a]][ind[b]];
  for (int k = 0; k < nsub; k++) {
    double p = A[k][k];
    for (int q = 0; q < k; q++) p -= A[k][q] * A[k][q];
    if (!(p > 0.0)) return false;
    p = sqrt(p);
    A[k][k] = p;
    for (int i = k + 1; i < nsub; i++) {
      double v = A[i][k];
      for (int q = 0; q < k; q++) v -= A[i][q] * A[k][q];
      A[i][k] = v / p;
    }
  }
  for (int i = 0; i < nsub; i++) {
    double v = rhs[i];
    for (int q = 0; q < i; q++) v -= A[i][q] * out[q];
    out[i] = v / A[i][i];
  }
  for (int i = nsub - 1; i >= 0; i--) {
    double v = out[i];
    for (int q = i + 1; q < nsub; q++) v -= A[q][i] * out[q];
    out[i] = v / A[i][i];
  }
  return true;
}

// Subspace minimisation (cmprlb + subsm, v3.0).  s.z enters as the Cauchy point
// and leaves as the subspace minimiser.  Returns false on a non-PD reduced matrix.
TLBO_HD bool subsm(LbfgsbState& s) {
  const int n = s.n;
  int ind[LB_NMAX];
  int nsub = 0;
  for (int i = 0; i < n; i++)
    if (s.iwhere[i] <= 0) ind[nsub++] = i;
  if (nsub == 0 || s.col == 0) return true;
  double zc[LB_NMAX], rr[LB_NMAX], dsub[LB_NMAX], xp[LB_NMAX];
  for (int i = 0; i < n; i++) zc[i] = s.z[i] - s.x[i];
  for (int a = 0; a < nsub; a++) {
    const int k = ind[a];
    double v = 0.0;
    for (int j = 0; j < n; j++) v += s.B[k][j] * zc[j];
    rr[a] = -(s.g[k] + v);
  }
  if (!solve_free(s, ind, nsub, rr, dsub)) return false;
  // projection step
  int iword = 0;
  for (int i = 0; i < n; i++) xp[i] = s.z[i];
  for (int a = 0; a < nsub; a++) {
    const int k = ind[a];
    double xk = dmax(s.l[k], s.z[k] + dsub[a]);
    xk = dmin(s.u[k], xk);
    s.z[k] = xk;
    if (xk == s.l[k] || xk == s.u[k]) iword = 1;
  }
  if (iword == 0) return true;
  double dd_p = 0.0;
  for (int i = 0; i < n; i++) dd_p += (s.z[i] - s.x[i]) * s.g[i];
  if (dd_p > 0.0) {
    for (int i = 0; i < n; i++) s.z[i] = xp[i];
    double alpha = 1.0, temp1 = 1.0;
    int ibd = -1;
    for (int a = 0; a < nsub; a++) {
      const int k = ind[a];
      const double dk = dsub[a];
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
      const double dk = dsub[ibd];
      const int k = ind[ibd];
      if (dk > 0.0) { s.z[k] = s.u[k]; dsub[ibd] = 0.0; }
      else if (dk < 0.0) { s.z[k] = s.l[k]; dsub[ibd] = 0.0; }
    }
    for (int a = 0; a < nsub; a++) s.z[ind[a]] += alpha * dsub[a];
  }
  return true;
}

// dcstep (MINPACK-2)
TLBO_HD void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                    double fp, double dp, int& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = dmax(fabs(theta), dmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = dmax(fabs(theta), dmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
    else stpf = stpq;
    brackt = 1;
  } else if (fabs(dp) < fabs(dx)) {
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = dmax(fabs(theta), dmax(fabs(dx), fabs(dp)));
    gamma = s * sqrt(dmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      if (stp > stx) stpf = dmin(stp + 0.66 * (sty - stp), stpf);
      else stpf = dmax(stp + 0.66 * (sty - stp), stpf);
    } else {
      if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
      else stpf = stpq;
      stpf = dmin(stpmax, stpf);
      stpf = dmax(stpmin, stpf);
    }
  } else {
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = dmax(fabs(theta), dmax(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > fx) { sty = stp; fy = fp; dy = dp; }
  else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// dcsrch (MINPACK-2).  task: 0 = START, 1 = FG (continue).  Returns 1 = FG (evaluate at
// s.stp), 2 = CONVERGENCE, 3 = WARNING, 4 = ERROR.
TLBO_HD int dcsrch(LbfgsbState& s, double f, double g, int start) {
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
    dcstep(s.ls_stx, fxm, gxm, s.ls_sty, fym, gym, s.stp, fm, gm, s.ls_brackt, s.ls_stmin, s.ls_stmax);
    s.ls_fx = fxm + s.ls_stx * s.ls_gtest;
    s.ls_fy = fym + s.ls_sty * s.ls_gtest;
    s.ls_gx = gxm + s.ls_gtest;
    s.ls_gy = gym + s.ls_gtest;
  } else {
    dcstep(s.ls_stx, s.ls_fx, s.ls_gx, s.ls_sty, s.ls_fy, s.ls_gy, s.stp, f, g, s.ls_brackt, s.ls_stmin,
           s.ls_stmax);
  }
  if (s.ls_brackt) {
    if (fabs(s.ls_sty - s.ls_stx) >= 0.66 * s.ls_width1) s.stp = s.ls_stx + 0.5 * (s.ls_sty - s.ls_stx);
    s.ls_width1 = s.ls_width;
    s.ls_width = fabs(s.ls_sty - s.ls_stx);
  }
  if (s.ls_brackt) {
    s.ls_stmin = dmin(s.ls_stx, s.ls_sty);
    s.ls_stmax = dmax(s.ls_stx, s.ls_sty);
  } else {
    s.ls_stmin = s.stp + xtrapl * (s.stp - s.ls_stx);
    s.ls_stmax = s.stp + xtrapu * (s.stp - s.ls_stx);
  }
  s.stp = dmax(s.stp, stpmin);
  s.stp = dmin(s.stp, stpmax);
  if ((s.ls_brackt && (s.stp <= s.ls_stmin || s.stp >= s.ls_stmax)) ||
      (s.ls_brackt && s.ls_stmax - s.ls_stmin <= xtol * s.ls_stmax))
    s.stp = s.ls_stx;
  return 1;
}

// Start a new iteration (label 222 .. first call of lnsrlb).  Returns LB_EVAL with the
// trial point in s.x, or LB_DONE.
TLBO_HD_NOINLINE int new_iteration(LbfgsbState& s) {
  const int n = s.n;
  for (int guard = 0; guard < 4; guard++) {
    cauchy(s);
    bool ok = true;
    int nfree = 0;
    for (int i = 0; i < n; i++) nfree += (s.iwhere[i] <= 0);
    if (nfree > 0 && s.col > 0) ok = subsm(s);
    if (!ok) { reset_memory(s); continue; }   // non-PD: refresh memory and restart the iteration
    for (int i = 0; i < n; i++) s.d[i] = s.z[i] - s.x[i];
    // ---- lnsrlb, first entry
    s.dtd = 0.0;
    for (int i = 0; i < n; i++) s.dtd += s.d[i] * s.d[i];
    s.dnorm = sqrt(s.dtd);
    s.stpmx = 1.0e10;
    if (s.iter == 0) {
      s.stpmx = 1.0;
    } else {
      for (int i = 0; i < n; i++) {
        const double a1 = s.d[i];
        if (a1 < 0.0) {
          const double a2 = s.l[i] - s.x[i];
          if (a2 >= 0.0) s.stpmx = 0.0;
          else if (a1 * s.stpmx < a2) s.stpmx = a2 / a1;
        } else if (a1 > 0.0) {
          const double a2 = s.u[i] - s.x[i];
          if (a2 <= 0.0) s.stpmx = 0.0;
          else if (a1 * s.stpmx > a2) s.stpmx = a2 / a1;
        }
      }
    }
    s.stp = 1.0;   // boxed problem: every hyper-parameter has both bounds
    for (int i = 0; i < n; i++) { s.t[i] = s.x[i]; s.r[i] = s.g[i]; }
    s.fold = s.f;
    s.ifun = 0;
    s.iback = 0;
    s.gd = 0.0;
    for (int i = 0; i < n; i++) s.gd += s.g[i] * s.d[i];
    s.gdold = s.gd;
    int lsret = 4;
    if (!(s.gd >= 0.0)) lsret = dcsrch(s, s.f, s.gd, 1);
    if (s.gd >= 0.0 /* info = -4 */ ) {
      if (s.col == 0) { s.status = LB_ABNORMAL; return LB_DONE; }
      reset_memory(s);
      continue;
    }
    (void)lsret;   // an ERROR return is treated like FG by lnsrlb (csave is neither CONV nor WARN)
    s.ifun = 1;
    s.nfgv += 1;
    s.iback = 0;
    if (s.stp == 1.0) { for (int i = 0; i < n; i++) s.x[i] = s.z[i]; }
    else { for (int i = 0; i < n; i++) s.x[i] = s.stp * s.d[i] + s.t[i]; }
    s.phase = 1;
    return LB_EVAL;
  }
  s.status = LB_ERROR;
  return LB_DONE;
}

}  // namespace lb

// Initialise: x0 is clipped into [l, u] (scipy clips before the first evaluation).
TLBO_HD void lbfgsb_init(LbfgsbState& s, int n, const double* x0, const double* l, const double* u) {
  s.n = n;
  for (int i = 0; i < n; i++) {
    s.l[i] = l[i]; s.u[i] = u[i];
    double v = x0[i];
    if (v < l[i]) v = l[i];
    if (v > u[i]) v = u[i];
    s.x[i] = v;
    s.iwhere[i] = 0;
  }
  s.factr = 1.0e7; s.pgtol = 1.0e-5; s.maxiter = 15000; s.maxfun = 15000; s.maxls = 20;
  s.col = 0; s.theta = 1.0;
  s.iter = 0; s.nfgv = 0; s.nskip = 0; s.phase = 0; s.status = LB_ERROR;
  s.ls_brackt = 0; s.ls_stage = 1;
  s.ls_ginit = s.ls_gtest = s.ls_gx = s.ls_gy = s.ls_finit = s.ls_fx = s.ls_fy = 0.0;
  s.ls_stx = s.ls_sty = s.ls_stmin = s.ls_stmax = s.ls_width = s.ls_width1 = 0.0;
  s.stp = 1.0; s.stpmx = 1.0; s.gd = s.gdold = 0.0; s.f = s.fold = s.f_last = 0.0;
  lb::rebuild_B(s);
}

// Feed f(s.x), g(s.x).  Returns LB_EVAL (evaluate at the new s.x and call again) or
// LB_DONE (s.x is the result, s.f_last the reported objective, s.status why).
TLBO_HD_NOINLINE int lbfgsb_feed(LbfgsbState& s, double f, const double* g) {
  const int n = s.n;
  s.f = f;
  s.f_last = f;
  for (int i = 0; i < n; i++) s.g[i] = g[i];
  if (s.phase == 0) {
    s.nfgv = 1;
    s.sbgnrm = lb::projgr(s);
    if (s.sbgnrm <= s.pgtol) { s.status = LB_CONV_PGTOL; return LB_DONE; }
    return lb::new_iteration(s);
  }
  // ---- inside the line search (lnsrlb re-entry, label 556)
  s.gd = 0.0;
  for (int i = 0; i < n; i++) s.gd += s.g[i] * s.d[i];
  const int lsret = lb::dcsrch(s, f, s.gd, 0);
  if (lsret != 2 && lsret != 3) {
    // another trial point is wanted
    s.ifun += 1;
    s.nfgv += 1;
    s.iback = s.ifun - 1;
    if (s.iback >= s.maxls) {
      // line search failed: restore the previous iterate
      for (int i = 0; i < n; i++) { s.x[i] = s.t[i]; s.g[i] = s.r[i]; }
      s.f = s.fold;
      if (s.col == 0) {
        s.nfgv -= 1;
        s.iter += 1;
        s.status = LB_ABNORMAL;
        return LB_DONE;
      }
      s.nfgv -= 1;
      lb::reset_memory(s);
      return lb::new_iteration(s);
    }
    if (s.stp == 1.0) { for (int i = 0; i < n; i++) s.x[i] = s.z[i]; }
    else { for (int i = 0; i < n; i++) s.x[i] = s.stp * s.d[i] + s.t[i]; }
    return LB_EVAL;
  }
  // ---- NEW_X
  s.iter += 1;
  s.sbgnrm = lb::projgr(s);
  // (scipy's python driver: maxiter / maxfun checks happen on NEW_X, before the tests below)
  if (s.iter >= s.maxiter) { s.status = LB_STOP_MAXITER; return LB_DONE; }
  if (s.nfgv > s.maxfun) { s.status = LB_STOP_MAXFUN; return LB_DONE; }
  if (s.sbgnrm <= s.pgtol) { s.status = LB_CONV_PGTOL; return LB_DONE; }
  const double ddum0 = lb::dmax(fabs(s.fold), lb::dmax(fabs(s.f), 1.0));
  if ((s.fold - s.f) <= LB_EPSMCH * s.factr * ddum0) { s.status = LB_CONV_FACTR; return LB_DONE; }
  // ---- BFGS pair
  double rr = 0.0, dr, ddum;
  for (int i = 0; i < n; i++) { s.r[i] = s.g[i] - s.r[i]; rr += s.r[i] * s.r[i]; }
  if (s.stp == 1.0) { dr = s.gd - s.gdold; ddum = -s.gdold; }
  else {
    dr = (s.gd - s.gdold) * s.stp;
    for (int i = 0; i < n; i++) s.d[i] *= s.stp;
    ddum = -s.gdold * s.stp;
  }
  if (dr <= LB_EPSMCH * ddum) {
    s.nskip += 1;
  } else {
    if (s.col == LB_M) {
      for (int k = 0; k + 1 < LB_M; k++)
        for (int i = 0; i < n; i++) { s.S[k][i] = s.S[k + 1][i]; s.Y[k][i] = s.Y[k + 1][i]; }
      s.col = LB_M - 1;
    }
    for (int i = 0; i < n; i++) { s.S[s.col][i] = s.d[i]; s.Y[s.col][i] = s.r[i]; }
    s.col += 1;
    s.theta = rr / dr;
    lb::rebuild_B(s);
  }
  return lb::new_iteration(s);
}
