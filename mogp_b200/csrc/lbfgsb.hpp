// lbfgsb.hpp — the optimiser behind GPy's `model.optimize()` (mogp/obsmodel.py:94, mogp_constrained.py:333) as a
// host-side state machine: paramz's default optimiser is SciPy's `fmin_l_bfgs_b(f_fp, x0, maxfun=1000, maxiter=1000)`
// = L-BFGS-B 3.0 (Byrd, Lu, Nocedal, Zhu; Morales & Nocedal) with m = 10, factr = 1e7, pgtol = 1e-5, maxls = 20 and NO
// bounds (the positivity of the hyper-parameters is handled by paramz's Logexp transform, not by the optimiser).
//
// Without bounds the algorithm reduces to: quasi-Newton direction d = -B^-1 g with B the limited-memory BFGS matrix
// (B0 = theta I, theta = y'y / s'y of the latest pair, at most m pairs), first step along -g with length
// min(1/|g|, 1e10), More'-Thuente line search (dcsrch / dcstep of MINPACK-2, ftol = 1e-3, gtol = 0.9, xtol = 0.1),
// update skipped when s'y <= eps * (-g_old'd * stp), memory dropped and the iteration restarted when the line search
// fails.  This file restates exactly that control flow (the labels in the comments are those of mainlb / lnsrlb) in
// reverse-communication form: the caller evaluates (f, g) when asked, so that one host thread can interleave the
// refits of many clusters while their evaluations run on the device.
//
// The generalized Cauchy point and the subspace minimisation of the bounded algorithm are not needed: with every
// variable free the subspace step IS the quasi-Newton step.  B is formed explicitly (n <= 16; n = 3 for the GP refit)
// by applying the stored BFGS updates oldest first - the compact representation L-BFGS-B uses is the same matrix.
//
// Pure C++ (no CUDA): tested on the CPU against scipy.optimize.fmin_l_bfgs_b, iterate by iterate
// (tests/test_lbfgsb.py, through the mogp_lbfgsb_* entry points).
#pragma once
#include <math.h>
#include <string.h>

#include <algorithm>

namespace mogp {

struct Dcsrch {
  bool brackt = false;
  int stage = 1;
  double ginit = 0, gtest = 0, gx = 0, gy = 0, finit = 0, fx = 0, fy = 0, stx = 0, sty = 0, stmin = 0, stmax = 0, width = 0,
         width1 = 0;
};
enum LsTask { LS_START = 0, LS_FG = 1, LS_CONVERGENCE = 2, LS_WARNING = 3, LS_ERROR = 4 };

// MINPACK-2 dcstep: safeguarded cubic / quadratic step and update of the interval of uncertainty.
inline void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
                   bool& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {  // case 1: higher function value, the minimum is bracketed
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = std::max(fabs(theta), std::max(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = fabs(stpc - stx) < fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (sgnd < 0.0) {  // case 2: lower value, derivatives of opposite sign
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = std::max(fabs(theta), std::max(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
    brackt = true;
  } else if (fabs(dp) < fabs(dx)) {  // case 3: lower value, same sign, the derivative decreases
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = std::max(fabs(theta), std::max(fabs(dx), fabs(dp)));
    gamma = s * sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else if (stp > stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq;
      if (stp > stx) stpf = std::min(stp + 0.66 * (sty - stp), stpf);
      else stpf = std::max(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
      stpf = std::min(stpmax, stpf);
      stpf = std::max(stpmin, stpf);
    }
  } else {  // case 4: lower value, same sign, the derivative does not decrease
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = std::max(fabs(theta), std::max(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpc = stp + r * (sty - stp);
      stpf = stpc;
    } else if (stp > stx) {
      stpf = stpmax;
    } else {
      stpf = stpmin;
    }
  }
  if (fp > fx) {
    sty = stp;
    fy = fp;
    dy = dp;
  } else {
    if (sgnd < 0.0) {
      sty = stx;
      fy = fx;
      dy = dx;
    }
    stx = stp;
    fx = fp;
    dx = dp;
  }
  stp = stpf;
}

// MINPACK-2 dcsrch: one reverse-communication call of the More'-Thuente search.
inline LsTask dcsrch(double f, double g, double& stp, double ftol, double gtol, double xtol, double stpmin, double stpmax,
                     LsTask task, Dcsrch& S) {
  const double xtrapl = 1.1, xtrapu = 4.0;
  if (task == LS_START) {
    if (stp < stpmin || stp > stpmax || g >= 0.0 || ftol < 0.0 || gtol < 0.0 || xtol < 0.0 || stpmin < 0.0 || stpmax < stpmin)
      return LS_ERROR;
    S.brackt = false;
    S.stage = 1;
    S.finit = f;
    S.ginit = g;
    S.gtest = ftol * S.ginit;
    S.width = stpmax - stpmin;
    S.width1 = S.width / 0.5;
    S.stx = 0.0;
    S.fx = S.finit;
    S.gx = S.ginit;
    S.sty = 0.0;
    S.fy = S.finit;
    S.gy = S.ginit;
    S.stmin = 0.0;
    S.stmax = stp + xtrapu * stp;
    return LS_FG;
  }
  const double ftest = S.finit + stp * S.gtest;
  if (S.stage == 1 && f <= ftest && g >= 0.0) S.stage = 2;
  LsTask out = LS_FG;
  if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) out = LS_WARNING;             // rounding errors prevent progress
  if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) out = LS_WARNING;            // xtol test satisfied
  if (stp == stpmax && f <= ftest && g <= S.gtest) out = LS_WARNING;                // stp = stpmax
  if (stp == stpmin && (f > ftest || g >= S.gtest)) out = LS_WARNING;               // stp = stpmin
  if (f <= ftest && fabs(g) <= gtol * (-S.ginit)) out = LS_CONVERGENCE;
  if (out != LS_FG) return out;
  if (S.stage == 1 && f <= S.fx && f > ftest) {
    // modified function: predict the step while a lower value has been found but the decrease is not sufficient
    const double fm = f - stp * S.gtest;
    double fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
    const double gm = g - S.gtest;
    double gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
    dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
    S.fx = fxm + S.stx * S.gtest;
    S.fy = fym + S.sty * S.gtest;
    S.gx = gxm + S.gtest;
    S.gy = gym + S.gtest;
  } else {
    dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
  }
  if (S.brackt) {  // bisection when the interval does not shrink fast enough
    if (fabs(S.sty - S.stx) >= 0.66 * S.width1) stp = S.stx + 0.5 * (S.sty - S.stx);
    S.width1 = S.width;
    S.width = fabs(S.sty - S.stx);
  }
  if (S.brackt) {
    S.stmin = std::min(S.stx, S.sty);
    S.stmax = std::max(S.stx, S.sty);
  } else {
    S.stmin = stp + xtrapl * (stp - S.stx);
    S.stmax = stp + xtrapu * (stp - S.stx);
  }
  stp = std::max(stp, stpmin);
  stp = std::min(stp, stpmax);
  if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) || (S.brackt && S.stmax - S.stmin <= xtol * S.stmax)) stp = S.stx;
  return LS_FG;
}

struct Lbfgsb {
  static constexpr int NMAX = 16, MMAX = 20;
  enum Task {
    T_START = 0,
    T_FG = 1,            // evaluate f, g at x and call step() again (FG_START / FG_LNSRCH)
    T_NEW_X = 2,         // an iteration has finished (the SciPy wrapper counts iterations here); call step() again
    T_CONV_PGTOL = 3,    // CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL
    T_CONV_FACTR = 4,    // CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH
    T_ABNORMAL = 5,      // ABNORMAL_TERMINATION_IN_LNSRCH
    T_STOP = 6           // iteration / evaluation limit (the wrapper's maxiter / maxfun)
  };
  // parameters
  int n = 0, m = 10, maxls = 20, maxiter = 1000, maxfun = 1000;
  double factr = 1e7, pgtol = 1e-5;
  // iterate (caller reads x, writes f and g)
  double x[NMAX], g[NMAX], f = 0.0;
  // counters
  int iter = 0, nfgv = 0, nskip = 0, n_restart = 0;
  Task task = T_START;

  void init(int n_, const double* x0, int m_ = 10, double factr_ = 1e7, double pgtol_ = 1e-5, int maxls_ = 20) {
    n = n_;
    m = std::min(m_, (int)MMAX);
    factr = factr_;
    pgtol = pgtol_;
    maxls = maxls_;
    memcpy(x, x0, sizeof(double) * n);
    col = 0;
    theta = 1.0;
    iter = nfgv = nskip = n_restart = 0;
    first_fg = true;
    task = T_START;
  }

  bool finished() const { return task >= T_CONV_PGTOL; }

  // Advance until the optimiser needs (f, g) at x, announces a new iterate, or stops.
  Task step() {
    const double epsmch = 2.220446049250313e-16;
    switch (task) {
      case T_START:
        first_fg = true;
        return task = T_FG;
      case T_FG:
        if (first_fg) {  // FG_START
          first_fg = false;
          nfgv = 1;
          sbgnrm = supnorm(g);
          if (sbgnrm <= pgtol) return task = T_CONV_PGTOL;
          goto L222;
        }
        goto L666;  // FG_LNSRCH
      case T_NEW_X:
        goto L777;
      default:
        return task;
    }
  L222:
    direction();
    ls_fresh = true;
  L666 : {
    int info = 0;
    const bool need_fg = lnsrlb(info);
    if (info != 0 || iback >= maxls) {
      // restore the previous iterate
      memcpy(x, t, sizeof(double) * n);
      memcpy(g, r, sizeof(double) * n);
      f = fold;
      if (col == 0) {  // abnormal termination
        if (info == 0) nfgv -= 1;
        iter += 1;
        return task = T_ABNORMAL;
      }
      if (info == 0) nfgv -= 1;  // refresh the memory and restart the iteration
      col = 0;
      theta = 1.0;
      n_restart += 1;
      goto L222;
    }
    if (need_fg) return task = T_FG;
    iter += 1;
    sbgnrm = supnorm(g);
    return task = T_NEW_X;
  }
  L777:
    if (iter >= maxiter || nfgv > maxfun) return task = T_STOP;  // (the SciPy wrapper's checks at NEW_X)
    if (sbgnrm <= pgtol) return task = T_CONV_PGTOL;
    {
      const double ddum = std::max(fabs(fold), std::max(fabs(f), 1.0));
      if (fold - f <= epsmch * factr * ddum) return task = T_CONV_FACTR;
    }
    {
      // r = g_new - g_old, s = stp * d
      double rr = 0.0, dr, ddum;
      for (int i = 0; i < n; ++i) {
        r[i] = g[i] - r[i];
        rr += r[i] * r[i];
      }
      if (stp == 1.0) {
        dr = gd - gdold;
        ddum = -gdold;
      } else {
        dr = (gd - gdold) * stp;
        for (int i = 0; i < n; ++i) d[i] *= stp;
        ddum = -gdold * stp;
      }
      if (dr <= epsmch * ddum) {
        nskip += 1;  // skip the update
      } else {
        // matupd: newest pair last; at most m pairs
        if (col == m) {
          for (int q = 1; q < m; ++q) {
            memcpy(ws[q - 1], ws[q], sizeof(double) * n);
            memcpy(wy[q - 1], wy[q], sizeof(double) * n);
          }
          col = m - 1;
        }
        memcpy(ws[col], d, sizeof(double) * n);
        memcpy(wy[col], r, sizeof(double) * n);
        col += 1;
        theta = rr / dr;
      }
    }
    goto L222;
  }

 private:
  int col = 0;
  double theta = 1.0;
  double ws[MMAX][NMAX], wy[MMAX][NMAX];
  double d[NMAX], z[NMAX], t[NMAX], r[NMAX];
  double fold = 0.0, gd = 0.0, gdold = 0.0, stp = 0.0, dnorm = 0.0, sbgnrm = 0.0;
  int ifun = 0, iback = 0;
  bool first_fg = true, ls_fresh = true;
  Dcsrch ls;
  LsTask ls_task = LS_START;

  double supnorm(const double* v) const {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = std::max(s, fabs(v[i]));
    return s;
  }

  // z = x + (quasi-Newton step), d = z - x   (cauchy + subsm of the unbounded problem; label 555)
  void direction() {
    double dq[NMAX];
    if (col == 0) {
      for (int i = 0; i < n; ++i) dq[i] = -g[i] / theta;  // x - g / theta: the Cauchy point of the model with B = theta I
    } else {
      // B = theta I, then the stored BFGS updates oldest first: B <- B - (B s)(B s)' / s'B s + y y' / y's
      double B[NMAX][NMAX];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) B[i][j] = i == j ? theta : 0.0;
      for (int q = 0; q < col; ++q) {
        double Bs[NMAX], sBs = 0.0, ys = 0.0;
        for (int i = 0; i < n; ++i) {
          double a = 0.0;
          for (int j = 0; j < n; ++j) a += B[i][j] * ws[q][j];
          Bs[i] = a;
        }
        for (int i = 0; i < n; ++i) {
          sBs += ws[q][i] * Bs[i];
          ys += wy[q][i] * ws[q][i];
        }
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) B[i][j] += -Bs[i] * Bs[j] / sBs + wy[q][i] * wy[q][j] / ys;
      }
      // solve B dq = -g (B is symmetric positive definite: Cholesky; falls back to steepest descent if it is not)
      double L[NMAX][NMAX];
      bool ok = true;
      for (int j = 0; j < n && ok; ++j) {
        double dj = B[j][j];
        for (int k = 0; k < j; ++k) dj -= L[j][k] * L[j][k];
        if (!(dj > 0.0)) {
          ok = false;
          break;
        }
        L[j][j] = sqrt(dj);
        for (int i = j + 1; i < n; ++i) {
          double a = B[i][j];
          for (int k = 0; k < j; ++k) a -= L[i][k] * L[j][k];
          L[i][j] = a / L[j][j];
        }
      }
      if (ok) {
        double w[NMAX];
        for (int i = 0; i < n; ++i) {
          double a = -g[i];
          for (int k = 0; k < i; ++k) a -= L[i][k] * w[k];
          w[i] = a / L[i][i];
        }
        for (int i = n - 1; i >= 0; --i) {
          double a = w[i];
          for (int k = i + 1; k < n; ++k) a -= L[k][i] * dq[k];
          dq[i] = a / L[i][i];
        }
      } else {  // (L-BFGS-B: "nonpositive definiteness in Cholesky factorization in formt; refresh the lbfgs memory")
        col = 0;
        theta = 1.0;
        n_restart += 1;
        for (int i = 0; i < n; ++i) dq[i] = -g[i];
      }
    }
    for (int i = 0; i < n; ++i) {
      z[i] = x[i] + dq[i];
      d[i] = z[i] - x[i];
    }
  }

  // lnsrlb: returns true when (f, g) are needed at the trial point now in x; false when the search has finished
  // (x, f, g are the accepted point).  info = -4: ascent direction.
  bool lnsrlb(int& info) {
    const double big = 1e10, ftol = 1e-3, gtol = 0.9, xtol = 0.1;
    if (ls_fresh) {
      ls_fresh = false;
      double dtd = 0.0;
      for (int i = 0; i < n; ++i) dtd += d[i] * d[i];
      dnorm = sqrt(dtd);
      stp = (iter == 0) ? std::min(1.0 / dnorm, big) : 1.0;
      memcpy(t, x, sizeof(double) * n);
      memcpy(r, g, sizeof(double) * n);
      fold = f;
      ifun = 0;
      iback = 0;
      ls_task = LS_START;
    }
    gd = 0.0;
    for (int i = 0; i < n; ++i) gd += g[i] * d[i];
    if (ifun == 0) {
      gdold = gd;
      if (gd >= 0.0) {  // the directional derivative >= 0: line search is impossible
        info = -4;
        return false;
      }
    }
    ls_task = dcsrch(f, gd, stp, ftol, gtol, xtol, 0.0, big, ls_task, ls);
    if (ls_task != LS_CONVERGENCE && ls_task != LS_WARNING) {
      ifun += 1;
      nfgv += 1;
      iback = ifun - 1;
      if (stp == 1.0) memcpy(x, z, sizeof(double) * n);
      else
        for (int i = 0; i < n; ++i) x[i] = stp * d[i] + t[i];
      return true;
    }
    return false;
  }
};

}  // namespace mogp
