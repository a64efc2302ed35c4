// pw_radau.cuh -- step-faithful replica of scipy.integrate.solve_ivp(method='Radau',
// t_eval=...) for systems of N = 1 or 2 equations: the default integrator of
// oxygen.ion_fraction (oxygen.py:186-191, :426-433) and helium.ion_fraction, and the one
// the documentation uses for carbon.ion_fraction.
//
// Follows scipy 1.18.1:
//   _ivp/radau.py:14-48     constants (Radau IIA order 5: C, E, MU_REAL, MU_COMPLEX, T, TI, P)
//   _ivp/radau.py:51-139    solve_collocation_system (simplified Newton, real + complex LU)
//   _ivp/radau.py:142-178   predict_factor (Gustafsson controller)
//   _ivp/radau.py:295-345   Radau.__init__ (select_initial_step with order 3, newton_tol)
//   _ivp/radau.py:409-535   _step_impl
//   _ivp/radau.py:537-571   dense output Q = Z^T P
//   _ivp/common.py          select_initial_step, norm, num_jac / _dense_num_jac (finite-
//                           difference Jacobian with its adaptive `factor`)
//   _ivp/ivp.py             t_eval bookkeeping (as in pw_rk45.cuh)
// The linear algebra is written out for N x N real and complex matrices with partial
// pivoting (LAPACK getrf/getrs order of operations; complex quotients by Smith's formula).
// Summation orders inside numpy.dot / BLAS are not specified, so values agree with scipy to
// rounding, amplified by the rtol = 1e-3 step controller to ~1e-8; the counters (nfev, njev,
// nlu, accepted steps) are held identical on the host build (tests/test_hostsim_solvers.py).
#pragma once
#include "pw_common.cuh"
#include "pw_rk45.cuh"  // Rk45Opts doubles as the solve_ivp options block (rtol, atol, first_step, max_step)

namespace pw {

struct RadauStats {
  int nfev, njev, nlu, naccept;
  int njac_fev;  // right-hand sides spent inside num_jac (scipy does not count them in nfev)
};

struct RadauC {  // complex scalar, just enough arithmetic for the 1 + 1 complex-pair systems
  double re, im;
};
PW_HD RadauC rc_mul(RadauC a, RadauC b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
PW_HD RadauC rc_sub(RadauC a, RadauC b) { return {a.re - b.re, a.im - b.im}; }
PW_HD RadauC rc_div(RadauC a, RadauC b) {  // Smith (1962), as compilers' complex division
  if (fabs(b.re) >= fabs(b.im)) {
    const double r = b.im / b.re, d = b.re + b.im * r;
    return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
  }
  const double r = b.re / b.im, d = b.re * r + b.im;
  return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
}

// LU with partial pivoting, row exchanges recorded in piv (getrf); solve in place (getrs)
template <int N>
PW_HD void radau_lu_real(double (*a)[N], int* piv) {
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < N; ++i)
      if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); p = i; }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < N; ++j) { const double t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
    for (int i = k + 1; i < N; ++i) {
      a[i][k] = a[i][k] / a[k][k];
      for (int j = k + 1; j < N; ++j) a[i][j] -= a[i][k] * a[k][j];
    }
  }
}
template <int N>
PW_HD void radau_solve_real(const double (*a)[N], const int* piv, double* b) {
  for (int k = 0; k < N; ++k) {
    if (piv[k] != k) { const double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
    for (int i = k + 1; i < N; ++i) b[i] -= a[i][k] * b[k];
  }
  for (int k = N - 1; k >= 0; --k) {
    for (int j = k + 1; j < N; ++j) b[k] -= a[k][j] * b[j];
    b[k] = b[k] / a[k][k];
  }
}
template <int N>
PW_HD void radau_lu_cplx(RadauC (*a)[N], int* piv) {
  for (int k = 0; k < N; ++k) {
    int p = k;
    double best = fabs(a[k][k].re) + fabs(a[k][k].im);  // izamax: |re| + |im|
    for (int i = k + 1; i < N; ++i) {
      const double m = fabs(a[i][k].re) + fabs(a[i][k].im);
      if (m > best) { best = m; p = i; }
    }
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < N; ++j) { const RadauC t = a[k][j]; a[k][j] = a[p][j]; a[p][j] = t; }
    for (int i = k + 1; i < N; ++i) {
      a[i][k] = rc_div(a[i][k], a[k][k]);
      for (int j = k + 1; j < N; ++j) a[i][j] = rc_sub(a[i][j], rc_mul(a[i][k], a[k][j]));
    }
  }
}
template <int N>
PW_HD void radau_solve_cplx(const RadauC (*a)[N], const int* piv, RadauC* b) {
  for (int k = 0; k < N; ++k) {
    if (piv[k] != k) { const RadauC t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
    for (int i = k + 1; i < N; ++i) b[i] = rc_sub(b[i], rc_mul(a[i][k], b[k]));
  }
  for (int k = N - 1; k >= 0; --k) {
    for (int j = k + 1; j < N; ++j) b[k] = rc_sub(b[k], rc_mul(a[k][j], b[j]));
    b[k] = rc_div(b[k], a[k][k]);
  }
}

// Solve y' = f(t, y) from t_eval[0] to t_eval[n-1] with scipy's Radau; the solution at all
// t_eval points goes to y_out[k*N + c].  Returns the number of points written (== n on
// success; < n is scipy's status -1, "Required step size is less than spacing between numbers").
// RHS signature: rhs(t, const double* y, double* dydt).
template <int N, class RHS>
PW_HD_NOINLINE int radau_solve(RHS& rhs, const double* __restrict__ t_eval, int n, const double* y0,
                               const Rk45Opts& o, double* __restrict__ y_out, RadauStats* st) {
  const double S6 = 2.449489742783178;  // 6 ** 0.5
  const double C[3] = {(4 - S6) / 10, (4 + S6) / 10, 1.0};
  const double E[3] = {(-13 - 7 * S6) / 3, (-13 + 7 * S6) / 3, -1.0 / 3};
  const double MU_REAL = 3.637834252744496;                                    // 3 + 3**(2/3) - 3**(1/3)
  const RadauC MU_COMPLEX = {2.6810828736277523, -3.050430199247411};           // radau.py:23-24
  const double T[3][3] = {{0.09443876248897524, -0.14125529502095421, 0.03002919410514742},
                          {0.25021312296533332, 0.20412935229379994, -0.38294211275726192},
                          {1, 1, 0}};
  const double TI[3][3] = {{4.17871859155190428, 0.32768282076106237, 0.52337644549944951},
                           {-4.17871859155190428, -0.32768282076106237, 0.47662355450055044},
                           {0.50287263494578682, -2.57192694985560522, 0.59603920482822492}};
  const double P[3][3] = {{13.0 / 3 + 7 * S6 / 3, -23.0 / 3 - 22 * S6 / 3, 10.0 / 3 + 5 * S6},
                          {13.0 / 3 - 7 * S6 / 3, -23.0 / 3 + 22 * S6 / 3, 10.0 / 3 - 5 * S6},
                          {1.0 / 3, -8.0 / 3, 10.0 / 3}};
  const int NEWTON_MAXITER = 6;
  const double MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
  const double INF = 1.0 / 0.0;
  const double t0 = t_eval[0], t_bound = t_eval[n - 1];
  const double max_step = (o.max_step > 0.0) ? o.max_step : INF;
  const double rtol = o.rtol, atol = o.atol;

  int nfev = 0, njev = 0, nlu = 0, nacc = 0, njac_fev = 0;
  double t = t0, y[N], f[N];
  for (int c = 0; c < N; ++c) y[c] = y0[c];
  rhs(t, y, f); ++nfev;

  auto rms = [&](const double* x, int m) {  // common.norm: np.linalg.norm(x) / x.size ** 0.5
    double s = 0.0;
    for (int c = 0; c < m; ++c) s += x[c] * x[c];
    return sqrt(s) / sqrt((double)m);
  };

  // ---- select_initial_step (common.py), order = 3, direction = +1
  double h_abs_self;
  if (o.first_step > 0.0) {
    h_abs_self = o.first_step;
  } else {
    const double interval = fabs(t_bound - t0);
    if (interval == 0.0) {
      h_abs_self = 0.0;
    } else {
      double scale[N], tmp[N], y1[N], f1[N];
      for (int c = 0; c < N; ++c) scale[c] = atol + fabs(y[c]) * rtol;
      for (int c = 0; c < N; ++c) tmp[c] = y[c] / scale[c];
      const double d0 = rms(tmp, N);
      for (int c = 0; c < N; ++c) tmp[c] = f[c] / scale[c];
      const double d1 = rms(tmp, N);
      double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, interval);
      for (int c = 0; c < N; ++c) y1[c] = y[c] + h0 * 1.0 * f[c];
      rhs(t0 + h0 * 1.0, y1, f1); ++nfev;
      for (int c = 0; c < N; ++c) tmp[c] = (f1[c] - f[c]) / scale[c];
      const double d2 = rms(tmp, N) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pw_pow(0.01 / fmax(d1, d2), 1.0 / (3 + 1));
      h_abs_self = fmin(fmin(100 * h0, h1), fmin(interval, max_step));
    }
  }
  bool have_old = false;          // h_abs_old / error_norm_old are not None
  double h_abs_old_self = 0.0, error_norm_old_self = 0.0;
  const double newton_tol = fmax(10 * kEps / rtol, fmin(0.03, sqrt(rtol)));

  // ---- num_jac (common.py): dense finite-difference Jacobian with adaptive factor
  double J[N][N], jac_factor[N];
  bool have_factor = false;
  auto num_jac = [&](double tj, const double* yj, const double* fj) {
    ++njev;
    const double DIFF_REJECT = 2.0097183471152322e-14, DIFF_SMALL = 1.8189894035458565e-12,
                 DIFF_BIG = 0.0001220703125, MIN_FAC = 2.220446049250313e-13;
    double factor[N], y_scale[N], h[N];
    for (int c = 0; c < N; ++c) factor[c] = have_factor ? jac_factor[c] : 1.4901161193847656e-08;  // EPS ** 0.5
    for (int c = 0; c < N; ++c) {
      const double sgn = (fj[c] >= 0) ? 1.0 : -1.0;
      y_scale[c] = sgn * fmax(atol, fabs(yj[c]));
      h[c] = (yj[c] + factor[c] * y_scale[c]) - yj[c];
      while (h[c] == 0) {
        factor[c] *= 10;
        h[c] = (yj[c] + factor[c] * y_scale[c]) - yj[c];
      }
    }
    double diff[N][N], max_diff[N], scale[N];  // diff[row][col]
    auto column = [&](int col, double hc, double* dcol, double* mdiff, double* sc) {
      double yp[N], fn[N];
      for (int c = 0; c < N; ++c) yp[c] = yj[c];
      yp[col] = yj[col] + hc;
      rhs(tj, yp, fn); ++njac_fev;
      int mi = 0;
      for (int r = 0; r < N; ++r) {
        dcol[r] = fn[r] - fj[r];
        if (fabs(dcol[r]) > fabs(dcol[mi])) mi = r;  // argmax: first maximum
      }
      *mdiff = fabs(dcol[mi]);
      *sc = fmax(fabs(fj[mi]), fabs(fn[mi]));
    };
    for (int col = 0; col < N; ++col) {
      double dcol[N];
      column(col, h[col], dcol, &max_diff[col], &scale[col]);
      for (int r = 0; r < N; ++r) diff[r][col] = dcol[r];
    }
    for (int col = 0; col < N; ++col) {
      if (max_diff[col] < DIFF_REJECT * scale[col]) {
        const double new_factor = 10 * factor[col];
        const double h_new = (yj[col] + new_factor * y_scale[col]) - yj[col];
        double dcol[N], md, sc;
        column(col, h_new, dcol, &md, &sc);
        if (max_diff[col] * sc < md * scale[col]) {
          factor[col] = new_factor;
          h[col] = h_new;
          for (int r = 0; r < N; ++r) diff[r][col] = dcol[r];
          scale[col] = sc;
          max_diff[col] = md;
        }
      }
    }
    for (int r = 0; r < N; ++r)
      for (int col = 0; col < N; ++col) J[r][col] = diff[r][col] / h[col];
    for (int c = 0; c < N; ++c) {
      if (max_diff[c] < DIFF_SMALL * scale[c]) factor[c] *= 10;
      if (max_diff[c] > DIFF_BIG * scale[c]) factor[c] *= 0.1;
      jac_factor[c] = fmax(factor[c], MIN_FAC);
    }
    have_factor = true;
  };
  num_jac(t, y, f);
  bool current_jac = true;

  bool have_lu = false;
  double LUr[N][N];
  RadauC LUc[N][N];
  int pivr[N], pivc[N];
  auto factorize = [&](double h) {
    const double mr = MU_REAL / h;
    const RadauC mc = {MU_COMPLEX.re / h, MU_COMPLEX.im / h};
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) {
        LUr[r][c] = ((r == c) ? mr : 0.0) - J[r][c];
        LUc[r][c] = {((r == c) ? mc.re : 0.0) - J[r][c], (r == c) ? mc.im : 0.0};
      }
    radau_lu_real<N>(LUr, pivr);
    radau_lu_cplx<N>(LUc, pivc);
    nlu += 2;
    have_lu = true;
  };

  // dense output of the last accepted step (RadauDenseOutput)
  bool have_sol = false;
  double sol_t_old = 0.0, sol_h = 0.0, sol_y_old[N], Q[N][3];
  auto sol_eval = [&](double tt, double* out) {
    const double x = (tt - sol_t_old) / sol_h;
    const double p1 = x, p2 = p1 * x, p3 = p2 * x;
    for (int c = 0; c < N; ++c) out[c] = (Q[c][0] * p1 + Q[c][1] * p2 + Q[c][2] * p3) + sol_y_old[c];
  };

  int n_out = 0;
  bool finished = (t == t_bound);
  if (finished) {
    while (n_out < n && t_eval[n_out] <= t) {
      for (int c = 0; c < N; ++c) y_out[n_out * N + c] = y[c];
      ++n_out;
    }
  }
  auto done = [&]() {
    if (st) { st->nfev = nfev; st->njev = njev; st->nlu = nlu; st->naccept = nacc; st->njac_fev = njac_fev; }
    return n_out;
  };

  while (!finished) {
    // ---- Radau._step_impl (radau.py:409-535)
    const double min_step = 10 * fabs(nextafter(t, INF) - t);
    double h_abs, h_abs_old = 0.0, error_norm_old = 0.0;
    bool have_old_l;
    if (h_abs_self > max_step) { h_abs = max_step; have_old_l = false; }
    else if (h_abs_self < min_step) { h_abs = min_step; have_old_l = false; }
    else { h_abs = h_abs_self; have_old_l = have_old; h_abs_old = h_abs_old_self; error_norm_old = error_norm_old_self; }

    auto predict_factor = [&](double ha, double en) {
      double multiplier;
      if (!have_old_l || en == 0) multiplier = 1;
      else multiplier = ha / h_abs_old * pw_pow(error_norm_old / en, 0.25);
      return fmin(1.0, multiplier) * pw_pow(en, -0.25);
    };

    bool rejected = false, step_accepted = false;
    double h = 0.0, t_new = t, y_new[N], Z[3][N], error_norm = 0.0, safety = 0.0, rate = 0.0;
    bool have_rate = false;
    int n_iter = 0;
    while (!step_accepted) {
      if (h_abs < min_step) return done();  // TOO_SMALL_STEP
      h = h_abs * 1.0;
      t_new = t + h;
      if (1.0 * (t_new - t_bound) > 0) t_new = t_bound;
      h = t_new - t;
      h_abs = fabs(h);

      double Z0[3][N];
      for (int i = 0; i < 3; ++i) {
        if (!have_sol) {
          for (int c = 0; c < N; ++c) Z0[i][c] = 0.0;
        } else {
          double ys[N];
          sol_eval(t + h * C[i], ys);
          for (int c = 0; c < N; ++c) Z0[i][c] = ys[c] - y[c];
        }
      }
      double scale[N];
      for (int c = 0; c < N; ++c) scale[c] = atol + fabs(y[c]) * rtol;

      bool converged = false;
      while (!converged) {
        if (!have_lu) factorize(h);
        // ---- solve_collocation_system (radau.py:51-139)
        const double M_real = MU_REAL / h;
        const RadauC M_complex = {MU_COMPLEX.re / h, MU_COMPLEX.im / h};
        double W[3][N], F[3][N], dW[3][N];
        for (int i = 0; i < 3; ++i)
          for (int c = 0; c < N; ++c) {
            W[i][c] = TI[i][0] * Z0[0][c] + TI[i][1] * Z0[1][c] + TI[i][2] * Z0[2][c];
            Z[i][c] = Z0[i][c];
          }
        double dW_norm_old = 0.0;
        bool have_norm_old = false;
        have_rate = false;
        int k = 0;
        for (k = 0; k < NEWTON_MAXITER; ++k) {
          bool finite = true;
          for (int i = 0; i < 3; ++i) {
            double yy[N];
            for (int c = 0; c < N; ++c) yy[c] = y[c] + Z[i][c];
            rhs(t + h * C[i], yy, F[i]); ++nfev;
            for (int c = 0; c < N; ++c)
              if (!(fabs(F[i][c]) <= 1.7976931348623157e308)) finite = false;
          }
          if (!finite) break;
          double f_real[N];
          RadauC f_cplx[N];
          for (int c = 0; c < N; ++c) {
            f_real[c] = (F[0][c] * TI[0][0] + F[1][c] * TI[0][1] + F[2][c] * TI[0][2]) - M_real * W[0][c];
            const RadauC fw = {F[0][c] * TI[1][0] + F[1][c] * TI[1][1] + F[2][c] * TI[1][2],
                               F[0][c] * TI[2][0] + F[1][c] * TI[2][1] + F[2][c] * TI[2][2]};
            const RadauC ww = {W[1][c], W[2][c]};
            f_cplx[c] = rc_sub(fw, rc_mul(M_complex, ww));
          }
          radau_solve_real<N>(LUr, pivr, f_real);
          radau_solve_cplx<N>(LUc, pivc, f_cplx);
          double nrm[3 * N];
          for (int c = 0; c < N; ++c) {
            dW[0][c] = f_real[c]; dW[1][c] = f_cplx[c].re; dW[2][c] = f_cplx[c].im;
            nrm[c] = dW[0][c] / scale[c]; nrm[N + c] = dW[1][c] / scale[c]; nrm[2 * N + c] = dW[2][c] / scale[c];
          }
          const double dW_norm = rms(nrm, 3 * N);
          if (have_norm_old) { rate = dW_norm / dW_norm_old; have_rate = true; }
          if (have_rate && (rate >= 1 || pw_pow(rate, (double)(NEWTON_MAXITER - k)) / (1 - rate) * dW_norm > newton_tol))
            break;
          for (int i = 0; i < 3; ++i)
            for (int c = 0; c < N; ++c) W[i][c] += dW[i][c];
          for (int i = 0; i < 3; ++i)
            for (int c = 0; c < N; ++c) Z[i][c] = T[i][0] * W[0][c] + T[i][1] * W[1][c] + T[i][2] * W[2][c];
          if (dW_norm == 0 || (have_rate && rate / (1 - rate) * dW_norm < newton_tol)) {
            converged = true;
            break;
          }
          dW_norm_old = dW_norm;
          have_norm_old = true;
        }
        n_iter = (k < NEWTON_MAXITER) ? k + 1 : NEWTON_MAXITER;
        if (!converged) {
          if (current_jac) break;
          num_jac(t, y, f);
          current_jac = true;
          have_lu = false;
        }
      }
      if (!converged) {
        h_abs *= 0.5;
        have_lu = false;
        continue;
      }
      double ZE[N], err[N], es[N];
      for (int c = 0; c < N; ++c) {
        y_new[c] = y[c] + Z[2][c];
        ZE[c] = (Z[0][c] * E[0] + Z[1][c] * E[1] + Z[2][c] * E[2]) / h;
        err[c] = f[c] + ZE[c];
      }
      radau_solve_real<N>(LUr, pivr, err);
      double scale2[N];
      for (int c = 0; c < N; ++c) {
        scale2[c] = atol + fmax(fabs(y[c]), fabs(y_new[c])) * rtol;
        es[c] = err[c] / scale2[c];
      }
      error_norm = rms(es, N);
      safety = 0.9 * (2 * NEWTON_MAXITER + 1) / (2 * NEWTON_MAXITER + n_iter);
      if (rejected && error_norm > 1) {
        double ye[N], fe[N];
        for (int c = 0; c < N; ++c) ye[c] = y[c] + err[c];
        rhs(t, ye, fe); ++nfev;
        for (int c = 0; c < N; ++c) err[c] = fe[c] + ZE[c];
        radau_solve_real<N>(LUr, pivr, err);
        for (int c = 0; c < N; ++c) es[c] = err[c] / scale2[c];
        error_norm = rms(es, N);
      }
      if (error_norm > 1) {
        const double factor = predict_factor(h_abs, error_norm);
        h_abs *= fmax(MIN_FACTOR, safety * factor);
        have_lu = false;
        rejected = true;
      } else {
        step_accepted = true;  // (a NaN error norm is not > 1: scipy accepts the step)
      }
    }
    const bool recompute_jac = n_iter > 2 && have_rate && rate > 1e-3;
    double factor = predict_factor(h_abs, error_norm);
    factor = fmin(MAX_FACTOR, safety * factor);
    if (!recompute_jac && factor < 1.2) factor = 1;
    else have_lu = false;
    double f_new[N];
    rhs(t_new, y_new, f_new); ++nfev;
    if (recompute_jac) {
      num_jac(t_new, y_new, f_new);
      current_jac = true;
    } else {
      current_jac = false;
    }
    h_abs_old_self = h_abs_self;
    error_norm_old_self = error_norm;
    have_old = true;
    h_abs_self = h_abs * factor;
    ++nacc;
    // dense output of this step: Q = Z^T P, y(t) = y_old + Q @ [x, x^2, x^3]
    for (int c = 0; c < N; ++c) {
      for (int q = 0; q < 3; ++q) Q[c][q] = Z[0][c] * P[0][q] + Z[1][c] * P[1][q] + Z[2][c] * P[2][q];
      sol_y_old[c] = y[c];
    }
    sol_t_old = t; sol_h = t_new - t; have_sol = true;
    int n_new = n_out;
    while (n_new < n && t_eval[n_new] <= t_new) ++n_new;  // searchsorted(side='right')
    for (int k = n_out; k < n_new; ++k) sol_eval(t_eval[k], &y_out[k * N]);
    n_out = n_new;
    t = t_new;
    for (int c = 0; c < N; ++c) { y[c] = y_new[c]; f[c] = f_new[c]; }
    if (1.0 * (t - t_bound) >= 0) finished = true;
  }
  return done();
}

}  // namespace pw
