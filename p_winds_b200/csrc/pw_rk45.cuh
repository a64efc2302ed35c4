// pw_rk45.cuh -- step-faithful replica of scipy.integrate.solve_ivp(method='RK45',
// t_eval=...) for systems of N equations (N = 1 for hydrogen.py:451, N = 2 when
// helium.population_fraction is called with method='RK45', helium.py:702).
//
// Follows scipy 1.18.1 line by line:
//   _ivp/common.py:68-134   select_initial_step
//   _ivp/rk.py:14-73        rk_step (Dormand-Prince tableau :374ff)
//   _ivp/rk.py:111-176      _step_impl (controller, min_step failure)
//   _ivp/rk.py:715-737      RkDenseOutput
//   _ivp/ivp.py             t_eval bookkeeping (points <= t of each accepted step,
//                           side='right'; a failed solve returns fewer points)
// Trace parity (identical nfev / accepted steps on the host build) is checked in
// tests/test_hostsim_solvers.py against scipy itself.
#pragma once
#include "pw_common.cuh"

namespace pw {

struct Rk45Opts {
  double rtol, atol;
  double first_step;  // <= 0: select automatically
  double max_step;    // <= 0: infinity
  // Work limit (right-hand-side evaluations per solve); <= 0: none, as in scipy.  scipy's RK45
  // has no step limit: on a stiff start it crawls at a step just above 10 ulp(t) -- e.g.
  // hydrogen.ion_fraction at T0 = 5000 K, Mdot = 7.2e9 g/s, h = 0.85, tidal takes 1.9e8
  // evaluations (hours in Python) before it fails.  The engine gives such a model up early and
  // says so (status PW_FAIL_WORK_LIMIT), instead of holding a GPU thread for minutes.
  long long max_nfev = 0;
};

struct Rk45Stats {
  int nfev, naccept, nreject;
  int limit = 0;  // 1: stopped by Rk45Opts::max_nfev
};

#define PW_RK_C2 (1.0 / 5)
#define PW_RK_C3 (3.0 / 10)
#define PW_RK_C4 (4.0 / 5)
#define PW_RK_C5 (8.0 / 9)

// Solve y' = f(t, y) from t_eval[0] to t_eval[n-1], writing the solution at all
// t_eval points into y_out[k*N + c].  Returns the number of t_eval points written
// (== n on success; < n means scipy's status -1 "step size too small").
// RHS signature: rhs(t, const double* y, double* dydt).
template <int N, class RHS>
PW_HD_NOINLINE int rk45_solve(RHS& rhs, const double* __restrict__ t_eval, int n, const double* y0,
                     const Rk45Opts& o, double* __restrict__ y_out, Rk45Stats* st) {
  const double A21 = 1.0 / 5;
  const double A31 = 3.0 / 40, A32 = 9.0 / 40;
  const double A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9;
  const double A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729;
  const double A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176,
               A65 = -5103.0 / 18656;
  const double B1 = 35.0 / 384, B2 = 0.0, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784,
               B6 = 11.0 / 84;
  const double E1 = -71.0 / 57600, E2 = 0.0, E3 = 71.0 / 16695, E4 = -71.0 / 1920,
               E5 = 17253.0 / 339200, E6 = -22.0 / 525, E7 = 1.0 / 40;
  // dense-output matrix P (7 x 4), rk.py RK45.P
  const double P[7][4] = {
      {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
      {0, 0, 0, 0},
      {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
      {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
      {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
      {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
      {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};
  const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
  const double INF = 1.0 / 0.0;
  const double t0 = t_eval[0], t_bound = t_eval[n - 1];
  const double max_step = (o.max_step > 0.0) ? o.max_step : INF;
  const double sqrtN = sqrt((double)N);

  double t = t0, y[N], f[N], K[7][N];
  long long nfev = 0;
  int nacc = 0, nrej = 0;
  for (int c = 0; c < N; ++c) y[c] = y0[c];
  rhs(t, y, f); ++nfev;

  auto rms = [&](const double* x) {  // common.norm: np.linalg.norm(x) / x.size ** 0.5
    double s = 0.0;
    for (int c = 0; c < N; ++c) s += x[c] * x[c];
    return sqrt(s) / sqrtN;
  };

  // ---- select_initial_step (common.py:68-134), order = 4, direction = +1
  double h_abs;
  if (o.first_step > 0.0) {
    h_abs = o.first_step;
  } else {
    const double interval = fabs(t_bound - t0);
    if (interval == 0.0) {
      h_abs = 0.0;
    } else {
      double scale[N], tmp[N], y1[N], f1[N];
      for (int c = 0; c < N; ++c) scale[c] = o.atol + fabs(y[c]) * o.rtol;
      for (int c = 0; c < N; ++c) tmp[c] = y[c] / scale[c];
      const double d0 = rms(tmp);
      for (int c = 0; c < N; ++c) tmp[c] = f[c] / scale[c];
      const double d1 = rms(tmp);
      double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
      h0 = fmin(h0, interval);
      for (int c = 0; c < N; ++c) y1[c] = y[c] + h0 * 1.0 * f[c];
      rhs(t0 + h0 * 1.0, y1, f1); ++nfev;
      for (int c = 0; c < N; ++c) tmp[c] = (f1[c] - f[c]) / scale[c];
      const double d2 = rms(tmp) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
      else h1 = pw_pow(0.01 / fmax(d1, d2), 1.0 / (4 + 1));
      h_abs = fmin(fmin(100 * h0, h1), fmin(interval, max_step));
    }
  }

  int n_out = 0;  // t_eval points written so far (ivp.py t_eval_i)
  // OdeSolver.step(): "if self.n == 0 or self.t == self.t_bound" -> finished, no step
  bool finished = (t == t_bound);
  if (finished) {
    // degenerate span: the first loop pass in solve_ivp emits t_eval points <= t
    while (n_out < n && t_eval[n_out] <= t) {
      for (int c = 0; c < N; ++c) y_out[n_out * N + c] = y[c];
      ++n_out;
    }
  }
  while (!finished) {
    // ---- RungeKutta._step_impl (rk.py:111-176)
    const double min_step = 10 * fabs(nextafter(t, INF) - t);
    double ha;
    if (h_abs > max_step) ha = max_step;
    else if (h_abs < min_step) ha = min_step;
    else ha = h_abs;
    bool accepted = false, rejected = false;
    double h = 0.0, t_new = t, y_new[N];
    while (!accepted) {
      if (ha < min_step) {  // TOO_SMALL_STEP -> solver.status = 'failed'
        if (st) { st->nfev = (int)(nfev < 2000000000 ? nfev : 2000000000); st->naccept = nacc; st->nreject = nrej; }
        return n_out;
      }
      if (o.max_nfev > 0 && nfev >= o.max_nfev) {  // work limit (no scipy counterpart, see Rk45Opts)
        if (st) { st->nfev = (int)(nfev < 2000000000 ? nfev : 2000000000); st->naccept = nacc; st->nreject = nrej; st->limit = 1; }
        return n_out;
      }
      h = ha * 1.0;
      t_new = t + h;
      if (1.0 * (t_new - t_bound) > 0) t_new = t_bound;
      h = t_new - t;
      ha = fabs(h);
      // rk_step: K[s] = fun(t + c*h, y + dot(K[:s].T, a[:s]) * h)
      double yt[N];
      for (int c = 0; c < N; ++c) K[0][c] = f[c];
      for (int c = 0; c < N; ++c) yt[c] = y[c] + (K[0][c] * A21) * h;
      rhs(t + PW_RK_C2 * h, yt, K[1]);
      for (int c = 0; c < N; ++c) yt[c] = y[c] + (K[0][c] * A31 + K[1][c] * A32) * h;
      rhs(t + PW_RK_C3 * h, yt, K[2]);
      for (int c = 0; c < N; ++c) yt[c] = y[c] + (K[0][c] * A41 + K[1][c] * A42 + K[2][c] * A43) * h;
      rhs(t + PW_RK_C4 * h, yt, K[3]);
      for (int c = 0; c < N; ++c)
        yt[c] = y[c] + (K[0][c] * A51 + K[1][c] * A52 + K[2][c] * A53 + K[3][c] * A54) * h;
      rhs(t + PW_RK_C5 * h, yt, K[4]);
      for (int c = 0; c < N; ++c)
        yt[c] = y[c] + (K[0][c] * A61 + K[1][c] * A62 + K[2][c] * A63 + K[3][c] * A64 + K[4][c] * A65) * h;
      rhs(t + 1.0 * h, yt, K[5]);
      for (int c = 0; c < N; ++c)
        y_new[c] = y[c] + h * (K[0][c] * B1 + K[1][c] * B2 + K[2][c] * B3 + K[3][c] * B4 +
                               K[4][c] * B5 + K[5][c] * B6);
      rhs(t + h, y_new, K[6]);
      nfev += 6;
      double en[N];
      for (int c = 0; c < N; ++c) {
        const double scale = o.atol + fmax(fabs(y[c]), fabs(y_new[c])) * o.rtol;
        const double err = (K[0][c] * E1 + K[1][c] * E2 + K[2][c] * E3 + K[3][c] * E4 +
                            K[4][c] * E5 + K[5][c] * E6 + K[6][c] * E7) * h;
        en[c] = err / scale;
      }
      const double error_norm = rms(en);
      if (error_norm < 1) {
        double factor;
        if (error_norm == 0) factor = MAX_FACTOR;
        else factor = fmin(MAX_FACTOR, SAFETY * pw_pow(error_norm, -0.2));
        if (rejected) factor = fmin(1.0, factor);
        ha *= factor;
        accepted = true;
      } else if (error_norm >= 1) {
        ha *= fmax(MIN_FACTOR, SAFETY * pw_pow(error_norm, -0.2));
        rejected = true;
        ++nrej;
      } else {
        // NaN error norm: Python's `nan < 1` is False, so scipy takes the reject
        // branch with max(MIN_FACTOR, nan) == MIN_FACTOR
        ha *= MIN_FACTOR;
        rejected = true;
        ++nrej;
      }
    }
    ++nacc;
    h_abs = ha;
    // ---- solve_ivp: emit t_eval points in (t_old, t_new] via dense output
    //      Q = K^T P ; y(t) = y_old + h * Q @ [x, x^2, x^3, x^4]
    int n_new = n_out;
    while (n_new < n && t_eval[n_new] <= t_new) ++n_new;  // searchsorted(side='right')
    if (n_new > n_out) {
      double Q[N][4];
      for (int c = 0; c < N; ++c)
        for (int q = 0; q < 4; ++q) {
          double s = 0.0;
          for (int k = 0; k < 7; ++k) s += K[k][c] * P[k][q];
          Q[c][q] = s;
        }
      for (int k = n_out; k < n_new; ++k) {
        const double x = (t_eval[k] - t) / h;
        const double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;
        for (int c = 0; c < N; ++c) {
          const double yy = h * (Q[c][0] * p1 + Q[c][1] * p2 + Q[c][2] * p3 + Q[c][3] * p4);
          y_out[k * N + c] = yy + y[c];
        }
      }
      n_out = n_new;
    }
    t = t_new;
    for (int c = 0; c < N; ++c) { y[c] = y_new[c]; f[c] = K[6][c]; }
    if (1.0 * (t - t_bound) >= 0) finished = true;
  }
  if (st) { st->nfev = (int)(nfev < 2000000000 ? nfev : 2000000000); st->naccept = nacc; st->nreject = nrej; }
  return n_out;
}

}  // namespace pw
