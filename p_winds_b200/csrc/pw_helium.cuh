// pw_helium.cuh -- He singlet / triplet population profile of one model
// (reference: p_winds/helium.py:413-785 `population_fraction`, :270-292
// `recombination`, :317-380 `collision`).  Oklopcic & Hirata 2018 Eq. 15 in
// sonic-point units, integrated with the LSODA replica (method='odeint', the
// reference default), LSODA at user tolerances, or the RK45 replica.
//
// One *thread* integrates one model: the solve is ~10^3 dependent implicit
// steps, there is nothing to share between lanes, and a retrieval grid supplies
// thousands of independent models.
#pragma once
#include "pw_common.cuh"
#include "pw_lsoda.cuh"
#include "pw_rk45.cuh"

namespace pw {

enum { PW_HE_ODEINT = 0, PW_HE_LSODA = 1, PW_HE_RK45 = 2 };

struct HeParams {
  unsigned e2tab = 0;  // device: shared-memory address of kExp2Tab (pw_exp_neg_sh)
  double T, h_fraction, vs, rs, rhos;  // per model
  double r_pl;
  double y0[2];
  int relax, max_n_relax;
  double convergence;
  double rates[6];  // phi_1, phi_3, a_1, a_3, a_h_1, a_h_3 (cgs; helium.radiative_processes[_mono])
  int method;
  double rtol, atol;  // for PW_HE_LSODA / PW_HE_RK45
  double first_step, max_step;
};

// helium.collision (helium.py:317-380)
PW_HD_NOINLINE void he_collision(double T, double* q13, double* q31a, double* q31b, double* bq_he, double* bq_hep) {
  const double lt[9] = {3.75, 4.00, 4.25, 4.50, 4.75, 5.00, 5.25, 5.50, 5.75};
  const double g13[9] = {6.198E-2, 6.458E-2, 6.387E-2, 6.157E-2, 5.832E-2, 5.320E-2, 4.787E-2, 4.018E-2, 3.167E-2};
  const double g31a[9] = {2.389, 2.456, 2.275, 1.916, 1.496, 1.111, 8.003E-1, 5.660E-1, 3.944E-1};
  const double g31b[9] = {7.965E-1, 9.579E-1, 1.042, 1.015, 8.950E-1, 7.265E-1, 5.516E-1, 3.948E-1, 2.677E-1};
  double tt[9];
  for (int i = 0; i < 9; ++i) tt[i] = pw_pow(10.0, lt[i]);
  const int j = bracket(tt, 9, T, 0);
  const double G13 = np_interp_at(tt, g13, 9, T, j);
  const double G31a = np_interp_at(tt, g31a, 9, T, j);
  const double G31b = np_interp_at(tt, g31b, 9, T, j);
  const double kt = 8.617333262145179e-05 * T;
  const double k1 = 2.10E-8 * sqrt(13.6 / kt);
  *q13 = k1 * G13 * exp(-19.81 / kt);
  *q31a = k1 * G31a / 3 * exp(-0.80 / kt);
  *q31b = k1 * G31b / 3 * exp(-1.40 / kt);
  *bq_he = 1.75E-11 * pw_pow(300 / T, 0.75) * exp(-128E3 / T);
  *bq_hep = 1.25E-15 * pw_pow(300 / T, -0.25);
}

// Right-hand side helium.py:624-686.  The eight density-weighted rate
// coefficients and the two attenuation factors depend on the radius only, so they
// are cached per abscissa: the corrector iterations and the finite-difference
// Jacobian of one LSODA step all evaluate at the same x.
struct HeRhs {
  const double *rn, *v, *rho, *fh, *tau1, *tau3;
  const double* slope;  // [5][n]: np.interp slopes of v, rho, fh, tau1, tau3 per cell (HeWork::slope)
  unsigned e2tab;       // device: shared-memory address of the CTA's copy of kExp2Tab
  int n;
  double lk1;         // log(k1)
  double lc[8];       // log of alpha_rec_1, alpha_rec_3, q_13, q_31a, q_31b, Q_31, Q_He, Q_He+
  double kc[8];       // k1 * the same eight coefficients
  double phi1, phi3, ba31;
  // cache
  int j;
  double xc;
  double c[8], e1, e3, vv, rvv;  // rvv = 1 / vv (PW_RDIV)
  double sl[5], x0, b0[5];
  int jc;

  // out-of-line copy for the callers that are not latency critical (RK45 stages, the
  // `return_rates` table): keeps one instance of the two exps and the table walk there
  PW_HD_NOINLINE void coeffs_out(double x) { coeffs(x); }
  PW_HD void coeffs(double x) {
    const int jj = bracket(rn, n, x, j);
    j = jj;
    double v_, rho_, fh_, t1, t3;
    if (jj < 0 || jj >= n - 1 || rn[jj] == x) {
      v_ = np_interp_at_out(rn, v, n, x, jj);
      rho_ = np_interp_at_out(rn, rho, n, x, jj);
      fh_ = np_interp_at_out(rn, fh, n, x, jj);
      t1 = np_interp_at_out(rn, tau1, n, x, jj);
      t3 = np_interp_at_out(rn, tau3, n, x, jj);
    } else {
      if (jj != jc) {  // the cell's base values and slopes, kept while x stays in the cell
        jc = jj;
        x0 = rn[jj];
        b0[0] = v[jj]; b0[1] = rho[jj]; b0[2] = fh[jj]; b0[3] = tau1[jj]; b0[4] = tau3[jj];
#pragma unroll
        for (int q = 0; q < 5; ++q) sl[q] = slope[q * n + jj];
      }
      const double d = x - x0;
      v_ = sl[0] * d + b0[0];
      rho_ = sl[1] * d + b0[1];
      fh_ = sl[2] * d + b0[2];
      t1 = sl[3] * d + b0[3];
      t3 = sl[4] * d + b0[4];
    }
#if defined(PW_HE_LOGEXP)
    // literal form of helium.py:637-652: exp(log(k1) + log(rho) + log(coef)) * f_h_ion
    const double lr = lk1 + log(rho_);
    c[0] = exp(lr + lc[0]) * fh_;
    c[1] = exp(lr + lc[1]) * fh_;
    c[2] = exp(lr + lc[2]) * fh_;
    c[3] = exp(lr + lc[3]) * fh_;
    c[4] = exp(lr + lc[4]) * fh_;
    c[5] = exp(lr + lc[5]) * (1 - fh_);
    c[6] = exp(lr + lc[6]) * fh_;
    c[7] = exp(lr + lc[7]) * (1 - fh_);
#else
    // exp(log k1 + log rho + log coef) == k1 * rho * coef.  The reference takes the
    // detour through logarithms to avoid overflow, which cannot happen in binary64
    // here (k1 * rho < 1e45); its round trip carries ~|log|*eps ~ 1e-14 relative
    // rounding noise, the plain product 2e-16.  The product is the default; build with
    // -DPW_HE_LOGEXP for the literal form (host build: tests/test_hostsim_stages.py compares both).
    c[0] = (kc[0] * rho_) * fh_;
    c[1] = (kc[1] * rho_) * fh_;
    c[2] = (kc[2] * rho_) * fh_;
    c[3] = (kc[3] * rho_) * fh_;
    c[4] = (kc[4] * rho_) * fh_;
    c[5] = (kc[5] * rho_) * (1 - fh_);
    c[6] = (kc[6] * rho_) * fh_;
    c[7] = (kc[7] * rho_) * (1 - fh_);
#endif
    // attenuation factors e^{-tau}: table + series exp (pw_exp_neg, <= 3e-16 relative)
#if defined(__CUDA_ARCH__)
    e1 = pw_exp_neg_sh(-t1, e2tab);
    e3 = pw_exp_neg_sh(-t3, e2tab);
#else
    e1 = pw_exp_neg(-t1, kExp2Tab);
    e3 = pw_exp_neg(-t3, kExp2Tab);
#endif
    vv = v_;
    rvv = PW_DIV(1., v_);
    xc = x;
  }

  // terms[11] in the reference's order (helium.py:684-686) when `terms` != null
  PW_HD void eval(const double* y, double* dydx, double* terms) const {
    const double f_1 = y[0], f_3 = y[1];
    const double term_11 = (1 - f_1 - f_3) * c[0];
    const double term_12 = f_3 * ba31;
    const double term_13 = f_1 * phi1 * e1;
    const double term_14 = f_1 * c[2];
    const double term_15 = f_3 * c[3];
    const double term_16 = f_3 * c[4];
    const double term_17 = f_3 * c[5];
    const double term_18 = f_1 * c[6];
    const double term_19 = (1 - f_1 - f_3) * c[7];
    const double term_31 = (1 - f_1 - f_3) * c[1];
    const double term_33 = f_3 * phi3 * e3;
    if (dydx) {
      dydx[0] = PW_RDIV(term_11 + term_12 - term_13 - term_14 + term_15 + term_16 + term_17 - term_18 + term_19, vv, rvv);
      dydx[1] = PW_RDIV(term_31 - term_12 - term_33 + term_14 - term_15 - term_16 - term_17, vv, rvv);
    }
    if (terms) {
      terms[0] = term_13; terms[1] = term_33; terms[2] = term_11; terms[3] = term_31;
      terms[4] = term_12; terms[5] = term_14; terms[6] = term_15; terms[7] = term_16;
      terms[8] = term_17; terms[9] = term_18; terms[10] = term_19;
    }
  }

  // RHS concept of the integrators: at(x) selects the abscissa, eval(y, dydx) evaluates there
  PW_HD void at(double x) {
    if (!(x == xc)) coeffs(x);
  }
  // same, for call sites off the hot path: works on a copy so that this object's address
  // never escapes into a call (it has to stay in registers for the hot loop)
  PW_HD void at_cold(double x) {
    if (x == xc) return;
    HeRhs t = *this;
    t.coeffs_out(x);
    *this = t;
  }
  PW_HD void eval(const double* y, double* dydx) const { eval(y, dydx, nullptr); }
  PW_HD_NOINLINE void eval_out(const double* y, double* dydx) const { eval(y, dydx, nullptr); }
  PW_HD void eval_cold(const double* y, double* dydx) const {
    const HeRhs t = *this;
    double yy[2] = {y[0], y[1]}, dd[2];
    t.eval_out(yy, dd);
    dydx[0] = dd[0]; dydx[1] = dd[1];
  }
  PW_HD void operator()(double x, const double* y, double* dydx) {
    if (!(x == xc)) coeffs_out(x);
    eval(y, dydx, nullptr);
  }
};

// Per-model scratch (global memory on the device), Nr doubles each.
struct HeWork {
  double *rn, *tau1, *tau3, *tau1h, *tau3h;
  double* w;      // 5*Nr: [0,Nr) dr*density, [Nr,3Nr) previous profiles, [3Nr,5Nr) RK45 output staging
  double* slope;  // 5*Nr: np.interp slopes (fp[j+1]-fp[j])/(xp[j+1]-xp[j]) of v, rho, fh, tau1, tau3
};
constexpr int kHeWorkPerRadius = 15;  // doubles of HeWork per radius and model

// slope table of one interpolated profile, exactly as numpy forms each slope
PW_HD_NOINLINE void he_slopes(const double* xp, const double* fp, int n, double* out) {
#pragma unroll 1
  for (int j = 0; j < n - 1; ++j) out[j] = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  out[n - 1] = 0.0;
}

struct HeOut {
  double* f1;      // [Nr]
  double* f3;      // [Nr]
  double* scal;    // [4]: n_relax, nfe_total, nst_total, nje_total
  double* rates;   // [11][Nr] or null (helium.py:770-784)
  int* status;
};

// Integrate once over the radius grid; returns false when the reference's solver
// would have reported a problem (odeint warning / solve_ivp short output).
// kRk45 selects the integrator at compile time so that each kernel carries one of them.
template <bool kRk45>
PW_HD bool he_solve(const HeParams& p, HeRhs& rhs, const LsodaTables& tab, int nr, double* f1, double* f3,
                    double* ytmp, int* nfe, int* nst, int* nje) {
  const double* rn = rhs.rn;
  rhs.j = 0; rhs.jc = -2; rhs.xc = nan("");
  if (kRk45) {
    // solve_ivp(method='RK45', t_eval=r) (helium.py:702-710); interleaved (f1, f3)
    // samples go to the spare scratch and are split afterwards
    Rk45Opts o{p.rtol, p.atol, p.first_step, p.max_step};
    Rk45Stats st{0, 0, 0};
    const int got = rk45_solve<2>(rhs, rn, nr, p.y0, o, ytmp, &st);
    for (int k = 0; k < got; ++k) { f1[k] = ytmp[2 * k]; f3[k] = ytmp[2 * k + 1]; }
    *nfe += st.nfev; *nst += st.naccept;
    return got == nr;
  }
  LsodaOpts lo;
  if (p.method == PW_HE_ODEINT) { lo.rtol = 1.49012e-8; lo.atol = 1.49012e-8; }
  else { lo.rtol = p.rtol; lo.atol = p.atol; }
  // odeint: mxstep = 500 per output interval.  solve_ivp's LSODA wrapper advances one
  // internal step per call (itask = 5), so no such limit applies there.
  lo.mxstep = (p.method == PW_HE_ODEINT) ? 500 : 2000000000;
  lo.hmax = (p.max_step > 0.) ? p.max_step : 0.;
  lo.h0 = (p.first_step > 0.) ? p.first_step : 0.;
  double hi_rows[Lsoda<2, HeRhs>::kHiDoubles];
  Lsoda<2, HeRhs> s(rhs, tab, lo, hi_rows);
  f1[0] = p.y0[0];
  f3[0] = p.y0[1];
  // odeint = one lsoda(itask = 1) call per output radius: integrate past r[k], interpolate
  // back (intdy), carry on from the internal state.  Written as one flat loop over steps --
  // every radius the last step has passed is emitted before the next step is taken -- so
  // that the lanes of a warp, each with its own model, all sit in the same loop body.
  int k = 1;
  bool ok = s.begin(rn[0], p.y0, rn[1]) >= 0;
  while (ok && k < nr) {
    if (s.advance_one() < 0) { ok = false; break; }
    while (k < nr && s.reached(rn[k])) {
      double yo[2];
      s.intdy(rn[k], yo);
      f1[k] = yo[0];
      f3[k] = yo[1];
      ++k;
      s.new_call();  // the next output radius is a new lsoda call: nslast = nst
    }
  }
  // odeint stops at the first failing output point (ODEintWarning).  scipy does not
  // initialise its result array, so rows k.. of the reference's output are stale heap: there
  // is nothing to reproduce, and the callers below discard f1 / f3 of a failed solve.
  *nfe += s.nfe; *nst += s.nst; *nje += s.nje;
  return ok;
}

// Failure classes of the reference (helium.py):
//   status 2  the first solve warns / comes back short -> RuntimeError (:691-697, :708-710).
//             A short solve_ivp output inside the relax loop (:741-745) ends the reference with
//             a shape-mismatch ValueError at :753; the engine files it under status 2 as well.
//             Profiles NaN.
//   status 3  an `odeint` of the relax loop warns (:736).  The reference does not check it
//             and carries on with the rows scipy left uninitialised, so its result for such a
//             model is not reproducible (tests/golden/he_failure_classes.npz records the
//             run-to-run spread of the unmodified reference).  What IS deterministic is the
//             class itself -- everything up to the first warning is.  The engine therefore
//             stops at that warning: status 3, n_relax = the pass that warned, and f1 / f3 =
//             the profiles of the last pass that completed (the state the failing pass started
//             from).  chi^2 / log-likelihood treat status 3 as a failed model unless asked
//             otherwise (include/pwinds_b200.h).
// `r` radius profile in planetary radii (shared), v / rho / f_h per model.
template <bool kRk45>
PW_HD void he_population_model(const HeParams& p, const LsodaTables& tab, const double* __restrict__ r,
                               const double* __restrict__ v, const double* __restrict__ rho,
                               const double* __restrict__ f_h, int nr, const HeWork& w, const HeOut& out) {
  const double rs_cm = p.rs * 7.1492E+09;
  const double alpha_unit = (rs_cm * rs_cm) * p.vs * 1E5;                 // helium.py:553
  const double a_r1 = 1.54E-13 * pw_pow(p.T / 1E4, -0.486) / alpha_unit;     // :290, :555
  const double a_r3 = 2.10E-13 * pw_pow(p.T / 1E4, -0.778) / alpha_unit;     // :291, :556
  const double m_h = 1.67262192E-24 / (p.rhos * (rs_cm * rs_cm * rs_cm)); // :559-560
  const double phi_unit = p.vs * 1E5 / p.rs / 7.1492E+09;                 // :570
  const double phi_1 = p.rates[0] / phi_unit, phi_3 = p.rates[1] / phi_unit;
  const double a_1 = p.rates[2] / (rs_cm * rs_cm), a_3 = p.rates[3] / (rs_cm * rs_cm);
  const double a_h_1 = p.rates[4] / (rs_cm * rs_cm), a_h_3 = p.rates[5] / (rs_cm * rs_cm);
  double q13, q31a, q31b, bq_he, bq_hep;
  he_collision(p.T, &q13, &q31a, &q31b, &bq_he, &bq_hep);
  q13 /= alpha_unit; q31a /= alpha_unit; q31b /= alpha_unit; bq_he /= alpha_unit; bq_hep /= alpha_unit;
  const double bq31 = 5E-10 / alpha_unit;      // :598
  const double ba31 = 1.272E-4 / phi_unit;     // :599
  const double he_fraction = 1 - p.h_fraction;
  const double k1 = p.h_fraction / (p.h_fraction + 4 * he_fraction) / m_h;   // :616
  const double k2 = he_fraction / (p.h_fraction + 4 * he_fraction) / m_h;    // :617

  // r, dr in sonic units; total and neutral-H column densities, rectangle rule with
  // the local cell included (:604-621)
  for (int i = 0; i < nr; ++i) w.rn[i] = r[i] * p.r_pl / p.rs;
  {
    double c = 0.0, ch = 0.0;
    for (int i = nr - 1; i >= 0; --i) {
      const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
      w.w[i] = dr * rho[i];
      c += w.w[i];
      ch += w.w[i] * (1 - f_h[i]);
      w.tau1h[i] = k1 * a_h_1 * ch;
      w.tau3h[i] = k1 * a_h_3 * ch;
      w.tau1[i] = (p.y0[0] * k2 * a_1 * c + w.tau1h[i]);
      w.tau3[i] = (p.y0[1] * k2 * a_3 * c + w.tau3h[i]);
    }
  }
  HeRhs rhs;
  rhs.rn = w.rn; rhs.v = v; rhs.rho = rho; rhs.fh = f_h; rhs.tau1 = w.tau1; rhs.tau3 = w.tau3;
  rhs.slope = w.slope;
  rhs.e2tab = p.e2tab;
  rhs.n = nr;
  he_slopes(w.rn, v, nr, w.slope);
  he_slopes(w.rn, rho, nr, w.slope + nr);
  he_slopes(w.rn, f_h, nr, w.slope + 2 * nr);
#if defined(PW_HE_LOGEXP)
  rhs.lk1 = log(k1);
  rhs.lc[0] = log(a_r1); rhs.lc[1] = log(a_r3); rhs.lc[2] = log(q13); rhs.lc[3] = log(q31a);
  rhs.lc[4] = log(q31b); rhs.lc[5] = log(bq31); rhs.lc[6] = log(bq_he); rhs.lc[7] = log(bq_hep);
#endif
  rhs.kc[0] = k1 * a_r1; rhs.kc[1] = k1 * a_r3; rhs.kc[2] = k1 * q13; rhs.kc[3] = k1 * q31a;
  rhs.kc[4] = k1 * q31b; rhs.kc[5] = k1 * bq31; rhs.kc[6] = k1 * bq_he; rhs.kc[7] = k1 * bq_hep;
  rhs.phi1 = phi_1; rhs.phi3 = phi_3; rhs.ba31 = ba31;

  int nfe = 0, nst = 0, nje = 0, n_relax = 0, status = PW_OK;
  double* f1 = out.f1;
  double* f3 = out.f3;
  auto clamp = [&]() {  // helium.py:715-718
    for (int i = 0; i < nr; ++i) {
      if (f1[i] < 0) f1[i] = 1E-15;
      if (f3[i] < 0) f3[i] = 1E-15;
      if (f1[i] > 1.0) f1[i] = 1.0;
      if (f3[i] > 1.0) f3[i] = 1.0;
    }
  };
  double* p1 = w.w + nr;
  double* p3 = w.w + 2 * nr;
  double* ytmp = w.w + 3 * nr;
  // pass 0 = the first solve (helium.py:688-718); passes 1..max_n_relax = the relax loop
  // (:723-763).  One call site keeps a single copy of the integrator in the kernel.
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  for (int it = 0; it < n_pass; ++it) {
    double s1 = 0.0, s3 = 0.0;
    if (it > 0) {
      ++n_relax;
      // new optical depths from the current populations (:729-732); keep the previous
      // profiles for the convergence test (:753-756)
      double c1 = 0.0, c3 = 0.0;
      for (int i = nr - 1; i >= 0; --i) {
        c1 += w.w[i] * f1[i];
        c3 += w.w[i] * f3[i];
        w.tau1[i] = k2 * a_1 * c1 + w.tau1h[i];
        w.tau3[i] = k2 * a_3 * c3 + w.tau3h[i];
      }
      for (int i = 0; i < nr; ++i) { s1 += f1[i]; s3 += f3[i]; p1[i] = f1[i]; p3[i] = f3[i]; }
    }
    he_slopes(w.rn, w.tau1, nr, w.slope + 3 * nr);
    he_slopes(w.rn, w.tau3, nr, w.slope + 4 * nr);
    const bool ok = he_solve<kRk45>(p, rhs, tab, nr, f1, f3, ytmp, &nfe, &nst, &nje);
    if (!ok) {
      if (it == 0 || p.method != PW_HE_ODEINT) {
        status = PW_FAIL_HE_SOLVER;  // helium.py:693-697, :708-710 (short solve_ivp output)
        break;
      }
      // helium.py:736 is outside the warnings->error guard: the reference's result is undefined
      // from here on (see above).  Stop, and hand back the last completed pass.
      status = PW_WARN_HE_RELAX;
      for (int i = 0; i < nr; ++i) { f1[i] = p1[i]; f3[i] = p3[i]; }
      break;
    }
    clamp();
    if (it > 0) {
      double d1 = 0.0, d3 = 0.0;
      for (int i = 0; i < nr; ++i) { d1 += f1[i] - p1[i]; d3 += f3[i] - p3[i]; }
      const double rel1 = fabs(d1) / s1, rel3 = fabs(d3) / s3;
      if (rel1 < p.convergence && rel3 < p.convergence) break;
    }
  }
  if (status == PW_FAIL_HE_SOLVER) {
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) { f1[i] = nan_; f3[i] = nan_; }
  } else if (out.rates) {
    // _fun(r, [f_1_r, f_3_r], rates=True) * phi_unit  (:770-773, :684-686)
    HeRhs rr = rhs;  // a copy, so that the solver's own object never has its address taken
    for (int i = 0; i < nr; ++i) {
      double y[2] = {f1[i], f3[i]}, t[11];
      rr.coeffs_out(w.rn[i]);
      rr.eval(y, nullptr, t);
      for (int k = 0; k < 11; ++k) out.rates[k * nr + i] = t[k] * phi_unit;
    }
  }
  out.scal[0] = (double)n_relax; out.scal[1] = (double)nfe; out.scal[2] = (double)nst; out.scal[3] = (double)nje;
  *out.status = status;
}


#if defined(__CUDACC__) && defined(PW_HE_PS_KERNEL)
// ---------------------------------------------------------------------------------------------
// EXPERIMENT (developer build -DPW_HE_PS_KERNEL, scripts/he_ps.py; not in the shipped library).
// The same model, scheduled by the warp (device only; method 'odeint' / 'LSODA').  Measured on the
// 64 x 64 grid: bit-identical to he_population_model and SLOWER -- 71 ms against
// 35 ms with the production schedule, 90 ms against 75 ms at 32 models per warp.  The control build
// (-DPW_PS_FREE: same state machine, no vote) takes 101 ms at one model per warp against 48 ms: cut into
// blocks behind a switch the step keeps ~110 values alive across every block and the compiler spills
// (1840 bytes of stack against 1280), which costs more than the alignment of the lanes recovers (2.5x over
// the unaligned state machine at 32 models per warp).  Kept as the record of the experiment.
//
// he_population_model runs one model per lane from start to end; the lanes of a warp drift apart
// (different numbers of corrector iterations, rejected steps, Jacobians, order changes) and the
// hardware serialises them: at eight models per warp an instruction is issued for two of them on
// average.  Here every lane keeps its model in the same registers, but the step is cut into blocks
// (Lsoda::ps_*) and the WARP picks which block runs next: the one most lanes are waiting in, or one a
// lane has waited `kPsMaxWait` picks for.  Lanes in that phase execute the block together from its
// first instruction, the others skip it; nothing is shared between lanes, and every lane executes
// exactly the statements of he_population_model in the same order: bit-identical results
// (scripts/he_ps.py compares f1, f3, counters and status of the whole grid with torch.equal).
//
// Phases of a lane: kPsPredict .. kPsOrder (the LSODA step), kHeEmit (outputs the last step passed),
// kHePass (end of a solve + beginning of the next relax pass: O(Nr) work), kHeDone.
enum { kHeEmit = 6, kHePass = 7, kHeDone = 8, kHeNPhase = 9 };
constexpr int kPsMaxWait = 6;

__device__ __forceinline__ void he_population_ps(const HeParams& p, const LsodaTables& tab, const double* __restrict__ r,
                                                 const double* __restrict__ v, const double* __restrict__ rho,
                                                 const double* __restrict__ f_h, int nr, const HeWork& w,
                                                 const HeOut& out, bool has_model) {
  // ---- per-model constants (he_population_model, same statements) ----
  const double rs_cm = p.rs * 7.1492E+09;
  const double alpha_unit = (rs_cm * rs_cm) * p.vs * 1E5;
  const double m_h = 1.67262192E-24 / (p.rhos * (rs_cm * rs_cm * rs_cm));
  const double phi_unit = p.vs * 1E5 / p.rs / 7.1492E+09;
  const double a_1 = p.rates[2] / (rs_cm * rs_cm), a_3 = p.rates[3] / (rs_cm * rs_cm);
  const double he_fraction = 1 - p.h_fraction;
  const double k2 = he_fraction / (p.h_fraction + 4 * he_fraction) / m_h;
  HeRhs rhs;
  int phase = kHeDone;
  if (has_model) {
    const double a_r1 = 1.54E-13 * pw_pow(p.T / 1E4, -0.486) / alpha_unit;
    const double a_r3 = 2.10E-13 * pw_pow(p.T / 1E4, -0.778) / alpha_unit;
    const double phi_1 = p.rates[0] / phi_unit, phi_3 = p.rates[1] / phi_unit;
    const double a_h_1 = p.rates[4] / (rs_cm * rs_cm), a_h_3 = p.rates[5] / (rs_cm * rs_cm);
    double q13, q31a, q31b, bq_he, bq_hep;
    he_collision(p.T, &q13, &q31a, &q31b, &bq_he, &bq_hep);
    q13 /= alpha_unit; q31a /= alpha_unit; q31b /= alpha_unit; bq_he /= alpha_unit; bq_hep /= alpha_unit;
    const double bq31 = 5E-10 / alpha_unit;
    const double ba31 = 1.272E-4 / phi_unit;
    const double k1 = p.h_fraction / (p.h_fraction + 4 * he_fraction) / m_h;
    for (int i = 0; i < nr; ++i) w.rn[i] = r[i] * p.r_pl / p.rs;
    {
      double c = 0.0, ch = 0.0;
      for (int i = nr - 1; i >= 0; --i) {
        const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
        w.w[i] = dr * rho[i];
        c += w.w[i];
        ch += w.w[i] * (1 - f_h[i]);
        w.tau1h[i] = k1 * a_h_1 * ch;
        w.tau3h[i] = k1 * a_h_3 * ch;
        w.tau1[i] = (p.y0[0] * k2 * a_1 * c + w.tau1h[i]);
        w.tau3[i] = (p.y0[1] * k2 * a_3 * c + w.tau3h[i]);
      }
    }
    rhs.rn = w.rn; rhs.v = v; rhs.rho = rho; rhs.fh = f_h; rhs.tau1 = w.tau1; rhs.tau3 = w.tau3;
    rhs.slope = w.slope;
    rhs.e2tab = p.e2tab;
    rhs.n = nr;
    he_slopes(w.rn, v, nr, w.slope);
    he_slopes(w.rn, rho, nr, w.slope + nr);
    he_slopes(w.rn, f_h, nr, w.slope + 2 * nr);
    rhs.kc[0] = k1 * a_r1; rhs.kc[1] = k1 * a_r3; rhs.kc[2] = k1 * q13; rhs.kc[3] = k1 * q31a;
    rhs.kc[4] = k1 * q31b; rhs.kc[5] = k1 * bq31; rhs.kc[6] = k1 * bq_he; rhs.kc[7] = k1 * bq_hep;
    rhs.phi1 = phi_1; rhs.phi3 = phi_3; rhs.ba31 = ba31;
    phase = kHePass;
  }
  LsodaOpts lo;
  if (p.method == PW_HE_ODEINT) { lo.rtol = 1.49012e-8; lo.atol = 1.49012e-8; }
  else { lo.rtol = p.rtol; lo.atol = p.atol; }
  lo.mxstep = (p.method == PW_HE_ODEINT) ? 500 : 2000000000;
  lo.hmax = (p.max_step > 0.) ? p.max_step : 0.;
  lo.h0 = (p.first_step > 0.) ? p.first_step : 0.;
  double hi_rows[Lsoda<2, HeRhs>::kHiDoubles];
  Lsoda<2, HeRhs> s(rhs, tab, lo, hi_rows);
  int nfe = 0, nst = 0, nje = 0, n_relax = 0, status = PW_OK;
  int it = -1;          // pass being integrated (-1: none yet)
  int k = 1;            // next output radius
  bool ok = true;       // the running solve has met no problem
  double s1 = 0.0, s3 = 0.0;
  double* f1 = out.f1;
  double* f3 = out.f3;
  double* p1 = w.w + nr;
  double* p3 = w.w + 2 * nr;
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  int wait = 0;
  const unsigned full = 0xffffffffu;
  for (;;) {
    // ---- the warp's choice: most populated phase, unless some lane has waited too long ----
#if defined(PW_PS_FREE)   // control experiment: no vote, every lane runs the block it is in (the code shape alone)
    if (__all_sync(full, phase == kHeDone)) break;
    const int run = phase;
    if (phase == kHeDone) continue;
#else
    const unsigned peers = __match_any_sync(full, phase);
    unsigned key = 0;
    if (phase != kHeDone)
      key = ((wait >= kPsMaxWait) ? (1u << 16) : 0u) | ((unsigned)__popc(peers) << 8) | (unsigned)(kHeNPhase - phase);
    key = __reduce_max_sync(full, key);
    if (key == 0) break;   // every lane is done
    const int run = kHeNPhase - (int)(key & 0xffu);
    if (phase != run) { ++wait; continue; }
    wait = 0;
#endif
    switch (run) {
      case Lsoda<2, HeRhs>::kPsPredict: phase = s.ps_predict(); break;
      case Lsoda<2, HeRhs>::kPsIter: phase = s.ps_iter(); break;
      case Lsoda<2, HeRhs>::kPsJac: phase = s.ps_jac(); break;
      case Lsoda<2, HeRhs>::kPsTest: phase = s.ps_test(); break;
      case Lsoda<2, HeRhs>::kPsFail: phase = s.ps_fail(); break;
      case Lsoda<2, HeRhs>::kPsOrder: phase = s.ps_order(); break;
      case kHeEmit: {
        // a step has ended (Lsoda::kPsStepDone == kHeEmit): he_solve's loop body after advance_one()
        if (s.ps_rc < 0) { ok = false; phase = kHePass; break; }
        while (k < nr && s.reached(w.rn[k])) {
          double yo[2];
          s.intdy(w.rn[k], yo);
          f1[k] = yo[0];
          f3[k] = yo[1];
          ++k;
          s.new_call();
        }
        if (k < nr) { s.ps_new_step(); phase = Lsoda<2, HeRhs>::kPsPredict; }
        else phase = kHePass;
      } break;
      default: {   // kHePass: end of pass `it` (if any), then the beginning of the next one
        bool finished = false;
        if (it >= 0) {
          nfe += s.nfe; nst += s.nst; nje += s.nje;
          if (!ok) {
            if (it == 0 || p.method != PW_HE_ODEINT) {
              status = PW_FAIL_HE_SOLVER;
            } else {
              status = PW_WARN_HE_RELAX;
              for (int i = 0; i < nr; ++i) { f1[i] = p1[i]; f3[i] = p3[i]; }
            }
            finished = true;
          } else {
            for (int i = 0; i < nr; ++i) {   // clamp, helium.py:715-718
              if (f1[i] < 0) f1[i] = 1E-15;
              if (f3[i] < 0) f3[i] = 1E-15;
              if (f1[i] > 1.0) f1[i] = 1.0;
              if (f3[i] > 1.0) f3[i] = 1.0;
            }
            if (it > 0) {
              double d1 = 0.0, d3 = 0.0;
              for (int i = 0; i < nr; ++i) { d1 += f1[i] - p1[i]; d3 += f3[i] - p3[i]; }
              const double rel1 = fabs(d1) / s1, rel3 = fabs(d3) / s3;
              if (rel1 < p.convergence && rel3 < p.convergence) finished = true;
            }
            if (it + 1 >= n_pass) finished = true;
          }
        }
        if (finished) {
          if (status == PW_FAIL_HE_SOLVER) {
            const double nan_ = nan("");
            for (int i = 0; i < nr; ++i) { f1[i] = nan_; f3[i] = nan_; }
          } else if (out.rates) {
            HeRhs rr = rhs;
            for (int i = 0; i < nr; ++i) {
              double y[2] = {f1[i], f3[i]}, t[11];
              rr.coeffs_out(w.rn[i]);
              rr.eval(y, nullptr, t);
              for (int q = 0; q < 11; ++q) out.rates[q * nr + i] = t[q] * phi_unit;
            }
          }
          out.scal[0] = (double)n_relax; out.scal[1] = (double)nfe; out.scal[2] = (double)nst; out.scal[3] = (double)nje;
          *out.status = status;
          phase = kHeDone;
          break;
        }
        ++it;
        s1 = 0.0; s3 = 0.0;
        if (it > 0) {
          ++n_relax;
          double c1 = 0.0, c3 = 0.0;
          for (int i = nr - 1; i >= 0; --i) {
            c1 += w.w[i] * f1[i];
            c3 += w.w[i] * f3[i];
            w.tau1[i] = k2 * a_1 * c1 + w.tau1h[i];
            w.tau3[i] = k2 * a_3 * c3 + w.tau3h[i];
          }
          for (int i = 0; i < nr; ++i) { s1 += f1[i]; s3 += f3[i]; p1[i] = f1[i]; p3[i] = f3[i]; }
        }
        he_slopes(w.rn, w.tau1, nr, w.slope + 3 * nr);
        he_slopes(w.rn, w.tau3, nr, w.slope + 4 * nr);
        // he_solve: first call of the pass
        rhs.j = 0; rhs.jc = -2; rhs.xc = nan("");
        f1[0] = p.y0[0];
        f3[0] = p.y0[1];
        k = 1;
        ok = s.begin(w.rn[0], p.y0, w.rn[1]) >= 0;
        if (ok && k < nr) { s.ps_new_step(); phase = Lsoda<2, HeRhs>::kPsPredict; }
        else phase = kHePass;   // (begin failed, or a one-point grid: the pass ends at once)
      } break;
    }
  }
}
#endif

}  // namespace pw
