// pw_oxygen.cuh -- O II fraction profile of one model (reference: p_winds/oxygen.py:186-491
// `ion_fraction`, :103-182 rate coefficients).  One equation, integrated with the Radau
// replica (the reference's default), the LSODA replica (method='odeint') or the RK45 replica.
// One thread per model, like the helium stage.  SURVEY.md 8(f-3).
#pragma once
#include "pw_common.cuh"
#include "pw_lsoda.cuh"
#include "pw_radau.cuh"
#include "pw_rk45.cuh"

namespace pw {

enum { PW_O_RADAU = 0, PW_O_ODEINT = 1, PW_O_RK45 = 2 };

struct OParams {
  double T, h_fraction, vs, rs, rhos;  // per model
  double r_pl, o_fraction, y0;
  int relax, max_n_relax;
  double convergence;
  double rates[4];  // phi_oi, a_oi, a_h_oi, a_he (cgs; oxygen.radiative_processes)
  int method;
  double rtol, atol, first_step, max_step;
};

// Right-hand side oxygen.py:364-411.  As in the helium stage the rate coefficients are
// formed as k * coef * rho instead of exp(log k + log rho + log coef) (algebraic identity;
// see pw_helium.cuh), literal form with -DPW_HE_LOGEXP.
struct ORhs {
  const double *rn, *v, *rho, *fh, *fhe, *tau;
  int n, hint;
  double k1, k2, ion, ct_oi_hp, rec, ct_oii_h, phi;
  PW_HD void terms(double x, double f, double* t, double* vv) {
    const int j = bracket(rn, n, x, hint);
    hint = j;
    const double v_ = np_interp_at(rn, v, n, x, j);
    const double rho_ = np_interp_at(rn, rho, n, x, j);
    const double fh_ = np_interp_at(rn, fh, n, x, j);
    const double fhe_ = np_interp_at(rn, fhe, n, x, j);
    const double t_oi = np_interp_at(rn, tau, n, x, j);
#if defined(PW_HE_LOGEXP)
    const double lt1 = log(k1) + log(rho_), lt2 = log(k2) + log(rho_);
    const double ion_ne = exp(lt1 + log(ion)) * fh_ + exp(lt2 + log(ion)) * fhe_;
    const double ct1 = exp(lt1 + log(ct_oi_hp)) * fh_;
    const double rec_ne = exp(lt1 + log(rec)) * fh_ + exp(lt2 + log(rec)) * fhe_;
    const double ct2 = exp(lt1 + log(ct_oii_h)) * (1 - fh_);
#else
    const double ion_ne = ((k1 * ion) * rho_) * fh_ + ((k2 * ion) * rho_) * fhe_;
    const double ct1 = ((k1 * ct_oi_hp) * rho_) * fh_;
    const double rec_ne = ((k1 * rec) * rho_) * fh_ + ((k2 * rec) * rho_) * fhe_;
    const double ct2 = ((k1 * ct_oii_h) * rho_) * (1 - fh_);
#endif
    t[0] = (1 - f) * phi * exp(-t_oi);  // photoionisation
    t[1] = (1 - f) * ion_ne;            // electron-impact ionisation
    t[2] = (1 - f) * ct1;               // charge exchange with H+
    t[3] = f * rec_ne;                  // recombination
    t[4] = f * ct2;                     // charge exchange of O II with H
    *vv = v_;
  }
  PW_HD void operator()(double x, const double* y, double* dydx) {
    double t[5], vv;
    terms(x, y[0], t, &vv);
    dydx[0] = (t[0] + t[1] + t[2] - t[3] - t[4]) / vv;
  }
  // RHS concept of the LSODA replica
  double xs;
  PW_HD void at(double x) { xs = x; }
  PW_HD void at_cold(double x) { xs = x; }
  PW_HD void eval(const double* y, double* dydx) {
    double t[5], vv;
    terms(xs, y[0], t, &vv);
    dydx[0] = (t[0] + t[1] + t[2] - t[3] - t[4]) / vv;
  }
  PW_HD void eval_cold(const double* y, double* dydx) { eval(y, dydx); }
};

constexpr int kOWorkPerRadius = 8;  // rn, w (dr*rho), tau_h, tau_he, tau, prev, ytmp, spare
struct OWork {
  double *rn, *w, *tau_h, *tau_he, *tau, *prev, *ytmp;
};
struct OOut {
  double* f_o;    // [Nr]
  double* scal;   // [4]: n_relax, nfev total, n_solves, -
  double* rates;  // [5][Nr] or null: photoionization, recombination, e_impact_ionization,
                  //                  charge_exchange_HII, charge_exchange_HI (oxygen.py:484-489)
  int* status;
};

PW_HD bool o_solve(const OParams& p, ORhs& rhs, const LsodaTables& tab, int nr, double* f, double* ytmp, int* nfev) {
  const double* rn = rhs.rn;
  rhs.hint = 0;
  const double y0 = p.y0;
  const Rk45Opts o{p.rtol, p.atol, p.first_step, p.max_step};
  if (p.method == PW_O_RADAU) {
    RadauStats st{0, 0, 0, 0, 0};
    const int got = radau_solve<1>(rhs, rn, nr, &y0, o, f, &st);
    *nfev += st.nfev;
    return got == nr;
  }
  if (p.method == PW_O_RK45) {
    Rk45Stats st{0, 0, 0};
    const int got = rk45_solve<1>(rhs, rn, nr, &y0, o, f, &st);
    *nfev += st.nfev;
    return got == nr;
  }
  // odeint (oxygen.py:413-423): LSODA at rtol = atol = 1.49012e-8, mxstep = 500 per output radius
  LsodaOpts lo{1.49012e-8, 1.49012e-8, 500, 0., 0.};
  double hi_rows[Lsoda<1, ORhs>::kHiDoubles];
  Lsoda<1, ORhs> s(rhs, tab, lo, hi_rows);
  f[0] = y0;
  bool ok = true;
  int k = 1;
  for (; k < nr; ++k) {
    double yo[1];
    if (s.call(rn[0], &y0, rn[k], yo) < 0) { ok = false; break; }
    f[k] = yo[0];
  }
  *nfev += s.nfe;
  (void)ytmp;
  return ok;
}

PW_HD void o_ion_fraction_model(const OParams& p, const LsodaTables& tab, const double* __restrict__ r,
                                const double* __restrict__ v, const double* __restrict__ rho,
                                const double* __restrict__ f_h, const double* __restrict__ f_he, int nr,
                                const OWork& w, const OOut& out) {
  const double rs_cm = p.rs * 7.1492E+09;
  const double alpha_unit = (rs_cm * rs_cm) * p.vs * 1E5;                        // oxygen.py:306
  const double rec = 3.25E-12 * pw_pow(300 / p.T, 0.66) / alpha_unit;            // :148, :308
  const double m_h = 1.67262192E-24 / (p.rhos * (rs_cm * rs_cm * rs_cm));        // :311-312
  const double phi_unit = p.vs * 1E5 / p.rs / 7.1492E+09;                        // :317
  const double phi_oi = p.rates[0] / phi_unit;
  const double a_oi = p.rates[1] / (rs_cm * rs_cm), a_h_oi = p.rates[2] / (rs_cm * rs_cm),
               a_he = p.rates[3] / (rs_cm * rs_cm);
  const double ratio = 11.3 / (8.617333262145179e-05 * p.T);                      // :121-125
  const double ion = 3.59E-8 * pw_pow(0.073 + ratio, -1.0) * pw_pow(ratio, 0.34) * exp(-ratio) / alpha_unit;
  const double ct_oii_h = 5.66E-10 * pw_pow(300 / p.T, -0.36) * exp(8.6 / p.T) / alpha_unit;   // :176-177
  const double ct_oi_hp = 7.31E-10 * pw_pow(300 / p.T, -0.23) * exp(-226 / p.T) / alpha_unit;  // :179-180
  const double he_fraction = 1 - p.h_fraction;
  const double den = p.h_fraction + 4 * he_fraction + 16 * p.o_fraction;
  const double k1 = p.h_fraction / den / m_h, k2 = he_fraction / den / m_h, k3 = p.o_fraction / den / m_h;

  for (int i = 0; i < nr; ++i) w.rn[i] = r[i] * p.r_pl / p.rs;
  {
    double c = 0.0, ch = 0.0, che = 0.0;
    for (int i = nr - 1; i >= 0; --i) {
      const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
      w.w[i] = dr * rho[i];
      c += w.w[i];
      ch += w.w[i] * (1 - f_h[i]);
      che += w.w[i] * he_fraction * (1 - f_he[i]);
      w.tau_h[i] = k1 * a_h_oi * ch;
      w.tau_he[i] = k2 * a_he * che;
      w.tau[i] = (1 - p.y0) * k3 * a_oi * c + w.tau_h[i] + w.tau_he[i];
    }
  }
  ORhs rhs;
  rhs.rn = w.rn; rhs.v = v; rhs.rho = rho; rhs.fh = f_h; rhs.fhe = f_he; rhs.tau = w.tau;
  rhs.n = nr; rhs.hint = 0; rhs.xs = 0.0;
  rhs.k1 = k1; rhs.k2 = k2; rhs.ion = ion; rhs.ct_oi_hp = ct_oi_hp; rhs.rec = rec; rhs.ct_oii_h = ct_oii_h;
  rhs.phi = phi_oi;

  int nfev = 0, n_relax = 0, n_solves = 0, status = PW_OK;
  double* f = out.f_o;
  auto clamp = [&]() {  // oxygen.py:437-438
    for (int i = 0; i < nr; ++i) {
      if (f[i] < 0) f[i] = 1E-15;
      if (f[i] > 1.0) f[i] = 1.0;
    }
  };
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  for (int it = 0; it < n_pass; ++it) {
    double sp = 0.0;
    if (it > 0) {
      ++n_relax;
      double c = 0.0;
      for (int i = nr - 1; i >= 0; --i) {  // :447-450
        c += w.w[i] * (1 - f[i]);
        w.tau[i] = k3 * a_oi * c + w.tau_h[i] + w.tau_he[i];
      }
      for (int i = 0; i < nr; ++i) { sp += f[i]; w.prev[i] = f[i]; }
    }
    ++n_solves;
    const bool ok = o_solve(p, rhs, tab, nr, f, w.ytmp, &nfev);
    if (!ok && (it == 0 || p.method != PW_O_ODEINT)) {
      // first solve: RuntimeError (:415-433); a short solve_ivp result inside the relax loop
      // cannot be used by the reference either (shape mismatch) -- reported as a failure
      status = PW_FAIL_HE_SOLVER;
      break;
    }
    if (!ok) {   // oxygen.py:455
      // An `odeint` of the relax loop warned.  The reference does not check it and carries on with rows scipy
      // left uninitialised (helium.py:736 has the same pattern; pw_helium.cuh): undefined from here on.  Stop,
      // hand back the last completed pass, status PW_WARN_HE_RELAX.
      status = PW_WARN_HE_RELAX;
      for (int i = 0; i < nr; ++i) f[i] = w.prev[i];
      break;
    }
    clamp();
    if (it > 0) {
      double d = 0.0;
      for (int i = 0; i < nr; ++i) d += f[i] - w.prev[i];
      if (fabs(d) / sp < p.convergence) break;  // :470-475
    }
  }
  if (status == PW_FAIL_HE_SOLVER) {
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) f[i] = nan_;
  } else if (out.rates) {
    rhs.hint = 0;
    for (int i = 0; i < nr; ++i) {
      double t[5], vv;
      rhs.terms(w.rn[i], f[i], t, &vv);
      out.rates[0 * nr + i] = t[0] * phi_unit;
      out.rates[1 * nr + i] = t[3] * phi_unit;
      out.rates[2 * nr + i] = t[1] * phi_unit;
      out.rates[3 * nr + i] = t[2] * phi_unit;
      out.rates[4 * nr + i] = t[4] * phi_unit;
    }
  }
  out.scal[0] = (double)n_relax; out.scal[1] = (double)nfev; out.scal[2] = (double)n_solves; out.scal[3] = 0.0;
  *out.status = status;
}

}  // namespace pw
