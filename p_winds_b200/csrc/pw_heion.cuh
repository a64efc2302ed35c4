// pw_heion.cuh -- He II fraction profile of one model without singlet / triplet bookkeeping
// (reference: p_winds/helium.py:789-1032 `ion_fraction`, :294-313 `recombination_all`,
// :384-409 `charge_transfer`).  One equation; Radau replica (reference default), LSODA replica
// ('odeint') or RK45 replica.  One thread per model.  SURVEY.md 8(f-3).
#pragma once
#include "pw_common.cuh"
#include "pw_lsoda.cuh"
#include "pw_radau.cuh"
#include "pw_rk45.cuh"
#include "pw_oxygen.cuh"  // method enum, work layout and the solver dispatch are shared

namespace pw {

struct HeIonParams {
  double T, h_fraction, vs, rs, rhos;  // per model
  double r_pl, y0;
  int relax, max_n_relax;
  double convergence;
  double rates[3];  // phi_he, a_he, a_h (helium.radiative_processes(combined_ionization=True))
  int method;       // PW_O_RADAU / PW_O_ODEINT / PW_O_RK45
  double rtol, atol, first_step, max_step;
};

// helium.py:944-975; product form of the rate coefficients (see pw_helium.cuh)
struct HeIonRhs {
  const double *rn, *v, *rho, *fh, *tau;
  int n, hint;
  double k1, k2, rec, ct_he_hp, ct_hep_h, phi;
  PW_HD void rhs_at(double x, const double* y, double* d) {
    const int j = bracket(rn, n, x, hint);
    hint = j;
    const double f = y[0];
    const double v_ = np_interp_at(rn, v, n, x, j);
    const double rho_ = np_interp_at(rn, rho, n, x, j);
    const double fh_ = np_interp_at(rn, fh, n, x, j);
    const double t_he = np_interp_at(rn, tau, n, x, j);
#if defined(PW_HE_LOGEXP)
    const double lt1 = log(k1) + log(rho_), lt2 = log(k2) + log(rho_);
    const double ct1 = exp(lt1 + log(ct_he_hp)) * fh_;
    const double rec_ne = exp(lt1 + log(rec)) * fh_ + exp(lt2 + log(rec)) * f;
    const double ct2 = exp(lt1 + log(ct_hep_h)) * (1 - fh_);
#else
    const double ct1 = ((k1 * ct_he_hp) * rho_) * fh_;
    const double rec_ne = ((k1 * rec) * rho_) * fh_ + ((k2 * rec) * rho_) * f;
    const double ct2 = ((k1 * ct_hep_h) * rho_) * (1 - fh_);
#endif
    const double term1 = (1 - f) * phi * exp(-t_he);
    const double term2 = (1 - f) * ct1;
    const double term3 = f * rec_ne;
    const double term4 = f * ct2;
    d[0] = (term1 + term2 - term3 - term4) / v_;
  }
  PW_HD void operator()(double x, const double* y, double* d) { rhs_at(x, y, d); }
  double xs;
  PW_HD void at(double x) { xs = x; }
  PW_HD void at_cold(double x) { xs = x; }
  PW_HD void eval(const double* y, double* d) { rhs_at(xs, y, d); }
  PW_HD void eval_cold(const double* y, double* d) { rhs_at(xs, y, d); }
};

// one solve over the radius grid with the chosen integrator (shared shape with o_solve)
template <class RHS>
PW_HD bool scalar_ode_solve(int method, double y0, const Rk45Opts& o, RHS& rhs, const LsodaTables& tab, int nr,
                            double* f, int* nfev) {
  const double* rn = rhs.rn;
  rhs.hint = 0;
  if (method == PW_O_RADAU) {
    RadauStats st{0, 0, 0, 0, 0};
    const int got = radau_solve<1>(rhs, rn, nr, &y0, o, f, &st);
    *nfev += st.nfev;
    return got == nr;
  }
  if (method == PW_O_RK45) {
    Rk45Stats st{0, 0, 0};
    const int got = rk45_solve<1>(rhs, rn, nr, &y0, o, f, &st);
    *nfev += st.nfev;
    return got == nr;
  }
  LsodaOpts lo{1.49012e-8, 1.49012e-8, 500, 0., 0.};
  double hi_rows[Lsoda<1, RHS>::kHiDoubles];
  Lsoda<1, RHS> s(rhs, tab, lo, hi_rows);
  f[0] = y0;
  bool ok = true;
  int k = 1;
  for (; k < nr; ++k) {
    double yo[1];
    if (s.call(rn[0], &y0, rn[k], yo) < 0) { ok = false; break; }
    f[k] = yo[0];
  }
  *nfev += s.nfe;
  return ok;
}

// work: rn, w (dr*rho), tau_h, tau, prev  (uses the first five slots of an OWork-sized block)
PW_HD void he_ion_fraction_model(const HeIonParams& p, const LsodaTables& tab, const double* __restrict__ r,
                                 const double* __restrict__ v, const double* __restrict__ rho,
                                 const double* __restrict__ f_h, int nr, const OWork& w, double* f, double* scal,
                                 int* status_out) {
  const double rs_cm = p.rs * 7.1492E+09;
  const double alpha_unit = (rs_cm * rs_cm) * p.vs * 1E5;                          // helium.py:887
  const double rec = 4.6E-12 * pw_pow(p.T / 300, -0.64) / alpha_unit;              // :312
  const double m_h = 1.67262192E-24 / (p.rhos * (rs_cm * rs_cm * rs_cm));          // :892-893
  const double phi_unit = p.vs * 1E5 / p.rs / 7.1492E+09;
  const double phi_he = p.rates[0] / phi_unit;
  const double a_he = p.rates[1] / (rs_cm * rs_cm), a_h = p.rates[2] / (rs_cm * rs_cm);
  const double ct_hep_h = 1.25E-15 * pw_pow(300 / p.T, -0.25) / alpha_unit;                       // :403
  const double ct_he_hp = 1.75E-11 * pw_pow(300 / p.T, 0.75) * exp(-128000 / p.T) / alpha_unit;   // :406-407
  const double he_fraction = 1 - p.h_fraction;
  const double k1 = p.h_fraction / (p.h_fraction + 4 * he_fraction) / m_h;
  const double k2 = he_fraction / (p.h_fraction + 4 * he_fraction) / m_h;
  double* tau_h = w.tau_h;
  for (int i = 0; i < nr; ++i) w.rn[i] = r[i] * p.r_pl / p.rs;
  {
    double c = 0.0, ch = 0.0;
    for (int i = nr - 1; i >= 0; --i) {
      const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
      w.w[i] = dr * rho[i];
      c += w.w[i];
      ch += w.w[i] * (1 - f_h[i]);
      tau_h[i] = k1 * a_h * ch;
      w.tau[i] = (p.y0 * k2 * a_he * c + tau_h[i]);  // sic: f, not 1 - f (helium.py:941)
    }
  }
  HeIonRhs rhs;
  rhs.rn = w.rn; rhs.v = v; rhs.rho = rho; rhs.fh = f_h; rhs.tau = w.tau; rhs.n = nr; rhs.hint = 0; rhs.xs = 0.0;
  rhs.k1 = k1; rhs.k2 = k2; rhs.rec = rec; rhs.ct_he_hp = ct_he_hp; rhs.ct_hep_h = ct_hep_h; rhs.phi = phi_he;
  const Rk45Opts o{p.rtol, p.atol, p.first_step, p.max_step};
  int nfev = 0, n_relax = 0, n_solves = 0, status = PW_OK;
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  for (int it = 0; it < n_pass; ++it) {
    double sp = 0.0;
    if (it > 0) {
      ++n_relax;
      double c = 0.0;
      for (int i = nr - 1; i >= 0; --i) {  // helium.py:1002-1004
        c += w.w[i] * (1 - f[i]);
        w.tau[i] = k2 * a_he * c + tau_h[i];
      }
      for (int i = 0; i < nr; ++i) { sp += f[i]; w.prev[i] = f[i]; }
    }
    ++n_solves;
    const bool ok = scalar_ode_solve(p.method, p.y0, o, rhs, tab, nr, f, &nfev);
    if (!ok && (it == 0 || p.method != PW_O_ODEINT)) {
      status = PW_FAIL_HE_SOLVER;
      break;
    }
    if (!ok) {   // helium.py:1007
      // An `odeint` of the relax loop warned.  The reference does not check it and carries on with rows scipy
      // left uninitialised (helium.py:736 has the same pattern; pw_helium.cuh): undefined from here on.  Stop,
      // hand back the last completed pass, status PW_WARN_HE_RELAX.
      status = PW_WARN_HE_RELAX;
      for (int i = 0; i < nr; ++i) f[i] = w.prev[i];
      break;
    }
    if (it > 0) {  // the reference clamps only inside the relax loop (:1015-1016)
      for (int i = 0; i < nr; ++i) {
        if (f[i] < 0) f[i] = 1E-15;
        if (f[i] > 1.0) f[i] = 1.0;
      }
      double d = 0.0;
      for (int i = 0; i < nr; ++i) d += f[i] - w.prev[i];
      if (fabs(d) / sp < p.convergence) break;
    }
  }
  if (status == PW_FAIL_HE_SOLVER) {
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) f[i] = nan_;
  }
  scal[0] = (double)n_relax; scal[1] = (double)nfev; scal[2] = (double)n_solves; scal[3] = 0.0;
  *status_out = status;
}

}  // namespace pw
