// pw_carbon.cuh -- C II / C III fraction profiles of one model (reference:
// p_winds/carbon.py:256-639 `ion_fraction`, :135-252 rate coefficients).  Two equations,
// integrated with the LSODA replica (method='odeint', the reference default), the Radau replica
// (what docs/source/exospheric_metals.ipynb uses) or the RK45 replica.  One thread per model.
// SURVEY.md 8(f-3).
#pragma once
#include "pw_common.cuh"
#include "pw_lsoda.cuh"
#include "pw_radau.cuh"
#include "pw_rk45.cuh"

namespace pw {

enum { PW_C_ODEINT = 0, PW_C_RADAU = 1, PW_C_RK45 = 2 };

struct CParams {
  double T, h_fraction, vs, rs, rhos;  // per model
  double r_pl, c_fraction, y0[2];
  int relax, max_n_relax;
  double convergence;
  double rates[7];  // phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he (cgs; carbon.radiative_processes)
  int method;
  double rtol, atol, first_step, max_step;
};

// Right-hand side carbon.py:452-529; rate coefficients in product form (see pw_helium.cuh),
// literal exp(log + log + log) with -DPW_HE_LOGEXP.
struct CRhs {
  const double *rn, *v, *rho, *fh, *fhe, *tau1, *tau2;
  int n, hint;
  double k1, k2, ion1, ion2, rec1, rec2, ct_ci_hp, ct_ci_hep, ct_cii_h, ct_ciii_h, ct_ciii_he, phi1, phi2;
  // t[11] = term11, 12, 13, 14, 15, 16, 17, 18, 19, 21, 22
  PW_HD void terms(double x, double f2, double f3, double* t, double* vv) {
    const int j = bracket(rn, n, x, hint);
    hint = j;
    const double v_ = np_interp_at(rn, v, n, x, j);
    const double rho_ = np_interp_at(rn, rho, n, x, j);
    const double fh_ = np_interp_at(rn, fh, n, x, j);
    const double fhe_ = np_interp_at(rn, fhe, n, x, j);
    const double t_ci = np_interp_at(rn, tau1, n, x, j);
    const double t_cii = np_interp_at(rn, tau2, n, x, j);
#if defined(PW_HE_LOGEXP)
    const double lt1 = log(k1) + log(rho_), lt2 = log(k2) + log(rho_);
#define PW_C_H(coef) exp(lt1 + log(coef))
#define PW_C_HE(coef) exp(lt2 + log(coef))
#else
#define PW_C_H(coef) ((k1 * (coef)) * rho_)
#define PW_C_HE(coef) ((k2 * (coef)) * rho_)
#endif
    const double ion1_ne = PW_C_H(ion1) * fh_ + PW_C_HE(ion1) * fhe_;
    const double cthp = PW_C_H(ct_ci_hp) * fh_;
    const double cthep = PW_C_HE(ct_ci_hep) * fhe_;
    const double rec1_ne = PW_C_H(rec1) * fh_ + PW_C_HE(rec1) * fhe_;
    const double ct2h = PW_C_H(ct_cii_h) * (1 - fh_);
    const double rec2_ne = PW_C_H(rec2) * fh_ + PW_C_HE(rec2) * fhe_;
    const double ct3h = PW_C_H(ct_ciii_h) * (1 - fh_);
    const double ct3he = PW_C_HE(ct_ciii_he) * (1 - fhe_);
    const double ion2_ne = PW_C_H(ion2) * fh_ + PW_C_HE(ion2) * fhe_;
#undef PW_C_H
#undef PW_C_HE
    const double f1 = 1 - f2 - f3;
    t[0] = f1 * phi1 * exp(-t_ci);
    t[1] = f1 * ion1_ne;
    t[2] = f1 * cthp;
    t[3] = f1 * cthep;
    t[4] = f2 * rec1_ne;
    t[5] = f2 * ct2h;
    t[6] = f3 * rec2_ne;
    t[7] = f3 * ct3h;
    t[8] = f3 * ct3he;
    t[9] = f2 * phi2 * exp(-t_cii);
    t[10] = f2 * ion2_ne;
    *vv = v_;
  }
  PW_HD void rhs_at(double x, const double* y, double* d) {
    double t[11], vv;
    terms(x, y[0], y[1], t, &vv);
    d[0] = (t[0] + t[1] + t[2] + t[3] - t[4] - t[5] + t[6] + t[7] + t[8] - t[9] - t[10]) / vv;
    d[1] = (t[9] + t[10] - t[6] - t[7] - t[8]) / vv;
  }
  PW_HD void operator()(double x, const double* y, double* d) { rhs_at(x, y, d); }
  // RHS concept of the LSODA replica
  double xs;
  PW_HD void at(double x) { xs = x; }
  PW_HD void at_cold(double x) { xs = x; }
  PW_HD void eval(const double* y, double* d) { rhs_at(xs, y, d); }
  PW_HD void eval_cold(const double* y, double* d) { rhs_at(xs, y, d); }
};

constexpr int kCWorkPerRadius = 11;  // the nine arrays below + 2*Nr of interleaved solver output
struct CWork {
  double *rn, *w, *tau_ci_h, *tau_cii_h, *tau_he, *tau1, *tau2, *prev2, *prev3;
  double* ytmp;  // 2*Nr
};
struct COut {
  double *f2, *f3;  // [Nr] each
  double* scal;     // [4]: n_relax, nfev total, n_solves, -
  double* rates;    // [11][Nr] or null, in the order of carbon.py:531-533
  int* status;
};

PW_HD bool c_solve(const CParams& p, CRhs& rhs, const LsodaTables& tab, int nr, double* f2, double* f3, double* ytmp,
                   int* nfev) {
  const double* rn = rhs.rn;
  rhs.hint = 0;
  const Rk45Opts o{p.rtol, p.atol, p.first_step, p.max_step};
  if (p.method == PW_C_RADAU || p.method == PW_C_RK45) {
    int got;
    if (p.method == PW_C_RADAU) {
      RadauStats st{0, 0, 0, 0, 0};
      got = radau_solve<2>(rhs, rn, nr, p.y0, o, ytmp, &st);
      *nfev += st.nfev;
    } else {
      Rk45Stats st{0, 0, 0};
      got = rk45_solve<2>(rhs, rn, nr, p.y0, o, ytmp, &st);
      *nfev += st.nfev;
    }
    for (int k = 0; k < got; ++k) { f2[k] = ytmp[2 * k]; f3[k] = ytmp[2 * k + 1]; }
    return got == nr;
  }
  // odeint (carbon.py:535-546): LSODA at rtol = atol = 1.49012e-8, mxstep = 500 per output radius
  LsodaOpts lo{1.49012e-8, 1.49012e-8, 500, 0., 0.};
  double hi_rows[Lsoda<2, CRhs>::kHiDoubles];
  Lsoda<2, CRhs> s(rhs, tab, lo, hi_rows);
  f2[0] = p.y0[0];
  f3[0] = p.y0[1];
  bool ok = true;
  int k = 1;
  for (; k < nr; ++k) {
    double yo[2];
    if (s.call(rn[0], p.y0, rn[k], yo) < 0) { ok = false; break; }
    f2[k] = yo[0];
    f3[k] = yo[1];
  }
  *nfev += s.nfe;
  return ok;
}

PW_HD void c_ion_fraction_model(const CParams& p, const LsodaTables& tab, const double* __restrict__ r,
                                const double* __restrict__ v, const double* __restrict__ rho,
                                const double* __restrict__ f_h, const double* __restrict__ f_he, int nr,
                                const CWork& w, const COut& out) {
  const double rs_cm = p.rs * 7.1492E+09;
  const double alpha_unit = (rs_cm * rs_cm) * p.vs * 1E5;                                    // carbon.py:393
  const double rec1 = 4.67E-12 * pw_pow(300 / p.T, 0.60) / alpha_unit;                       // :191
  const double rec2 = 2.3E-12 * pw_pow(1000 / p.T, 0.645) / alpha_unit;                      // :192
  const double m_h = 1.67262192E-24 / (p.rhos * (rs_cm * rs_cm * rs_cm));                    // :399-400
  const double phi_unit = p.vs * 1E5 / p.rs / 7.1492E+09;                                    // :405
  const double phi1 = p.rates[0] / phi_unit, phi2 = p.rates[1] / phi_unit;
  const double a_ci = p.rates[2] / (rs_cm * rs_cm), a_cii = p.rates[3] / (rs_cm * rs_cm);
  const double a_h_ci = p.rates[4] / (rs_cm * rs_cm), a_h_cii = p.rates[5] / (rs_cm * rs_cm);
  const double a_he = p.rates[6] / (rs_cm * rs_cm);
  const double e_el = 8.617333262145179e-05 * p.T;                                            // :157-164
  const double r1 = 11.3 / e_el, r2 = 24.4 / e_el;
  const double ion1 = 6.85E-8 * pw_pow(0.193 + r1, -1.0) * pw_pow(r1, 0.25) * exp(-r1) / alpha_unit;
  const double ion2 = 1.86E-8 * pw_pow(0.286 + r2, -1.0) * pw_pow(r2, 0.24) * exp(-r2) / alpha_unit;
  const double ct_cii_h = 6.30E-17 * pw_pow(300 / p.T, -1.96) * exp(-1.7E5 / p.T) / alpha_unit;      // :237-238
  const double ct_ci_hp = 1.31E-15 * pw_pow(300 / p.T, -0.213) / alpha_unit;                          // :242
  const double ct_ci_hep = 2.5E-15 * pw_pow(300 / p.T, -1.597) / alpha_unit;                          // :243
  const double ct_ciii_h = 1.67E-4 * pw_pow(p.T / 10000, 2.79) * (1 + 304.72 * exp(-4.07 * p.T / 10000)) / alpha_unit;
  const double ct_ciii_he = 1.23E-9 / alpha_unit;                                                     // :248
  const double he_fraction = 1 - p.h_fraction;
  const double den = p.h_fraction + 4 * he_fraction + 12 * p.c_fraction;
  const double k1 = p.h_fraction / den / m_h, k2 = he_fraction / den / m_h, k3 = p.c_fraction / den / m_h;

  for (int i = 0; i < nr; ++i) w.rn[i] = r[i] * p.r_pl / p.rs;
  {
    double c = 0.0, ch = 0.0, che = 0.0;
    for (int i = nr - 1; i >= 0; --i) {
      const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
      w.w[i] = dr * rho[i];
      c += w.w[i];
      ch += w.w[i] * (1 - f_h[i]);
      che += w.w[i] * he_fraction * (1 - f_he[i]);
      w.tau_ci_h[i] = k1 * a_h_ci * ch;
      w.tau_cii_h[i] = k1 * a_h_cii * ch;
      w.tau_he[i] = k2 * a_he * che;
      w.tau1[i] = (1 - p.y0[0] - p.y0[1]) * k3 * a_ci * c + w.tau_ci_h[i] + w.tau_he[i];
      w.tau2[i] = p.y0[0] * k3 * a_cii * c + w.tau_cii_h[i] + w.tau_he[i];
    }
  }
  CRhs rhs;
  rhs.rn = w.rn; rhs.v = v; rhs.rho = rho; rhs.fh = f_h; rhs.fhe = f_he; rhs.tau1 = w.tau1; rhs.tau2 = w.tau2;
  rhs.n = nr; rhs.hint = 0; rhs.xs = 0.0;
  rhs.k1 = k1; rhs.k2 = k2; rhs.ion1 = ion1; rhs.ion2 = ion2; rhs.rec1 = rec1; rhs.rec2 = rec2;
  rhs.ct_ci_hp = ct_ci_hp; rhs.ct_ci_hep = ct_ci_hep; rhs.ct_cii_h = ct_cii_h; rhs.ct_ciii_h = ct_ciii_h;
  rhs.ct_ciii_he = ct_ciii_he; rhs.phi1 = phi1; rhs.phi2 = phi2;

  int nfev = 0, n_relax = 0, n_solves = 0, status = PW_OK;
  double* f2 = out.f2;
  double* f3 = out.f3;
  auto clamp = [&]() {  // carbon.py:563-566
    for (int i = 0; i < nr; ++i) {
      if (f2[i] < 0) f2[i] = 1E-15;
      if (f3[i] < 0) f3[i] = 1E-15;
      if (f2[i] > 1.0) f2[i] = 1.0;
      if (f3[i] > 1.0) f3[i] = 1.0;
    }
  };
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  for (int it = 0; it < n_pass; ++it) {
    double s2 = 0.0, s3 = 0.0;
    if (it > 0) {
      ++n_relax;
      double c1 = 0.0, c2 = 0.0;
      for (int i = nr - 1; i >= 0; --i) {  // :576-582
        c1 += w.w[i] * (1 - f2[i] - f3[i]);
        c2 += w.w[i] * f2[i];
        w.tau1[i] = k3 * a_ci * c1 + w.tau_ci_h[i] + w.tau_he[i];
        w.tau2[i] = k3 * a_cii * c2 + w.tau_cii_h[i] + w.tau_he[i];
      }
      for (int i = 0; i < nr; ++i) { s2 += f2[i]; s3 += f3[i]; w.prev2[i] = f2[i]; w.prev3[i] = f3[i]; }
    }
    ++n_solves;
    const bool ok = c_solve(p, rhs, tab, nr, f2, f3, w.ytmp, &nfev);
    if (!ok && (it == 0 || p.method != PW_C_ODEINT)) {
      status = PW_FAIL_HE_SOLVER;  // RuntimeError of carbon.py:541-557
      break;
    }
    if (!ok) {   // carbon.py:587
      // An `odeint` of the relax loop warned.  The reference does not check it and carries on with rows scipy
      // left uninitialised (helium.py:736 has the same pattern; pw_helium.cuh): undefined from here on.  Stop,
      // hand back the last completed pass, status PW_WARN_HE_RELAX.
      status = PW_WARN_HE_RELAX;
      for (int i = 0; i < nr; ++i) { f2[i] = w.prev2[i]; f3[i] = w.prev3[i]; }
      break;
    }
    clamp();
    if (it > 0) {
      double d2 = 0.0, d3 = 0.0;
      for (int i = 0; i < nr; ++i) { d2 += f2[i] - w.prev2[i]; d3 += f3[i] - w.prev3[i]; }
      if (fabs(d2) / s2 < p.convergence && fabs(d3) / s3 < p.convergence) break;  // :607-616
    }
  }
  if (status == PW_FAIL_HE_SOLVER) {
    const double nan_ = nan("");
    for (int i = 0; i < nr; ++i) { f2[i] = nan_; f3[i] = nan_; }
  } else if (out.rates) {
    // _fun(r, [f_cii_r, f_ciii_r], rates=True): term11, 21, 15, 17, 12, 22, 13, 14, 16, 18, 19
    const int order[11] = {0, 9, 4, 6, 1, 10, 2, 3, 5, 7, 8};
    rhs.hint = 0;
    for (int i = 0; i < nr; ++i) {
      double t[11], vv;
      rhs.terms(w.rn[i], f2[i], f3[i], t, &vv);
      for (int k = 0; k < 11; ++k) out.rates[k * nr + i] = t[order[k]] * phi_unit;
    }
  }
  out.scal[0] = (double)n_relax; out.scal[1] = (double)nfev; out.scal[2] = (double)n_solves; out.scal[3] = 0.0;
  *out.status = status;
}

}  // namespace pw
