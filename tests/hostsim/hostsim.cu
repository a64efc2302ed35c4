// tests/hostsim/hostsim.cu -- TEST INFRASTRUCTURE.
// Host build of the engine's `__host__ __device__` numerical core so that the CPU
// test-suite can step exactly the code the sm_100a kernels run (parity tier P1,
// SURVEY.md 8(c)) against scipy traces.  Never loaded by the product package.
#include <vector>
#include <cstring>
#include "../../p_winds_b200/csrc/pw_common.cuh"
#include "../../p_winds_b200/csrc/pw_parker.cuh"
#include "../../p_winds_b200/csrc/pw_sed.cuh"
#include "../../p_winds_b200/csrc/pw_rk45.cuh"
#include "../../p_winds_b200/csrc/pw_hydrogen.cuh"

using namespace pw;

#ifndef PW_SRC_HASH
#define PW_SRC_HASH "unknown"
#endif
extern "C" const char* hs_version(void) { return "hostsim src:" PW_SRC_HASH; }

extern "C" {

long long hs_h_max_nfev = 0;  // work limit of the hydrogen RK45 solve (0 = none, like scipy); set by tests

void hs_parker(const double* r, int n, double* v, double* rho) {
  for (int i = 0; i < n; ++i) parker_point(r[i], &v[i], &rho[i]);
}

void hs_parker_tidal(const double* r, int n, double cs, double rs, double mp, double ms, double a,
                     double* v, double* rho) {
  TidalCoef c = tidal_coef(cs, rs, mp, ms, a);
  for (int i = 0; i < n; ++i) parker_point_tidal(c, r[i], &v[i], &rho[i]);
}

void hs_parker_scalars(double T, double mu, double mp, double ms, double a, double mdot, double* out) {
  out[0] = sound_speed(T, mu);
  out[1] = radius_sonic_point(mp, out[0]);
  out[2] = radius_sonic_point_tidal(mp, out[0], ms, a);
  out[3] = density_sonic_point(mdot, out[2], out[0]);
}

void hs_sed_setup(const double* wl, const double* flux, int n, double* gw, double* s_h, double* s_he,
                  double* scalars) {
  double red[40];
  sed_setup(wl, flux, n, gw, s_h, s_he, scalars, red);
}

// opts: [r_pl, m_pl, m_star, a_au, initial_f_ion, convergence, max_n_relax, relax, exact_phi,
//        phi_abs, a0, rtol, atol, first_step, max_step, out_tidal]
int hs_h_ion_fraction(const double* model /*T, mdot, h, mu0*/, const double* opts, const double* r, int nr,
                      const double* gw, const double* s_h, const double* s_he, int n_h, double* f_r,
                      double* v, double* rho, double* scal) {
  HParams p;
  p.T = model[0]; p.mdot = model[1]; p.h_fraction = model[2]; p.mu0 = model[3];
  p.r_pl = opts[0]; p.m_pl = opts[1]; p.m_star = opts[2]; p.a_au = opts[3];
  p.initial_f_ion = opts[4]; p.convergence = opts[5]; p.max_n_relax = (int)opts[6];
  p.relax = (int)opts[7]; p.exact_phi = (int)opts[8]; p.phi_abs = opts[9]; p.a0 = opts[10];
  p.ivp.rtol = opts[11]; p.ivp.atol = opts[12]; p.ivp.first_step = opts[13]; p.ivp.max_step = opts[14];
  p.ivp.max_nfev = hs_h_max_nfev;
  p.out_tidal = (int)opts[15];
  SedTables t{n_h, 0, 0, gw, s_h, s_he};
  std::vector<double> buf(10 * nr + 40);
  HWork w;
  double* b = buf.data();
  w.rn = b; w.v = b + nr; w.rho = b + 2 * nr; w.phi = b + 3 * nr; w.tau = b + 4 * nr; w.f = b + 5 * nr;
  w.fprev = b + 6 * nr; w.colh = b + 7 * nr; w.colhe = b + 8 * nr; w.rcm = b + 9 * nr; w.red = b + 10 * nr;
  int status = 0;
  HOut o{f_r, v, rho, scal, &status};
  h_ion_fraction_model(p, t, r, nr, w, o);
  return status;
}

}  // extern "C"

#include "../../p_winds_b200/csrc/pw_helium.cuh"

extern "C" {

// model: [T, h_fraction, vs, rs, rhos]; opts: [r_pl, y0_0, y0_1, relax, max_n_relax, convergence,
//   method, rtol, atol, first_step, max_step]; rates[6]
int hs_he_population(const double* model, const double* opts, const double* rates, const double* r,
                     const double* v, const double* rho, const double* f_h, int nr, double* f1, double* f3,
                     double* scal, double* rates_out) {
  static LsodaTables tab;
  static bool init = false;
  if (!init) { lsoda_tables_init(&tab); init = true; }
  HeParams p;
  p.T = model[0]; p.h_fraction = model[1]; p.vs = model[2]; p.rs = model[3]; p.rhos = model[4];
  p.r_pl = opts[0]; p.y0[0] = opts[1]; p.y0[1] = opts[2]; p.relax = (int)opts[3];
  p.max_n_relax = (int)opts[4]; p.convergence = opts[5]; p.method = (int)opts[6];
  p.rtol = opts[7]; p.atol = opts[8]; p.first_step = opts[9]; p.max_step = opts[10];
  for (int i = 0; i < 6; ++i) p.rates[i] = rates[i];
  std::vector<double> buf((size_t)kHeWorkPerRadius * nr);
  double* b = buf.data();
  HeWork w{b, b + nr, b + 2 * nr, b + 3 * nr, b + 4 * nr, b + 5 * nr, b + 10 * nr};
  int status = 0;
  HeOut o{f1, f3, scal, rates_out, &status};
  if (p.method == PW_HE_RK45) he_population_model<true>(p, tab, r, v, rho, f_h, nr, w, o);
  else he_population_model<false>(p, tab, r, v, rho, f_h, nr, w, o);
  return status;
}

void hs_lsoda_tables(double* elco, double* tesco) {
  LsodaTables t;
  lsoda_tables_init(&t);
  memcpy(elco, t.elco, sizeof(t.elco));
  memcpy(tesco, t.tesco, sizeof(t.tesco));
}

}  // extern "C"

// Pure-arithmetic test problems for trace parity of the LSODA / RK45 replicas
// against scipy (no libm in the RHS, so both sides evaluate identical doubles).
struct VdpRhs {
  double mu;
  PW_HD void at(double) {}
  PW_HD void at_cold(double) {}
  PW_HD void eval_cold(const double* y, double* d) const { eval(y, d); }
  PW_HD void eval(const double* y, double* d) const {
    d[0] = y[1];
    d[1] = mu * ((1.0 - y[0] * y[0]) * y[1]) - y[0];
  }
  PW_HD void operator()(double, const double* y, double* d) { eval(y, d); }
};

#include "../../p_winds_b200/csrc/pw_radau.cuh"

// scalar stiff test problem for the N = 1 Radau path: y' = -lam * (y - cos(t)) - sin(t)
struct CosRhs {
  double lam;
  PW_HD void operator()(double t, const double* y, double* d) { d[0] = -lam * (y[0] - cos(t)) - sin(t); }
};

extern "C" {
// counters: [nfev, njev, nlu, naccept, njac_fev]
int hs_radau_vdp(double mu, const double* y0, const double* t, int n, double rtol, double atol, double* yout,
                 int* counters) {
  VdpRhs rhs{mu};
  Rk45Opts o{rtol, atol, 0., 0.};
  RadauStats st{0, 0, 0, 0, 0};
  const int got = radau_solve<2>(rhs, t, n, y0, o, yout, &st);
  counters[0] = st.nfev; counters[1] = st.njev; counters[2] = st.nlu; counters[3] = st.naccept; counters[4] = st.njac_fev;
  return got;
}
int hs_radau_cos(double lam, double y0, const double* t, int n, double rtol, double atol, double* yout,
                 int* counters) {
  CosRhs rhs{lam};
  Rk45Opts o{rtol, atol, 0., 0.};
  RadauStats st{0, 0, 0, 0, 0};
  const int got = radau_solve<1>(rhs, t, n, &y0, o, yout, &st);
  counters[0] = st.nfev; counters[1] = st.njev; counters[2] = st.nlu; counters[3] = st.naccept; counters[4] = st.njac_fev;
  return got;
}
}

extern "C" {
// stats per output point: [nst, nfe, nje, nqu, mused, hu, tcur, tsw]
int hs_lsoda_vdp(double mu, const double* y0, const double* t, int n, double rtol, double atol, int mxstep,
                 double* yout, double* stats) {
  static LsodaTables tab;
  lsoda_tables_init(&tab);
  VdpRhs rhs{mu};
  LsodaOpts lo{rtol, atol, mxstep, 0., 0.};
  double hi_rows[Lsoda<2, VdpRhs>::kHiDoubles];
  Lsoda<2, VdpRhs> s(rhs, tab, lo, hi_rows);
  yout[0] = y0[0]; yout[1] = y0[1];
  for (int k = 1; k < n; ++k) {
    const int istate = s.call(t[0], y0, t[k], &yout[2 * k]);
    double* st = stats + 8 * k;
    st[0] = s.nst; st[1] = s.nfe; st[2] = s.nje; st[3] = s.nqu; st[4] = s.mused; st[5] = s.hu; st[6] = s.tn; st[7] = s.tsw;
    if (istate < 0) return istate;
  }
  return 2;
}
int hs_rk45_vdp(double mu, const double* y0, const double* t, int n, double rtol, double atol, double* yout,
                int* counters) {
  VdpRhs rhs{mu};
  Rk45Opts o{rtol, atol, 0., 0.};
  Rk45Stats st{0, 0, 0};
  const int got = rk45_solve<2>(rhs, t, n, y0, o, yout, &st);
  counters[0] = st.nfev; counters[1] = st.naccept; counters[2] = st.nreject;
  return got;
}
}

#include "../../p_winds_b200/csrc/pw_oxygen.cuh"

extern "C" {
// model: [T, h_fraction, vs, rs, rhos]; opts: [r_pl, o_fraction, y0, relax, max_n_relax, convergence,
//   method, rtol, atol, first_step, max_step]; rates[4]
int hs_o_ion_fraction(const double* model, const double* opts, const double* rates, const double* r,
                      const double* v, const double* rho, const double* f_h, const double* f_he, int nr,
                      double* f_o, double* scal, double* rates_out) {
  static LsodaTables tab;
  static bool init = false;
  if (!init) { lsoda_tables_init(&tab); init = true; }
  OParams p;
  p.T = model[0]; p.h_fraction = model[1]; p.vs = model[2]; p.rs = model[3]; p.rhos = model[4];
  p.r_pl = opts[0]; p.o_fraction = opts[1]; p.y0 = opts[2]; p.relax = (int)opts[3];
  p.max_n_relax = (int)opts[4]; p.convergence = opts[5]; p.method = (int)opts[6];
  p.rtol = opts[7]; p.atol = opts[8]; p.first_step = opts[9]; p.max_step = opts[10];
  for (int i = 0; i < 4; ++i) p.rates[i] = rates[i];
  std::vector<double> buf(kOWorkPerRadius * nr);
  double* b = buf.data();
  OWork w{b, b + nr, b + 2 * nr, b + 3 * nr, b + 4 * nr, b + 5 * nr, b + 6 * nr};
  int status = 0;
  OOut o{f_o, scal, rates_out, &status};
  o_ion_fraction_model(p, tab, r, v, rho, f_h, f_he, nr, w, o);
  return status;
}
}

#include "../../p_winds_b200/csrc/pw_heion.cuh"

extern "C" {
// model: [T, h_fraction, vs, rs, rhos]; opts: [r_pl, y0, relax, max_n_relax, convergence, method, rtol, atol,
//   first_step, max_step]; rates[3]
int hs_he_ion_fraction(const double* model, const double* opts, const double* rates, const double* r,
                       const double* v, const double* rho, const double* f_h, int nr, double* f, double* scal) {
  static LsodaTables tab;
  static bool init = false;
  if (!init) { lsoda_tables_init(&tab); init = true; }
  HeIonParams p;
  p.T = model[0]; p.h_fraction = model[1]; p.vs = model[2]; p.rs = model[3]; p.rhos = model[4];
  p.r_pl = opts[0]; p.y0 = opts[1]; p.relax = (int)opts[2]; p.max_n_relax = (int)opts[3]; p.convergence = opts[4];
  p.method = (int)opts[5]; p.rtol = opts[6]; p.atol = opts[7]; p.first_step = opts[8]; p.max_step = opts[9];
  for (int i = 0; i < 3; ++i) p.rates[i] = rates[i];
  std::vector<double> buf(kOWorkPerRadius * nr);
  double* b = buf.data();
  OWork w{b, b + nr, b + 2 * nr, b + 3 * nr, b + 4 * nr, b + 5 * nr, b + 6 * nr};
  int status = 0;
  he_ion_fraction_model(p, tab, r, v, rho, f_h, nr, w, f, scal, &status);
  return status;
}
}

#include "../../p_winds_b200/csrc/pw_carbon.cuh"

extern "C" {
// model: [T, h_fraction, vs, rs, rhos]; opts: [r_pl, c_fraction, y0_0, y0_1, relax, max_n_relax,
//   convergence, method, rtol, atol, first_step, max_step]; rates[7]
int hs_c_ion_fraction(const double* model, const double* opts, const double* rates, const double* r,
                      const double* v, const double* rho, const double* f_h, const double* f_he, int nr,
                      double* f2, double* f3, double* scal, double* rates_out) {
  static LsodaTables tab;
  static bool init = false;
  if (!init) { lsoda_tables_init(&tab); init = true; }
  CParams p;
  p.T = model[0]; p.h_fraction = model[1]; p.vs = model[2]; p.rs = model[3]; p.rhos = model[4];
  p.r_pl = opts[0]; p.c_fraction = opts[1]; p.y0[0] = opts[2]; p.y0[1] = opts[3]; p.relax = (int)opts[4];
  p.max_n_relax = (int)opts[5]; p.convergence = opts[6]; p.method = (int)opts[7];
  p.rtol = opts[8]; p.atol = opts[9]; p.first_step = opts[10]; p.max_step = opts[11];
  for (int i = 0; i < 7; ++i) p.rates[i] = rates[i];
  std::vector<double> buf(kCWorkPerRadius * nr);
  double* b = buf.data();
  CWork w{b, b + nr, b + 2 * nr, b + 3 * nr, b + 4 * nr, b + 5 * nr, b + 6 * nr, b + 7 * nr, b + 8 * nr, b + 9 * nr};
  int status = 0;
  COut o{f2, f3, scal, rates_out, &status};
  c_ion_fraction_model(p, tab, r, v, rho, f_h, f_he, nr, w, o);
  return status;
}
}

#include "../../p_winds_b200/csrc/pw_transit.cuh"

extern "C" {

void hs_wofz_re(const double* x, const double* y, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = wofz_re(x[i], y[i]);
}

void hs_voigt(const double* x, int n, double sigma, double gamma, double* out) {
  for (int i = 0; i < n; ++i) out[i] = voigt_profile(x[i], sigma, gamma);
}

// 'average' optical depth table tau[nr][nw] exactly as the kernels form it
// (k_rt_los + k_rt_sum phase B): ncol[i] * Phi(lambda)
void hs_tau_average(const double* r, const double* n, const double* v, const double* tprof, int nr,
                    const double* w0, const double* fosc, const double* aij, int nlines, double mass,
                    double v_bulk, double rv, int nz, int turbulence, double T, const double* wl, int nw,
                    double* tau, double* sums_out) {
  std::vector<double> ncol(nr);
  double sums[4], red[40];
  std::vector<double> row(nz);
  rt_los_average(r, n, v, tprof, nr, nz, ncol.data(), sums, red, row.data());
  RtParams p;
  p.n_lines = nlines;
  for (int k = 0; k < nlines; ++k) { p.w0[k] = w0[k]; p.fosc[k] = fosc[k]; p.aij[k] = aij[k]; }
  p.mass = mass; p.v_bulk = v_bulk; p.rv_planet = rv; p.nz = nz; p.turbulence = turbulence;
  p.t_is_profile = tprof != nullptr;
  const double vw2 = sums[1] / sums[0];
  const double temp = tprof ? sums[2] / sums[0] : T;
  for (int k = 0; k < nw; ++k) {
    const double phi = rt_phi_average(p, wl[k], temp, vw2);
    for (int i = 0; i < nr; ++i) tau[(size_t)i * nw + k] = ncol[i] * phi;
  }
  if (sums_out) { sums_out[0] = sums[0]; sums_out[1] = sums[1]; sums_out[2] = sums[2]; }
}

}  // extern "C"

#include "../../p_winds_b200/csrc/pw_draw.cuh"

extern "C" void hs_exp_neg(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = pw::pw_exp_neg(x[i], pw::kExp2Tab);
}
extern "C" void hs_exp_neg_sum(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = pw::pw_exp_neg_sum(x[i], pw::kExp2Tab64);
}

extern "C" void hs_root_newton(const double* a, const int* k, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = pw::pw_root_newton(a[i], k[i]);
}

extern "C" void hs_pil_ellipse_rows(double x0, double y0, double x1, double y1, int n, int* lo, int* hi) {
  pw::pil_ellipse_rows(x0, y0, x1, y1, n, lo, hi);
}
