// pw_hydrogen.cuh -- H+ fraction profile of one model, evaluated by one CTA
// (reference: p_winds/hydrogen.py:227-573 `ion_fraction`, :22-114
// `radiative_processes_exact`).
//
// Data-parallel pieces (Parker structure over radii, the Nr x N_lambda
// photoionisation integrals) are spread over the CTA; the adaptive RK45 solve and
// the five mu-bar quadratures are inherently serial and run on thread 0 while the
// other warps wait at the barrier.
#pragma once
#include "pw_common.cuh"
#include "pw_parker.cuh"
#include "pw_rk45.cuh"
#include "pw_sed.cuh"

namespace pw {

struct HParams {
  // per model
  double T, mdot, h_fraction, mu0;
  // per batch
  double r_pl, m_pl, m_star, a_au;
  double initial_f_ion, convergence;
  int max_n_relax, relax, exact_phi;
  double phi_abs, a0;   // used when !exact_phi: hydrogen.radiative_processes[_mono] scalars
  Rk45Opts ivp;
  // the Parker structure handed to the helium stage uses parker.structure
  // (plain) or parker.structure_tidal (docs/source/tidal_effects.ipynb)
  int out_tidal;
};

// Shared-memory work arrays of one CTA, Nr doubles each.
struct HWork {
  double *rn, *v, *rho, *phi, *tau, *f, *fprev, *colh, *colhe, *rcm;
  double* red;  // 40 doubles: block reduction scratch + broadcast slots
};

struct HOut {
  double* f_r;     // [Nr]
  double* v;       // [Nr] final structure for the helium stage (sonic units)
  double* rho;     // [Nr]
  double* scal;    // [8]: mu_bar, vs, rs, rhos, n_relax, nfev_total, n_solves, -
  int* status;
};

#if defined(__CUDA_ARCH__)
#define PW_LANE ((int)(threadIdx.x & 31))
#define PW_WARP ((int)(threadIdx.x >> 5))
#define PW_NWARP ((int)(blockDim.x >> 5))
#define PW_NLANE 32
#else
#define PW_LANE 0
#define PW_WARP 0
#define PW_NWARP 1
#define PW_NLANE 1
#endif

// hydrogen.radiative_processes_exact for all radii; f_h == nullptr means the
// scalar f_h_r = 0.0 of the first pass (hydrogen.py:359-363).
PW_HD void phi_exact(const SedTables& t, const HWork& w, int nr, const double* rho_abs_scale,
                     double rhos, const double* f_h, double h_fraction, unsigned e2tab = 0) {
  const double he_to_h = (1 - h_fraction) / h_fraction;
  // number densities (hydrogen.py:84-94); staged in tau/phi as scratch
  double* n_h = w.tau;
  double* n_he = w.phi;
  for (int i = PW_TID; i < nr; i += PW_NT) {
    const double f = f_h ? f_h[i] : 0.0;
    const double rho = rho_abs_scale[i] * rhos;
    const double mu = (1 + 4 * he_to_h) / (1 + f + he_to_h);
    const double n_tot = rho / mu / kMH;
    const double n_htot = 1 / (1 + f + he_to_h) * n_tot;
    n_h[i] = n_htot * (1 - f);
    n_he[i] = (n_htot * he_to_h) * (1 - f);
  }
  PW_SYNC();
  // column densities: cumulative_trapezoid on the reversed arrays, sign flipped
  // (hydrogen.py:96-106); sequential like numpy.cumsum
  if (PW_TID == 0) {
    double ch = 0.0, che = 0.0;
    w.colh[nr - 1] = 0.0;
    w.colhe[nr - 1] = 0.0;
    for (int i = nr - 2; i >= 0; --i) {
      const double d = w.rcm[i] - w.rcm[i + 1];
      ch += d * (n_h[i] + n_h[i + 1]) / 2.0;
      che += d * (n_he[i] + n_he[i + 1]) / 2.0;
      w.colh[i] = -ch;
      w.colhe[i] = -che;
    }
  }
  PW_SYNC();
  // phi'(r_i) = | sum_j w_j F_j sigma_j / E_j exp(-tau_ij) |  (hydrogen.py:111-112):
  // one warp per radius, lanes over wavelength bins, shuffle-tree reduction
  for (int i = PW_WARP; i < nr; i += PW_NWARP) {
    const double ch = w.colh[i], che = w.colhe[i];
    double acc = 0.0;
    for (int j = PW_LANE; j < t.n_h; j += PW_NLANE) {
      const double tau = ch * t.s_h[j] + che * t.s_he[j];
#if defined(__CUDA_ARCH__)
      acc += t.gw[j] * pw_exp_neg_sh(-tau, e2tab);
#else
      acc += t.gw[j] * pw_exp_neg(-tau, kExp2Tab);
#endif
    }
    acc = warp_sum(acc);
    if (PW_LANE == 0) w.phi[i] = fabs(acc);
  }
  PW_SYNC();
}

struct HRhs {
  const double *rn, *phi, *v, *rho, *tau;
  int n, exact, hint;
  double k2, phi0;
  // np.interp slopes (fp[j+1] - fp[j]) / (xp[j+1] - xp[j]) of phi (exact) or tau, of v and of rho,
  // formed once per pass by h_pass_fields exactly as numpy forms them at every call: three of the
  // five divisions of a right-hand side gone, bit-identical.  Null: evaluate np_interp_at literally.
  const double *sp = nullptr, *sv = nullptr, *srho = nullptr;
  // hydrogen.py:431-448
  PW_HD void operator()(double x, const double* y, double* dydx) {
    const int j = bracket(rn, n, x, hint);
    hint = j;
    double pi_, v_, rho_;   // interpolated phi (or tau), v, rho
    const double* pa = exact ? phi : tau;
    bool done = false;
    if (sp && j >= 0 && j < n - 1) {
      const double x0 = rn[j];
      if (x0 != x) {
        const double d = x - x0;
        pi_ = sp[j] * d + pa[j];
        v_ = sv[j] * d + v[j];
        rho_ = srho[j] * d + rho[j];
        done = (pi_ == pi_) && (v_ == v_) && (rho_ == rho_);  // a NaN takes numpy's fall-back path below
      }
    }
    if (!done) {
      pi_ = np_interp_at(rn, pa, n, x, j);
      v_ = np_interp_at(rn, v, n, x, j);
      rho_ = np_interp_at(rn, rho, n, x, j);
    }
    const double pp = exact ? pi_ : exp(-pi_) * phi0;
    const double f = y[0];
    const double term1 = PW_DIV(1. - f, v_) * pp;
    const double term2 = PW_DIV(k2 * rho_ * (f * f), v_);
    dydx[0] = term1 - term2;
  }
};

// ---------------------------------------------------------------------------------------
// hydrogen.ion_fraction (hydrogen.py:227-573) cut into the pieces the engine schedules:
//   h_state_init   before the first pass
//   h_pass_fields  block-cooperative: sonic-point normalisation, Parker structure over the
//                  radii, exact photoionisation integrals (Nr x N_lambda exps) or the
//                  rectangle-rule optical depth                       (:387-425, :480-501)
//   h_pass_solve   one thread: solve_ivp RK45, mu-bar, convergence test (:451-476, :504-538)
//   h_finish       block-cooperative: outputs + the structure handed to the helium stage
// A batch runs them as separate kernels -- fields: one CTA per model, every lane busy;
// solve: one *thread* per model -- instead of parking 127 lanes of a CTA at a barrier while
// lane 0 integrates.  h_ion_fraction_model strings them together for one model (host
// build of the tests; same code, same arithmetic).
// ---------------------------------------------------------------------------------------
struct HState {
  double mu;       // mean molecular weight the next pass normalises with
  double mu_pass;  // ... and the one the last executed pass was normalised with (return_rates)
  double mu_bar;   // of the last solution
  double vs_amw;   // `vs` average_molecular_weight is called with (hydrogen.py:351, rebound at :480 if exact)
  int n_relax, nfev_total, n_solves, status;
  int done;        // no further passes: converged, single pass, or failed
};

struct HPassScal {
  double vs, rs, rhos, phi_unit, k1, k2;
};

PW_HD void h_state_init(const HParams& p, HState* s) {
  s->mu = p.mu0;
  s->mu_pass = p.mu0;
  s->mu_bar = 0.0;
  s->vs_amw = sound_speed(p.T, p.mu0);
  s->n_relax = 0; s->nfev_total = 0; s->n_solves = 0; s->status = PW_OK; s->done = 0;
}

// _normalize(phi_abs, k1_abs, k2_abs, r, mu)  (hydrogen.py:387-412)
PW_HD HPassScal h_pass_scalars(const HParams& p, double mu) {
  const double alpha_rec = 2.59E-13 * pw_pow(p.T / 1E4, -0.7);  // hydrogen.py:207-223
  const double he_fraction = 1 - p.h_fraction;
  const double a_0 = p.exact_phi ? 0. : p.a0;
  const double k1_abs = p.h_fraction * a_0 / (p.h_fraction + 4 * he_fraction) / kMH;
  const double k2_abs = p.h_fraction / (p.h_fraction + 4 * he_fraction) * alpha_rec / kMH;
  HPassScal q;
  q.vs = sound_speed(p.T, mu);
  q.rs = radius_sonic_point_tidal(p.m_pl, q.vs, p.m_star, p.a_au);
  q.rhos = density_sonic_point(p.mdot, q.rs, q.vs);
  q.phi_unit = q.vs * 1E5 / q.rs / 7.1492E+09;
  const double k1_unit = 1 / (q.rhos * q.rs * 7.1492E+09);
  const double k2_unit = q.vs * 1E5 / q.rs / 7.1492E+09 / q.rhos;
  q.k1 = k1_abs / k1_unit;
  q.k2 = k2_abs / k2_unit;
  return q;
}

// All threads of the CTA.  Reads w.f (solution of the previous pass) when pass > 0; leaves
// rn, v, rho and phi (exact) or tau (approximate) in w.
// `e2tab`: 32-bit shared-memory address of the CTA's copy of kExp2Tab (device build).
PW_HD void h_pass_fields(const HParams& p, const SedTables& sed, const double* __restrict__ r, int nr,
                         const HWork& w, int pass, double mu, unsigned e2tab = 0) {
  for (int i = PW_TID; i < nr; i += PW_NT)
    w.rcm[i] = r[i] * p.r_pl * (7.1492e7 / 1e-2);  // (r * R_pl * u.Rjup).to(u.cm)
  const HPassScal q = h_pass_scalars(p, mu);
  const TidalCoef tc = tidal_coef(q.vs, q.rs, p.m_pl, p.m_star, p.a_au);
  for (int i = PW_TID; i < nr; i += PW_NT) {
    const double x = r[i] * p.r_pl / q.rs;
    w.rn[i] = x;
    parker_point_tidal(tc, x, &w.v[i], &w.rho[i]);
  }
  PW_SYNC();
  if (p.exact_phi) {
    // hydrogen.py:355-363 (pass 0, f = 0) and :480-490 (relax passes, f = f_r)
    phi_exact(sed, w, nr, w.rho, q.rhos, pass == 0 ? nullptr : w.f, p.h_fraction, e2tab);
    for (int i = PW_TID; i < nr; i += PW_NT) w.phi[i] = w.phi[i] / q.phi_unit;
  } else if (PW_TID == 0) {
    // rectangle-rule column density, local cell included (hydrogen.py:424-425, :499-501)
    double c = 0.0;
    for (int i = nr - 1; i >= 0; --i) {
      const double dr = (i < nr - 1) ? (w.rn[i + 1] - w.rn[i]) : (w.rn[nr - 1] - w.rn[nr - 2]);
      const double wgt = (pass == 0) ? dr * w.rho[i] : dr * w.rho[i] * (1 - w.f[i]);
      c += wgt;
      w.tau[i] = q.k1 * c;
    }
  }
  PW_SYNC();
  // np.interp slopes for the right-hand side of this pass (HRhs), in the scratch the column
  // densities no longer need: colh <- phi or tau, colhe <- v, rcm <- rho
  {
    const double* pa = p.exact_phi ? w.phi : w.tau;
    for (int i = PW_TID; i < nr; i += PW_NT) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      if (i < nr - 1) {
        const double dx = w.rn[i + 1] - w.rn[i];
        s0 = (pa[i + 1] - pa[i]) / dx;
        s1 = (w.v[i + 1] - w.v[i]) / dx;
        s2 = (w.rho[i + 1] - w.rho[i]) / dx;
      }
      w.colh[i] = s0; w.colhe[i] = s1; w.rcm[i] = s2;
    }
  }
  PW_SYNC();
}

// One thread.  solve_ivp(_fun, (r[0], r[-1]), [initial_f_ion], t_eval=r) (:451, :504) on the
// fields of this pass, then mu-bar and the relax-loop bookkeeping (:462-476, :515-538).
PW_HD void h_pass_solve(const HParams& p, const double* __restrict__ r, int nr, const HWork& w, int pass,
                        HState* s) {
  const HPassScal q = h_pass_scalars(p, s->mu);
  s->mu_pass = s->mu;
  if (p.exact_phi) s->vs_amw = q.vs;
  const double he_h = (1 - p.h_fraction) / p.h_fraction;
  for (int i = 0; i < nr; ++i) w.fprev[i] = w.f[i];
  HRhs rhs{w.rn, w.phi, w.v, w.rho, w.tau, nr, p.exact_phi, 0, q.k2, p.phi_abs / q.phi_unit};
  rhs.sp = w.colh; rhs.sv = w.colhe; rhs.srho = w.rcm;   // slopes left by h_pass_fields (may be null)
  Rk45Stats st{0, 0, 0};
  const double y0 = p.initial_f_ion;
  const int got = rk45_solve<1>(rhs, w.rn, nr, &y0, p.ivp, w.f, &st);
  s->nfev_total += st.nfev;
  ++s->n_solves;
  if (pass > 0) s->n_relax = pass;
  if (got != nr) {
    // len(f_r) != len(r) -> RuntimeError (:458-460, :511-513); or the engine's own work limit
    s->status = st.limit ? PW_FAIL_WORK_LIMIT : PW_FAIL_H_SOLVER;
    s->done = 1;
    return;
  }
  s->mu_bar = average_molecular_weight(w.f, r, p.r_pl, w.v, s->vs_amw, nr, p.m_pl, p.T, he_h);
  if (pass > 0) {
    double sd = 0.0, sp = 0.0;
    for (int i = 0; i < nr; ++i) { sd += w.f[i] - w.fprev[i]; sp += w.fprev[i]; }
    if (fabs(sd / sp) < p.convergence) s->done = 1;
  }
  if (!p.relax || pass >= p.max_n_relax) s->done = 1;
  s->mu = s->mu_bar;
}

// All threads of the CTA: outputs + the structure the helium stage needs (quickstart.ipynb cell 7)
PW_HD void h_finish(const HParams& p, const double* __restrict__ r, int nr, const double* __restrict__ f,
                    const HState& s, const HOut& out) {
  const double mu_bar = s.mu_bar;
  int status = s.status;
  const double vs_f = sound_speed(p.T, mu_bar);
  double rs_f, rhos_f;
  if (status == PW_OK && !(mu_bar == mu_bar)) status = PW_FAIL_NAN;
  if (p.out_tidal) rs_f = radius_sonic_point_tidal(p.m_pl, vs_f, p.m_star, p.a_au);
  else rs_f = radius_sonic_point(p.m_pl, vs_f);
  rhos_f = density_sonic_point(p.mdot, rs_f, vs_f);
  const TidalCoef tcf = tidal_coef(vs_f, rs_f, p.m_pl, p.m_star, p.a_au);
  const double nan_ = nan("");
  for (int i = PW_TID; i < nr; i += PW_NT) {
    if (status == PW_OK) {
      const double x = r[i] * p.r_pl / rs_f;
      double vv, rr;
      if (p.out_tidal) parker_point_tidal(tcf, x, &vv, &rr);
      else parker_point(x, &vv, &rr);
      out.f_r[i] = f[i];
      out.v[i] = vv;
      out.rho[i] = rr;
    } else {
      out.f_r[i] = nan_; out.v[i] = nan_; out.rho[i] = nan_;
    }
  }
  if (PW_TID == 0) {
    out.scal[0] = (status == PW_OK) ? mu_bar : nan_;
    out.scal[1] = vs_f; out.scal[2] = rs_f; out.scal[3] = rhos_f;
    out.scal[4] = (double)s.n_relax; out.scal[5] = (double)s.nfev_total; out.scal[6] = (double)s.n_solves;
    out.scal[7] = 0.0;
    *out.status = status;
  }
}

// return_rates=True of hydrogen.ion_fraction (hydrogen.py:538-561), all threads of the CTA:
//   photoionization = radiative_processes_exact(r, rho_norm * rhos, f_r) * (1 - f_r) with the
//                     density structure of the last normalisation (w.rho + s.mu_pass),
//   recombination   = k2_abs * final_rho * f_r^2 with the structure of the final mu_bar.
// w.rho must hold the last pass's rho_norm and w.f the final f_r; rates = [2][Nr].
PW_HD void h_rates(const HParams& p, const SedTables& sed, const double* __restrict__ r, int nr, const HWork& w,
                   const HState& s, double* rates, unsigned e2tab = 0) {
  for (int i = PW_TID; i < nr; i += PW_NT) w.rcm[i] = r[i] * p.r_pl * (7.1492e7 / 1e-2);
  const HPassScal q = h_pass_scalars(p, s.mu_pass);
  PW_SYNC();
  phi_exact(sed, w, nr, w.rho, q.rhos, w.f, p.h_fraction, e2tab);
  const double alpha_rec = 2.59E-13 * pw_pow(p.T / 1E4, -0.7);
  const double he_fraction = 1 - p.h_fraction;
  const double k2_abs = p.h_fraction / (p.h_fraction + 4 * he_fraction) * alpha_rec / kMH;
  const double vs_f = sound_speed(p.T, s.mu_bar);
  const double rs_f = radius_sonic_point_tidal(p.m_pl, vs_f, p.m_star, p.a_au);
  const double rhos_f = density_sonic_point(p.mdot, rs_f, vs_f);
  const TidalCoef tcf = tidal_coef(vs_f, rs_f, p.m_pl, p.m_star, p.a_au);
  for (int i = PW_TID; i < nr; i += PW_NT) {
    double vv, rr;
    parker_point_tidal(tcf, r[i] * p.r_pl / rs_f, &vv, &rr);
    const double f = w.f[i];
    rates[i] = w.phi[i] * (1 - f);
    rates[nr + i] = k2_abs * (rr * rhos_f) * (f * f);
  }
}

// One model, one CTA (or one host thread).  `r` is the radius profile in planetary radii.
PW_HD void h_ion_fraction_model(const HParams& p, const SedTables& sed, const double* __restrict__ r,
                                int nr, const HWork& w, const HOut& out) {
  HState s;
  h_state_init(p, &s);
  const int n_pass = p.relax ? p.max_n_relax + 1 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    h_pass_fields(p, sed, r, nr, w, pass, s.mu);
    if (PW_TID == 0) {
      h_pass_solve(p, r, nr, w, pass, &s);
      w.red[34] = s.mu; w.red[35] = s.mu_bar; w.red[36] = s.vs_amw;
      w.red[37] = (double)s.n_relax; w.red[38] = (double)s.nfev_total; w.red[39] = (double)s.n_solves;
      w.red[33] = (double)(2 * s.status + s.done);
    }
    PW_SYNC();
    s.mu = w.red[34]; s.mu_bar = w.red[35]; s.vs_amw = w.red[36];
    s.n_relax = (int)w.red[37]; s.nfev_total = (int)w.red[38]; s.n_solves = (int)w.red[39];
    s.status = (int)w.red[33] >> 1; s.done = (int)w.red[33] & 1;
    PW_SYNC();
    if (s.done) break;
  }
  h_finish(p, r, nr, w.f, s, out);
}

}  // namespace pw
