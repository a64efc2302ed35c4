// pw_sed.cuh -- per-context SED tables and flux-averaged scalars
// (reference: microphysics.py:27-220, hydrogen.py:118-169, helium.py:24-154,
// tools.py:27-48).  Model independent, so it runs once per context (SURVEY.md
// 8(a) a8-a10, a12, a14) -- but still on the device: one CTA builds the
// cross-section tables and does the six Simpson integrals as block reductions.
#pragma once
#include "pw_common.cuh"

namespace pw {

// (c.h * c.c).to(erg * angstrom).value  (hydrogen.py:61)
#define PW_HC_ERG_A ((6.62607015e-34 * 299792458.0) * (1.0 / (1e-7 * 1e-10)))

// tools.nearest_index (tools.py:27-48)
PW_HD int nearest_index(const double* a, int n, double target) {
  int lo = 0, hi = n;  // searchsorted side='left'
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < target) lo = mid + 1; else hi = mid;
  }
  int i = lo;
  if (i < 1) i = 1;
  if (i > n - 1) i = n - 1;
  const double left = a[i - 1], right = a[i];
  if (target - left < right - target) i -= 1;
  return i;
}

// microphysics.hydrogen_cross_section (microphysics.py:53-62), NaN -> 0
PW_HD double sigma_h(double wl) {
  const double eps = sqrt(911.65 / wl - 1);
  const double q = wl / 911.65;
  const double a = 6.3E-18 * exp(4 - (4 * atan(eps)) / eps) / (1 - exp(-2 * kPi / eps)) *
                   ((q * q) * (q * q));
  return (a != a) ? 0.0 : a;
}

// microphysics.helium_total_cross_section (microphysics.py:78-110), Yan+ 1998
PW_HD double sigma_he_total(double wl) {
  const double hz = 299792458.0 / (wl * 1e-10);
  const double e_ev = hz * 6.62607015e-34 / 1.602176634e-19;
  if (e_ev < 24.58) return 0.0;
  const double x = e_ev / 24.58;
  const double e_term = pow(e_ev / 1.0 / 1000., 3.5);
  const double sx = sqrt(x);
  // sum_i coef_i / x**((i+1)/2), i = 0..5
  double s = 0.;
  s = s + -4.7416 / sx;
  s = s + 14.8200 / x;
  s = s + -30.8678 / (x * sx);
  s = s + 37.3584 / (x * x);
  s = s + -23.4585 / (x * x * sx);
  s = s + 5.9133 / (x * x * x);
  return (733 * (1e-28 / (1e-2 * 1e-2))) / e_term * (1 + s);
}

// microphysics.helium_singlet_cross_section (microphysics.py:114-143)
PW_HD double sigma_he_singlet(double wl) {
  const double energy = 12398.41984332 / wl;
  double scale = 37.0 - 19.1 * pow(energy / 65.4, -0.76);
  if (scale < 0) scale = 0.0;
  return sigma_h(wl) * scale;
}

// He 2^3S photoionisation cross-section (microphysics.py:147-193, Norcross 1971),
// ascending wavelength; value = 8.0670E-18 * differential oscillator strength.
PW_HD double he3_table_wl(int i) {
  const double t[23] = {209.49, 219.59, 230.71, 243.01, 256.70, 271.21, 271.94, 331.36,
                        357.34, 387.75, 423.81, 467.27, 520.65, 587.81, 674.86, 792.18,
                        958.87, 1214.41, 1655.63, 2023.15, 2275.74, 2528.27, 2593.01};
  return t[i];
}
PW_HD double he3_table_sigma(int i) {
  const double t[23] = {0.1537, 0.1750, 0.200, 0.231, 0.274, 0.338, 0.343, 0.0520,
                        0.0325, 0.0310, 0.0358, 0.0461, 0.0557, 0.0620, 0.0780, 0.1138,
                        0.1572, 0.247, 0.435, 0.501, 0.537, 0.589, 0.605};
  return 8.0670E-18 * t[i];
}

// Device-resident SED tables.  Arrays of length n_h are over the hydrogen cut
// wavelength[:i_break+1] (hydrogen.py:67-72).
struct SedTables {
  int n_h;            // i_break + 1
  int i1, i0;         // helium cuts (exclusive), helium.py:97-105
  const double* gw;   // simpson weight * flux * sigma_H / E   (erg^-1 ... -> 1/s after sum)
  const double* s_h;  // sigma_H(lambda)       [cm^2]
  const double* s_he; // sigma_He,total(lambda) [cm^2]
};

// index of the scalar outputs
enum {
  PW_SC_H_PHI = 0, PW_SC_H_A0 = 1,
  PW_SC_HE_PHI1 = 2, PW_SC_HE_PHI3 = 3, PW_SC_HE_A1 = 4, PW_SC_HE_A3 = 5,
  PW_SC_HE_AH1 = 6, PW_SC_HE_AH3 = 7,
  PW_SC_N_H = 8, PW_SC_I1 = 9, PW_SC_I0 = 10,
  PW_SC_O_PHI = 11, PW_SC_O_A = 12, PW_SC_O_AH = 13, PW_SC_O_AHE = 14,  // oxygen.radiative_processes
  // carbon.radiative_processes: phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he
  PW_SC_C_PHI1 = 16, PW_SC_C_PHI2 = 17, PW_SC_C_A1 = 18, PW_SC_C_A2 = 19, PW_SC_C_AH1 = 20, PW_SC_C_AH2 = 21,
  PW_SC_C_AHE = 22,
  // helium.radiative_processes(combined_ionization=True): phi, a_he (a_h is PW_SC_HE_AH1)
  PW_SC_HEI_PHI = 24, PW_SC_HEI_A = 25,
  PW_SC_COUNT = 32
};

// microphysics.general_cross_section (microphysics.py:224-281) for one parameter row of
// sigma_properties_v1996: E_thr, E_max, E_0, sigma_0, y_a, p, y_w, y_0, y_1
PW_HD double sigma_verner(double wl, double e_thr, double e_0, double sigma_0, double y_a, double p, double y_w,
                          double y_0, double y_1) {
  const double energy = ((6.62607015e-34 * 299792458.0) / wl) * ((1.0 / 1e-10) / 1.602176634e-19);
  if (energy < e_thr) return 0.0;
  const double x = energy / e_0 - y_0;
  const double y = sqrt(x * x + y_1 * y_1);
  const double term1 = (x - 1) * (x - 1) + y_w * y_w;
  const double term2 = pow(y, 0.5 * p - 5.5);
  const double term3 = pow(1 + sqrt(y / y_a), -p);
  return sigma_0 * (term1 * term2 * term3) * 1E-18;
}
PW_HD double sigma_verner_ci(double wl) {   // 'C I'  (microphysics.py:300-301)
  return sigma_verner(wl, 1.126E1, 2.144E0, 5.027E2, 6.126E1, 5.101E0, 9.157E-2, 1.133E0, 1.607E0);
}
PW_HD double sigma_verner_cii(double wl) {  // 'C II' (microphysics.py:302-303)
  return sigma_verner(wl, 2.438E1, 4.058E-1, 8.709E0, 1.261E2, 8.578E0, 2.093E0, 4.929E1, 3.234E0);
}

// microphysics.general_cross_section(wavelength, 'O I') (microphysics.py:224-281; Verner et al.
// 1996 parameters of microphysics.py:307-308): E_thr, E_max, E_0, sigma_0, y_a, p, y_w, y_0, y_1
PW_HD double sigma_verner_oi(double wl) {
  const double e_thr = 1.362E1, e_0 = 1.240E0, sigma_0 = 1.745E3, y_a = 3.784E0, p = 1.764E1, y_w = 7.589E-2,
               y_0 = 8.698E0, y_1 = 1.271E-1;
  const double energy = ((6.62607015e-34 * 299792458.0) / wl) * ((1.0 / 1e-10) / 1.602176634e-19);
  if (energy < e_thr) return 0.0;
  const double x = energy / e_0 - y_0;
  const double y = sqrt(x * x + y_1 * y_1);
  const double term1 = (x - 1) * (x - 1) + y_w * y_w;
  const double term2 = pow(y, 0.5 * p - 5.5);
  const double term3 = pow(1 + sqrt(y / y_a), -p);
  return sigma_0 * (term1 * term2 * term3) * 1E-18;
}

// Block-cooperative (one CTA) setup.  `gw/s_h/s_he` must hold n doubles each,
// scalars PW_SC_COUNT doubles, `red` 33 doubles of shared scratch (device).
// Integrals follow scipy.integrate.simpson per panel (pw::simpson_weight).
PW_HD void sed_setup(const double* __restrict__ wl, const double* __restrict__ flux, int n,
                     double* __restrict__ gw, double* __restrict__ s_h, double* __restrict__ s_he,
                     double* __restrict__ scalars, double* red) {
  const int ib = nearest_index(wl, n, 911.65);
  const int n_h = ib + 1;
  const int i1 = nearest_index(wl, n, 504.0);
  const int i0 = ib;
  auto X = [&](int i) { return wl[i]; };

  // ---- hydrogen tables + hydrogen.radiative_processes (hydrogen.py:139-169)
  double p_fa = 0, p_f = 0, p_fae = 0;
  for (int j = PW_TID; j < n_h; j += PW_NT) {
    const double a = sigma_h(wl[j]);
    const double en = PW_HC_ERG_A / wl[j];
    const double w = simpson_weight(n_h, j, X);
    const double g = flux[j] * a / en;
    s_h[j] = a;
    s_he[j] = sigma_he_total(wl[j]);
    gw[j] = w * g;
    p_fa += w * (flux[j] * a);
    p_f += w * flux[j];
    p_fae += w * g;
  }
  const double s_fa = block_sum(p_fa, red), s_f = block_sum(p_f, red), s_fae = block_sum(p_fae, red);

  // ---- helium.radiative_processes (helium.py:85-154)
  // singlet cut [:i1]
  double q_fa1 = 0, q_f1 = 0, q_fah1 = 0, q_phi1 = 0, q_fat = 0, q_phit = 0;
  for (int j = PW_TID; j < i1; j += PW_NT) {
    const double w = simpson_weight(i1, j, X);
    const double a1 = sigma_he_singlet(wl[j]);
    // energy = ((c.h * (c.c / wl / angstrom).to(Hz)).to(eV)).value ; (energy*eV).to(erg)
    const double e_ev = (6.62607015e-34 * ((299792458.0 / wl[j]) / 1.0 * (1.0 / 1e-10 / 1.0))) *
                        (1.0 / 1.602176634e-19);
    const double e_erg = e_ev * (1.602176634e-19 / 1e-7);
    q_fa1 += w * (flux[j] * a1);
    q_f1 += w * flux[j];
    q_fah1 += w * (flux[j] * sigma_h(wl[j]));
    q_phi1 += w * (flux[j] * a1 / e_erg);
    const double at = sigma_he_total(wl[j]);  // combined_ionization=True branch (helium.py:160-177)
    q_fat += w * (flux[j] * at);
    q_phit += w * (flux[j] * at / e_erg);
  }
  const double s_fat = block_sum(q_fat, red), s_phit = block_sum(q_phit, red);
  const double s_fa1 = block_sum(q_fa1, red), s_f1 = block_sum(q_f1, red),
               s_fah1 = block_sum(q_fah1, red), s_phi1 = block_sum(q_phi1, red);
  // hydrogen cut [:i0] for a_h_3
  double q_fah3 = 0;
  for (int j = PW_TID; j < i0; j += PW_NT) {
    const double w = simpson_weight(i0, j, X);
    q_fah3 += w * (flux[j] * sigma_h(wl[j]));
  }
  const double s_fah3 = block_sum(q_fah3, red);
  // triplet: 23-point table with the flux resampled by np.interp (helium.py:116-121)
  double q_fa3 = 0, q_f3 = 0, q_phi3 = 0;
  auto X3 = [&](int i) { return he3_table_wl(i); };
  for (int j = PW_TID; j < 23; j += PW_NT) {
    const double w3 = he3_table_wl(j), a3 = he3_table_sigma(j);
    const double f3 = np_interp_at(wl, flux, n, w3, bracket(wl, n, w3, 0));
    const double en3 = 1.98644586e-08 / w3;
    const double w = simpson_weight(23, j, X3);
    q_fa3 += w * (f3 * a3);
    q_f3 += w * f3;
    q_phi3 += w * (f3 * a3 / en3);
  }
  const double s_fa3 = block_sum(q_fa3, red), s_f3 = block_sum(q_f3, red),
               s_phi3 = block_sum(q_phi3, red);

  // ---- oxygen.radiative_processes (oxygen.py:26-99): O I cut [:io1] at 12398.42 / 13.62 A,
  //      He cut [:io0] at 504 A
  const int io1 = nearest_index(wl, n, 12398.42 / 1.362E1);
  const int io0 = i1;
  double o_f = 0, o_fa = 0, o_fah = 0, o_phi = 0;
  for (int j = PW_TID; j < io1; j += PW_NT) {
    const double w = simpson_weight(io1, j, X);
    const double a = sigma_verner_oi(wl[j]);
    const double e_ev = (6.62607015e-34 * ((299792458.0 / wl[j]) / 1.0 * (1.0 / 1e-10 / 1.0))) *
                        (1.0 / 1.602176634e-19);
    const double e_erg = e_ev * (1.602176634e-19 / 1e-7);
    o_f += w * flux[j];
    o_fa += w * (flux[j] * a);
    o_fah += w * (flux[j] * sigma_h(wl[j]));
    o_phi += w * (flux[j] * a / e_erg);
  }
  const double so_f = block_sum(o_f, red), so_fa = block_sum(o_fa, red), so_fah = block_sum(o_fah, red),
               so_phi = block_sum(o_phi, red);
  double o_f0 = 0, o_fahe = 0;
  for (int j = PW_TID; j < io0; j += PW_NT) {
    const double w = simpson_weight(io0, j, X);
    o_f0 += w * flux[j];
    o_fahe += w * (flux[j] * sigma_he_total(wl[j]));
  }
  const double so_f0 = block_sum(o_f0, red), so_fahe = block_sum(o_fahe, red);

  // ---- carbon.radiative_processes (carbon.py:30-131): C I cut [:ic1] at 12398.42 / 11.26 A,
  //      C II cut [:ic2] at 12398.42 / 24.38 A, He cut [:io0]
  const int ic1 = nearest_index(wl, n, 12398.42 / 1.126E1);
  const int ic2 = nearest_index(wl, n, 12398.42 / 2.438E1);
  double c_res[2][4];
  for (int sp = 0; sp < 2; ++sp) {
    const int icut = sp == 0 ? ic1 : ic2;
    double q_f = 0, q_fa = 0, q_fah = 0, q_phi = 0;
    for (int j = PW_TID; j < icut; j += PW_NT) {
      const double w = simpson_weight(icut, j, X);
      const double a = sp == 0 ? sigma_verner_ci(wl[j]) : sigma_verner_cii(wl[j]);
      const double e_ev = (6.62607015e-34 * ((299792458.0 / wl[j]) / 1.0 * (1.0 / 1e-10 / 1.0))) *
                          (1.0 / 1.602176634e-19);
      const double e_erg = e_ev * (1.602176634e-19 / 1e-7);
      q_f += w * flux[j];
      q_fa += w * (flux[j] * a);
      q_fah += w * (flux[j] * sigma_h(wl[j]));
      q_phi += w * (flux[j] * a / e_erg);
    }
    c_res[sp][0] = block_sum(q_f, red); c_res[sp][1] = block_sum(q_fa, red);
    c_res[sp][2] = block_sum(q_fah, red); c_res[sp][3] = block_sum(q_phi, red);
  }

  if (PW_TID == 0) {
    scalars[PW_SC_HEI_PHI] = fabs(s_phit);
    scalars[PW_SC_HEI_A] = fabs(s_fat / s_f1);
    scalars[PW_SC_C_PHI1] = fabs(c_res[0][3]);
    scalars[PW_SC_C_PHI2] = fabs(c_res[1][3]);
    scalars[PW_SC_C_A1] = fabs(c_res[0][1] / c_res[0][0]);
    scalars[PW_SC_C_A2] = fabs(c_res[1][1] / c_res[1][0]);
    scalars[PW_SC_C_AH1] = fabs(c_res[0][2] / c_res[0][0]);
    scalars[PW_SC_C_AH2] = fabs(c_res[1][2] / c_res[1][0]);
    scalars[PW_SC_C_AHE] = fabs(so_fahe / so_f0);
    scalars[PW_SC_O_PHI] = fabs(so_phi);
    scalars[PW_SC_O_A] = fabs(so_fa / so_f);
    scalars[PW_SC_O_AH] = fabs(so_fah / so_f);
    scalars[PW_SC_O_AHE] = fabs(so_fahe / so_f0);
    scalars[PW_SC_H_PHI] = fabs(s_fae);
    scalars[PW_SC_H_A0] = fabs(s_fa / s_f);
    scalars[PW_SC_HE_PHI1] = fabs(s_phi1);
    scalars[PW_SC_HE_PHI3] = fabs(s_phi3);
    scalars[PW_SC_HE_A1] = fabs(s_fa1 / s_f1);
    scalars[PW_SC_HE_A3] = fabs(s_fa3 / s_f3);
    scalars[PW_SC_HE_AH1] = fabs(s_fah1 / s_f1);
    scalars[PW_SC_HE_AH3] = fabs(s_fah3 / s_f3);  // sic: mixed grids, helium.py:145-146
    scalars[PW_SC_N_H] = (double)n_h;
    scalars[PW_SC_I1] = (double)i1;
    scalars[PW_SC_I0] = (double)i0;
  }
  PW_SYNC();
}

}  // namespace pw
