// pw_parker.cuh -- isothermal Parker wind (reference: p_winds/parker.py).
//
// The reference finds v(r) by a secant iteration on v exp(-v^2/2) = D(r)
// (parker.py:191-219, :321-347).  With u = v^2 that is u - 1 - ln u = s,
// s = -ln(D^2) - 1 >= 0, i.e. v = sqrt(-W_k(-D^2)) with the Lambert-W branch
// k = 0 below the sonic point and k = -1 above it.  We solve it by a monotone
// (safeguard-free) Newton iteration on the convex function u - 1 - ln u, started on
// the outer side of the root; agreement with the reference's secant result is
// <= 2e-13 relative (tests/test_parity_parker.py).
#pragma once
#include "pw_common.cuh"

namespace pw {

// parker.py:83-106  (km/s)
PW_HD double sound_speed(double T, double mu) {
  return sqrt(1.380649e-29 * T / mu / 1.67262192369e-27);
}
// parker.py:110-131  (R_Jup)
PW_HD double radius_sonic_point(double m_pl, double cs) {
  return 1772.0378503888546 * m_pl / 2 / (cs * cs);
}
// parker.py:226-267
PW_HD double radius_sonic_point_tidal(double m_pl, double cs, double m_star, double a) {
  const double grav = 1772.0378503888546;
  const double ms = 1047.56551466 * m_star;
  const double aa = 2092.51203911 * a;
  const double vk = sqrt(grav * ms / aa);
  const double q = cs / vk;
  const double q2 = q * q;
  const double m1 = sqrt(m_pl * m_pl / 4 + 8 * (ms * ms) / 81 * (q2 * q2 * q2));
  return aa * (cbrt((m1 + m_pl / 2) / (3 * ms)) - cbrt((m1 - m_pl / 2) / (3 * ms)));
}
// parker.py:135-159  (g cm^-3)
PW_HD double density_sonic_point(double mdot, double rs, double cs) {
  const double r = rs * kRjupCm;
  return mdot / 4 / kPi / (r * r) / (cs * 1E5);
}

// Root of u - 1 - ln(u) = s on the requested side of u = 1.
PW_HD double parker_u(double s, bool supersonic) {
  if (!(s > 0.0)) return 1.0;
  const double p = sqrt(2.0 * s);
  double u;
  if (supersonic) {
    u = 1.0 + s + p;  // right of the root: Newton on a convex increasing branch is monotone
    for (int it = 0; it < 60; ++it) {
      const double g = u - 1.0 - log(u) - s;
      const double du = g / (1.0 - 1.0 / u);
      u -= du;
      if (fabs(du) <= 4e-16 * u) break;
    }
  } else {
    // left of the root; iterate on h(u) = ln u - u + 1 + s (concave, increasing)
    u = exp(-1.0 - s);
    if (p < 1.0 && 1.0 - p > u) u = 1.0 - p;
    for (int it = 0; it < 60; ++it) {
      const double hval = log(u) - u + 1.0 + s;
      const double du = hval / (1.0 / u - 1.0);
      u -= du;
      if (fabs(du) <= 4e-16 * u) break;
    }
  }
  return u;
}

// parker.structure (parker.py:163-223), one radius (r in units of the sonic radius)
PW_HD void parker_point(double r, double* v, double* rho) {
  const double s = 4.0 * (log(r) + 1.0 / r - 1.0);
  const double u = parker_u(s, r > 1.0);
  const double vv = sqrt(u);
  *v = vv;
  *rho = exp(2 * 1 / r - 3.0 / 2 - 0.5 * (vv * vv));
}

// Coefficients of parker.structure_tidal (parker.py:271-358) that do not depend
// on r: exponent(r) = -A/r + A - B r^2 + B - 0.5 with r in sonic-radius units.
struct TidalCoef {
  double gm_k2;   // G m_p / (c_s^2 r_s)
  double tid_k3;  // 3 G m_* r_s^2 / (2 a^3 c_s^2)
  // spelled-out pieces for the density expression of the reference (:350-356)
  double cs2, rs, gmp, gms3, k3;
};

PW_HD TidalCoef tidal_coef(double cs_kms, double rs_rjup, double m_pl, double m_star, double a_au) {
  TidalCoef c;
  const double cs = 100000. * cs_kms;
  const double rs = 7.1492e+09 * rs_rjup;
  const double mp = 1.8981246e+30 * m_pl;
  const double ms = 1.98840987e+33 * m_star;
  const double aa = 1.49597871e+13 * a_au;
  const double g = 6.6743e-11 * (1.0 / ((1e-2 * 1e-2 * 1e-2) / 1e-3));  // astropy G in cgs
  c.cs2 = cs * cs;
  c.rs = rs;
  c.gmp = g * mp;
  c.gms3 = 3 * g * ms;
  c.k3 = 2 * (aa * aa * aa) * c.cs2;
  c.gm_k2 = c.gmp / (c.cs2 * rs);
  c.tid_k3 = c.gms3 * (rs * rs) / c.k3;
  return c;
}

PW_HD void parker_point_tidal(const TidalCoef& c, double r, double* v, double* rho) {
  // exponent of the right-hand side of parker.py:322-327
  const double rr = r * c.rs;
  const double e = -c.gmp / (c.cs2 * rr) + c.gmp / (c.cs2 * c.rs) - c.gms3 * (rr * rr) / c.k3 +
                   c.gms3 * (c.rs * c.rs) / c.k3 - 0.5;
  // s = -ln(D^2) - 1 with D = r^-2 exp(e)
  const double s = 4.0 * log(r) - 2.0 * e - 1.0;
  const double u = parker_u(s, r > 1.0);
  const double vv = sqrt(u);
  *v = vv;
  // parker.py:350-356
  const double k1 = c.cs2 * rr, k2 = c.cs2 * c.rs;
  *rho = exp(c.gmp / k1 - c.gmp / k2 + c.gms3 * (rr * rr) / c.k3 - c.gms3 * (c.rs * c.rs) / c.k3 +
             0.5 - (vv * vv) / 2);
}

// parker.average_molecular_weight (parker.py:20-79), serial; radius profile is
// r_pl_units * r_pl [R_Jup], velocity profile v_norm * vs_kms [km/s].  Lampon et al. 2020 Eq. A.3.
PW_HD double average_molecular_weight(const double* f, const double* r_pl_units, double r_pl,
                                      const double* v_norm, double vs_kms, int n, double m_pl,
                                      double T, double he_h) {
  const double mp = m_pl * 1.8981246e+27;
  const double k_b = 1.380649e-23, grav = 6.6743e-11;
  auto R = [&](int i) { return (r_pl_units[i] * r_pl) * 71492000.0; };
  auto V = [&](int i) { return (v_norm[i] * vs_kms) * 1000; };
  auto MU = [&](int i) { return (1 + 4 * he_h) / (1 + he_h + f[i]); };
  const double i1 = simpson_serial(n, [&](int i) { const double r = R(i); return MU(i) / (r * r); }, R);
  const double i2 = simpson_serial(n, [&](int i) { return MU(i) * V(i); }, V);
  double i3 = 0.0;  // numpy.trapezoid(mu_r, 1 / mu_r)
  for (int i = 0; i + 1 < n; ++i) {
    const double d = 1 / MU(i + 1) - 1 / MU(i);
    i3 += d * (MU(i + 1) + MU(i)) / 2.0;
  }
  const double i4 = simpson_serial(n, [&](int i) { const double r = R(i); return 1 / (r * r); }, R);
  const double i5 = simpson_serial(n, V, V);
  const double i6 = 1 / MU(n - 1) - 1 / MU(0);
  const double term_1 = grav * mp * i1 + i2 + k_b * T * i3;
  const double term_2 = grav * mp * i4 + i5 + k_b * T * i6;
  return term_1 / term_2;
}

}  // namespace pw
