// pw_transit.cuh -- line-of-sight projection, Voigt optical depth and the pixel-wise
// radiative transfer of one model (reference: p_winds/transit.py:120-230
// `radiative_transfer_2d`, :234-307 `profile_los`, :312-536 `optical_depth_2d`).
//
// scipy.special.voigt_profile -> wofz is S. G. Johnson's Faddeeva package (third
// party, compiled into scipy, not in /root/reference): continued fraction for large
// |z|, Zaghloul & Ali's Algorithm 916 series otherwise.  Restated here for the
// real part only (the Voigt profile needs nothing else), FP64 throughout.
#pragma once
#include "pw_common.cuh"

namespace pw {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double erfcx_pos(double y) { return erfcx(y); }
#else
inline double erfcx_pos(double y) {  // host build (tests only): glibc has no erfcx
  if (y < 25.0) return exp(y * y) * erfc(y);
  const double y2 = y * y;
  return 0.5641895835477563 / y * (1.0 - 0.5 / y2 + 0.75 / (y2 * y2));
}
#endif

// Re w(x + i y), y >= 0, w(z) = exp(-z^2) erfc(-i z).  Relative accuracy ~1e-15
// (checked against scipy.special.wofz in tests/test_hostsim_stages.py::test_wofz_against_scipy and on the GPU).
// exp(-a^2 n^2) and a^2 n^2 for n = 0 .. 52 (a = pi / sqrt(-log(eps / 2)); beyond n = 52 the factor underflows):
// the coefficients of the Zaghloul-Ali series, correctly rounded (made with 400-bit arithmetic).  S. G. Johnson's
// implementation behind scipy.special.wofz tabulates the same numbers; evaluating exp(-a2 * n * n) per term, as
// round 1 did, cost one exponential per term of the series -- a third of the 'formal' Voigt kernel.
#define PW_EXPA2N2_VALUES { 1, 0.76440528167122157, 0.34142452716654842, 0.089107264692941252, 0.013588729905546009, 0.0012108545525343747, 6.3045261393344934e-05, 1.9180515657711467e-06, 3.4096944771483239e-08, 3.5417508909946941e-10, 2.149650795832607e-12, 7.6236891183372437e-15, 1.5798279711068111e-17, 1.9129418910358267e-20, 1.3534465676420535e-23, 5.5953571242858869e-27, 1.3516425797240177e-30, 1.9078458284350117e-34, 1.5735192029144294e-38, 7.5831243232803286e-43, 2.1353627543869708e-47, 3.5135206378719578e-52, 3.3780083026639692e-57, 1.8976943946830099e-62, 6.2292992607266884e-68, 1.1948117200693873e-73, 1.3390818113300596e-79, 8.769243034832239e-86, 3.35555576166255e-92, 7.5026411068817307e-99, 9.8019220074541035e-106, 7.4826541282226891e-113, 3.337701225668094e-120, 8.6993459815986112e-128, 1.3248695148408885e-135, 1.1789814420131525e-143, 6.1303912023618002e-152, 1.8625878595082211e-160, 3.3066840820143276e-169, 3.4301728088794625e-178, 2.0791539777580822e-187, 7.3638454532398496e-197, 1.5239476039408574e-206, 1.8428193504653209e-216, 1.3020955380299293e-226, 5.3758890352108054e-237, 1.2968958459976315e-247, 1.8281307802286657e-258, 1.5057635534868423e-269, 7.246923207992942e-281, 2.0379705131472682e-292, 3.3488021592787382e-304, 3.2153514072389716e-316 }
#define PW_A2N2_VALUES { 0, 0.26865715707523596, 1.0746286283009439, 2.4179144136771238, 4.2985145132037754, 6.7164289268808988, 9.6716576547084951, 13.164200696686562, 17.194058052815102, 21.761229723094111, 26.865715707523595, 32.50751600610355, 38.686630618833981, 45.403059545714875, 52.656802786746248, 60.447860341928092, 68.776232211260407, 77.641918394743186, 87.044918892376444, 96.98523370416018, 107.46286283009438, 118.47780627017906, 130.0300640244142, 142.11963609279982, 154.74652247533592, 167.91072317202247, 181.6122381828595, 195.85106750784701, 210.62721114698499, 225.94066910027342, 241.79144136771237, 258.17952794930176, 275.10492884504163, 292.56764405493198, 310.56767357897274, 329.10501741716405, 348.17967556950578, 367.79164803599804, 387.94093481664072, 408.62753591143388, 429.85145132037752, 451.61268104347164, 473.91122508071624, 496.74708343211125, 520.12025609765681, 544.03074307735278, 568.47854437119929, 593.46365997919622, 618.98608990134369, 645.04583413764146, 671.64289268808989, 698.77726555268873, 726.448952731438 }
constexpr int kWofzTab = 53;
constexpr double kExpA2N2[kWofzTab] = PW_EXPA2N2_VALUES;
constexpr double kA2N2[kWofzTab] = PW_A2N2_VALUES;
#if defined(__CUDACC__)
__constant__ double kExpA2N2Dev[kWofzTab] = PW_EXPA2N2_VALUES;
__constant__ double kA2N2Dev[kWofzTab] = PW_A2N2_VALUES;
#endif
#if defined(__CUDA_ARCH__)
#define PW_WOFZ_E(n) kExpA2N2Dev[n]
#define PW_WOFZ_A(n) kA2N2Dev[n]
#else
#define PW_WOFZ_E(n) kExpA2N2[n]
#define PW_WOFZ_A(n) kA2N2[n]
#endif

PW_HD double wofz_re(double xin, double y) {
  const double x = fabs(xin);
  const double a = 0.518321480430085929872;   // pi / sqrt(-log(eps*0.5))
  const double c = 0.329973702884629072537;   // (2/pi) * a
  const double a2 = 0.268657157075235951582;  // a^2
  const double relerr = 2.220446049250313e-16;
  const double ispi = 0.56418958354775628694807945156;  // 1 / sqrt(pi)
  if (y > 7 || (x > 6 && (y > 0.1 || (x > 8 && y > 1e-10) || x > 28))) {
    // continued fraction (Gautschi / Poppe & Wijers), nu terms
    if (x + y > 4000) {
      if (x + y > 1e7) {  // w(z) ~ i / (sqrt(pi) z)
        if (x > y) {
          const double yax = y / x;
          const double denom = ispi / (x + yax * y);
          return denom * yax;
        }
        const double xya = x / y;
        return ispi / (xya * x + y);
      }
      const double dr = x * x - y * y - 0.5, di = 2 * x * y;  // w ~ i z / (sqrt(pi) (z^2 - 1/2))
      const double denom = ispi / (dr * dr + di * di);
      return denom * (x * di - y * dr);
    }
    const double c0 = 3.9, c1 = 11.398, c2 = 0.08254, c3 = 0.1421, c4 = 0.2023;
    double nu = floor(c0 + c1 / (c2 * x + c3 * y + c4));
    double wr = x, wi = y;
    for (nu = 0.5 * (nu - 1); nu > 0.4; nu -= 0.5) {
      const double denom = nu / (wr * wr + wi * wi);
      wr = x - wr * denom;
      wi = y + wi * denom;
    }
    const double denom = ispi / (wr * wr + wi * wi);
    return denom * wi;
  }
  double sum1 = 0, sum2 = 0, sum3 = 0, sum5 = 0;
  double ret;
  if (x < 10) {
    double prod2ax = 1, prodm2ax = 1, expx2;
    if (x < 5e-4) {
      const double x2 = x * x;
      expx2 = 1 - x2 * (1 - 0.5 * x2);
      const double ax2 = 1.036642960860171859744 * x;  // 2 a x
      const double exp2ax = 1 + ax2 * (1 + ax2 * (0.5 + 0.166666666666666666667 * ax2));
      const double expm2ax = 1 - ax2 * (1 - ax2 * (0.5 - 0.166666666666666666667 * ax2));
      for (int n = 1; n < kWofzTab; ++n) {
        const double coef = PW_WOFZ_E(n) * expx2 / (PW_WOFZ_A(n) + y * y);
        prod2ax *= exp2ax;
        prodm2ax *= expm2ax;
        sum1 += coef;
        sum2 += coef * prodm2ax;
        sum3 += coef * prod2ax;
        if (coef * prod2ax < relerr * sum3) break;
      }
    } else {
      expx2 = exp(-x * x);
      const double exp2ax = exp((2 * a) * x), expm2ax = 1 / exp2ax;
      for (int n = 1; n < kWofzTab; ++n) {
        const double coef = PW_WOFZ_E(n) * expx2 / (PW_WOFZ_A(n) + y * y);
        prod2ax *= exp2ax;
        prodm2ax *= expm2ax;
        sum1 += coef;
        sum2 += coef * prodm2ax;
        sum3 += coef * prod2ax;
        sum5 += (coef * prod2ax) * (a * n);
        if ((coef * prod2ax) * (a * n) < relerr * sum5) break;
      }
    }
    const double expx2erfcxy = expx2 * erfcx_pos(y);
    const double xy = x * y;
    if (y > 5) {
      const double sinxy = sin(xy);
      const double sc = (fabs(xy) < 1e-4) ? 1 - (0.1666666666666666666667) * xy * xy : sinxy / xy;
      ret = (expx2erfcxy - c * y * sum1) * cos(2 * xy) + (c * x * expx2) * sinxy * sc;
    } else {
      const double sinxy = sin(xy);
      const double cos2xy = cos(2 * xy);
      const double coef1 = expx2erfcxy - c * y * sum1;
      const double coef2 = c * x * expx2;
      const double sc = (fabs(xy) < 1e-4) ? 1 - (0.1666666666666666666667) * xy * xy : sinxy / xy;
      ret = coef1 * cos2xy + coef2 * sinxy * sc;
    }
  } else {
    // x >= 10 with tiny y: exp(-x^2) underflows relative to the sum; only the terms
    // around n0 = x / a matter
    ret = 0.0;
    const double n0 = floor(x / a + 0.5);
    const double dx = a * n0 - x;
    sum3 = exp(-dx * dx) / (a2 * (n0 * n0) + y * y);
    sum5 = a * n0 * sum3;
    const double exp1 = exp(4 * a * dx);
    double exp1dn = 1;
    int dn;
    bool done = false;
    for (dn = 1; n0 - dn > 0; ++dn) {
      const double np = n0 + dn, nm = n0 - dn;
      double tp = exp(-(a * dn + dx) * (a * dn + dx));
      double tm = tp * (exp1dn *= exp1);
      tp /= (a2 * (np * np) + y * y);
      tm /= (a2 * (nm * nm) + y * y);
      sum3 += tp + tm;
      sum5 += a * (np * tp + nm * tm);
      if (a * (np * tp + nm * tm) < relerr * sum5) { done = true; break; }
    }
    while (!done && dn < 400) {
      const double np = n0 + dn++;
      const double tp = exp(-(a * dn + dx) * (a * dn + dx)) / (a2 * (np * np) + y * y);
      sum3 += tp;
      sum5 += a * np * tp;
      if (a * np * tp < relerr * sum5) break;
    }
  }
  return ret + (0.5 * c) * y * (sum2 + sum3);
}

// scipy.special.voigt_profile(x, sigma, gamma)  (xsf voigt_profile)
PW_HD double voigt_profile(double x, double sigma, double gamma) {
  const double INV_SQRT_2 = 0.707106781186547524401;
  const double SQRT_2PI = 2.5066282746310002416123552393401042;
  if (sigma == 0) {
    if (gamma == 0) {
      if (x != x) return x;
      if (x == 0) return 1.0 / 0.0;
      return 0;
    }
    return gamma / kPi / (x * x + gamma * gamma);
  }
  if (gamma == 0) return 1 / SQRT_2PI / sigma * exp(-(x / sigma) * (x / sigma) / 2);
  const double zreal = x / sigma * INV_SQRT_2;
  const double zimag = gamma / sigma * INV_SQRT_2;
  return wofz_re(zreal, zimag) / sigma / SQRT_2PI;
}

struct RtParams {
  // per batch
  int n_lines;
  double w0[8], fosc[8], aij[8];  // central wavelength [m], oscillator strength, Einstein A [1/s]
  double mass;                    // particle mass [kg]
  double v_bulk, rv_planet;       // [m/s]
  int nz;                         // z_grid_size
  int turbulence;                 // turbulence_broadening
  int t_is_profile;               // gas_temperature given per radius
};

// One LOS cell (transit.py:273-296): distance, interpolated density/velocity
// (/temperature), line-of-sight velocity.  r, n, v in SI.
PW_HD void los_cell(const double* __restrict__ r, const double* __restrict__ n, const double* __restrict__ v,
                    const double* __restrict__ tprof, int nr, double b, double z, double* n_c, double* vlos,
                    double* t_c) {
  const double d = sqrt(b * b + z * z);
  double nc = 0.0, vc = 0.0, tc = 0.0;
  if (!(d < r[0] || d > r[nr - 1])) {
    const int hi = interp1d_hi(r, nr, d);
    const int lo = hi - 1;
    const double dx = r[hi] - r[lo];
    const double whi = PW_DIV_INLINE(d - r[lo], dx), wlo = PW_DIV_INLINE(r[hi] - d, dx);
    nc = whi * n[hi] + wlo * n[lo];
    vc = whi * v[hi] + wlo * v[lo];
    if (tprof) tc = whi * tprof[hi] + wlo * tprof[lo];
  }
  *n_c = nc;
  *vlos = PW_DIV_INLINE(vc * z, sqrt(z * z + b * b));
  if (t_c) *t_c = tc;
}

// numpy.linspace(-r_top, r_top, nz)[j]
PW_HD double los_z(double r_top, int nz, int j) {
  if (j == nz - 1) return r_top;
  const double step = (r_top - (-r_top)) / (double)(nz - 1);
  return (double)j * step + (-r_top);
}

// Phase A of the 'average' method, block-cooperative: column density per impact
// parameter ncol[i] = trapezoid_z(n_los[:, i]) and the density-weighted moments
// sums[0] = sum n, sums[1] = sum v_los^2 n, sums[2] = sum T n over the whole
// (Nz, Nr) LOS grid (transit.py:482-488, :517-519).
PW_HD void rt_los_average(const double* __restrict__ r, const double* __restrict__ n,
                          const double* __restrict__ v, const double* __restrict__ tprof, int nr, int nz,
                          double* __restrict__ ncol, double* sums, double* red, double* rows) {
  // `rows`: nz doubles of scratch per warp (device: shared memory; host: one row)
  const double r_top = r[nr - 1];
  double s_n = 0.0, s_v2n = 0.0, s_tn = 0.0;
#if defined(__CUDA_ARCH__)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, nlane = 32;
#else
  const int lane = 0, warp = 0, nwarp = 1, nlane = 1;
#endif
  double* row = rows + (size_t)warp * nz;
  for (int i = warp; i < nr; i += nwarp) {
    const double b = r[i];
    // every cell of the line of sight once: the lane that owns z_j keeps n(z_j) in the warp's
    // row for its neighbour's trapezoid and adds the cell to the moments
    for (int j = lane; j < nz; j += nlane) {
      const double z0 = los_z(r_top, nz, j);
      double n0, vl0, t0;
      los_cell(r, n, v, tprof, nr, b, z0, &n0, &vl0, tprof ? &t0 : nullptr);
      row[j] = n0;
      s_n += n0;
      s_v2n += (vl0 * vl0) * n0;
      if (tprof) s_tn += t0 * n0;
    }
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    double acc = 0.0;
    // each lane owns z-intervals j -> (z_j, z_{j+1}); trapezoid d * (y1 + y0) / 2
    for (int j = lane; j + 1 < nz; j += nlane) {
      const double z0 = los_z(r_top, nz, j), z1 = los_z(r_top, nz, j + 1);
      acc += (z1 - z0) * (row[j + 1] + row[j]) / 2.0;
    }
    acc = warp_sum(acc);
    if (lane == 0) ncol[i] = acc;
#if defined(__CUDA_ARCH__)
    __syncwarp();   // the row is reused for the next impact parameter
#endif
  }
  const double t_n = block_sum(s_n, red);
  const double t_v2n = block_sum(s_v2n, red);
  const double t_tn = block_sum(s_tn, red);
  if (PW_TID == 0) { sums[0] = t_n; sums[1] = t_v2n; sums[2] = t_tn; }
  PW_SYNC();
}

// Phase B: Phi(lambda) = sum_k sigma_k * voigt(dnu_k + shift_k, alpha_k, gamma_k)
// for one wavelength (transit.py:436-453, :479-513, :523-534), 'average' method.
// The line list comes in through pointers so that a kernel can hand over the arrays of its
// (constant-bank) parameter block as they are: copying them into a local struct and indexing that
// with the line number puts the copy in local memory.
PW_HD double rt_phi_average_raw(const double* __restrict__ w0, const double* __restrict__ fosc,
                                const double* __restrict__ aij, int n_lines, double mass, double v_bulk,
                                double rv_planet, int turbulence, double wl, double temp, double vw2) {
  const double c_speed = 2.99792458e+08, k_b = 1.380649e-23;
  double phi = 0.0;
  const double nu_rest = c_speed / wl;
  const double vt2 = turbulence ? (5.0 / 6 * k_b * temp / mass) : 0.0;  // sqrt(.)**2, :495
  for (int k = 0; k < n_lines; ++k) {
    const double nu0 = c_speed / w0[k];
    const double gamma = aij[k] / 4 / kPi;
    const double alpha = nu0 / c_speed * sqrt(k_b * temp / mass + vw2 + vt2);
    const double shift = (v_bulk + rv_planet) / c_speed * nu0;
    const double prof = voigt_profile((nu_rest - nu0) + shift, alpha, gamma);
    phi += prof * (2.654008854574474e-06 * fosc[k]);
  }
  return phi;
}
PW_HD double rt_phi_average(const RtParams& p, double wl, double temp, double vw2) {
  return rt_phi_average_raw(p.w0, p.fosc, p.aij, p.n_lines, p.mass, p.v_bulk, p.rv_planet, p.turbulence, wl, temp,
                            vw2);
}

}  // namespace pw
