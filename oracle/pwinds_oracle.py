"""CPU oracle for the p-winds retrieval hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy/scipy *restatement* of the
reference's algorithm for the path named in BASELINE.json (`hydrogen.ion_fraction`
-> `parker.structure[_tidal]` -> `helium.population_fraction` ->
`transit.draw_transit` -> `transit.radiative_transfer_2d`).  It exists so that the
CUDA engine can be checked on the GPU box, where /root/reference does not exist.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it; the product package `p_winds_b200` never does.

Parity pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (imported
from /root/reference with the stand-ins in `oracle/standins/`) and this module on
the same inputs and stores the reference outputs in `tests/golden/`;
`tests/test_oracle_vs_golden.py` holds this module to those vectors (bit-exact or
<= 1e-13 relative for every stage) and to the reference's own golden numbers
(tests/test_hydrogen.py:35,42; tests/test_transit.py:35-48; tests/test_parker.py).

The numerical work is delegated to the same scipy/numpy routines the reference
calls (`solve_ivp` RK45, `odeint` LSODA, `simpson`, `cumulative_trapezoid`,
`optimize.newton`, `interp1d`, `voigt_profile`); scipy/numpy are third-party,
unpinned in the reference (`requirements.txt:1-2`), versions in this image:
scipy 1.18.1 / numpy 2.3.5.  `flatstar` (requirements.txt:4, absent) is restated
from the reference's golden transit depth: Pillow's integer ellipse rasteriser
(see `pil_ellipse_spans`).

All citations are to files under /root/reference/p_winds/.
"""
import warnings

import numpy as np
from scipy.integrate import (simpson, solve_ivp, odeint, cumulative_trapezoid,
                             trapezoid)
from scipy.interpolate import interp1d
from scipy.optimize import newton
from scipy.special import voigt_profile

# ----------------------------------------------------------------------------
# Constants exactly as the reference spells them
# ----------------------------------------------------------------------------
RJUP_CM = 7.1492E+09          # parker.py:157
M_H_G = 1.67262192E-24        # hydrogen.py:347, helium.py:560
_H_SI = 6.62607015e-34        # astropy.constants.h  (CODATA 2018)
_C_SI = 299792458.0           # astropy.constants.c
_EV_J = 1.602176634e-19
# (c.h * c.c).to(erg * angstrom).value, hydrogen.py:61
HC_ERG_A = (_H_SI * _C_SI) * (1.0 / (1e-7 * 1e-10))
G_CGS = 6.6743e-11 * (1.0 / ((1e-2 ** 3) / 1e-3 / (1.0 ** 2)))  # parker.py:318


class Sed:
    """Stellar spectrum at the planet: wavelength [angstrom], flux
    [erg s-1 cm-2 A-1].  Replaces the reference's spectrum dict
    (tools.py:405-408) after unit conversion (hydrogen.py:57-60)."""

    def __init__(self, wavelength, flux_lambda):
        self.wl = np.ascontiguousarray(wavelength, dtype=float)
        self.flux = np.ascontiguousarray(flux_lambda, dtype=float)

    @classmethod
    def from_file(cls, path, scale_flux=1.0):
        t = np.loadtxt(path, usecols=(0, 1), dtype=float)   # tools.py:385
        return cls(t[:, 0], t[:, 1] * scale_flux * 1.0)


# ----------------------------------------------------------------------------
# tools / microphysics
# ----------------------------------------------------------------------------
def nearest_index(array, target):
    """tools.py:27-48."""
    i = int(np.clip(np.searchsorted(array, target), 1, len(array) - 1))
    left, right = array[i - 1], array[i]
    return i - int(target - left < right - target)


def sigma_h(wl):
    """microphysics.py:53-62 (wavelength branch), NaN -> 0."""
    wl = np.asarray(wl, dtype=float)
    with np.errstate(all='ignore'):
        eps = (911.65 / wl - 1) ** 0.5
        a = (6.3E-18 * np.exp(4 - (4 * np.arctan(eps)) / eps) /
             (1 - np.exp(-2 * np.pi / eps)) * (wl / 911.65) ** 4)
    if isinstance(a, np.ndarray):
        a[np.isnan(a)] = 0.0
    return a


def sigma_he_total(wl):
    """microphysics.py:78-110 (Yan et al. 1998) with the unit algebra of the
    astropy stand-in written out."""
    wl = np.asarray(wl, dtype=float)
    hz = _C_SI / (wl * 1e-10)
    e_ev = hz * _H_SI / _EV_J
    x = e_ev / 24.58
    e_term = (e_ev / 1.0 / 1000.) ** (7 / 2)
    coefs = [-4.7416, 14.8200, -30.8678, 37.3584, -23.4585, 5.9133]
    s = 0.
    for i, c in enumerate(coefs):
        s = s + c / x ** ((i + 1) / 2)
    out = (733 * (1e-28 / (1e-2 * 1e-2))) / e_term * (1 + s)
    out = np.array(out, dtype=float)
    out[e_ev < 24.58] = 0.
    return out


def sigma_he_singlet(wl):
    """microphysics.py:114-143."""
    wl = np.asarray(wl, dtype=float)
    energy = 12398.41984332 / wl
    scale = np.ones_like(wl)
    scale *= 37.0 - 19.1 * (energy / 65.4) ** (-0.76)
    scale[scale < 0] = 0.0
    return sigma_h(wl) * scale


_HE3_TABLE = np.array([
    [2593.01, 0.605], [2528.27, 0.589], [2275.74, 0.537], [2023.15, 0.501],
    [1655.63, 0.435], [1214.41, 0.247], [958.87, 0.1572], [792.18, 0.1138],
    [674.86, 0.0780], [587.81, 0.0620], [520.65, 0.0557], [467.27, 0.0461],
    [423.81, 0.0358], [387.75, 0.0310], [357.34, 0.0325], [331.36, 0.0520],
    [271.94, 0.343], [271.21, 0.338], [256.70, 0.274], [243.01, 0.231],
    [230.71, 0.200], [219.59, 0.1750], [209.49, 0.1537]])


def sigma_he_triplet_table():
    """microphysics.py:147-193 (Norcross 1971): ascending wavelength, cm^2."""
    return np.flip(_HE3_TABLE[:, 0]), 8.0670E-18 * np.flip(_HE3_TABLE[:, 1])


_HE_GAMMA = np.array([
    [3.75, 6.198E-2, 2.389, 7.965E-1], [4.00, 6.458E-2, 2.456, 9.579E-1],
    [4.25, 6.387E-2, 2.275, 1.042], [4.50, 6.157E-2, 1.916, 1.015],
    [4.75, 5.832E-2, 1.496, 8.950E-1], [5.00, 5.320E-2, 1.111, 7.265E-1],
    [5.25, 4.787E-2, 8.003E-1, 5.516E-1], [5.50, 4.018E-2, 5.660E-1, 3.948E-1],
    [5.75, 3.167E-2, 3.944E-1, 2.677E-1]])   # microphysics.py:197-220


def he3_lines():
    """lines.py:14-55: (w0[3] m, f[3], A[3] 1/s)."""
    return (np.array([1.082909114, 1.083025010, 1.083033977]) * 1E-6,
            np.array([5.9902e-02, 1.7974e-01, 2.9958e-01]),
            np.array([1.0216e+07] * 3))


# ----------------------------------------------------------------------------
# parker
# ----------------------------------------------------------------------------
def sound_speed(T, mu=1.0):
    """parker.py:83-106 (km/s)."""
    return (1.380649e-29 * T / mu / 1.67262192369e-27) ** 0.5


def radius_sonic_point(m_pl, cs):
    """parker.py:110-131 (R_Jup)."""
    return 1772.0378503888546 * m_pl / 2 / cs ** 2


def radius_sonic_point_tidal(m_pl, cs, m_star, a):
    """parker.py:226-267."""
    grav = 1772.0378503888546
    ms = 1047.56551466 * m_star
    aa = 2092.51203911 * a
    v_k = np.sqrt(grav * ms / aa)
    m1 = np.sqrt(m_pl ** 2 / 4 + 8 * ms ** 2 / 81 * (cs / v_k) ** 6)
    return aa * (((m1 + m_pl / 2) / (3 * ms)) ** (1 / 3) -
                 ((m1 - m_pl / 2) / (3 * ms)) ** (1 / 3))


def density_sonic_point(mdot, rs, cs):
    """parker.py:135-159 (g cm-3)."""
    return mdot / 4 / np.pi / (rs * RJUP_CM) ** 2 / (cs * 1E5)


def structure(r, v_guess=None):
    """parker.py:163-223: secant solve of v exp(-v^2/2) = r^-2 exp(-2/r + 3/2)."""
    def eq(v, rk):
        return v * np.exp(-0.5 * v ** 2) - (1 / rk) ** 2 * np.exp(-2 * 1 / rk + 3 / 2)
    if v_guess is not None:
        v = newton(eq, x0=v_guess, args=(r,))
    elif isinstance(r, np.ndarray):
        v = newton(eq, x0=np.array(r > 1, dtype=int) * 2 + 0.1, args=(r,))
    else:
        v = newton(eq, x0=1E-1 if r <= 1.0 else 2.0, args=(r,))
    return v, np.exp(2 * 1 / r - 3 / 2 - 0.5 * v ** 2)


def structure_tidal(r, cs, rs, m_pl, m_star, a):
    """parker.py:271-358."""
    cs = 100000. * cs
    rs = 7.1492e+09 * rs
    mp = 1.8981246e+30 * m_pl
    ms = 1.98840987e+33 * m_star
    aa = 1.49597871e+13 * a
    g = G_CGS

    def eq(v, rk):
        return v * np.exp(-v ** 2 / 2) - (1 / rk) ** 2 * np.exp(
            - g * mp / (cs ** 2 * (rk * rs)) + g * mp / (cs ** 2 * rs)
            - 3 * g * ms * (rk * rs) ** 2 / (2 * aa ** 3 * cs ** 2)
            + 3 * g * ms * rs ** 2 / (2 * aa ** 3 * cs ** 2) - 0.5)
    if isinstance(r, np.ndarray):
        v = newton(eq, x0=(np.array(r > 1, dtype=int) * 2 + 0.1), args=(r,),
                   maxiter=1000)
    else:
        v = newton(eq, x0=1E-1 if r <= 1.0 else 2.0, args=(r,))
    k1 = cs ** 2 * (r * rs)
    k2 = cs ** 2 * rs
    k3 = 2 * aa ** 3 * cs ** 2
    rho = np.exp(g * mp / k1 - g * mp / k2 + 3 * g * ms * (r * rs) ** 2 / k3 -
                 3 * g * ms * rs ** 2 / k3 + 0.5 - np.array(v) ** 2 / 2)
    return v, rho


def average_molecular_weight(f_r, r_rjup, v_kms, m_pl, T, he_h):
    """parker.py:20-79 (Lampon et al. 2020 Eq. A.3)."""
    mp = m_pl * 1.8981246e+27
    r = r_rjup * 71492000.0
    v = v_kms * 1000
    k_b, grav = 1.380649e-23, 6.6743e-11
    mu_r = (1 + 4 * he_h) / (1 + he_h + f_r)
    i1 = simpson(mu_r / r ** 2, x=r)
    i2 = simpson(mu_r * v, x=v)
    i3 = trapezoid(mu_r, 1 / mu_r)
    i4 = simpson(1 / r ** 2, x=r)
    i5 = simpson(v, x=v)
    i6 = 1 / mu_r[-1] - 1 / mu_r[0]
    return (grav * mp * i1 + i2 + k_b * T * i3) / (grav * mp * i4 + i5 + k_b * T * i6)


# ----------------------------------------------------------------------------
# hydrogen
# ----------------------------------------------------------------------------
def h_radiative_processes(sed):
    """hydrogen.py:118-169 -> (phi [1/s], a_0 [cm2])."""
    ib = nearest_index(sed.wl, 911.65)
    wl, fl = sed.wl[:ib + 1], sed.flux[:ib + 1]
    en = (HC_ERG_A / sed.wl)[:ib + 1]
    a = sigma_h(wl)
    a_0 = abs(simpson(fl * a, x=wl) / simpson(fl, x=wl))
    phi = abs(simpson(fl * a / en, x=wl))
    return phi, a_0


def h_radiative_processes_mono(flux_euv, average_photon_energy=20.):
    """hydrogen.py:173-203."""
    a_0 = 6.3E-18 * (average_photon_energy / 13.6) ** (-3)
    return flux_euv * 6.24150907E+11 * a_0 / average_photon_energy, a_0


def h_recombination(T):
    """hydrogen.py:207-223."""
    return 2.59E-13 * (T / 1E4) ** (-0.7)


def phi_exact(sed, r_cm, rho, f_h, h_fraction, f_he=None):
    """hydrogen.py:22-114: tau-dependent photoionisation rate per radius."""
    ib = nearest_index(sed.wl, 911.65)
    wl, fl = sed.wl[:ib + 1], sed.flux[:ib + 1]
    en = (HC_ERG_A / sed.wl)[:ib + 1]
    xx, _ = np.meshgrid(wl, r_cm)
    a_l = sigma_h(xx)
    r_rev = r_cm[::-1]
    he_to_h = (1 - h_fraction) / h_fraction
    mu = (1 + 4 * he_to_h) / (1 + f_h + he_to_h)
    n_tot = rho / mu / M_H_G
    n_htot = 1 / (1 + f_h + he_to_h) * n_tot
    n_h = n_htot * (1 - f_h)
    n_hetot = n_htot * he_to_h
    n_he = n_hetot * (1 - (f_h if f_he is None else f_he))
    col_h = -cumulative_trapezoid(n_h[::-1], r_rev, initial=0)[::-1]
    tau = col_h[:, None] * a_l
    col_he = -cumulative_trapezoid(n_he[::-1], r_rev, initial=0)[::-1]
    tau += col_he[:, None] * sigma_he_total(xx)
    return abs(simpson(y=fl * a_l / en * np.exp(-tau), x=wl, axis=-1))


def h_ion_fraction(r, r_pl, T, h_fraction, mdot, m_pl, mu_0=1.0, m_star=1.0,
                   a=1.0, sed=None, flux_euv=None, initial_f_ion=0.0,
                   relax_solution=False, convergence=0.01, max_n_relax=10,
                   exact_phi=False, **ivp):
    """hydrogen.py:227-573.  Returns a dict: f_r, mu_bar, status (0 ok, 1 solver
    failed = the reference's RuntimeError at :458/:511), n_relax, nfev (list)."""
    alpha = h_recombination(T)
    vs = sound_speed(T, mu_0)
    rs = radius_sonic_point_tidal(m_pl, vs, m_star, a)
    r_cm = r * r_pl * (7.1492e7 / 1e-2)      # (r * R_pl * u.Rjup).to(u.cm)
    if exact_phi and sed is not None:
        rhos = density_sonic_point(mdot, rs, vs)
        _, rho_n = structure_tidal(r * r_pl / rs, vs, rs, m_pl, m_star, a)
        phi_abs = phi_exact(sed, r_cm, rho_n * rhos, 0.0, h_fraction)
        a_0 = 0.
    elif sed is not None:
        phi_abs, a_0 = h_radiative_processes(sed)
    elif flux_euv is not None:
        phi_abs, a_0 = h_radiative_processes_mono(flux_euv)
    else:
        raise ValueError('Either `spectrum_at_planet` or `flux_euv` must be '
                         'provided.')
    he_fraction = 1 - h_fraction
    he_h = he_fraction / h_fraction
    k1_abs = h_fraction * a_0 / (h_fraction + 4 * he_fraction) / M_H_G
    k2_abs = h_fraction / (h_fraction + 4 * he_fraction) * alpha / M_H_G

    def normalize(phi_, mu_):
        cs_ = sound_speed(T, mu_)
        rs_ = radius_sonic_point_tidal(m_pl, cs_, m_star, a)
        rhos_ = density_sonic_point(mdot, rs_, cs_)
        phi_unit = cs_ * 1E5 / rs_ / 7.1492E+09
        k1_unit = 1 / (rhos_ * rs_ * 7.1492E+09)
        k2_unit = cs_ * 1E5 / rs_ / 7.1492E+09 / rhos_
        rn = r * r_pl / rs_
        drn = np.diff(rn)
        drn = np.concatenate((drn, np.array([drn[-1], ])))
        vn, rhon = structure_tidal(rn, cs_, rs_, m_pl, m_star, a)
        return phi_ / phi_unit, k1_abs / k1_unit, k2_abs / k2_unit, rn, drn, vn, rhon

    out = {'status': 0, 'n_relax': 0, 'nfev': [], 'f_r': None, 'mu_bar': np.nan}

    def solve(phi, k2, rn, vn, rhon, tau):
        def fun(x, f):
            if exact_phi:
                pp = np.interp(x, rn, phi)
            else:
                pp = np.exp(-np.interp(x, rn, tau)) * phi
            v_ = np.interp(x, rn, vn)
            rho_ = np.interp(x, rn, rhon)
            return (1. - f) / v_ * pp - k2 * rho_ * f ** 2 / v_
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            sol = solve_ivp(fun, (rn[0], rn[-1]), np.array([initial_f_ion, ]),
                            t_eval=rn, **ivp)
        out['nfev'].append(sol.nfev)
        return sol['y'][0]

    phi, k1, k2, rn, drn, vn, rhon = normalize(phi_abs, mu_0)
    tau = None if exact_phi else k1 * np.flip(np.cumsum(np.flip(drn * rhon)))
    f_r = solve(phi, k2, rn, vn, rhon, tau)
    if len(f_r) != len(rn):
        out['status'] = 1
        return out
    mu_bar = average_molecular_weight(f_r, r * r_pl, vn * vs, m_pl, T, he_h)
    if relax_solution:
        for _ in range(max_n_relax):
            prev = np.copy(f_r)
            out['n_relax'] += 1
            if exact_phi:
                vs = sound_speed(T, mu_bar)
                rs = radius_sonic_point_tidal(m_pl, vs, m_star, a)
                rhos = density_sonic_point(mdot, rs, vs)
                _, rho_n = structure_tidal(r * r_pl / rs, vs, rs, m_pl, m_star, a)
                phi_abs = phi_exact(sed, r_cm, rho_n * rhos, f_r, h_fraction)
            phi, k1, k2, rn, drn, vn, rhon = normalize(phi_abs, mu_bar)
            if not exact_phi:
                tau = k1 * np.flip(np.cumsum(np.flip(drn * rhon * (1 - f_r))))
            f_r = solve(phi, k2, rn, vn, rhon, tau)
            if len(f_r) != len(rn):
                out['status'] = 1
                return out
            mu_bar = average_molecular_weight(f_r, r * r_pl, vn * vs, m_pl, T, he_h)
            if abs(np.sum(f_r - prev) / np.sum(prev)) < convergence:
                break
    out['f_r'] = f_r
    out['mu_bar'] = mu_bar
    if exact_phi and sed is not None:
        # return_rates=True (hydrogen.py:538-561): photoionisation with the structure of the last
        # normalisation, recombination with the final structure
        fin_phi = phi_exact(sed, r_cm, rho_n * rhos, f_r, h_fraction) * (1 - f_r)
        fvs = sound_speed(T, mu_bar)
        frs = radius_sonic_point_tidal(m_pl, fvs, m_star, a)
        frhos = density_sonic_point(mdot, frs, fvs)
        _, frho_n = structure_tidal(r * r_pl / frs, fvs, frs, m_pl, m_star, a)
        out['rates'] = {'photoionization': fin_phi, 'recombination': k2_abs * (frho_n * frhos) * f_r ** 2}
    return out


# ----------------------------------------------------------------------------
# helium
# ----------------------------------------------------------------------------
def he_radiative_processes(sed):
    """helium.py:24-154 -> phi_1, phi_3, a_1, a_3, a_h_1, a_h_3."""
    wl, fl = sed.wl, sed.flux
    e_ev = (_H_SI * ((_C_SI / wl) / 1.0 * (1.0 / 1e-10 / 1.0))) * (1.0 / _EV_J)
    e_erg = e_ev * (_EV_J / 1e-7)
    i1 = nearest_index(wl, 504)
    i0 = nearest_index(wl, 911.65)
    wl1, fl1, en1 = wl[:i1], fl[:i1], e_erg[:i1]
    wl0, fl0 = wl[:i0], fl[:i0]
    a_l1 = sigma_he_singlet(wl1)
    wl3, a_l3 = sigma_he_triplet_table()
    en3 = 1.98644586e-08 / wl3
    fl3 = np.interp(wl3, wl, fl)
    a_1 = abs(simpson(fl1 * a_l1, x=wl1) / simpson(fl1, x=wl1))
    a_3 = abs(simpson(fl3 * a_l3, x=wl3) / simpson(fl3, x=wl3))
    a_h_1 = abs(simpson(fl1 * sigma_h(wl1), x=wl1) / simpson(fl1, x=wl1))
    a_h_3 = abs(simpson(fl0 * sigma_h(wl0), x=wl0) / simpson(fl3, x=wl3))
    phi_1 = abs(simpson(fl1 * a_l1 / en1, x=wl1))
    phi_3 = abs(simpson(fl3 * a_l3 / en3, x=wl3))
    return phi_1, phi_3, a_1, a_3, a_h_1, a_h_3


def he_recombination(T):
    """helium.py:270-292."""
    return 1.54E-13 * (T / 1E4) ** (-0.486), 2.10E-13 * (T / 1E4) ** (-0.778)


def he_collision(T):
    """helium.py:317-380."""
    tt = 10 ** _HE_GAMMA[:, 0]
    g13 = np.interp(T, tt, _HE_GAMMA[:, 1])
    g31a = np.interp(T, tt, _HE_GAMMA[:, 2])
    g31b = np.interp(T, tt, _HE_GAMMA[:, 3])
    kt = 8.617333262145179e-05 * T
    k1 = 2.10E-8 * (13.6 / kt) ** 0.5
    return (k1 * g13 * np.exp(-19.81 / kt), k1 * g31a / 3 * np.exp(-0.80 / kt),
            k1 * g31b / 3 * np.exp(-1.40 / kt),
            1.75E-11 * (300 / T) ** 0.75 * np.exp(-128E3 / T),
            1.25E-15 * (300 / T) ** (-0.25))


def he_population(r, v, rho, f_h, r_pl, T, h_fraction, vs, rs, rhos, sed=None,
                  rates=None, initial_state=(0.5, 0.5), relax_solution=False,
                  convergence=0.01, max_n_relax=10, method='odeint', **ivp):
    """helium.py:413-785.  `rates` optionally carries the 6 SED scalars of
    `he_radiative_processes` (they are model independent).  Returns dict: f_1, f_3,
    status (0 ok, 2 = first-solve failure -> RuntimeError at :696/:709,
    3 = a relax-loop odeint warned, :736), n_relax, nfe (list).

    Status 3.  The reference does not check the relax loop's ``odeint`` (:736) and
    carries on with the result array, whose rows after the failing output point
    scipy leaves UNINITIALISED -- the unmodified reference differs from itself run
    to run on such models (tests/golden/he_failure_classes.npz, key
    ``f_1_run_spread``).  Only the class is defined: everything up to the first
    warning is deterministic.  This restatement (and the engine) stop there:
    status 3, ``n_relax`` = the pass that warned, profiles = those of the last
    pass that completed."""
    y0 = np.array(initial_state, dtype=float)
    rs_cm = rs * 7.1492E+09
    alpha_unit = (rs_cm ** 2 * vs * 1E5)
    a_r1, a_r3 = he_recombination(T)
    a_r1, a_r3 = a_r1 / alpha_unit, a_r3 / alpha_unit
    m_h = 1.67262192E-24 / (rhos * rs_cm ** 3)
    phi_unit = vs * 1E5 / rs / 7.1492E+09
    if rates is None:
        if sed is None:
            raise ValueError('Either `spectrum_at_planet` must be provided, or '
                             '`flux_euv` and `flux_fuv` must be provided.')
        rates = he_radiative_processes(sed)
    phi_1, phi_3, a_1, a_3, a_h_1, a_h_3 = rates
    phi_1, phi_3 = phi_1 / phi_unit, phi_3 / phi_unit
    a_1, a_3 = a_1 / rs_cm ** 2, a_3 / rs_cm ** 2
    a_h_1, a_h_3 = a_h_1 / rs_cm ** 2, a_h_3 / rs_cm ** 2
    q13, q31a, q31b, bq_he, bq_hep = (q / alpha_unit for q in he_collision(T))
    bq31 = 5E-10 / alpha_unit
    ba31 = 1.272E-4 / phi_unit
    rn = r * r_pl / rs
    dr = np.diff(rn)
    dr = np.concatenate((dr, np.array([dr[-1], ])))
    col = np.flip(np.cumsum(np.flip(dr * rho)))
    col_h0 = np.flip(np.cumsum(np.flip(dr * rho * (1 - f_h))))
    he_fraction = 1 - h_fraction
    k1 = h_fraction / (h_fraction + 4 * he_fraction) / m_h
    k2 = he_fraction / (h_fraction + 4 * he_fraction) / m_h
    tau_1_h = k1 * a_h_1 * col_h0
    tau_3_h = k1 * a_h_3 * col_h0
    tau = [y0[0] * k2 * a_1 * col + tau_1_h, y0[1] * k2 * a_3 * col + tau_3_h]

    def fun(x, y):
        f1, f3 = y[0], y[1]
        v_ = np.interp(x, rn, v)
        rho_ = np.interp(x, rn, rho)
        fh = np.interp(x, rn, f_h)
        lk = np.log(k1) + np.log(rho_)

        def ne(c):
            return np.exp(np.log(k1) + np.log(rho_) + np.log(c))
        ar1 = ne(a_r1) * fh
        ar3 = ne(a_r3) * fh
        c13 = ne(q13) * fh
        c31a = ne(q31a) * fh
        c31b = ne(q31b) * fh
        cq31 = ne(bq31) * (1 - fh)
        cqhe = ne(bq_he) * fh
        cqhep = ne(bq_hep) * (1 - fh)
        t11 = (1 - f1 - f3) * ar1
        t12 = f3 * ba31
        t1 = np.interp(x, rn, tau[0])
        t3 = np.interp(x, rn, tau[1])
        t13 = f1 * phi_1 * np.exp(-t1)
        t14 = f1 * c13
        t15 = f3 * c31a
        t16 = f3 * c31b
        t17 = f3 * cq31
        t18 = f1 * cqhe
        t19 = (1 - f1 - f3) * cqhep
        t31 = (1 - f1 - f3) * ar3
        t33 = f3 * phi_3 * np.exp(-t3)
        d1 = (t11 + t12 - t13 - t14 + t15 + t16 + t17 - t18 + t19) / v_
        d3 = (t31 - t12 - t33 + t14 - t15 - t16 - t17) / v_
        return np.array([d1, d3])

    out = {'status': 0, 'n_relax': 0, 'nfe': [], 'nst': [], 'nje': [],
           'f_1': None, 'f_3': None}

    def solve(first):
        if method == 'odeint':
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter('always')
                sol, info = odeint(fun, y0=y0, t=rn, tfirst=True, full_output=True)
            out['nfe'].append(int(info['nfe'][-1]))
            out['nst'].append(int(info['nst'][-1]))
            out['nje'].append(int(info['nje'][-1]))
            if len(w) > 0:
                return None if first else 'warned'
            return np.copy(sol).T[0], np.copy(sol).T[1]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            sol = solve_ivp(fun, (rn[0], rn[-1]), y0, t_eval=rn, method=method, **ivp)
        out['nfe'].append(sol.nfev)
        if len(sol['y'][0]) != len(rn):
            return None
        return sol['y'][0], sol['y'][1]

    def clamp(f):
        f[f < 0] = 1E-15
        f[f > 1.0] = 1.0
        return f

    s = solve(True)
    if s is None:
        out['status'] = 2
        return out
    f1, f3 = clamp(s[0]), clamp(s[1])
    if relax_solution:
        for _ in range(max_n_relax):
            p1, p3 = np.copy(f1), np.copy(f3)
            out['n_relax'] += 1
            tau[0] = k2 * a_1 * np.flip(np.cumsum(np.flip(dr * rho * f1))) + tau_1_h
            tau[1] = k2 * a_3 * np.flip(np.cumsum(np.flip(dr * rho * f3))) + tau_3_h
            s = solve(False)
            if s is None:
                out['status'] = 2
                return out
            if isinstance(s, str):  # odeint warned inside the relax loop: see the docstring
                out['status'] = 3
                f1, f3 = p1, p3
                break
            f1, f3 = clamp(s[0]), clamp(s[1])
            d1 = abs(np.sum(f1 - p1)) / np.sum(p1)
            d3 = abs(np.sum(f3 - p3)) / np.sum(p3)
            if d1 < convergence and d3 < convergence:
                break
    out['f_1'], out['f_3'] = f1, f3
    return out


# ----------------------------------------------------------------------------
# transit
# ----------------------------------------------------------------------------
def pil_ellipse_spans(x0, y0, x1, y1):
    """Rows and inclusive column spans of a filled ellipse exactly as Pillow's
    ``ImageDraw.ellipse(fill=...)`` rasterises it (integer Bresenham-like walk in
    half-pixel units over a quarter arc; Pillow `src/libImaging/Draw.c`
    quarter_next / ellipse_next).  `flatstar` (absent third-party dependency of
    transit.py:94-100) draws its disks with this call; pinned by the reference's
    golden transit depth (tests/test_transit.py:37).
    Arguments are the float bounding box; Pillow truncates them with (int).
    Returns (rows, col_lo, col_hi), unclipped."""
    x0, y0, x1, y1 = int(x0), int(y0), int(x1), int(y1)
    a, b = x1 - x0, y1 - y0
    if a < 0 or b < 0:
        return (np.zeros(0, int),) * 3
    a2, b2 = a * a, b * b
    a2b2 = a2 * b2

    def delta(x, y):
        return abs(a2 * y * y + b2 * x * x - a2b2)
    cx, cy, ex, ey = a, b % 2, a % 2, b
    first_x = {cy: cx}
    while not (cx == ex and cy == ey):
        nx, ny = cx, cy + 2
        nd = delta(nx, ny)
        if nx > 1:
            d2 = delta(cx - 2, cy + 2)
            if nd > d2:
                nx, ny, nd = cx - 2, cy + 2, d2
            d3 = delta(cx - 2, cy)
            if nd > d3:
                nx, ny = cx - 2, cy
        cx, cy = nx, ny
        if cy not in first_x:
            first_x[cy] = cx
    rows, lo, hi = [], [], []
    for y, xr in first_x.items():
        for yy in ((y, -y) if y > 0 else (y,)):
            rows.append(y0 + (yy + b) // 2)
            lo.append(x0 + (a - xr) // 2)
            hi.append(x0 + (a + xr) // 2)
    o = np.argsort(rows)
    return np.array(rows)[o], np.array(lo)[o], np.array(hi)[o]


def _disk(center, radius, n):
    m = np.zeros((n, n))
    rows, lo, hi = pil_ellipse_spans(center[0] - radius, center[1] - radius,
                                     center[0] + radius, center[1] + radius)
    for rr, l, h in zip(rows, lo, hi):
        if 0 <= rr < n:
            l, h = max(l, 0), min(h, n - 1)
            if h >= l:
                m[rr, l:h + 1] = 1.0
    return m


def limb_darkening(mu, law, c):
    """Standard laws named at transit.py:62-66 (flatstar convention, unpinned)."""
    if law is None:
        return np.ones_like(mu)
    if law == 'linear':
        return 1 - c * (1 - mu)
    c = np.asarray(c, dtype=float)
    if law == 'quadratic':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu) ** 2
    if law == 'square-root':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu ** 0.5)
    if law == 'log':
        with np.errstate(all='ignore'):
            t = np.where(mu > 0, mu * np.log(np.where(mu > 0, mu, 1.0)), 0.0)
        return 1 - c[0] * (1 - mu) - c[1] * t
    if law == 'exp':
        return 1 - c[0] * (1 - mu) - c[1] / (1 - np.exp(mu))
    if law == 's3':
        return 1 - c[0] * (1 - mu) - c[1] * (1 - mu ** 1.5) - c[2] * (1 - mu ** 2)
    if law == 'c4':
        return (1 - c[0] * (1 - mu ** 0.5) - c[1] * (1 - mu)
                - c[2] * (1 - mu ** 1.5) - c[3] * (1 - mu ** 2))
    raise ValueError('unknown limb-darkening law')


def draw_transit(k, r_phys, b=0.0, phase=0.0, grid_size=100, supersampling=None,
                 resample_method=None, limb_darkening_law=None, ld_coefficient=None):
    """transit.py:20-116 with flatstar restated (box resampling only)."""
    if resample_method not in (None, 'box'):
        raise NotImplementedError('only box resampling is restated')
    n = int(round(grid_size * supersampling)) if supersampling is not None else grid_size
    r_star = n * 0.5
    cen = n / 2
    mask = _disk((cen, cen), r_star, n)
    yy, xx = np.mgrid[0:n, 0:n]
    rho2 = ((xx - cen) ** 2 + (yy - cen) ** 2) / r_star ** 2
    mu = np.sqrt(np.clip(1 - rho2, 0.0, 1.0))
    inten = mask * limb_darkening(mu, limb_darkening_law, ld_coefficient)
    inten = inten / np.sum(inten)
    r_p = r_star * k
    x_p = n / 2 + 2 * phase * r_star
    y_p = n / 2 + b * r_star
    tr = inten * (1 - _disk((x_p, y_p), r_p, n))
    depth = 1 - np.sum(tr) / np.sum(inten)
    if supersampling is not None:
        f = 1 / supersampling
        m = int(round(n * f))
        ss = n // m
        tr = tr[:m * ss, :m * ss].reshape(m, ss, m, ss).sum(axis=(1, 3))
        x_p, y_p, r_p = x_p * f, y_p * f, r_p * f
    ny, nx = tr.shape
    yy, xx = np.mgrid[0:ny, 0:nx]
    r_map = np.sqrt((xx - x_p) ** 2 + (yy - y_p) ** 2) * (r_phys / r_p)
    return tr, depth, r_map


def profile_los(r, n, v, nz, T=None):
    """transit.py:234-307."""
    z = np.linspace(-r[-1], r[-1], nz)
    coords = np.array(np.meshgrid(r, z))
    d = np.sum(coords ** 2, axis=0) ** 0.5
    vv = interp1d(r, v, bounds_error=False, fill_value=0.0)(d)
    ze = np.expand_dims(z, axis=1)
    v_los = vv * ze / (ze ** 2 + r ** 2) ** 0.5
    n_los = interp1d(r, n, bounds_error=False, fill_value=0.0)(d)
    if T is not None:
        return n_los, v_los, interp1d(r, T, bounds_error=False, fill_value=0.0)(d), z
    return n_los, v_los, z


def optical_depth_2d(r, n, v, w0, f, a_ij, wl, T, mass, nz=200, v_bulk=0.0,
                     rv_planet=0.0, method='average', turbulence=False):
    """transit.py:312-536 -> tau[Nr, Nw]."""
    if isinstance(T, (float, int)):
        n_los, v_los, z = profile_los(r, n, v, nz)
        temp = T
    elif isinstance(T, np.ndarray):
        n_los, v_los, temp, z = profile_los(r, n, v, nz, T)
    else:
        raise ValueError('`gas_temperature` has to be either `float` or '
                         '`np.ndarray`.')
    shape = np.shape(n_los)
    w0, f, a_ij = (np.atleast_1d(np.asarray(q, dtype=float)) for q in (w0, f, a_ij))
    c, k_b = 2.99792458e+08, 1.380649e-23
    nu0 = c / w0
    nu_rest = np.reshape(c / wl, (len(wl), 1))
    dnu = (nu_rest - nu0).T
    sigma = 2.654008854574474e-06 * f
    tau = []
    for i in range(len(w0)):
        gamma = a_ij[i] / 4 / np.pi
        if method == 'formal':
            alpha = nu0[i] / c * (k_b * temp / mass) ** 0.5
            if isinstance(T, np.ndarray):
                alpha = np.reshape(alpha, shape + (1,))
            add = np.reshape((v_los + v_bulk + rv_planet) / c * nu0[i], shape + (1,))
        elif method == 'average':
            vw = (np.sum(v_los ** 2 * n_los) / np.sum(n_los)) ** 0.5
            if isinstance(temp, np.ndarray):
                t_avg = np.sum(temp * n_los) / np.sum(n_los)
            else:
                t_avg = temp
            vt = np.sqrt(5 / 6 * k_b * t_avg / mass) if turbulence is True else 0.0
            alpha = nu0[i] / c * (k_b * t_avg / mass + vw ** 2 + vt ** 2) ** 0.5
            add = (v_bulk + rv_planet) / c * nu0[i]
        else:
            raise ValueError('The chosen ``wind_broadening_method`` is not '
                             'implemented.')
        prof = voigt_profile(dnu[i] + add, alpha, gamma)
        tau.append(trapezoid(prof * np.expand_dims(n_los, axis=-1), z, axis=0)
                   * sigma[i])
    return np.sum(np.array(tau), axis=0)


def radiative_transfer_2d(i0, r_map, r, n, v, w0, f, a_ij, wl, T, mass,
                          v_bulk=0.0, rv_planet=0.0, method='average', nz=200,
                          turbulence=False):
    """transit.py:120-230 -> spectrum[Nw]."""
    tau = optical_depth_2d(r, n, v, w0, f, a_ij, wl, T, mass, nz, v_bulk,
                           rv_planet, method, turbulence)
    t = interp1d(r, tau, axis=0, bounds_error=False, fill_value=0.0)(r_map)
    i0 = np.reshape(i0, np.shape(i0) + (1,))
    return np.sum(i0 * np.exp(-t), axis=(0, 1))


# ----------------------------------------------------------------------------
# helium.ion_fraction: helium.py:789-1032
# ----------------------------------------------------------------------------
def he_radiative_processes_combined(sed):
    """helium.radiative_processes(..., combined_ionization=True), helium.py:160-177 -> phi, a_he, a_h."""
    wl, flux = sed.wl, sed.flux
    energy = (_H_SI * ((_C_SI / wl) / 1.0 * (1.0 / 1e-10 / 1.0))) * (1.0 / _EV_J)
    energy_erg = energy * (_EV_J / 1e-7)
    i1 = nearest_index(wl, 504)
    w1, f1, e1 = wl[:i1], flux[:i1], energy_erg[:i1]
    a_l = sigma_he_total(w1)
    a_he = abs(simpson(f1 * a_l, x=w1) / simpson(f1, x=w1))
    a_h = abs(simpson(f1 * sigma_h(w1), x=w1) / simpson(f1, x=w1))
    phi = abs(simpson(f1 * a_l / e1, x=w1))
    return phi, a_he, a_h


def he_recombination_all(T):
    """helium.py:294-313."""
    return 4.6E-12 * (T / 300) ** (-0.64)


def he_charge_transfer(T):
    """helium.py:384-409 -> (He + H+ -> He+, He+ + H -> He)."""
    ct_hep_h = 1.25E-15 * (300 / T) ** (-0.25)
    ct_he_hp = 1.75E-11 * (300 / T) ** 0.75 * np.exp(-128000 / T)
    return ct_he_hp, ct_hep_h


def he_ion_fraction(r, v, rho, f_h, r_pl, T, h_fraction, vs, rs, rhos, sed=None, rates=None,
                    initial_f_he_ion=0.0, relax_solution=False, convergence=0.01, max_n_relax=10,
                    method='Radau', **options):
    """helium.py:789-1032 -> dict(f_he_ion, status, n_relax, nfev)."""
    alpha_unit = ((rs * 7.1492E+09) ** 2 * vs * 1E5)
    alpha_rec = he_recombination_all(T) / alpha_unit
    m_h = 1.67262192E-24 / (rhos * (rs * 7.1492E+09) ** 3)
    phi_unit = vs * 1E5 / rs / 7.1492E+09
    phi_he, a_he, a_h = rates if rates is not None else he_radiative_processes_combined(sed)
    phi_he = phi_he / phi_unit
    a_he = a_he / (rs * 7.1492E+09) ** 2
    a_h = a_h / (rs * 7.1492E+09) ** 2
    ct_he_hp, ct_hep_h = he_charge_transfer(T)
    ct_hep_h = ct_hep_h / alpha_unit
    ct_he_hp = ct_he_hp / alpha_unit
    rn = r * r_pl / rs
    dr = np.diff(rn)
    dr = np.concatenate((dr, np.array([dr[-1], ])))
    col = np.flip(np.cumsum(np.flip(dr * rho)))
    col_h0 = np.flip(np.cumsum(np.flip(dr * rho * (1 - f_h))))
    he_fraction = 1 - h_fraction
    k1 = h_fraction / (h_fraction + 4 * he_fraction) / m_h
    k2 = he_fraction / (h_fraction + 4 * he_fraction) / m_h
    tau_h = k1 * a_h * col_h0
    st = {'tau': (initial_f_he_ion * k2 * a_he * col + tau_h)}

    def fun(x, y):
        f = y
        _v = np.interp(x, rn, v)
        _rho = np.interp(x, rn, rho)
        fh = np.interp(x, rn, f_h)
        lt1 = np.log(k1) + np.log(_rho)
        lt2 = np.log(k2) + np.log(_rho)
        ct1 = np.exp(lt1 + np.log(ct_he_hp)) * fh
        rec = np.exp(lt1 + np.log(alpha_rec)) * fh + np.exp(lt2 + np.log(alpha_rec)) * f
        ct2 = np.exp(lt1 + np.log(ct_hep_h)) * (1 - fh)
        t_he = np.interp(x, rn, st['tau'])
        term1 = (1 - f) * phi_he * np.exp(-t_he)
        term2 = (1 - f) * ct1
        term3 = f * rec
        term4 = f * ct2
        return (term1 + term2 - term3 - term4) / _v

    out = {'status': 0, 'n_relax': 0, 'nfev': []}
    nan = np.full(len(rn), np.nan)

    def solve(first):
        if method == 'odeint':
            sol = _odeint_relax(fun, initial_f_he_ion, rn)
            if sol is None:
                return None if first else 'warned'
            return sol.T[0]
        sol = solve_ivp(fun, (rn[0], rn[-1],), np.array([initial_f_he_ion, ]), t_eval=rn, method=method,
                        **options)
        out['nfev'].append(sol.nfev)
        return sol['y'][0] if len(sol['y'][0]) == len(rn) else None

    f = solve(True)
    if f is None:
        out.update(status=2, f_he_ion=nan)
        return out
    if relax_solution is True:     # (no clamp after the first solve, helium.py:985-997)
        for i in range(max_n_relax):
            prev = np.copy(f)
            st['tau'] = k2 * a_he * np.flip(np.cumsum(np.flip(dr * rho * (1 - f)))) + tau_h
            f = solve(False)
            if f is None:
                out.update(status=2, f_he_ion=nan)
                return out
            out['n_relax'] = i + 1
            if isinstance(f, str):      # odeint warned: stop, last completed pass (status 3)
                out['status'] = 3
                f = prev
                break
            f[f < 0] = 1E-15
            f[f > 1.0] = 1.0
            if abs(np.sum(f - prev)) / np.sum(prev) < convergence:
                break
    out['f_he_ion'] = f
    return out


def _odeint_relax(fun, y0, rn):
    """`odeint` as the relax loops of helium.ion_fraction / oxygen / carbon call it (unchecked).  Returns None
    when it warned: the reference carries on with rows scipy left uninitialised, i.e. its result is undefined
    from there on; the oracle (and the engine) stop at the warning -- status 3, last completed pass (see
    he_population)."""
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        sol = odeint(fun, y0=y0, t=rn, tfirst=True)
    return None if len(w) > 0 else np.copy(sol)


# ----------------------------------------------------------------------------
# Oxygen: oxygen.py, microphysics.general_cross_section
# ----------------------------------------------------------------------------
# microphysics.sigma_properties_v1996 (microphysics.py:300-318), Verner et al. 1996:
# E_threshold, E_max, E_0, sigma_0, y_a, p, y_w, y_0, y_1
_VERNER = {'C I': [1.126E1, 2.910E2, 2.144E0, 5.027E2, 6.126E1, 5.101E0, 9.157E-2, 1.133E0, 1.607E0],
           'C II': [2.438E1, 3.076E2, 4.058E-1, 8.709E0, 1.261E2, 8.578E0, 2.093E0, 4.929E1, 3.234E0],
           'O I': [1.362E1, 5.380E2, 1.240E0, 1.745E3, 3.784E0, 1.764E1, 7.589E-2, 8.698E0, 1.271E-1]}
SOLAR_OXYGEN_FRACTION = 10 ** (8.69 - 12.00)      # oxygen.py:21-22


def general_cross_section(wl, species):
    """microphysics.py:224-281 with the unit algebra of the astropy stand-in written out:
    energy = (c.h * c.c / wavelength / u.angstrom).to(u.eV).value."""
    wl = np.asarray(wl, dtype=float)
    energy = (_H_SI * _C_SI / wl) * ((1.0 / 1e-10) / _EV_J)
    e_thr, e_max, e_0, sigma_0, y_a, p, y_w, y_0, y_1 = _VERNER[species]
    x = energy / e_0 - y_0
    y = (x ** 2 + y_1 ** 2) ** 0.5
    term1 = (x - 1) ** 2 + y_w ** 2
    term2 = y ** (0.5 * p - 5.5)
    term3 = (1 + (y / y_a) ** 0.5) ** (-p)
    cs = sigma_0 * (term1 * term2 * term3) * 1E-18
    cs = np.array(cs, dtype=float)
    cs[energy < e_thr] = 0.0
    return cs


def o_radiative_processes(sed):
    """oxygen.py:26-99 -> phi_oi, a_oi, a_h_oi, a_he."""
    wl, flux = sed.wl, sed.flux
    energy = (_H_SI * ((_C_SI / wl) / 1.0 * (1.0 / 1e-10 / 1.0))) * (1.0 / _EV_J)
    energy_erg = energy * (_EV_J / 1e-7)
    wl_break_oi = 12398.42 / _VERNER['O I'][0]
    i0 = nearest_index(wl, 504)
    i1 = nearest_index(wl, wl_break_oi)
    w0, f0 = wl[:i0], flux[:i0]
    w1, f1, e1 = wl[:i1], flux[:i1], energy_erg[:i1]
    a_l = general_cross_section(w1, 'O I')
    a_oi = abs(simpson(f1 * a_l, x=w1) / simpson(f1, x=w1))
    a_h_oi = abs(simpson(f1 * sigma_h(w1), x=w1) / simpson(f1, x=w1))
    a_he = abs(simpson(f0 * sigma_he_total(w0), x=w0) / simpson(f0, x=w0))
    phi_oi = abs(simpson(f1 * a_l / e1, x=w1))
    return phi_oi, a_oi, a_h_oi, a_he


def o_electron_impact_ionization(T):
    """oxygen.py:103-126 (Voronov 1997)."""
    ratio = 11.3 / (8.617333262145179e-05 * T)
    return 3.59E-8 * (0.073 + ratio) ** (-1) * ratio ** 0.34 * np.exp(-ratio)


def o_recombination(T):
    """oxygen.py:130-149."""
    return 3.25E-12 * (300 / T) ** 0.66


def o_charge_transfer(T):
    """oxygen.py:153-182 -> (O I + H+ -> O II, O II + H -> O I)."""
    ct_oii_h = 5.66E-10 * (300 / T) ** (-0.36) * np.exp(8.6 / T)
    ct_oi_hp = 7.31E-10 * (300 / T) ** (-0.23) * np.exp(-226 / T)
    return ct_oi_hp, ct_oii_h


def o_ion_fraction(r, v, rho, f_h, f_he_ion, r_pl, T, h_fraction, vs, rs, rhos, sed=None, rates=None,
                   o_fraction=SOLAR_OXYGEN_FRACTION, initial_f_o_ion=0.0, relax_solution=False,
                   convergence=0.01, max_n_relax=10, method='Radau', **options):
    """oxygen.py:186-491 -> dict(f_o=..., status=0 ok / 2 solver failed, n_relax, nfev)."""
    alpha_unit = ((rs * 7.1492E+09) ** 2 * vs * 1E5)
    alpha_rec = o_recombination(T) / alpha_unit
    m_h = 1.67262192E-24 / (rhos * (rs * 7.1492E+09) ** 3)
    phi_unit = vs * 1E5 / rs / 7.1492E+09
    phi_oi, a_oi, a_h_oi, a_he = rates if rates is not None else o_radiative_processes(sed)
    phi_oi = phi_oi / phi_unit
    a_oi = a_oi / (rs * 7.1492E+09) ** 2
    a_h_oi = a_h_oi / (rs * 7.1492E+09) ** 2
    a_he = a_he / (rs * 7.1492E+09) ** 2
    ion_rate = o_electron_impact_ionization(T) / alpha_unit
    ct_oi_hp, ct_oii_h = o_charge_transfer(T)
    ct_oii_h = ct_oii_h / alpha_unit
    ct_oi_hp = ct_oi_hp / alpha_unit
    rn = r * r_pl / rs
    dr = np.diff(rn)
    dr = np.concatenate((dr, np.array([dr[-1], ])))
    col = np.flip(np.cumsum(np.flip(dr * rho)))
    col_h0 = np.flip(np.cumsum(np.flip(dr * rho * (1 - f_h))))
    he_fraction = 1 - h_fraction
    col_he0 = np.flip(np.cumsum(np.flip(dr * rho * he_fraction * (1 - f_he_ion))))
    den = h_fraction + 4 * he_fraction + 16 * o_fraction
    k1, k2, k3 = h_fraction / den / m_h, he_fraction / den / m_h, o_fraction / den / m_h
    tau_h = k1 * a_h_oi * col_h0
    tau_he = k2 * a_he * col_he0
    state = {'tau': (1 - initial_f_o_ion) * k3 * a_oi * col + tau_h + tau_he}

    def fun(x, y):
        f_oii = y
        _v = np.interp(x, rn, v)
        _rho = np.interp(x, rn, rho)
        fh = np.interp(x, rn, f_h)
        fhe = np.interp(x, rn, f_he_ion)
        lt1 = np.log(k1) + np.log(_rho)
        lt2 = np.log(k2) + np.log(_rho)
        ion_ne = np.exp(lt1 + np.log(ion_rate)) * fh + np.exp(lt2 + np.log(ion_rate)) * fhe
        ct1 = np.exp(lt1 + np.log(ct_oi_hp)) * fh
        rec_ne = np.exp(lt1 + np.log(alpha_rec)) * fh + np.exp(lt2 + np.log(alpha_rec)) * fhe
        ct2 = np.exp(lt1 + np.log(ct_oii_h)) * (1 - fh)
        t_oi = np.interp(x, rn, state['tau'])
        term1 = (1 - f_oii) * phi_oi * np.exp(-t_oi)
        term2 = (1 - f_oii) * ion_ne
        term3 = (1 - f_oii) * ct1
        term4 = f_oii * rec_ne
        term5 = f_oii * ct2
        return (term1 + term2 + term3 - term4 - term5) / _v

    out = {'status': 0, 'n_relax': 0, 'nfev': []}

    def solve():
        if method == 'odeint':
            with warnings.catch_warnings():
                warnings.filterwarnings('error')
                try:
                    sol = odeint(fun, y0=initial_f_o_ion, t=rn, tfirst=True)
                except Warning:
                    return None
            return np.copy(sol).T[0]
        sol = solve_ivp(fun, (rn[0], rn[-1],), np.array([initial_f_o_ion, ]), t_eval=rn, method=method,
                        **options)
        out['nfev'].append(sol.nfev)
        f = sol['y'][0]
        return f if len(f) == len(rn) else None

    f = solve()
    if f is None:
        out['status'] = 2
        out['f_o'] = np.full(len(rn), np.nan)
        return out
    f[f < 0] = 1E-15
    f[f > 1.0] = 1.0
    if relax_solution is True:
        for i in range(max_n_relax):
            prev = np.copy(f)
            state['tau'] = k3 * a_oi * np.flip(np.cumsum(np.flip(dr * rho * (1 - f)))) + tau_h + tau_he
            if method == 'odeint':
                sol = _odeint_relax(fun, initial_f_o_ion, rn)
                if sol is None:          # odeint warned: stop, last completed pass (status 3)
                    out['status'] = 3
                    out['n_relax'] = i + 1
                    f = prev
                    break
                f = sol.T[0]
            else:
                sol = solve_ivp(fun, (rn[0], rn[-1],), np.array([initial_f_o_ion, ]), t_eval=rn,
                                method=method, **options)
                out['nfev'].append(sol.nfev)
                f = sol['y'][0]
                if len(f) != len(rn):     # the reference would crash on the next line (shape mismatch)
                    out['status'] = 2
                    out['f_o'] = np.full(len(rn), np.nan)
                    return out
            f[f < 0] = 1E-15
            f[f > 1.0] = 1.0
            out['n_relax'] = i + 1
            if abs(np.sum(f - prev)) / np.sum(prev) < convergence:
                break
    out['f_o'] = f
    return out


# ----------------------------------------------------------------------------
# Carbon: carbon.py
# ----------------------------------------------------------------------------
SOLAR_CARBON_FRACTION = 10 ** (8.43 - 12.00)      # carbon.py:21-22


def c_radiative_processes(sed):
    """carbon.py:30-131 -> phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he."""
    wl, flux = sed.wl, sed.flux
    energy = (_H_SI * ((_C_SI / wl) / 1.0 * (1.0 / 1e-10 / 1.0))) * (1.0 / _EV_J)
    energy_erg = energy * (_EV_J / 1e-7)
    i0 = nearest_index(wl, 504)
    i1 = nearest_index(wl, 12398.42 / _VERNER['C I'][0])
    i2 = nearest_index(wl, 12398.42 / _VERNER['C II'][0])
    w0, f0 = wl[:i0], flux[:i0]
    w1, f1, e1 = wl[:i1], flux[:i1], energy_erg[:i1]
    w2, f2, e2 = wl[:i2], flux[:i2], energy_erg[:i2]
    a_l1 = general_cross_section(w1, 'C I')
    a_l2 = general_cross_section(w2, 'C II')
    a_ci = abs(simpson(f1 * a_l1, x=w1) / simpson(f1, x=w1))
    a_cii = abs(simpson(f2 * a_l2, x=w2) / simpson(f2, x=w2))
    a_h_ci = abs(simpson(f1 * sigma_h(w1), x=w1) / simpson(f1, x=w1))
    a_h_cii = abs(simpson(f2 * sigma_h(w2), x=w2) / simpson(f2, x=w2))
    a_he = abs(simpson(f0 * sigma_he_total(w0), x=w0) / simpson(f0, x=w0))
    phi_ci = abs(simpson(f1 * a_l1 / e1, x=w1))
    phi_cii = abs(simpson(f2 * a_l2 / e2, x=w2))
    return phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he


def c_electron_impact_ionization(T):
    """carbon.py:135-165 (Voronov 1997)."""
    e = 8.617333262145179e-05 * T
    r1, r2 = 11.3 / e, 24.4 / e
    return (6.85E-8 * (0.193 + r1) ** (-1) * r1 ** 0.25 * np.exp(-r1),
            1.86E-8 * (0.286 + r2) ** (-1) * r2 ** 0.24 * np.exp(-r2))


def c_recombination(T):
    """carbon.py:169-194."""
    return 4.67E-12 * (300 / T) ** 0.60, 2.3E-12 * (1000 / T) ** 0.645


def c_charge_transfer(T):
    """carbon.py:198-252 -> ct_ci_hp, ct_cii_h, ct_ci_hep, ct_cii_sii, ct_ciii_h, ct_ciii_he."""
    ct_cii_h = 6.30E-17 * (300 / T) ** (-1.96) * np.exp(-1.7E5 / T)
    ct_cii_sii = 2.1E-9
    ct_ci_hp = 1.31E-15 * (300 / T) ** (-0.213)
    ct_ci_hep = 2.5E-15 * (300 / T) ** (-1.597)
    ct_ciii_h = 1.67E-4 * (T / 10000) ** 2.79 * (1 + 304.72 * np.exp(-4.07 * T / 10000))
    ct_ciii_he = 1.23E-9
    return ct_ci_hp, ct_cii_h, ct_ci_hep, ct_cii_sii, ct_ciii_h, ct_ciii_he


def c_ion_fraction(r, v, rho, f_h, f_he_ion, r_pl, T, h_fraction, vs, rs, rhos, sed=None, rates=None,
                   c_fraction=SOLAR_CARBON_FRACTION, initial_f_c_ion=np.array([0.0, 0.0]),
                   relax_solution=False, convergence=0.01, max_n_relax=10, method='odeint', **options):
    """carbon.py:256-639 -> dict(f_cii, f_ciii, status=0 ok / 2 solver failed, n_relax, nfev)."""
    alpha_unit = ((rs * 7.1492E+09) ** 2 * vs * 1E5)
    rec_ci, rec_cii = c_recombination(T)
    rec_ci, rec_cii = rec_ci / alpha_unit, rec_cii / alpha_unit
    m_h = 1.67262192E-24 / (rhos * (rs * 7.1492E+09) ** 3)
    phi_unit = vs * 1E5 / rs / 7.1492E+09
    phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he = rates if rates is not None else c_radiative_processes(sed)
    phi_ci, phi_cii = phi_ci / phi_unit, phi_cii / phi_unit
    a_ci = a_ci / (rs * 7.1492E+09) ** 2
    a_h_ci = a_h_ci / (rs * 7.1492E+09) ** 2
    a_cii = a_cii / (rs * 7.1492E+09) ** 2
    a_h_cii = a_h_cii / (rs * 7.1492E+09) ** 2
    a_he = a_he / (rs * 7.1492E+09) ** 2
    ion_ci, ion_cii = c_electron_impact_ionization(T)
    ion_ci, ion_cii = ion_ci / alpha_unit, ion_cii / alpha_unit
    ct_ci_hp, ct_cii_h, ct_ci_hep, _, ct_ciii_h, ct_ciii_he = c_charge_transfer(T)
    ct_cii_h, ct_ci_hp, ct_ci_hep = ct_cii_h / alpha_unit, ct_ci_hp / alpha_unit, ct_ci_hep / alpha_unit
    ct_ciii_h, ct_ciii_he = ct_ciii_h / alpha_unit, ct_ciii_he / alpha_unit
    rn = r * r_pl / rs
    dr = np.diff(rn)
    dr = np.concatenate((dr, np.array([dr[-1], ])))
    col = np.flip(np.cumsum(np.flip(dr * rho)))
    col_h0 = np.flip(np.cumsum(np.flip(dr * rho * (1 - f_h))))
    he_fraction = 1 - h_fraction
    col_he0 = np.flip(np.cumsum(np.flip(dr * rho * he_fraction * (1 - f_he_ion))))
    den = h_fraction + 4 * he_fraction + 12 * c_fraction
    k1, k2, k3 = h_fraction / den / m_h, he_fraction / den / m_h, c_fraction / den / m_h
    tau_ci_h = k1 * a_h_ci * col_h0
    tau_cii_h = k1 * a_h_cii * col_h0
    tau_c_he = k2 * a_he * col_he0
    y0 = np.asarray(initial_f_c_ion, dtype=float)
    st = {'t1': (1 - y0[0] - y0[1]) * k3 * a_ci * col + tau_ci_h + tau_c_he,
          't2': y0[0] * k3 * a_cii * col + tau_cii_h + tau_c_he}

    def fun(x, y):
        f2, f3 = y[0], y[1]
        _v = np.interp(x, rn, v)
        _rho = np.interp(x, rn, rho)
        fh = np.interp(x, rn, f_h)
        fhe = np.interp(x, rn, f_he_ion)
        lt1 = np.log(k1) + np.log(_rho)
        lt2 = np.log(k2) + np.log(_rho)
        ion1 = np.exp(lt1 + np.log(ion_ci)) * fh + np.exp(lt2 + np.log(ion_ci)) * fhe
        cthp = np.exp(lt1 + np.log(ct_ci_hp)) * fh
        cthep = np.exp(lt2 + np.log(ct_ci_hep)) * fhe
        rec1 = np.exp(lt1 + np.log(rec_ci)) * fh + np.exp(lt2 + np.log(rec_ci)) * fhe
        ct2h = np.exp(lt1 + np.log(ct_cii_h)) * (1 - fh)
        rec2 = np.exp(lt1 + np.log(rec_cii)) * fh + np.exp(lt2 + np.log(rec_cii)) * fhe
        ct3h = np.exp(lt1 + np.log(ct_ciii_h)) * (1 - fh)
        ct3he = np.exp(lt2 + np.log(ct_ciii_he)) * (1 - fhe)
        ion2 = np.exp(lt1 + np.log(ion_cii)) * fh + np.exp(lt2 + np.log(ion_cii)) * fhe
        t_ci = np.interp(x, rn, st['t1'])
        t11 = (1 - f2 - f3) * phi_ci * np.exp(-t_ci)
        t12 = (1 - f2 - f3) * ion1
        t13 = (1 - f2 - f3) * cthp
        t14 = (1 - f2 - f3) * cthep
        t15 = f2 * rec1
        t16 = f2 * ct2h
        t17 = f3 * rec2
        t18 = f3 * ct3h
        t19 = f3 * ct3he
        t_cii = np.interp(x, rn, st['t2'])
        t21 = f2 * phi_cii * np.exp(-t_cii)
        t22 = f2 * ion2
        d2 = (t11 + t12 + t13 + t14 - t15 - t16 + t17 + t18 + t19 - t21 - t22) / _v
        d3 = (t21 + t22 - t17 - t18 - t19) / _v
        return np.array([d2, d3])

    out = {'status': 0, 'n_relax': 0, 'nfev': []}
    nan = np.full(len(rn), np.nan)

    def solve(first):
        if method == 'odeint':
            sol = _odeint_relax(fun, y0, rn)
            if sol is None:
                return None if first else 'warned'
            return sol.T[0].copy(), sol.T[1].copy()
        sol = solve_ivp(fun, (rn[0], rn[-1],), y0, t_eval=rn, method=method, **options)
        out['nfev'].append(sol.nfev)
        if len(sol['y'][0]) != len(rn):
            return None
        return sol['y'][0], sol['y'][1]

    def clamp(a):
        a[a < 0] = 1E-15
        a[a > 1.0] = 1.0

    res = solve(True)
    if res is None:
        out.update(status=2, f_cii=nan, f_ciii=nan.copy())
        return out
    f2, f3 = res
    clamp(f2)
    clamp(f3)
    if relax_solution is True:
        for i in range(max_n_relax):
            p2, p3 = np.copy(f2), np.copy(f3)
            st['t1'] = k3 * a_ci * np.flip(np.cumsum(np.flip(dr * rho * (1 - f2 - f3)))) + tau_ci_h + tau_c_he
            st['t2'] = k3 * a_cii * np.flip(np.cumsum(np.flip(dr * rho * f2))) + tau_cii_h + tau_c_he
            res = solve(False)
            if res is None:          # the reference would crash on the next lines (shape mismatch)
                out.update(status=2, f_cii=nan, f_ciii=nan.copy())
                return out
            if isinstance(res, str):  # odeint warned: stop, last completed pass (status 3)
                out['status'] = 3
                out['n_relax'] = i + 1
                f2, f3 = p2, p3
                break
            f2, f3 = res
            clamp(f2)
            clamp(f3)
            out['n_relax'] = i + 1
            with np.errstate(all='ignore'):
                d2 = abs(np.sum(f2 - p2)) / np.sum(p2)
                d3 = abs(np.sum(f3 - p3)) / np.sum(p3)
            if d2 < convergence and d3 < convergence:
                break
    out.update(f_cii=f2, f_ciii=f3)
    return out


# ----------------------------------------------------------------------------
# The retrieval loop around the path (docs/source/advanced_tutorial.ipynb cells
# 11, 13, 15): phase-averaged transmission spectrum, instrumental broadening,
# log-likelihood.  Notebook code, restated; `astropy.convolution` is absent from
# this image, so `convolve_extend` restates astropy's direct 1-D convolution
# (convolve.py + src/convolve.c: kernel divided by its sum, array padded with its
# edge values by half the kernel width, out[i] = sum_ii pad[i + ii] * g[nk-1-ii]):
# PARITY UNPINNED against astropy itself; tests hold it to numpy.convolve on the
# edge-padded array, which is the same mathematical definition.
# ----------------------------------------------------------------------------
def transmission_model(maps, r_si, n_si, v_si, w0, f, a_ij, wl, T, mass, v_wind=0.0):
    """advanced_tutorial.ipynb cell 11: `maps` = [(flux_map, r_map, transit_depth), ...]
    for the sampled phases -> mean over phases of (in-transit spectrum + depth)."""
    spectra = []
    for i0, r_map, depth in maps:
        spec = radiative_transfer_2d(i0, r_map, r_si, n_si, v_si, w0, f, a_ij, wl, T, mass,
                                     v_bulk=v_wind, method='average')
        spectra.append(spec + depth)
    return np.mean(np.array(spectra), axis=0)


def convolve_extend(array, kernel):
    """astropy.convolution.convolve(array, kernel, boundary='extend') for 1-D inputs
    without NaNs (normalize_kernel=True)."""
    array = np.asarray(array, dtype=float)
    kernel = np.asarray(kernel, dtype=float)
    if kernel.size % 2 == 0:
        raise ValueError('Kernel size must be odd in all axes.')
    g = kernel / kernel.sum()
    nk, w = g.size, g.size // 2
    pad = np.pad(array, w, mode='edge')
    out = np.zeros_like(array)
    for i in range(array.size):
        top = 0.0
        for ii in range(nk):
            top += pad[i + ii] * g[nk - 1 - ii]
        out[i] = top
    return out


def log_likelihood(model, y, yerr):
    """advanced_tutorial.ipynb cell 15 (the RuntimeError -> -inf branch is the caller's)."""
    sigma2 = yerr ** 2
    return -0.5 * np.sum((y - model) ** 2 / sigma2 + np.log(sigma2))


# ----------------------------------------------------------------------------
# One forward model = the docs' cascading_model
# (docs/source/quickstart.ipynb cells 4-12, advanced_tutorial.ipynb:226-345)
# ----------------------------------------------------------------------------
class Setup:
    """Everything that is fixed over a retrieval grid."""

    def __init__(self, sed, r=None, r_pl=1.39, m_pl=0.73, tidal=False,
                 m_star=1.07, a=0.04634, i0=None, r_map=None, lines=None, wl=None,
                 mass=4 * 1.67262192369e-27, v_bulk=0.0, nz=200,
                 h_tol=None, he_method='odeint', he_tol=None, species='he3'):
        self.sed = sed
        self.r = np.logspace(0, np.log10(20), 100) if r is None else r
        self.r_pl, self.m_pl = r_pl, m_pl
        self.tidal, self.m_star, self.a = tidal, m_star, a
        self.i0, self.r_map = i0, r_map
        self.lines = he3_lines() if lines is None else lines
        self.wl = np.linspace(1.0827, 1.0832, 200) * 1E-6 if wl is None else wl
        self.mass, self.v_bulk, self.nz = mass, v_bulk, nz
        self.h_tol = {} if h_tol is None else h_tol
        self.he_method = he_method
        self.he_tol = {} if he_tol is None else he_tol
        self.species = species      # 'he3' (He 10830) or 'h1' (Ly-alpha, cfg 2)
        self.he_rates = he_radiative_processes(sed)


def forward_model(s, T0, mdot, h_fraction):
    """One (T0, Mdot, h_fraction) model.  Returns dict with status
    (0 ok / 1 H solver failed / 2 He solver failed / 3 He relax warning)."""
    he_h = (1 - h_fraction) / h_fraction
    mu_0 = (1 + 4 * he_h) / (1 + he_h + 0.0)
    star = (s.m_star, s.a) if s.tidal else (1.0, 1.0)
    out = {'status': 0, 'spectrum': None}
    h = h_ion_fraction(s.r, s.r_pl, T0, h_fraction, mdot, s.m_pl, mu_0, star[0],
                       star[1], sed=s.sed, initial_f_ion=0.0, relax_solution=True,
                       exact_phi=True, **s.h_tol)
    out['h'] = h
    if h['status'] != 0:
        out['status'] = 1
        return out
    f_r, mu_bar = h['f_r'], h['mu_bar']
    vs = sound_speed(T0, mu_bar)
    if s.tidal:
        rs = radius_sonic_point_tidal(s.m_pl, vs, s.m_star, s.a)
        rhos = density_sonic_point(mdot, rs, vs)
        v, rho = structure_tidal(s.r * s.r_pl / rs, vs, rs, s.m_pl, s.m_star, s.a)
    else:
        rs = radius_sonic_point(s.m_pl, vs)
        rhos = density_sonic_point(mdot, rs, vs)
        v, rho = structure(s.r * s.r_pl / rs)
    out.update(vs=vs, rs=rs, rhos=rhos, v=v, rho=rho, f_r=f_r, mu_bar=mu_bar)
    r_si = s.r * (s.r_pl * 71492000)
    v_si = v * vs * 1000
    if s.species == 'h1':
        # cfg 2: neutral H number density, m^-3
        n = (1 - f_r) * rho * rhos * h_fraction / (h_fraction + 4 * (1 - h_fraction)) \
            / 1.67262192369e-24 * 1E6
    else:
        he = he_population(s.r, v, rho, f_r, s.r_pl, T0, h_fraction, vs, rs, rhos,
                           rates=s.he_rates, initial_state=np.array([1.0, 0.0]),
                           relax_solution=True, method=s.he_method, **s.he_tol)
        out['he'] = he
        if he['status'] == 2:
            out['status'] = 2
            return out
        out['status'] = he['status']
        out['f_1'], out['f_3'] = he['f_1'], he['f_3']
        he_fraction = 1 - h_fraction
        n_he = rho * rhos * he_fraction / (h_fraction + 4 * he_fraction) / 1.67262192369e-24
        n = he['f_3'] * n_he * 1E6
    out['n'] = n
    if s.i0 is not None:
        w0, f, a_ij = s.lines
        out['spectrum'] = radiative_transfer_2d(
            s.i0, s.r_map, r_si, n, v_si, w0, f, a_ij, s.wl, float(T0), s.mass,
            v_bulk=s.v_bulk, nz=s.nz)
    return out
