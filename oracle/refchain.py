"""The docs' cascading forward model (quickstart.ipynb cells 4-12;
advanced_tutorial.ipynb `cascading_model`) driven through the UNMODIFIED reference's own
public functions: hydrogen.ion_fraction -> parker.structure[_tidal] ->
helium.population_fraction -> transit.radiative_transfer_2d.

TEST / BENCH INFRASTRUCTURE: the fixture generators (make_golden.py, make_anchors.py) and the
reference arm of bench.py call this; nothing in the product does.  The reference is imported
by refloader (from /root/reference, or from the oracle/_ref copy on the GPU box)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refloader  # noqa: E402

refloader.load()
import astropy.units as u  # noqa: E402  (stand-in)
from p_winds import parker, hydrogen, helium, transit, tools, lines  # noqa: E402

DATA = os.path.join(refloader.REFERENCE_ROOT, 'data')

R_PL, M_PL = 1.39, 0.73
M_STAR, A_AU = 1.07, 0.04634
R = np.logspace(0, np.log10(20), 100)
M_HE = 4 * 1.67262192369e-27
TIGHT_H = dict(rtol=1e-10, atol=1e-13)
TIGHT_HE = dict(method='LSODA', rtol=1e-10, atol=1e-16)


def spectrum_dict():
    units = {'wavelength': u.angstrom,
             'flux': u.erg / u.s / u.cm ** 2 / u.angstrom}
    return tools.make_spectrum_from_file(
        os.path.join(DATA, 'solar_spectrum_scaled_lambda.dat'), units)


def one_model(sp, geom, T0, mdot, h, tidal, tight):
    """The docs' cascading model (quickstart.ipynb cells 4-12)."""
    he_h = (1 - h) / h
    mu_0 = (1 + 4 * he_h) / (1 + he_h + 0.0)
    kw = dict(star_mass=M_STAR, semimajor_axis=A_AU) if tidal else {}
    o = dict(status=0, f_r=np.full(100, np.nan), mu_bar=np.nan,
             v=np.full(100, np.nan), rho=np.full(100, np.nan),
             f_1=np.full(100, np.nan), f_3=np.full(100, np.nan),
             spectrum=np.full(200, np.nan), vs=np.nan, rs=np.nan, rhos=np.nan)
    try:
        f_r, mu = hydrogen.ion_fraction(
            R, R_PL, T0, h, mdot, M_PL, mu_0, spectrum_at_planet=sp,
            exact_phi=True, initial_f_ion=0.0, relax_solution=True,
            return_mu=True, **kw, **(TIGHT_H if tight else {}))
    except RuntimeError:
        o['status'] = 1
        return o
    vs = parker.sound_speed(T0, mu)
    if tidal:
        rs = parker.radius_sonic_point_tidal(M_PL, vs, M_STAR, A_AU)
        rhos = parker.density_sonic_point(mdot, rs, vs)
        v, rho = parker.structure_tidal(R * R_PL / rs, vs, rs, M_PL, M_STAR, A_AU)
    else:
        rs = parker.radius_sonic_point(M_PL, vs)
        rhos = parker.density_sonic_point(mdot, rs, vs)
        v, rho = parker.structure(R * R_PL / rs)
    o.update(f_r=f_r, mu_bar=mu, v=v, rho=rho, vs=vs, rs=rs, rhos=rhos)
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter('always')
            f1, f3 = helium.population_fraction(
                R, v, rho, f_r, R_PL, T0, h, vs, rs, rhos, sp,
                initial_state=np.array([1.0, 0.0]), relax_solution=True,
                **(TIGHT_HE if tight else {}))
        if any('odeint' in str(x.message).lower() or 'lsoda' in str(x.message).lower()
               or 'Excess' in str(x.message) for x in w):
            o['status'] = 3
    except RuntimeError:
        o['status'] = 2
        return o
    o.update(f_1=f1, f_3=f3)
    he_fraction = 1 - h
    n_he = rho * rhos * he_fraction / (h + 4 * he_fraction) / 1.67262192369e-24
    w0, w1, w2, f0, f1_, f2, a_ij = lines.he_3_properties()
    wl = np.linspace(1.0827, 1.0832, 200) * 1E-6
    i0, r_map = geom
    o['spectrum'] = transit.radiative_transfer_2d(
        i0, r_map, R * (R_PL * 71492000), f3 * n_he * 1E6, v * vs * 1000,
        np.array([w0, w1, w2]), np.array([f0, f1_, f2]), np.array([a_ij] * 3),
        wl, float(T0), M_HE, wind_broadening_method='average')
    return o




def reference_geometry(grid_size=100, supersampling=10):
    """transit.draw_transit of the reference (quickstart geometry: k = 0.12086, b = 0.499),
    hoisted out of the model loop exactly as the docs' retrieval does."""
    i0, _depth, r_map = transit.draw_transit(0.12086, R_PL * 71492000, 0.499, 0.0,
                                             supersampling=supersampling, grid_size=grid_size)
    return i0, r_map
