"""C II / C III fractions (reference: p_winds/carbon.py), reference signatures; the work is done
by k_carbon (csrc/pw_carbon.cuh) with the LSODA replica ('odeint', the reference default) or the
Radau IIA replica (the method the exospheric-metals tutorial uses)."""
import warnings

import numpy as np

from . import tools
from .engine import default_engine, STATUS_MESSAGES, Engine

__all__ = ['radiative_processes', 'electron_impact_ionization', 'recombination', 'charge_transfer',
           'ion_fraction']

_SOLAR_CARBON_ABUNDANCE_ = 8.43  # Asplund et al. 2009 (carbon.py:21)
_SOLAR_CARBON_FRACTION_ = 10 ** (_SOLAR_CARBON_ABUNDANCE_ - 12.00)
# order in which _fun(..., rates=True) returns the terms (carbon.py:531-533)
_TERM_ORDER = ['ion_rate_ci', 'ion_rate_cii', 'recomb_rate_cii', 'recomb_rate_ciii', 'e_imp_ion_ci',
               'e_imp_ion_cii', 'ch_exchange_ci_hii', 'ch_exchange_ci_heii', 'ch_exchange_cii_hi',
               'ch_exchange_ciii_hi', 'ch_exchange_ciii_hei']


def radiative_processes(spectrum_at_planet):
    """phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he (carbon.py:30-131), computed on the GPU
    when the SED is uploaded."""
    eng = default_engine()
    wl, fl = tools.spectrum_to_cgs(spectrum_at_planet)
    eng.set_sed_cached(wl, fl)
    s = eng.sed_scalars()
    return (s['c_phi_1'], s['c_phi_2'], s['c_a_1'], s['c_a_2'], s['c_a_h_1'], s['c_a_h_2'], s['c_a_he'])


def electron_impact_ionization(electron_temperature):
    """(C I -> C II, C II -> C III) by electron impact in cm3/s (carbon.py:135-165, Voronov 1997)."""
    e = 8.617333262145179e-05 * electron_temperature
    r1, r2 = 11.3 / e, 24.4 / e
    return (6.85E-8 * (0.193 + r1) ** (-1) * r1 ** 0.25 * np.exp(-r1),
            1.86E-8 * (0.286 + r2) ** (-1) * r2 ** 0.24 * np.exp(-r2))


def recombination(electron_temperature):
    """(C II -> C I, C III -> C II) in cm3/s (carbon.py:169-194)."""
    return (4.67E-12 * (300 / electron_temperature) ** 0.60,
            2.3E-12 * (1000 / electron_temperature) ** 0.645)


def charge_transfer(temperature):
    """ct_rate_ci_hp, ct_rate_cii_h, ct_rate_ci_hep, ct_rate_cii_sii, ct_rate_ciii_h, ct_rate_ciii_he
    in cm3/s (carbon.py:198-252)."""
    ct_rate_cii_h = 6.30E-17 * (300 / temperature) ** (-1.96) * np.exp(-1.7E5 / temperature)
    ct_rate_cii_sii = 2.1E-9
    ct_rate_ci_hp = 1.31E-15 * (300 / temperature) ** (-0.213)
    ct_rate_ci_hep = 2.5E-15 * (300 / temperature) ** (-1.597)
    ct_rate_ciii_h = 1.67E-4 * (temperature / 10000) ** 2.79 * (1 + 304.72 * np.exp(-4.07 * temperature / 10000))
    ct_rate_ciii_he = 1.23E-9
    return ct_rate_ci_hp, ct_rate_cii_h, ct_rate_ci_hep, ct_rate_cii_sii, ct_rate_ciii_h, ct_rate_ciii_he


def ion_fraction(radius_profile, velocity, density, hydrogen_ion_fraction, helium_ion_fraction,
                 planet_radius, temperature, h_fraction, speed_sonic_point, radius_sonic_point,
                 density_sonic_point, spectrum_at_planet, c_fraction=_SOLAR_CARBON_FRACTION_,
                 initial_f_c_ion=np.array([0.0, 0.0]), relax_solution=False, convergence=0.01,
                 max_n_relax=10, method='odeint', return_rates=False, **options_solve_ivp):
    """Fractions of singly and doubly ionised carbon versus radius (carbon.py:256-639): same
    arguments, units, return arity and exceptions as the reference.  `method` may be 'odeint'
    (default), 'Radau' or 'RK45' (with rtol / atol / first_step / max_step)."""
    rates = radiative_processes(spectrum_at_planet)
    opt = dict(options_solve_ivp)
    extra = set(opt) - {'rtol', 'atol', 'first_step', 'max_step'}
    if extra:
        raise ValueError('unsupported solve_ivp options on the GPU path: %s' % sorted(extra))
    eng = default_engine()
    opts = Engine.c_opts(planet_radius, rates, c_fraction, initial_f_c_ion, relax_solution is True, convergence,
                         max_n_relax, method, opt.get('rtol', 1e-3), opt.get('atol', 1e-6), opt.get('first_step'),
                         opt.get('max_step'))
    r = np.ascontiguousarray(radius_profile, dtype=float)
    o = eng.c_ion_fraction_batch([temperature], [h_fraction], [speed_sonic_point], [radius_sonic_point],
                                 [density_sonic_point], r, velocity, density, hydrogen_ion_fraction,
                                 helium_ion_fraction, opts, return_rates=return_rates is True)
    if int(o['status'].cpu()[0]) == 2:
        raise RuntimeError(STATUS_MESSAGES[2] if method == 'odeint' else STATUS_MESSAGES[1])
    if int(o['status'].cpu()[0]) == 3:
        # an `odeint` of the relax loop warned: the reference goes on with rows its solver never wrote (undefined
        # result); the engine returns the last relax pass that completed
        warnings.warn('odeint warned inside the relax loop: the fractions of the last completed pass are returned; '
                      'the reference result is undefined for this model.', RuntimeWarning, stacklevel=2)
    f2, f3 = o['f_cii'][0].cpu().numpy(), o['f_ciii'][0].cpu().numpy()
    if return_rates is True:
        t = dict(zip(_TERM_ORDER, o['rates'][0].cpu().numpy()))
        # the reference fills 'ionization_CII' with the electron-impact term (carbon.py:627-628); kept
        return f2, f3, {'ionization_CI': t['ion_rate_ci'], 'ionization_CII': t['e_imp_ion_cii'],
                        'recombination_CII': t['recomb_rate_cii'], 'recombination_CIII': t['recomb_rate_ciii'],
                        'e_impact_ion_CI': t['e_imp_ion_ci'], 'e_impact_ion_CII': t['e_imp_ion_cii'],
                        'charge_exchange_CI_HII': t['ch_exchange_ci_hii'],
                        'charge_exchange_CI_HeII': t['ch_exchange_ci_heii'],
                        'charge_exchange_CII_HI': t['ch_exchange_cii_hi'],
                        'charge_exchange_CIII_HI': t['ch_exchange_ciii_hi'],
                        'charge_exchange_CIII_HeI': t['ch_exchange_ciii_hei']}
    return f2, f3
