"""O II fraction (reference: p_winds/oxygen.py), reference signatures; the work is done by
k_oxygen (csrc/pw_oxygen.cuh) with the Radau IIA replica of csrc/pw_radau.cuh (the reference's
default integrator for this species)."""
import warnings

import numpy as np

from . import tools
from .engine import default_engine, STATUS_MESSAGES, Engine

__all__ = ['radiative_processes', 'electron_impact_ionization', 'recombination', 'charge_transfer',
           'ion_fraction']

_SOLAR_OXYGEN_ABUNDANCE_ = 8.69  # Asplund et al. 2009 (oxygen.py:21)
_SOLAR_OXYGEN_FRACTION_ = 10 ** (_SOLAR_OXYGEN_ABUNDANCE_ - 12.00)
_RATE_KEYS = ['photoionization', 'recombination', 'e_impact_ionization', 'charge_exchange_HII',
              'charge_exchange_HI']   # oxygen.py:484-489


def radiative_processes(spectrum_at_planet):
    """phi_oi, a_oi, a_h_oi, a_he (oxygen.py:26-99), computed on the GPU when the SED is uploaded."""
    eng = default_engine()
    wl, fl = tools.spectrum_to_cgs(spectrum_at_planet)
    eng.set_sed_cached(wl, fl)
    s = eng.sed_scalars()
    return s['o_phi'], s['o_a'], s['o_a_h'], s['o_a_he']


def electron_impact_ionization(electron_temperature):
    """O I + e -> O II in cm3/s (oxygen.py:103-126, Voronov 1997)."""
    ratio = 11.3 / (8.617333262145179e-05 * electron_temperature)
    return 3.59E-8 * (0.073 + ratio) ** (-1) * ratio ** 0.34 * np.exp(-ratio)


def recombination(electron_temperature):
    """O II + e -> O I in cm3/s (oxygen.py:130-149)."""
    return 3.25E-12 * (300 / electron_temperature) ** 0.66


def charge_transfer(temperature):
    """(O I + H+ -> O II, O II + H -> O I) in cm3/s (oxygen.py:153-182)."""
    ct_rate_oii_h = 5.66E-10 * (300 / temperature) ** (-0.36) * np.exp(8.6 / temperature)
    ct_rate_oi_hp = 7.31E-10 * (300 / temperature) ** (-0.23) * np.exp(-226 / temperature)
    return ct_rate_oi_hp, ct_rate_oii_h


def ion_fraction(radius_profile, velocity, density, hydrogen_ion_fraction, helium_ion_fraction,
                 planet_radius, temperature, h_fraction, speed_sonic_point, radius_sonic_point,
                 density_sonic_point, spectrum_at_planet, o_fraction=_SOLAR_OXYGEN_FRACTION_,
                 initial_f_o_ion=0.0, relax_solution=False, convergence=0.01, max_n_relax=10,
                 method='Radau', return_rates=False, **options_solve_ivp):
    """Fraction of singly ionised oxygen versus radius (oxygen.py:186-491): same arguments,
    units, return arity and exceptions as the reference.  `method` may be 'Radau' (default),
    'odeint' or 'RK45' (with rtol / atol / first_step / max_step)."""
    rates = radiative_processes(spectrum_at_planet)
    opt = dict(options_solve_ivp)
    extra = set(opt) - {'rtol', 'atol', 'first_step', 'max_step'}
    if extra:
        raise ValueError('unsupported solve_ivp options on the GPU path: %s' % sorted(extra))
    eng = default_engine()
    opts = Engine.o_opts(planet_radius, rates, o_fraction, initial_f_o_ion, relax_solution is True,
                         convergence, max_n_relax, method, opt.get('rtol', 1e-3), opt.get('atol', 1e-6),
                         opt.get('first_step'), opt.get('max_step'))
    r = np.ascontiguousarray(radius_profile, dtype=float)
    o = eng.o_ion_fraction_batch([temperature], [h_fraction], [speed_sonic_point], [radius_sonic_point],
                                 [density_sonic_point], r, velocity, density, hydrogen_ion_fraction,
                                 helium_ion_fraction, opts, return_rates=return_rates is True)
    if int(o['status'].cpu()[0]) == 2:
        raise RuntimeError(STATUS_MESSAGES[2] if method == 'odeint' else STATUS_MESSAGES[1])
    if int(o['status'].cpu()[0]) == 3:
        # an `odeint` of the relax loop warned: the reference goes on with rows its solver never wrote (undefined
        # result); the engine returns the last relax pass that completed
        warnings.warn('odeint warned inside the relax loop: the fractions of the last completed pass are returned; '
                      'the reference result is undefined for this model.', RuntimeWarning, stacklevel=2)
    f = o['f_o'][0].cpu().numpy()
    if return_rates is True:
        rr = o['rates'][0].cpu().numpy()
        return f, {k: rr[i] for i, k in enumerate(_RATE_KEYS)}
    return f
