"""He singlet / triplet populations (reference: p_winds/helium.py), reference
signatures; the work is done by k_helium (csrc/pw_helium.cuh, LSODA replica)."""
import warnings

import numpy as np

from . import tools
from .engine import default_engine, STATUS_MESSAGES, Engine

__all__ = ['radiative_processes', 'radiative_processes_mono', 'recombination', 'recombination_all',
           'collision', 'charge_transfer', 'population_fraction', 'ion_fraction']

_RATE_KEYS = ['ionization_1', 'ionization_3', 'recombination_1', 'recombination_3',
              'radiative_transition', 'transition_1_to_3', 'transition_3_to_21s',
              'transition_3_to_21p', 'other_ionization', 'charge_exchange_1',
              'charge_exchange_he_ion']     # helium.py:774-784

# He 2^3S cross-section table (microphysics.py:166-192, Norcross 1971), ascending wavelength
_HE3_WL = np.array([209.49, 219.59, 230.71, 243.01, 256.70, 271.21, 271.94, 331.36, 357.34,
                    387.75, 423.81, 467.27, 520.65, 587.81, 674.86, 792.18, 958.87, 1214.41,
                    1655.63, 2023.15, 2275.74, 2528.27, 2593.01])
_HE3_F = np.array([0.1537, 0.1750, 0.200, 0.231, 0.274, 0.338, 0.343, 0.0520, 0.0325, 0.0310,
                   0.0358, 0.0461, 0.0557, 0.0620, 0.0780, 0.1138, 0.1572, 0.247, 0.435, 0.501,
                   0.537, 0.589, 0.605])
# effective collision strengths (microphysics.py:209-219, Bray et al. 2000)
_GAMMA = np.array([[3.75, 6.198E-2, 2.389, 7.965E-1], [4.00, 6.458E-2, 2.456, 9.579E-1],
                   [4.25, 6.387E-2, 2.275, 1.042], [4.50, 6.157E-2, 1.916, 1.015],
                   [4.75, 5.832E-2, 1.496, 8.950E-1], [5.00, 5.320E-2, 1.111, 7.265E-1],
                   [5.25, 4.787E-2, 8.003E-1, 5.516E-1], [5.50, 4.018E-2, 5.660E-1, 3.948E-1],
                   [5.75, 3.167E-2, 3.944E-1, 2.677E-1]])


def _sigma_h(wl):
    eps = (911.65 / wl - 1) ** 0.5
    return (6.3E-18 * np.exp(4 - (4 * np.arctan(eps)) / eps) / (1 - np.exp(-2 * np.pi / eps)) *
            (wl / 911.65) ** 4)


def radiative_processes(spectrum_at_planet, combined_ionization=False):
    """phi_1, phi_3, a_1, a_3, a_h_1, a_h_3 (helium.py:24-154), computed on the GPU
    when the SED is uploaded."""
    eng = default_engine()
    wl, fl = tools.spectrum_to_cgs(spectrum_at_planet)
    eng.set_sed_cached(wl, fl)
    s = eng.sed_scalars()
    if combined_ionization:   # phi, a_he, a_h (helium.py:160-177)
        return s['hei_phi'], s['hei_a_he'], s['hei_a_h']
    return (s['he_phi_1'], s['he_phi_3'], s['he_a_1'], s['he_a_3'], s['he_a_h_1'], s['he_a_h_3'])


def radiative_processes_mono(flux_euv, flux_fuv, average_euv_photon_wavelength=242.0,
                             average_fuv_photon_wavelength=2348.0):
    """Monochromatic channels (helium.py:181-266), including the reference's use of a
    wavelength in place of an energy for a_h_1 (helium.py:251)."""
    w1 = average_euv_photon_wavelength
    scale = max(37.0 - 19.1 * ((12398.41984332 / w1) / 65.4) ** (-0.76), 0.0)
    a_1 = _sigma_h(w1) * scale
    a_3 = np.interp(average_fuv_photon_wavelength, _HE3_WL, 8.0670E-18 * _HE3_F)
    a_h_1 = 6.3E-18 * (average_euv_photon_wavelength / 13.6) ** (-3)
    a_h_3 = _sigma_h(average_fuv_photon_wavelength) if average_fuv_photon_wavelength < 911.0 else 0.0
    energy_1 = 12398.419843320025 / average_euv_photon_wavelength
    energy_3 = 12398.419843320025 / average_fuv_photon_wavelength
    phi_1 = flux_euv * 6.24150907e+11 * a_1 / energy_1
    phi_3 = flux_fuv * 6.24150907e+11 * a_3 / energy_3
    return phi_1, phi_3, a_1, a_3, a_h_1, a_h_3


def recombination(temperature):
    """alpha_rec_1, alpha_rec_3 in cm3/s (helium.py:270-292)."""
    return 1.54E-13 * (temperature / 1E4) ** (-0.486), 2.10E-13 * (temperature / 1E4) ** (-0.778)


def collision(temperature):
    """q_13, q_31a, q_31b, Q_He, Q_He+ in cm3/s (helium.py:317-380)."""
    tt = 10 ** _GAMMA[:, 0]
    g13, g31a, g31b = (np.interp(temperature, tt, _GAMMA[:, k]) for k in (1, 2, 3))
    kt = 8.617333262145179e-05 * temperature
    k1 = 2.10E-8 * (13.6 / kt) ** 0.5
    return (k1 * g13 * np.exp(-19.81 / kt), k1 * g31a / 3 * np.exp(-0.80 / kt),
            k1 * g31b / 3 * np.exp(-1.40 / kt),
            1.75E-11 * (300 / temperature) ** 0.75 * np.exp(-128E3 / temperature),
            1.25E-15 * (300 / temperature) ** (-0.25))


def population_fraction(radius_profile, velocity, density, hydrogen_ion_fraction, planet_radius,
                        temperature, h_fraction, speed_sonic_point, radius_sonic_point,
                        density_sonic_point, spectrum_at_planet=None, flux_euv=None, flux_fuv=None,
                        initial_state=np.array([0.5, 0.5]), relax_solution=False, convergence=0.01,
                        max_n_relax=10, method='odeint', return_rates=False, **options_solve_ivp):
    """He singlet and triplet fractions versus radius (helium.py:413-785): same
    arguments, units, return arity and exceptions as the reference.  `method` may be
    'odeint' (default, LSODA at rtol = atol = 1.49012e-8), 'LSODA' or 'RK45' (with
    rtol / atol / first_step / max_step)."""
    if spectrum_at_planet is not None:
        rates = radiative_processes(spectrum_at_planet)
    elif flux_euv is not None and flux_fuv is not None:
        rates = radiative_processes_mono(flux_euv, flux_fuv)
    else:
        raise ValueError('Either `spectrum_at_planet` must be provided, or `flux_euv` and '
                         '`flux_fuv` must be provided.')
    opt = dict(options_solve_ivp)
    extra = set(opt) - {'rtol', 'atol', 'first_step', 'max_step'}
    if extra:
        raise ValueError('unsupported solve_ivp options on the GPU path: %s' % sorted(extra))
    eng = default_engine()
    opts = Engine.he_opts(planet_radius, rates, initial_state, relax_solution is True, convergence,
                          max_n_relax, method, opt.get('rtol', 1e-3), opt.get('atol', 1e-6),
                          opt.get('first_step'), opt.get('max_step'))
    r = np.ascontiguousarray(radius_profile, dtype=float)
    o = eng.he_population_batch([temperature], [h_fraction], [speed_sonic_point],
                                [radius_sonic_point], [density_sonic_point], r, velocity, density,
                                hydrogen_ion_fraction, opts, return_rates=return_rates is True)
    status = int(o['status'].cpu()[0])
    if status == 2:
        msg = STATUS_MESSAGES[2] if method == 'odeint' else STATUS_MESSAGES[1]
        raise RuntimeError(msg)
    if status == 3:
        # helium.py:736: the reference's user sees an ODEintWarning here and gets an array whose rows after the
        # failing radius were never written; the engine returns the last relax pass that completed
        warnings.warn('odeint warned inside the relax loop (pass %d): the populations of the last completed pass '
                      'are returned; the reference result is undefined for this model.' % int(o['scal'][0, 0].cpu()),
                      RuntimeWarning, stacklevel=2)
    f1, f3 = o['f_1'][0].cpu().numpy(), o['f_3'][0].cpu().numpy()
    if return_rates is True:
        rr = o['rates'][0].cpu().numpy()
        return f1, f3, {k: rr[i] for i, k in enumerate(_RATE_KEYS)}
    return f1, f3


def recombination_all(temperature):
    """He II -> He I without singlet / triplet distinction, cm3/s (helium.py:294-313)."""
    return 4.6E-12 * (temperature / 300) ** (-0.64)


def charge_transfer(temperature):
    """(He + H+ -> He+, He+ + H -> He) in cm3/s (helium.py:384-409)."""
    ct_rate_hep_h = 1.25E-15 * (300 / temperature) ** (-0.25)
    ct_rate_he_hp = 1.75E-11 * (300 / temperature) ** 0.75 * np.exp(-128000 / temperature)
    return ct_rate_he_hp, ct_rate_hep_h


def ion_fraction(radius_profile, velocity, density, hydrogen_ion_fraction, planet_radius, temperature,
                 h_fraction, speed_sonic_point, radius_sonic_point, density_sonic_point,
                 spectrum_at_planet, initial_f_he_ion=0.0, relax_solution=False, convergence=0.01,
                 max_n_relax=10, method='Radau', **options_solve_ivp):
    """Fraction of ionised helium versus radius (helium.py:789-1032), same arguments, units and
    exceptions as the reference.  `method`: 'Radau' (default), 'odeint' or 'RK45'."""
    phi, a_he, a_h = radiative_processes(spectrum_at_planet, combined_ionization=True)
    opt = dict(options_solve_ivp)
    extra = set(opt) - {'rtol', 'atol', 'first_step', 'max_step'}
    if extra:
        raise ValueError('unsupported solve_ivp options on the GPU path: %s' % sorted(extra))
    eng = default_engine()
    opts = Engine.o_opts(planet_radius, (phi, a_he, a_h, 0.0), 0.0, initial_f_he_ion, relax_solution is True,
                         convergence, max_n_relax, method, opt.get('rtol', 1e-3), opt.get('atol', 1e-6),
                         opt.get('first_step'), opt.get('max_step'))
    r = np.ascontiguousarray(radius_profile, dtype=float)
    o = eng.he_ion_fraction_batch([temperature], [h_fraction], [speed_sonic_point], [radius_sonic_point],
                                  [density_sonic_point], r, velocity, density, hydrogen_ion_fraction, opts)
    if int(o['status'].cpu()[0]) == 2:
        raise RuntimeError(STATUS_MESSAGES[2] if method == 'odeint' else STATUS_MESSAGES[1])
    if int(o['status'].cpu()[0]) == 3:
        # an `odeint` of the relax loop warned: the reference goes on with rows its solver never wrote (undefined
        # result); the engine returns the last relax pass that completed
        warnings.warn('odeint warned inside the relax loop: the fractions of the last completed pass are returned; '
                      'the reference result is undefined for this model.', RuntimeWarning, stacklevel=2)
    return o['f_he_ion'][0].cpu().numpy()
