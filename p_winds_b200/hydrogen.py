"""Neutral / ionised hydrogen (reference: p_winds/hydrogen.py), reference signatures;
the work is done by k_hydrogen (csrc/pw_hydrogen.cuh) with a batch of one."""
import numpy as np

from . import tools
from .engine import default_engine, STATUS_MESSAGES, Engine

__all__ = ['radiative_processes', 'radiative_processes_mono', 'recombination', 'ion_fraction']


def _sed(spectrum_at_planet):
    eng = default_engine()
    wl, fl = tools.spectrum_to_cgs(spectrum_at_planet)
    eng.set_sed_cached(wl, fl)
    return eng


def radiative_processes(spectrum_at_planet):
    """(phi [1/s], a_0 [cm2]) at null optical depth (hydrogen.py:118-169)."""
    s = _sed(spectrum_at_planet).sed_scalars()
    return s['h_phi'], s['h_a0']


def radiative_processes_mono(flux_euv, average_photon_energy=20.):
    """Monochromatic channel (hydrogen.py:173-203); does not modify its argument."""
    a_0 = 6.3E-18 * (average_photon_energy / 13.6) ** (-3)
    return flux_euv * 6.24150907E+11 * a_0 / average_photon_energy, a_0


def recombination(temperature):
    """Case-B recombination rate, cm3/s (hydrogen.py:207-223)."""
    return 2.59E-13 * (temperature / 1E4) ** (-0.7)


def _ivp_options(options):
    method = options.pop('method', 'RK45')
    if method != 'RK45':
        raise ValueError("solve_ivp method %r is not available on the GPU path (RK45 only); "
                         "there is no CPU fallback" % (method,))
    allowed = {'rtol', 'atol', 'first_step', 'max_step'}
    extra = set(options) - allowed
    if extra:
        raise ValueError('unsupported solve_ivp options on the GPU path: %s' % sorted(extra))
    return options


def ion_fraction(radius_profile, planet_radius, temperature, h_fraction, mass_loss_rate,
                 planet_mass, mean_molecular_weight_0=1.0, star_mass=1.0, semimajor_axis=1.0,
                 spectrum_at_planet=None, flux_euv=None, initial_f_ion=0.0, relax_solution=False,
                 convergence=0.01, max_n_relax=10, exact_phi=False, return_mu=False,
                 return_rates=False, **options_solve_ivp):
    """H+ fraction versus radius (hydrogen.py:227-573): same arguments, units, return
    arity and exceptions as the reference."""
    ivp = _ivp_options(dict(options_solve_ivp))
    phi_abs = a_0 = 0.0
    exact = bool(exact_phi and spectrum_at_planet is not None)
    if exact:
        eng = _sed(spectrum_at_planet)
    elif spectrum_at_planet is not None:
        eng = _sed(spectrum_at_planet)
        phi_abs, a_0 = radiative_processes(spectrum_at_planet)
    elif flux_euv is not None:
        eng = default_engine()
        phi_abs, a_0 = radiative_processes_mono(flux_euv)
    else:
        raise ValueError('Either `spectrum_at_planet` or `flux_euv` must be provided.')
    if return_rates is True and not exact:
        # the reference reads variables that only the exact_phi branch defines (hydrogen.py:540-543)
        raise ValueError('`return_rates` needs `exact_phi=True` and a `spectrum_at_planet`.')
    opts = Engine.h_opts(planet_radius, planet_mass, star_mass, semimajor_axis, initial_f_ion,
                         relax_solution is True, convergence, max_n_relax, exact, phi_abs, a_0,
                         ivp.get('rtol', 1e-3), ivp.get('atol', 1e-6), ivp.get('first_step'),
                         ivp.get('max_step'))
    r = np.ascontiguousarray(radius_profile, dtype=float)
    o = eng.h_ion_fraction_batch([temperature], [mass_loss_rate], [h_fraction],
                                 [mean_molecular_weight_0], r, opts, return_rates=return_rates is True)
    status = int(o['status'].cpu()[0])
    if status != 0:
        raise RuntimeError(STATUS_MESSAGES.get(status, 'The solver failed (status %d).' % status))
    f_r = o['f_r'][0].cpu().numpy()
    mu_bar = float(o['scal'][0, 0].cpu())
    if return_rates is True:
        rr = o['rates'][0].cpu().numpy()
        rates = {'photoionization': rr[0], 'recombination': rr[1]}
        return (f_r, mu_bar, rates) if return_mu is True else (f_r, rates)
    if return_mu is True:
        return f_r, mu_bar
    return f_r
