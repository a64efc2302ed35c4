"""The retrieval loop of docs/source/advanced_tutorial.ipynb for a whole parameter grid.

The notebook evaluates one `theta = (log_m_dot, log_T, v_wind)` per call of
`cascading_model` / `log_likelihood` (cells 9-15) inside a Nelder-Mead minimisation; a
grid search or an MCMC ensemble evaluates thousands of them per step.  These functions
keep the notebook's names and argument meaning and take arrays of models instead.
Everything runs in the CUDA engine (no CPU fallback).
"""
import numpy as np

from .engine import Engine, default_engine

__all__ = ['set_phases', 'cascading_model', 'log_likelihood']


def set_phases(planet_to_star_ratio, planet_physical_radius, radius_profile_si, impact_parameter=0.0,
               sample_phases=(0.0,), grid_size=100, supersampling=None, engine=None):
    """Cell 11, first loop: one `transit.draw_transit` map per sampled phase, kept on the
    device; afterwards the engine's spectra are the mean over these phases of
    (in-transit spectrum + geometric transit depth).  Returns the transit depths."""
    eng = engine or default_engine()
    maps = []
    for ph in sample_phases:
        i0, depth, r_map = eng.draw_transit(planet_to_star_ratio, planet_physical_radius, impact_parameter,
                                            float(ph), grid_size, supersampling)
        maps.append((i0, r_map, depth))
    eng.set_geometry_phases(maps, radius_profile_si)
    return [m[2] for m in maps]


def cascading_model(T, mass_loss_rate, h_fraction, radius_profile, h_opts, he_opts, rt_opts,
                    wavelength_array, instrumental_profile=None, engine=None):
    """Cells 9-13 for arrays of models: atmospheric model (H, Parker, He), phase-averaged
    transmission spectrum, convolution with the instrumental profile
    (`astropy.convolution.convolve(..., boundary='extend')`).  Returns (spectra[B, Nw] on
    the device, status[B]); `v_wind` is `bulk_los_velocity` of `rt_opts`."""
    eng = engine or default_engine()
    out = eng.forward_batch(T, mass_loss_rate, h_fraction, radius_profile, h_opts, he_opts, rt_opts,
                            wavelength_array)
    spec = out['spectrum']
    if instrumental_profile is not None:
        spec = eng.convolve_extend_batch(spec, instrumental_profile)
    return spec, out['status']


def log_likelihood(model_spectra, status, y, yerr, engine=None):
    """Cell 15 per model: -0.5 * sum((y - model)^2 / yerr^2 + log(yerr^2)); -inf where the
    model failed (the notebook's `except RuntimeError`)."""
    eng = engine or default_engine()
    return eng.loglike_batch(model_spectra, np.asarray(y, dtype=float), np.asarray(yerr, dtype=float), status)
