"""Transit geometry and radiative transfer (reference: p_winds/transit.py),
reference signatures."""
import numpy as np

from .engine import default_engine, Engine

__all__ = ['draw_transit', 'radiative_transfer_2d', 'optical_depth_2d']


def draw_transit(planet_to_star_ratio, planet_physical_radius, impact_parameter=0.0, phase=0.0,
                 grid_size=100, supersampling=None, resample_method=None,
                 limb_darkening_law=None, ld_coefficient=None):
    """Normalised intensity map, geometric transit depth and planet-centric radius map
    (transit.py:20-116; flatstar's disk rasteriser restated in csrc/pw_draw.cuh)."""
    if resample_method not in (None, 'box'):
        raise ValueError("only the reference's default 'box' resampling is available on the GPU")
    i0, depth, rm = default_engine().draw_transit(
        planet_to_star_ratio, planet_physical_radius, impact_parameter, phase, grid_size,
        supersampling, limb_darkening_law, ld_coefficient)
    return i0.cpu().numpy(), depth, rm.cpu().numpy()


def _prepare(eng, intensity_0, r_from_planet, radius_profile, central_wavelength,
             oscillator_strength, einstein_coefficient, gas_temperature, particle_mass,
             bulk_los_velocity, planet_radial_velocity, wind_broadening_method, z_grid_size,
             turbulence_broadening):
    if isinstance(gas_temperature, (float, int)):
        t_prof = False
    elif isinstance(gas_temperature, np.ndarray):
        t_prof = True
    else:
        raise ValueError('`gas_temperature` has to be either `float` or `np.ndarray`.')
    if not isinstance(central_wavelength, (float, list, np.ndarray)):
        raise ValueError('The spectral line properties must be float, list or numpy.ndarray, and '
                         'their format must be consistent with each other.')
    opts = Engine.rt_opts(central_wavelength, oscillator_strength, einstein_coefficient,
                          particle_mass, bulk_los_velocity, planet_radial_velocity,
                          wind_broadening_method, z_grid_size, turbulence_broadening, t_prof)
    r_map = np.asarray(r_from_planet, dtype=float)
    i0 = np.asarray(intensity_0, dtype=float)
    if i0.ndim == 0:   # scalar intensity: broadcast over the map (transit.py:225-228)
        i0 = np.full(r_map.shape, float(i0))
    eng.set_geometry_cached(i0, r_map, np.asarray(radius_profile, dtype=float))
    return opts


def radiative_transfer_2d(intensity_0, r_from_planet, radius_profile, density_profile,
                          velocity_profile, central_wavelength, oscillator_strength,
                          einstein_coefficient, wavelength_grid, gas_temperature, particle_mass,
                          bulk_los_velocity=0.0, planet_radial_velocity=0.0,
                          wind_broadening_method='average', z_grid_size=200,
                          turbulence_broadening=False):
    """In-transit spectrum sum_pixels I0 exp(-tau) (transit.py:120-230)."""
    eng = default_engine()
    opts = _prepare(eng, intensity_0, r_from_planet, radius_profile, central_wavelength,
                    oscillator_strength, einstein_coefficient, gas_temperature, particle_mass,
                    bulk_los_velocity, planet_radial_velocity, wind_broadening_method, z_grid_size,
                    turbulence_broadening)
    spec = eng.rt2d_batch(density_profile, velocity_profile, gas_temperature, wavelength_grid, opts)
    return spec[0].cpu().numpy()


def optical_depth_2d(radius_profile, density_profile, velocity_profile, central_wavelength,
                     oscillator_strength, einstein_coefficient, wavelength_grid, gas_temperature,
                     particle_mass, z_grid_size, bulk_los_velocity=0.0, planet_radial_velocity=0.0,
                     wind_broadening_method='average', turbulence_broadening=False):
    """tau[Nr, Nw] versus impact parameter and wavelength (transit.py:312-536)."""
    eng = default_engine()
    r = np.asarray(radius_profile, dtype=float)
    opts = _prepare(eng, 0.0, np.array([[r[0]]]), r, central_wavelength, oscillator_strength,
                    einstein_coefficient, gas_temperature, particle_mass, bulk_los_velocity,
                    planet_radial_velocity, wind_broadening_method, z_grid_size,
                    turbulence_broadening)
    _, tau = eng.rt2d_batch(density_profile, velocity_profile, gas_temperature, wavelength_grid,
                            opts, return_tau=True)
    return tau[0].cpu().numpy()
