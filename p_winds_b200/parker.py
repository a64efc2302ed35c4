"""Isothermal Parker wind (reference: p_winds/parker.py), reference signatures.

The scalar helpers are the reference's one-line closed forms; `structure` /
`structure_tidal` evaluate the velocity and density profiles on the GPU
(Lambert-W form of the transcendental equation, csrc/pw_parker.cuh)."""
import numpy as np

from .engine import default_engine

__all__ = ['sound_speed', 'radius_sonic_point', 'density_sonic_point', 'structure',
           'radius_sonic_point_tidal', 'structure_tidal']


def sound_speed(temperature, mean_molecular_weight=1.0):
    """Isothermal sound speed in km/s (parker.py:83-106)."""
    return (1.380649e-29 * temperature / mean_molecular_weight / 1.67262192369e-27) ** 0.5


def radius_sonic_point(planet_mass, sound_speed_0):
    """Sonic radius in R_Jup (parker.py:110-131)."""
    return 1772.0378503888546 * planet_mass / 2 / sound_speed_0 ** 2


def density_sonic_point(mass_loss_rate, radius_sp, sound_speed_0):
    """Density at the sonic point in g cm-3 (parker.py:135-159)."""
    return mass_loss_rate / 4 / np.pi / (radius_sp * 7.1492E+09) ** 2 / (sound_speed_0 * 1E5)


def radius_sonic_point_tidal(planet_mass, sound_speed_0, star_mass, semi_major_axis):
    """Sonic radius with stellar tidal gravity, R_Jup (parker.py:226-267)."""
    grav = 1772.0378503888546
    star_mass = 1047.56551466 * star_mass
    semi_major_axis = 2092.51203911 * semi_major_axis
    v_k = np.sqrt(grav * star_mass / semi_major_axis)
    m1 = np.sqrt(planet_mass ** 2 / 4 + 8 * star_mass ** 2 / 81 * (sound_speed_0 / v_k) ** 6)
    return semi_major_axis * (((m1 + planet_mass / 2) / (3 * star_mass)) ** (1 / 3) -
                              ((m1 - planet_mass / 2) / (3 * star_mass)) ** (1 / 3))


def _out(t, scalar):
    a = t.cpu().numpy()
    return float(a[0]) if scalar else a


def structure(r, v_guess=None):
    """Velocity and density in sonic-point units (parker.py:163-223).  `v_guess` is
    accepted for signature compatibility; the GPU solver needs no first guess (it
    always converges to the transonic branch the reference's guesses select)."""
    scalar = np.ndim(r) == 0
    v, rho = default_engine().parker_structure(np.atleast_1d(np.asarray(r, dtype=float)))
    return _out(v, scalar), _out(rho, scalar)


def structure_tidal(r, sound_speed_0, r_sonic_point, planet_mass, star_mass, semi_major_axis):
    """As `structure`, with the tidal potential of the host star (parker.py:271-358)."""
    scalar = np.ndim(r) == 0
    v, rho = default_engine().parker_structure_tidal(
        np.atleast_1d(np.asarray(r, dtype=float)), sound_speed_0, r_sonic_point, planet_mass,
        star_mass, semi_major_axis)
    return _out(v, scalar), _out(rho, scalar)
