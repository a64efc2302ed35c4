"""Host-side helpers (reference: p_winds/tools.py).

Only what the forward-model path touches: `nearest_index` and the spectrum dict
(`make_spectrum_from_file`, tools.py:332-409).  The MUSCLES / standard-spectrum
loaders are one-off host I/O outside the hot path."""
import numpy as np

__all__ = ['nearest_index', 'make_spectrum_from_file', 'spectrum_to_cgs']

# scale to SI for the unit names the reference's docs use
_LENGTH = {'angstrom': 1e-10, 'AA': 1e-10, 'Angstrom': 1e-10, 'nm': 1e-9, 'micron': 1e-6,
           'um': 1e-6, 'cm': 1e-2, 'm': 1.0}
_FLUX_LAMBDA = {  # -> W m-2 m-1
    'erg / (s cm2 angstrom)': 1e-7 / 1e-4 / 1e-10, 'erg/s/cm2/angstrom': 1e-7 / 1e-4 / 1e-10,
    'erg s-1 cm-2 angstrom-1': 1e-7 / 1e-4 / 1e-10, 'erg/s/cm**2/AA': 1e-7 / 1e-4 / 1e-10,
    'W / (m2 angstrom)': 1.0 / 1e-10, 'W/m2/angstrom': 1.0 / 1e-10, 'W / (m2 nm)': 1.0 / 1e-9,
    'W/m2/nm': 1.0 / 1e-9, 'W / (m3)': 1.0, 'W/m3': 1.0}


def nearest_index(array, target_value):
    """Index of the element nearest to `target_value` in a sorted array, with the
    reference's tie rule (tools.py:27-48)."""
    array = np.asarray(array)
    index = array.searchsorted(target_value)
    index = np.clip(index, 1, len(array) - 1)
    left = array[index - 1]
    right = array[index]
    index -= target_value - left < right - target_value
    return index


def make_spectrum_from_file(filename, units, path='', skiprows=0, scale_flux=1.0,
                            star_distance=None, semi_major_axis=None):
    """Spectrum dict from a two-column text file (tools.py:332-409).  Unlike the
    reference this does not pop keys out of the caller's `units` dict."""
    table = np.loadtxt(path + filename, usecols=(0, 1), skiprows=skiprows, dtype=float)
    if 'wavelength' in units:
        x_axis, y_axis = 'wavelength', 'flux_lambda'
    else:
        x_axis, y_axis = 'frequency', 'flux_nu'
    conv_pc_to_au = 206264.8062471
    if star_distance is not None and semi_major_axis is not None:
        scale_to_planet = (star_distance * conv_pc_to_au / semi_major_axis) ** 2
    else:
        scale_to_planet = 1.0
    return {x_axis: table[:, 0], y_axis: table[:, 1] * scale_flux * scale_to_planet,
            '{}_unit'.format(x_axis): units[x_axis], 'flux_unit': units['flux']}


def _scale(unit, table, si_unit_builder):
    if unit is None:
        return None
    if isinstance(unit, str):
        if unit not in table:
            raise ValueError('unknown unit string %r' % unit)
        return table[unit]
    if hasattr(unit, 'scale') and hasattr(unit, 'dims'):   # oracle stand-in objects
        return unit.scale
    try:  # a real astropy unit
        import astropy.units as u
        return (1.0 * unit).to(si_unit_builder(u)).value
    except ImportError:
        raise ValueError('cannot interpret unit %r without astropy; pass a string' % (unit,))


def spectrum_to_cgs(spectrum):
    """(wavelength [angstrom], flux [erg s-1 cm-2 A-1]) from a spectrum dict
    (the conversion at hydrogen.py:57-60 / helium.py:85-88)."""
    if 'wavelength' not in spectrum:
        raise ValueError('only wavelength / flux_lambda spectra are supported on the GPU path')
    wl = np.asarray(spectrum['wavelength'], dtype=float)
    fl = np.asarray(spectrum['flux_lambda'], dtype=float)
    s_wl = _scale(spectrum.get('wavelength_unit', 'angstrom'), _LENGTH, lambda u: u.m)
    s_fl = _scale(spectrum.get('flux_unit', 'erg / (s cm2 angstrom)'), _FLUX_LAMBDA,
                  lambda u: u.W / u.m ** 3)
    f_wl = s_wl / 1e-10
    f_fl = s_fl / (1e-7 / 1e-4 / 1e-10)
    if abs(f_wl - 1.0) < 1e-12:
        f_wl = 1.0
    if abs(f_fl - 1.0) < 1e-12:
        f_fl = 1.0
    return wl * f_wl, fl * f_fl
