"""The C-ABI shared library loads without a GPU and exports every symbol that
include/pwinds_b200.h declares; the product refuses to run without CUDA and never
routes through the oracle."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'pwinds_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pwb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(built):
    from p_winds_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n
    assert set(names) == set(_lib.SYMBOLS), 'ctypes table and header disagree'
    assert b'sm_100a' in lib.pwb_version()


def test_option_blocks_match_the_compiled_layout(built):
    """ctypes mirrors of the option structs vs sizeof in the library (no GPU needed)."""
    from p_winds_b200 import _lib
    lib = _lib.load()
    for which, cls in enumerate((_lib.HOpts, _lib.HeOpts, _lib.OOpts, _lib.COpts, _lib.RtOpts)):
        assert lib.pwb_sizeof_opts(which) == C.sizeof(cls), cls.__name__


def test_library_reports_the_hash_of_its_sources(built):
    """pwb_version() carries the sha256 of csrc/ + the header it was built from; build() rebuilds on a
    content mismatch (not on mtimes), so the binary that ships to the GPU box is checkable there."""
    from p_winds_b200 import _lib
    v = _lib.load().pwb_version().decode()
    assert v.endswith('src:' + built.source_hash()), v


def test_library_contains_sm100a_code(built):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', built.LIB], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from p_winds_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.pwb_create(0, C.byref(ctx)) < 0 and not ctx.value
    from p_winds_b200.engine import Engine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Engine(0)
    from p_winds_b200 import hydrogen
    with pytest.raises(RuntimeError):
        hydrogen.ion_fraction([1.0, 2.0], 1.39, 9000., 0.9, 1e10, 0.73, flux_euv=1000.)


def test_missing_library_is_an_import_error(monkeypatch, built):
    from p_winds_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libpwinds_b200.so')
    with pytest.raises(ImportError, match='no CPU fallback'):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'p_winds_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'pwinds_oracle' not in txt and 'import oracle' not in txt, f
                assert 'scipy' not in txt or f.endswith(('.cu', '.cuh')), f   # scipy only named in comments of kernels
                assert 'hostsim' not in txt or f.endswith(('.cuh',)), f


def test_reference_signatures_kept():
    """Argument names and defaults of the five functions named by BASELINE.json."""
    import inspect
    import importlib
    # importable without a GPU: the CUDA library is only touched when a function runs
    hydrogen = importlib.import_module('p_winds_b200.hydrogen')
    helium = importlib.import_module('p_winds_b200.helium')
    transit = importlib.import_module('p_winds_b200.transit')
    parker = importlib.import_module('p_winds_b200.parker')
    sig = lambda f: list(inspect.signature(f).parameters)  # noqa: E731
    assert sig(parker.structure) == ['r', 'v_guess']
    assert sig(hydrogen.ion_fraction) == [
        'radius_profile', 'planet_radius', 'temperature', 'h_fraction', 'mass_loss_rate', 'planet_mass',
        'mean_molecular_weight_0', 'star_mass', 'semimajor_axis', 'spectrum_at_planet', 'flux_euv',
        'initial_f_ion', 'relax_solution', 'convergence', 'max_n_relax', 'exact_phi', 'return_mu',
        'return_rates', 'options_solve_ivp']
    assert sig(helium.population_fraction) == [
        'radius_profile', 'velocity', 'density', 'hydrogen_ion_fraction', 'planet_radius', 'temperature',
        'h_fraction', 'speed_sonic_point', 'radius_sonic_point', 'density_sonic_point',
        'spectrum_at_planet', 'flux_euv', 'flux_fuv', 'initial_state', 'relax_solution', 'convergence',
        'max_n_relax', 'method', 'return_rates', 'options_solve_ivp']
    assert sig(transit.draw_transit) == [
        'planet_to_star_ratio', 'planet_physical_radius', 'impact_parameter', 'phase', 'grid_size',
        'supersampling', 'resample_method', 'limb_darkening_law', 'ld_coefficient']
    assert sig(transit.radiative_transfer_2d) == [
        'intensity_0', 'r_from_planet', 'radius_profile', 'density_profile', 'velocity_profile',
        'central_wavelength', 'oscillator_strength', 'einstein_coefficient', 'wavelength_grid',
        'gas_temperature', 'particle_mass', 'bulk_los_velocity', 'planet_radial_velocity',
        'wind_broadening_method', 'z_grid_size', 'turbulence_broadening']
    d = inspect.signature(helium.population_fraction).parameters
    assert d['method'].default == 'odeint' and d['max_n_relax'].default == 10
    d = inspect.signature(transit.radiative_transfer_2d).parameters
    assert d['z_grid_size'].default == 200 and d['wind_broadening_method'].default == 'average'


def test_host_helpers_match_reference_formulas(oracle):
    from p_winds_b200 import parker, tools, lines
    assert parker.sound_speed(1e4, 1.0) == oracle.sound_speed(1e4, 1.0)
    assert parker.radius_sonic_point(0.73, 9.0) == oracle.radius_sonic_point(0.73, 9.0)
    assert parker.radius_sonic_point_tidal(0.73, 9.0, 1.07, 0.04634) == \
        oracle.radius_sonic_point_tidal(0.73, 9.0, 1.07, 0.04634)
    assert parker.density_sonic_point(1e11, 7.0, 9.0) == oracle.density_sonic_point(1e11, 7.0, 9.0)
    import numpy as np
    assert tools.nearest_index(np.linspace(1, 10, 10), np.pi) == 2      # reference tests/test_tools.py:14-17
    w0, f, a = oracle.he3_lines()
    l = lines.he_3_properties()
    assert list(l[:3]) == list(w0) and list(l[3:6]) == list(f) and l[6] == a[0]


def test_spectrum_dict_with_unit_objects():
    """tools.spectrum_to_cgs: the spectrum dict of tools.make_spectrum_from_file carries unit *objects*
    (tools.py:405-408).  astropy is absent here, so the stand-in's objects (scale + dims) exercise that
    branch; strings and objects must agree, and a non-cgs unit must be converted."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'standins'))
    import astropy.units as u     # the stand-in
    from p_winds_b200 import tools
    wl = np.array([100.0, 500.0, 911.0])
    fl = np.array([1.0, 2.0, 3.0])
    a = tools.spectrum_to_cgs({'wavelength': wl, 'flux_lambda': fl, 'wavelength_unit': 'angstrom',
                               'flux_unit': 'erg / (s cm2 angstrom)'})
    b = tools.spectrum_to_cgs({'wavelength': wl, 'flux_lambda': fl, 'wavelength_unit': u.angstrom,
                               'flux_unit': u.erg / u.s / u.cm ** 2 / u.angstrom})
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], wl)
    c = tools.spectrum_to_cgs({'wavelength': wl / 10, 'flux_lambda': fl * 1e-2, 'wavelength_unit': u.nm,
                               'flux_unit': u.W / u.m ** 2 / u.nm})
    assert np.allclose(c[0], wl, rtol=1e-14) and np.allclose(c[1], fl * 1e-2 * 1e3 / 10, rtol=1e-14)
    with pytest.raises(ValueError):
        tools.spectrum_to_cgs({'wavelength': wl, 'flux_lambda': fl, 'wavelength_unit': 'parsec-ish'})
