"""ctypes binding of libpwinds_b200.so (C ABI: include/pwinds_b200.h).

Fails loudly when the CUDA library has not been built: there is no CPU path.
"""
import ctypes as C
import os

LIB_PATH = os.environ.get('PWB_LIB_PATH') or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), 'libpwinds_b200.so')

c_double_p = C.POINTER(C.c_double)


class HOpts(C.Structure):
    _fields_ = [('planet_radius', C.c_double), ('planet_mass', C.c_double),
                ('star_mass', C.c_double), ('semimajor_axis', C.c_double),
                ('initial_f_ion', C.c_double), ('convergence', C.c_double),
                ('max_n_relax', C.c_int), ('relax_solution', C.c_int),
                ('exact_phi', C.c_int), ('phi_abs', C.c_double), ('a_0', C.c_double),
                ('rtol', C.c_double), ('atol', C.c_double), ('first_step', C.c_double),
                ('max_step', C.c_double), ('structure_tidal', C.c_int), ('max_nfev', C.c_longlong)]


class HeOpts(C.Structure):
    _fields_ = [('planet_radius', C.c_double), ('initial_state', C.c_double * 2),
                ('relax_solution', C.c_int), ('max_n_relax', C.c_int),
                ('convergence', C.c_double), ('rates', C.c_double * 6),
                ('method', C.c_int), ('rtol', C.c_double), ('atol', C.c_double),
                ('first_step', C.c_double), ('max_step', C.c_double)]


class OOpts(C.Structure):
    _fields_ = [('planet_radius', C.c_double), ('o_fraction', C.c_double), ('initial_f_o_ion', C.c_double),
                ('relax_solution', C.c_int), ('max_n_relax', C.c_int), ('convergence', C.c_double),
                ('rates', C.c_double * 4), ('method', C.c_int), ('rtol', C.c_double), ('atol', C.c_double),
                ('first_step', C.c_double), ('max_step', C.c_double)]


class COpts(C.Structure):
    _fields_ = [('planet_radius', C.c_double), ('c_fraction', C.c_double), ('initial_f_c_ion', C.c_double * 2),
                ('relax_solution', C.c_int), ('max_n_relax', C.c_int), ('convergence', C.c_double),
                ('rates', C.c_double * 7), ('method', C.c_int), ('rtol', C.c_double), ('atol', C.c_double),
                ('first_step', C.c_double), ('max_step', C.c_double)]


class RtOpts(C.Structure):
    _fields_ = [('n_lines', C.c_int), ('central_wavelength', C.c_double * 8),
                ('oscillator_strength', C.c_double * 8),
                ('einstein_coefficient', C.c_double * 8), ('particle_mass', C.c_double),
                ('bulk_los_velocity', C.c_double), ('planet_radial_velocity', C.c_double),
                ('wind_broadening_method', C.c_int), ('z_grid_size', C.c_int),
                ('turbulence_broadening', C.c_int), ('temperature_is_profile', C.c_int)]


# every symbol declared in include/pwinds_b200.h: (restype, argtypes)
_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
SYMBOLS = {
    'pwb_create': (_I, [_I, C.POINTER(_VP)]),
    'pwb_destroy': (None, [_VP]),
    'pwb_last_error': (C.c_char_p, [_VP]),
    'pwb_version': (C.c_char_p, []),
    'pwb_sizeof_opts': (_I, [_I]),
    'pwb_launch_count': (C.c_longlong, [_VP]),
    'pwb_n_active_pixels': (_I, [_VP]),
    'pwb_set_sed': (_I, [_VP, _VP, _VP, _I, _VP]),
    'pwb_get_sed_scalars': (_I, [_VP, c_double_p]),
    'pwb_parker_structure': (_I, [_VP, _VP, _I, _VP, _VP, _VP]),
    'pwb_parker_structure_tidal': (_I, [_VP, _VP, _I, _D, _D, _D, _D, _D, _VP, _VP, _VP]),
    'pwb_h_ion_fraction_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, C.POINTER(HOpts),
                                      _VP, _VP, _VP, _VP, _VP, _VP]),
    'pwb_h_rates_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, C.POINTER(HOpts), _VP, _VP]),
    'pwb_he_population_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP,
                                     C.POINTER(HeOpts), _VP, _VP, _VP, _VP, _VP, _VP]),
    'pwb_o_ion_fraction_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP, _VP,
                                      C.POINTER(OOpts), _VP, _VP, _VP, _VP, _VP]),
    'pwb_he_ion_fraction_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP,
                                       C.POINTER(OOpts), _VP, _VP, _VP, _VP]),
    'pwb_c_ion_fraction_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP, _VP, _VP,
                                      C.POINTER(COpts), _VP, _VP, _VP, _VP, _VP, _VP]),
    'pwb_set_geometry': (_I, [_VP, _VP, _VP, _I, _VP, _I, _VP]),
    'pwb_set_geometry_phase': (_I, [_VP, _I, _VP, _VP, _I, _VP, _I, _D, _VP]),
    'pwb_set_phase_count': (_I, [_VP, _I]),
    'pwb_convolve_extend_batch': (_I, [_VP, _I, _VP, _I, _VP, _I, _VP, _VP]),
    'pwb_loglike_batch': (_I, [_VP, _I, _VP, _I, _VP, _VP, _VP, _VP, _VP]),
    'pwb_rt2d_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, C.POINTER(RtOpts), _VP, _VP, _VP, _VP]),
    'pwb_forward_batch': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _I, C.POINTER(HOpts), C.POINTER(HeOpts),
                               C.POINTER(RtOpts), _I, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                               _VP, _VP, _VP]),
    'pwb_set_relax_warning_policy': (_I, [_VP, _I]),
    'pwb_chi2_batch': (_I, [_VP, _I, _VP, _I, _VP, _VP, _D, _VP, _VP, _VP]),
    'pwb_draw_transit': (_I, [_VP, _D, _D, _D, _D, _I, _I, _I, c_double_p, _VP, _VP, c_double_p, _VP]),
    'pwb_fp64_peak': (_I, [_VP, c_double_p, _VP]),
    'pwb_selftest_exp': (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    'pwb_selftest_arith': (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
}

_lib = None


def load():
    """Load the engine library, binding every symbol of the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'p_winds_b200: %s is missing. Build the sm_100a engine first '
            '(`./build.sh` or `python -c "import __graft_entry__ as g; g.build()"`). '
            'There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the ABI and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
