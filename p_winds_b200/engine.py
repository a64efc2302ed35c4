"""Batched GPU engine: the interface a retrieval uses (one call = a whole
(T0, Mdot, h_fraction) grid).  PyTorch supplies device buffers and streams only;
every computation happens in libpwinds_b200.so (C ABI: include/pwinds_b200.h).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

STATUS_MESSAGES = {
    1: 'The solver ``solve_ivp`` failed to obtain a solution.',   # hydrogen.py:459
    2: 'The solver ``odeint`` failed to obtain a solution.',      # helium.py:696
    # no reference counterpart: the reference's RK45 has no step limit and would crawl for minutes to hours
    5: 'The solver ``solve_ivp`` was given up after max_nfev right-hand-side evaluations '
       '(stiff start; the reference itself needs ~1e8 evaluations here before it fails).',
}
HE_METHODS = {'odeint': 0, 'LSODA': 1, 'RK45': 2}
O_METHODS = {'Radau': 0, 'odeint': 1, 'RK45': 2}
C_METHODS = {'odeint': 0, 'Radau': 1, 'RK45': 2}
LD_LAWS = {None: 0, 'linear': 1, 'quadratic': 2, 'square-root': 3, 'log': 4, 'exp': 5,
           's3': 6, 'c4': 7}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One engine = one GPU context (one per process / rank)."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError('p_winds_b200 needs a CUDA device (sm_100a); there is no CPU '
                               'fallback.')
        self.lib = _lib.load()
        self.device = torch.device('cuda', device)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        ctx = C.c_void_p()
        rc = self.lib.pwb_create(device, C.byref(ctx))
        if rc != 0:
            raise RuntimeError('pwb_create failed (%d)' % rc)
        self.ctx = ctx
        self._sed_key = None
        self._geo_key = None
        self._keep = {}

    def __del__(self):
        try:
            if getattr(self, 'ctx', None):
                self.lib.pwb_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.pwb_last_error(self.ctx).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype=torch.float64):
        """numpy / torch / scalar -> contiguous device tensor."""
        if isinstance(a, torch.Tensor):
            return a.to(self.device, dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def launch_count(self):
        return int(self.lib.pwb_launch_count(self.ctx))

    def n_active_pixels(self):
        """Terms of the pixel sum after merging equal-radius pixels (set_geometry)."""
        return int(self.lib.pwb_n_active_pixels(self.ctx))

    # ---------------------------------------------------------------------- SED
    def set_sed(self, wavelength_angstrom, flux_cgs):
        wl, fl = self.dev(wavelength_angstrom), self.dev(flux_cgs)
        self._check(self.lib.pwb_set_sed(self.ctx, _ptr(wl), _ptr(fl), wl.numel(), self._stream()))
        self._keep['sed'] = (wl, fl)
        self._sed_key = None   # an uncached upload invalidates what set_sed_cached remembers

    def set_sed_cached(self, wavelength_angstrom, flux_cgs):
        wl = np.ascontiguousarray(wavelength_angstrom, dtype=float)
        fl = np.ascontiguousarray(flux_cgs, dtype=float)
        key = (wl.shape, hash(wl.tobytes()), hash(fl.tobytes()))
        if key != self._sed_key:
            self.set_sed(wl, fl)
            self._sed_key = key

    def sed_scalars(self):
        out = (C.c_double * 32)()
        self._check(self.lib.pwb_get_sed_scalars(self.ctx, out))
        v = list(out)
        return {'h_phi': v[0], 'h_a0': v[1], 'he_phi_1': v[2], 'he_phi_3': v[3], 'he_a_1': v[4],
                'he_a_3': v[5], 'he_a_h_1': v[6], 'he_a_h_3': v[7], 'n_cut_h': int(v[8]),
                'i1': int(v[9]), 'i0': int(v[10]), 'o_phi': v[11], 'o_a': v[12], 'o_a_h': v[13],
                'o_a_he': v[14], 'c_phi_1': v[16], 'c_phi_2': v[17], 'c_a_1': v[18], 'c_a_2': v[19],
                'c_a_h_1': v[20], 'c_a_h_2': v[21], 'c_a_he': v[22], 'hei_phi': v[24], 'hei_a_he': v[25],
                'hei_a_h': v[6]}

    # ------------------------------------------------------------------- parker
    def parker_structure(self, r):
        r = self.dev(r)
        v, rho = torch.empty_like(r), torch.empty_like(r)
        self._check(self.lib.pwb_parker_structure(self.ctx, _ptr(r), r.numel(), _ptr(v), _ptr(rho),
                                                  self._stream()))
        return v, rho

    def parker_structure_tidal(self, r, sound_speed, r_sonic, planet_mass, star_mass, semimajor_axis):
        r = self.dev(r)
        v, rho = torch.empty_like(r), torch.empty_like(r)
        self._check(self.lib.pwb_parker_structure_tidal(
            self.ctx, _ptr(r), r.numel(), float(sound_speed), float(r_sonic), float(planet_mass),
            float(star_mass), float(semimajor_axis), _ptr(v), _ptr(rho), self._stream()))
        return v, rho

    # ----------------------------------------------------------------- hydrogen
    @staticmethod
    def h_opts(planet_radius, planet_mass, star_mass=1.0, semimajor_axis=1.0, initial_f_ion=0.0,
               relax_solution=False, convergence=0.01, max_n_relax=10, exact_phi=False,
               phi_abs=0.0, a_0=0.0, rtol=1e-3, atol=1e-6, first_step=None, max_step=None,
               structure_tidal=False, max_nfev=0):
        """`max_nfev`: RK45 right-hand-side evaluations per solve before a model is given up with
        status 5 (0 = the default of 2 000 000; < 0 = no limit, like the reference -- a GPU thread may then run
        for minutes on a model the reference itself needs hours to fail on)."""
        return _lib.HOpts(float(planet_radius), float(planet_mass), float(star_mass),
                          float(semimajor_axis), float(initial_f_ion), float(convergence),
                          int(max_n_relax), int(bool(relax_solution)), int(bool(exact_phi)),
                          float(phi_abs), float(a_0), float(rtol), float(atol),
                          float(first_step) if first_step else 0.0,
                          float(max_step) if (max_step and np.isfinite(max_step)) else 0.0,
                          int(bool(structure_tidal)), int(max_nfev))

    def h_ion_fraction_batch(self, T, mdot, h_fraction, mu0, r, opts, return_rates=False):
        T, mdot, hf, mu0, r = (self.dev(x).reshape(-1) for x in (T, mdot, h_fraction, mu0, r))
        B, nr = T.numel(), r.numel()
        f_r = torch.empty((B, nr), dtype=torch.float64, device=self.device)
        v, rho = torch.empty_like(f_r), torch.empty_like(f_r)
        scal = torch.empty((B, 8), dtype=torch.float64, device=self.device)
        status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.pwb_h_ion_fraction_batch(
            self.ctx, B, _ptr(T), _ptr(mdot), _ptr(hf), _ptr(mu0), _ptr(r), nr, C.byref(opts),
            _ptr(f_r), _ptr(v), _ptr(rho), _ptr(scal), _ptr(status), self._stream()))
        out = {'f_r': f_r, 'v': v, 'rho': rho, 'scal': scal, 'status': status}
        if return_rates:
            rates = torch.empty((B, 2, nr), dtype=torch.float64, device=self.device)
            self._check(self.lib.pwb_h_rates_batch(self.ctx, B, _ptr(T), _ptr(mdot), _ptr(hf), _ptr(mu0), _ptr(r),
                                                   nr, C.byref(opts), _ptr(rates), self._stream()))
            out['rates'] = rates
        return out

    # ------------------------------------------------------------------- helium
    @staticmethod
    def he_opts(planet_radius, rates, initial_state=(0.5, 0.5), relax_solution=False,
                convergence=0.01, max_n_relax=10, method='odeint', rtol=1e-3, atol=1e-6,
                first_step=None, max_step=None):
        if method not in HE_METHODS:
            raise ValueError("method %r is not available on the GPU path (supported: 'odeint', "
                             "'LSODA', 'RK45'); there is no CPU fallback" % (method,))
        return _lib.HeOpts(float(planet_radius), (C.c_double * 2)(*[float(x) for x in initial_state]),
                           int(bool(relax_solution)), int(max_n_relax), float(convergence),
                           (C.c_double * 6)(*[float(x) for x in rates]), HE_METHODS[method],
                           float(rtol), float(atol), float(first_step) if first_step else 0.0,
                           float(max_step) if (max_step and np.isfinite(max_step)) else 0.0)

    def he_population_batch(self, T, h_fraction, vs, rs, rhos, r, v, rho, f_h, opts,
                            return_rates=False, status=None):
        T, hf, vs, rs, rhos, r = (self.dev(x).reshape(-1) for x in (T, h_fraction, vs, rs, rhos, r))
        B, nr = T.numel(), r.numel()
        v, rho, f_h = (self.dev(x).reshape(B, nr) for x in (v, rho, f_h))
        f1 = torch.empty((B, nr), dtype=torch.float64, device=self.device)
        f3 = torch.empty_like(f1)
        scal = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        rates = torch.empty((B, 11, nr), dtype=torch.float64, device=self.device) if return_rates else None
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.pwb_he_population_batch(
            self.ctx, B, _ptr(T), _ptr(hf), _ptr(vs), _ptr(rs), _ptr(rhos), 1, _ptr(r), nr, _ptr(v),
            _ptr(rho), _ptr(f_h), C.byref(opts), _ptr(f1), _ptr(f3), _ptr(scal), _ptr(rates),
            _ptr(status), self._stream()))
        return {'f_1': f1, 'f_3': f3, 'scal': scal, 'rates': rates, 'status': status}

    # ------------------------------------------------------------------- oxygen
    @staticmethod
    def o_opts(planet_radius, rates, o_fraction=10 ** (8.69 - 12.00), initial_f_o_ion=0.0,
               relax_solution=False, convergence=0.01, max_n_relax=10, method='Radau', rtol=1e-3,
               atol=1e-6, first_step=None, max_step=None):
        if method not in O_METHODS:
            raise ValueError("method %r is not available on the GPU path (supported: 'Radau', 'odeint', "
                             "'RK45'); there is no CPU fallback" % (method,))
        return _lib.OOpts(float(planet_radius), float(o_fraction), float(initial_f_o_ion),
                          int(bool(relax_solution)), int(max_n_relax), float(convergence),
                          (C.c_double * 4)(*[float(x) for x in rates]), O_METHODS[method], float(rtol),
                          float(atol), float(first_step) if first_step else 0.0,
                          float(max_step) if (max_step and np.isfinite(max_step)) else 0.0)

    def o_ion_fraction_batch(self, T, h_fraction, vs, rs, rhos, r, v, rho, f_h, f_he_ion, opts,
                             return_rates=False, status=None):
        """oxygen.ion_fraction for B models (oxygen.py:186-491)."""
        T, hf, vs, rs, rhos, r = (self.dev(x).reshape(-1) for x in (T, h_fraction, vs, rs, rhos, r))
        B, nr = T.numel(), r.numel()
        v, rho, f_h, f_he = (self.dev(x).reshape(B, nr) for x in (v, rho, f_h, f_he_ion))
        f_o = torch.empty((B, nr), dtype=torch.float64, device=self.device)
        scal = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        rates = torch.empty((B, 5, nr), dtype=torch.float64, device=self.device) if return_rates else None
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.pwb_o_ion_fraction_batch(
            self.ctx, B, _ptr(T), _ptr(hf), _ptr(vs), _ptr(rs), _ptr(rhos), 1, _ptr(r), nr, _ptr(v),
            _ptr(rho), _ptr(f_h), _ptr(f_he), C.byref(opts), _ptr(f_o), _ptr(scal), _ptr(rates),
            _ptr(status), self._stream()))
        return {'f_o': f_o, 'scal': scal, 'rates': rates, 'status': status}

    def he_ion_fraction_batch(self, T, h_fraction, vs, rs, rhos, r, v, rho, f_h, opts, status=None):
        """helium.ion_fraction for B models (helium.py:789-1032); `opts` = Engine.o_opts(planet_radius,
        (phi_he, a_he, a_h, 0), initial_f_o_ion=initial_f_he_ion, ...)."""
        T, hf, vs, rs, rhos, r = (self.dev(x).reshape(-1) for x in (T, h_fraction, vs, rs, rhos, r))
        B, nr = T.numel(), r.numel()
        v, rho, f_h = (self.dev(x).reshape(B, nr) for x in (v, rho, f_h))
        f = torch.empty((B, nr), dtype=torch.float64, device=self.device)
        scal = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.pwb_he_ion_fraction_batch(
            self.ctx, B, _ptr(T), _ptr(hf), _ptr(vs), _ptr(rs), _ptr(rhos), 1, _ptr(r), nr, _ptr(v),
            _ptr(rho), _ptr(f_h), C.byref(opts), _ptr(f), _ptr(scal), _ptr(status), self._stream()))
        return {'f_he_ion': f, 'scal': scal, 'status': status}

    # ------------------------------------------------------------------- carbon
    @staticmethod
    def c_opts(planet_radius, rates, c_fraction=10 ** (8.43 - 12.00), initial_f_c_ion=(0.0, 0.0),
               relax_solution=False, convergence=0.01, max_n_relax=10, method='odeint', rtol=1e-3,
               atol=1e-6, first_step=None, max_step=None):
        if method not in C_METHODS:
            raise ValueError("method %r is not available on the GPU path (supported: 'odeint', 'Radau', "
                             "'RK45'); there is no CPU fallback" % (method,))
        return _lib.COpts(float(planet_radius), float(c_fraction),
                          (C.c_double * 2)(*[float(x) for x in initial_f_c_ion]), int(bool(relax_solution)),
                          int(max_n_relax), float(convergence), (C.c_double * 7)(*[float(x) for x in rates]),
                          C_METHODS[method], float(rtol), float(atol), float(first_step) if first_step else 0.0,
                          float(max_step) if (max_step and np.isfinite(max_step)) else 0.0)

    def c_ion_fraction_batch(self, T, h_fraction, vs, rs, rhos, r, v, rho, f_h, f_he_ion, opts,
                             return_rates=False, status=None):
        """carbon.ion_fraction for B models (carbon.py:256-639)."""
        T, hf, vs, rs, rhos, r = (self.dev(x).reshape(-1) for x in (T, h_fraction, vs, rs, rhos, r))
        B, nr = T.numel(), r.numel()
        v, rho, f_h, f_he = (self.dev(x).reshape(B, nr) for x in (v, rho, f_h, f_he_ion))
        f2 = torch.empty((B, nr), dtype=torch.float64, device=self.device)
        f3 = torch.empty_like(f2)
        scal = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        rates = torch.empty((B, 11, nr), dtype=torch.float64, device=self.device) if return_rates else None
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._check(self.lib.pwb_c_ion_fraction_batch(
            self.ctx, B, _ptr(T), _ptr(hf), _ptr(vs), _ptr(rs), _ptr(rhos), 1, _ptr(r), nr, _ptr(v),
            _ptr(rho), _ptr(f_h), _ptr(f_he), C.byref(opts), _ptr(f2), _ptr(f3), _ptr(scal), _ptr(rates),
            _ptr(status), self._stream()))
        return {'f_cii': f2, 'f_ciii': f3, 'scal': scal, 'rates': rates, 'status': status}

    # ------------------------------------------------------------------ transit
    def set_geometry(self, intensity_0, r_from_planet, radius_profile_si):
        i0, rm, r = self.dev(intensity_0).reshape(-1), self.dev(r_from_planet).reshape(-1), \
            self.dev(radius_profile_si).reshape(-1)
        if i0.numel() != rm.numel():
            raise ValueError('intensity map and radius map must have the same shape')
        self._check(self.lib.pwb_set_geometry(self.ctx, _ptr(i0), _ptr(rm), i0.numel(), _ptr(r),
                                              r.numel(), self._stream()))
        self._geo_nr = r.numel()
        self._geo_key = None   # an uncached upload invalidates what set_geometry_cached remembers

    def set_geometry_phases(self, maps, radius_profile_si):
        """Phase-averaged transit (advanced_tutorial.ipynb cell 11): `maps` is a sequence of
        (intensity_0, r_from_planet, transit_depth), one per sampled phase.  Afterwards the
        spectra of rt2d_batch / forward_batch are mean_phase(sum_pixels I0 e^-tau + depth)."""
        r = self.dev(radius_profile_si).reshape(-1)
        if not 1 <= len(maps) <= 16:
            raise ValueError('1 to 16 phases')
        for slot, (i0, rm, depth) in enumerate(maps):
            i0, rm = self.dev(i0).reshape(-1), self.dev(rm).reshape(-1)
            if i0.numel() != rm.numel():
                raise ValueError('intensity map and radius map must have the same shape')
            self._check(self.lib.pwb_set_geometry_phase(self.ctx, slot, _ptr(i0), _ptr(rm), i0.numel(),
                                                        _ptr(r), r.numel(), float(depth), self._stream()))
        self._check(self.lib.pwb_set_phase_count(self.ctx, len(maps)))
        self._geo_nr = r.numel()
        self._geo_key = None

    def convolve_extend_batch(self, spectrum, kernel):
        """astropy.convolution.convolve(spec, kernel, boundary='extend') for every row of
        `spectrum` (advanced_tutorial.ipynb cell 13).  The kernel is normalised to unit sum
        here, with numpy, as astropy does before its C loop."""
        spectrum = self.dev(spectrum)
        if spectrum.ndim == 1:
            spectrum = spectrum[None, :]
        k = np.ascontiguousarray(kernel, dtype=float).ravel()
        if k.size % 2 == 0:
            raise ValueError('Kernel size must be odd in all axes.')   # astropy's KernelSizeError text
        k = k / k.sum()
        kd = self.dev(k)
        out = torch.empty_like(spectrum)
        B, nw = spectrum.shape
        self._check(self.lib.pwb_convolve_extend_batch(self.ctx, B, _ptr(spectrum), nw, _ptr(kd), k.size,
                                                       _ptr(out), self._stream()))
        return out

    def loglike_batch(self, model, y, yerr, status=None):
        """-0.5 * sum((y - model)^2 / yerr^2 + log(yerr^2)) per model (advanced_tutorial.ipynb
        cell 15); -inf for failed models."""
        model = self.dev(model)
        B, nw = model.shape
        y, yerr = self.dev(y).reshape(-1), self.dev(yerr).reshape(-1)
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        self._check(self.lib.pwb_loglike_batch(self.ctx, B, _ptr(model), nw, _ptr(y), _ptr(yerr), _ptr(status),
                                               _ptr(out), self._stream()))
        return out

    def set_geometry_cached(self, intensity_0, r_from_planet, radius_profile_si):
        a, b, c = (np.ascontiguousarray(x, dtype=float) for x in
                   (intensity_0, r_from_planet, radius_profile_si))
        key = (a.shape, hash(a.tobytes()), hash(b.tobytes()), hash(c.tobytes()))
        if key != self._geo_key:
            self.set_geometry(a, b, c)
            self._geo_key = key

    @staticmethod
    def rt_opts(central_wavelength, oscillator_strength, einstein_coefficient, particle_mass,
                bulk_los_velocity=0.0, planet_radial_velocity=0.0, wind_broadening_method='average',
                z_grid_size=200, turbulence_broadening=False, temperature_is_profile=False):
        w0 = np.atleast_1d(np.asarray(central_wavelength, dtype=float))
        f = np.atleast_1d(np.asarray(oscillator_strength, dtype=float))
        a = np.atleast_1d(np.asarray(einstein_coefficient, dtype=float))
        if not (len(w0) == len(f) == len(a)) or len(w0) > 8:
            raise ValueError('The spectral line properties must be float, list or numpy.ndarray, '
                             'and their format must be consistent with each other (<= 8 lines).')
        if wind_broadening_method not in ('average', 'formal'):
            raise ValueError('The chosen ``wind_broadening_method`` is not implemented.')
        pad = lambda x: (C.c_double * 8)(*(list(x) + [0.0] * (8 - len(x))))  # noqa: E731
        return _lib.RtOpts(len(w0), pad(w0), pad(f), pad(a), float(particle_mass),
                           float(bulk_los_velocity), float(planet_radial_velocity),
                           0 if wind_broadening_method == 'average' else 1, int(z_grid_size),
                           int(turbulence_broadening is True), int(bool(temperature_is_profile)))

    def rt2d_batch(self, density, velocity, temperature, wavelength_grid, opts, status=None,
                   return_tau=False):
        wl = self.dev(wavelength_grid).reshape(-1)
        nr = self._geo_nr
        n = self.dev(density).reshape(-1, nr)
        v = self.dev(velocity).reshape(-1, nr)
        B, nw = n.shape[0], wl.numel()
        T = self.dev(temperature).reshape(B, -1)
        spec = torch.empty((B, nw), dtype=torch.float64, device=self.device)
        tau = torch.empty((B, nr, nw), dtype=torch.float64, device=self.device) if return_tau else None
        self._check(self.lib.pwb_rt2d_batch(self.ctx, B, _ptr(n), _ptr(v), _ptr(T), _ptr(wl), nw,
                                            C.byref(opts), _ptr(status), _ptr(spec), _ptr(tau),
                                            self._stream()))
        return (spec, tau) if return_tau else spec

    # -------------------------------------------------------------- whole model
    def forward_batch(self, T, mdot, h_fraction, r, h_opts, he_opts, rt_opts, wavelength_grid,
                      species='he3', out=None):
        """The docs' cascading model for a batch.  Returns a dict of device tensors."""
        T, mdot, hf, r, wl = (self.dev(x).reshape(-1) for x in (T, mdot, h_fraction, r, wavelength_grid))
        B, nr, nw = T.numel(), r.numel(), wl.numel()
        o = out if out is not None else self.alloc_forward(B, nr, nw)
        self._check(self.lib.pwb_forward_batch(
            self.ctx, B, _ptr(T), _ptr(mdot), _ptr(hf), _ptr(r), nr, C.byref(h_opts),
            C.byref(he_opts), C.byref(rt_opts), 0 if species == 'he3' else 1, _ptr(wl), nw,
            _ptr(o['spectrum']), _ptr(o['f_r']), _ptr(o['f_1']), _ptr(o['f_3']), _ptr(o['v']),
            _ptr(o['rho']), _ptr(o['hscal']), _ptr(o['hescal']), _ptr(o['status']), self._stream()))
        return o

    def alloc_forward(self, B, nr, nw):
        e = lambda *s: torch.empty(s, dtype=torch.float64, device=self.device)  # noqa: E731
        return {'spectrum': e(B, nw), 'f_r': e(B, nr), 'f_1': e(B, nr), 'f_3': e(B, nr),
                'v': e(B, nr), 'rho': e(B, nr), 'hscal': e(B, 8), 'hescal': e(B, 4),
                'status': torch.zeros(B, dtype=torch.int32, device=self.device)}

    def set_relax_warning_policy(self, accept):
        """How chi2_batch / loglike_batch score models of status 3 (an odeint of the helium relax
        loop warned; the reference's own result is undefined there).  False (default): failed
        model, +inf / -inf.  True: score the spectrum of the last completed relax pass."""
        self._check(self.lib.pwb_set_relax_warning_policy(self.ctx, 1 if accept else 0))

    def chi2_batch(self, spectrum, data, sigma, norm=1.0, status=None):
        spectrum = self.dev(spectrum)
        B, nw = spectrum.shape
        data, sigma = self.dev(data).reshape(-1), self.dev(sigma).reshape(-1)
        if status is not None:
            status = self.dev(status, torch.int32).reshape(-1)
        chi2 = torch.empty(B, dtype=torch.float64, device=self.device)
        self._check(self.lib.pwb_chi2_batch(self.ctx, B, _ptr(spectrum), nw, _ptr(data), _ptr(sigma),
                                            float(norm), _ptr(status), _ptr(chi2), self._stream()))
        return chi2

    def draw_transit(self, planet_to_star_ratio, planet_physical_radius, impact_parameter=0.0,
                     phase=0.0, grid_size=100, supersampling=None, limb_darkening_law=None,
                     ld_coefficient=None):
        ss = 1 if supersampling is None else int(round(supersampling))
        if supersampling is not None and abs(supersampling - ss) > 1e-12:
            raise ValueError('the GPU rasteriser supports integer supersampling factors')
        if limb_darkening_law not in LD_LAWS:
            raise ValueError('unknown limb-darkening law %r' % (limb_darkening_law,))
        c = np.zeros(4)
        if ld_coefficient is not None:
            cc = np.atleast_1d(np.asarray(ld_coefficient, dtype=float))
            c[:len(cc)] = cc
        i0 = torch.empty((grid_size, grid_size), dtype=torch.float64, device=self.device)
        rm = torch.empty_like(i0)
        depth = C.c_double()
        self._check(self.lib.pwb_draw_transit(
            self.ctx, float(planet_to_star_ratio), float(planet_physical_radius),
            float(impact_parameter), float(phase), int(grid_size), ss,
            LD_LAWS[limb_darkening_law], (C.c_double * 4)(*c), _ptr(i0), _ptr(rm), C.byref(depth),
            self._stream()))
        return i0, depth.value, rm

    def selftest_exp(self, x):
        """(pw_exp_neg_wide(x), pw_exp_neg_sum_sh(x)) of the pixel-sum kernel (test hook)."""
        x = self.dev(x).reshape(-1)
        w, g = torch.empty_like(x), torch.empty_like(x)
        self._check(self.lib.pwb_selftest_exp(self.ctx, x.numel(), _ptr(x), _ptr(w), _ptr(g), self._stream()))
        return w, g

    def selftest_arith(self, a, b, k):
        """(a / b, a ** (1 / k)) through the device routines of the LSODA step (test hook)."""
        a = self.dev(a); b = self.dev(b)
        kk = torch.as_tensor(np.ascontiguousarray(k, dtype=np.int32), device=self.device)
        q = torch.empty_like(a); r = torch.empty_like(a)
        self._check(self.lib.pwb_selftest_arith(self.ctx, a.numel(), _ptr(a), _ptr(b), _ptr(kk), _ptr(q), _ptr(r),
                                                self._stream()))
        return q, r

    def fp64_peak_tflops(self):
        t = C.c_double()
        self._check(self.lib.pwb_fp64_peak(self.ctx, C.byref(t), self._stream()))
        return t.value


_default = {}


def default_engine(device=None):
    """Process-wide engine used by the reference-signature functions."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
