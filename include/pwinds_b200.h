/* pwinds_b200.h -- C ABI of the B200 forward-model engine for the p-winds
 * retrieval hot path.
 *
 * The reference (ladsantos/p-winds v1.4.7) is pure Python: there is no FFI to
 * inherit, the "plugin boundary" is the set of module functions named in
 * BASELINE.json.  Each entry point below is what a ctypes binding placed behind one
 * of those functions would call (see INTEGRATION.md for the stubs); the citation on
 * each declaration is the reference function it replaces.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in the signatures: device buffers are passed
 *     as `void*`-compatible `double*` / `int*` device addresses (e.g. from
 *     torch.Tensor.data_ptr()), streams as `void*` (cudaStream_t, NULL = default);
 *   - every function returns 0 on success, < 0 on error; pwb_last_error() returns
 *     the message;  per-model solver outcomes are NOT errors: they are reported in
 *     the `status` output array (enum pwb_status) so that a retrieval grid can
 *     reproduce *which* models the reference would have raised RuntimeError for;
 *   - all arithmetic is FP64; arrays are C-contiguous; batched arrays are
 *     [B][Nr] / [B][Nw] row-major;
 *   - nothing here runs on the CPU except argument checking and launches.
 */
#ifndef PWINDS_B200_H
#define PWINDS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pwb_ctx pwb_ctx;

enum pwb_status {
  PWB_OK = 0,
  PWB_FAIL_H_SOLVER = 1,   /* hydrogen.py:458-460, :511-513 RuntimeError */
  PWB_FAIL_HE_SOLVER = 2,  /* helium.py:693-697, :708-710 RuntimeError */
  /* helium.py:736: an odeint of the relax loop warned.  The reference does not check that call and goes on
   * with rows scipy left uninitialised (its result is not reproducible); the engine stops at the warning:
   * profiles / spectrum of the last completed pass, n_relax = the pass that warned.  chi^2 and
   * log-likelihood treat the model as failed unless pwb_set_relax_warning_policy(ctx, 1). */
  PWB_WARN_HE_RELAX = 3,
  PWB_FAIL_NAN = 4,
  /* The hydrogen RK45 solve was given up after pwb_h_opts.max_nfev right-hand-side evaluations.
   * No reference equivalent: scipy's RK45 has no step limit; on a stiff start (cold, dense winds,
   * e.g. T0 = 5000 K, Mdot = 7e9 g/s, h = 0.85, tidal) the reference crawls for 1.9e8
   * evaluations -- hours -- before it raises its RuntimeError. */
  PWB_FAIL_WORK_LIMIT = 5
};

enum pwb_he_method { PWB_HE_ODEINT = 0, PWB_HE_LSODA = 1, PWB_HE_RK45 = 2 };

/* Options of hydrogen.ion_fraction (hydrogen.py:227-232). */
typedef struct {
  double planet_radius;        /* R_Jup */
  double planet_mass;          /* M_Jup */
  double star_mass;            /* M_sun (1.0 = reference default) */
  double semimajor_axis;       /* au    (1.0 = reference default) */
  double initial_f_ion;
  double convergence;
  int max_n_relax;
  int relax_solution;
  int exact_phi;
  double phi_abs, a_0;         /* used when !exact_phi (radiative_processes[_mono]) */
  double rtol, atol;           /* solve_ivp RK45 options (1e-3, 1e-6 = scipy defaults) */
  double first_step, max_step; /* <= 0: scipy default */
  int structure_tidal;         /* Parker structure written for the He stage: 0 plain, 1 tidal */
  long long max_nfev;          /* RK45 right-hand-side evaluations per solve before the model is given
                                * up with PWB_FAIL_WORK_LIMIT; 0 = the default (2 000 000, ~1000x a
                                * typical solve); < 0 = no limit, like scipy */
} pwb_h_opts;

/* Options of helium.population_fraction (helium.py:413-421). */
typedef struct {
  double planet_radius;
  double initial_state[2];
  int relax_solution;
  int max_n_relax;
  double convergence;
  double rates[6];             /* phi_1, phi_3, a_1, a_3, a_h_1, a_h_3 (helium.py:154 / :266) */
  int method;                  /* enum pwb_he_method */
  double rtol, atol, first_step, max_step;
} pwb_he_opts;

/* Options of oxygen.ion_fraction (oxygen.py:186-191). */
enum pwb_o_method { PWB_O_RADAU = 0, PWB_O_ODEINT = 1, PWB_O_RK45 = 2 };
typedef struct {
  double planet_radius;
  double o_fraction;           /* 10 ** (8.69 - 12) = solar (oxygen.py:21-22) */
  double initial_f_o_ion;
  int relax_solution;
  int max_n_relax;
  double convergence;
  double rates[4];             /* phi_oi, a_oi, a_h_oi, a_he (oxygen.radiative_processes, oxygen.py:99) */
  int method;                  /* enum pwb_o_method; 'Radau' is the reference default */
  double rtol, atol, first_step, max_step;
} pwb_o_opts;

/* Options of helium.ion_fraction (helium.py:789-793): the pwb_o_opts block with
 * rates[0..2] = phi_he, a_he, a_h of helium.radiative_processes(combined_ionization=True),
 * initial_f_o_ion = initial_f_he_ion; o_fraction is ignored. */
typedef pwb_o_opts pwb_heion_opts;

/* Options of carbon.ion_fraction (carbon.py:256-262). */
enum pwb_c_method { PWB_C_ODEINT = 0, PWB_C_RADAU = 1, PWB_C_RK45 = 2 };
typedef struct {
  double planet_radius;
  double c_fraction;           /* 10 ** (8.43 - 12) = solar (carbon.py:21-22) */
  double initial_f_c_ion[2];
  int relax_solution;
  int max_n_relax;
  double convergence;
  double rates[7];             /* phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he (carbon.py:131) */
  int method;                  /* enum pwb_c_method; 'odeint' is the reference default */
  double rtol, atol, first_step, max_step;
} pwb_c_opts;

/* Options of transit.radiative_transfer_2d (transit.py:120-126). */
typedef struct {
  int n_lines;                       /* <= 8 */
  double central_wavelength[8];      /* m */
  double oscillator_strength[8];
  double einstein_coefficient[8];    /* 1/s */
  double particle_mass;              /* kg */
  double bulk_los_velocity;          /* m/s */
  double planet_radial_velocity;     /* m/s */
  int wind_broadening_method;        /* 0 'average', 1 'formal' */
  int z_grid_size;
  int turbulence_broadening;
  int temperature_is_profile;        /* gas_temperature is an ndarray[Nr] per model */
} pwb_rt_opts;

/* ---- context ----------------------------------------------------------------- */
int pwb_create(int device, pwb_ctx** out);
void pwb_destroy(pwb_ctx* ctx);
const char* pwb_last_error(pwb_ctx* ctx);
/* "p_winds_b200 <version> (sm_100a) src:<sha256[:16] of csrc/ and this header>" */
const char* pwb_version(void);
/* sizeof the option blocks as compiled (0 pwb_h_opts, 1 pwb_he_opts, 2 pwb_o_opts, 3 pwb_c_opts,
 * 4 pwb_rt_opts): lets a binding check its struct layout against the library without a GPU */
int pwb_sizeof_opts(int which);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
long long pwb_launch_count(pwb_ctx* ctx);
/* terms of the pixel sum of the current geometry (lit, in-range pixels after equal-radius
 * pixels have been merged by pwb_set_geometry); 0 before pwb_set_geometry */
int pwb_n_active_pixels(pwb_ctx* ctx);

/* ---- per-context SED (spectrum dict after unit conversion: tools.py:405-408,
 *      hydrogen.py:57-61).  Builds sigma_H, sigma_He tables and Simpson weights on
 *      the device and evaluates hydrogen.radiative_processes (hydrogen.py:118-169)
 *      and helium.radiative_processes (helium.py:24-154).
 *      d_wavelength [angstrom], d_flux [erg s-1 cm-2 A-1], device pointers. */
int pwb_set_sed(pwb_ctx* ctx, const double* d_wavelength, const double* d_flux, int n, void* stream);
/* host out[32]: phi_H, a_0, phi_1, phi_3, a_1, a_3, a_h_1, a_h_3, n_cut_H, i1, i0,
 *      [11..14] oxygen.radiative_processes (phi_oi, a_oi, a_h_oi, a_he; oxygen.py:99),
 *      [16..22] carbon.radiative_processes (phi_ci, phi_cii, a_ci, a_cii, a_h_ci, a_h_cii, a_he; carbon.py:131),
 *      [24..25] helium.radiative_processes(combined_ionization=True): phi, a_he (a_h = out[6]) */
int pwb_get_sed_scalars(pwb_ctx* ctx, double* out32);

/* ---- parker.structure (parker.py:163-223) / structure_tidal (:271-358) --------- */
int pwb_parker_structure(pwb_ctx* ctx, const double* d_r, int n, double* d_v, double* d_rho, void* stream);
int pwb_parker_structure_tidal(pwb_ctx* ctx, const double* d_r, int n, double sound_speed,
                               double r_sonic, double planet_mass, double star_mass,
                               double semimajor_axis, double* d_v, double* d_rho, void* stream);

/* ---- hydrogen.ion_fraction, batched (hydrogen.py:227-573) -----------------------
 * inputs : d_T, d_mdot, d_hfrac, d_mu0 [B]; d_r [Nr] radius profile in planetary radii
 * outputs: d_f_r [B][Nr]; d_v, d_rho [B][Nr] final Parker structure in sonic units
 *          (what parker.structure[_tidal] returns for mu_bar; quickstart cell 7);
 *          d_scal [B][8] = mu_bar, vs, rs, rhos, n_relax, nfev, n_solves, 0;
 *          d_status [B] (enum pwb_status) */
int pwb_h_ion_fraction_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_mdot,
                             const double* d_hfrac, const double* d_mu0, const double* d_r, int Nr,
                             const pwb_h_opts* opts, double* d_f_r, double* d_v, double* d_rho,
                             double* d_scal, int* d_status, void* stream);

/* ---- hydrogen.ion_fraction(return_rates=True) (hydrogen.py:538-561): call right after
 *      pwb_h_ion_fraction_batch with the same batch and options (exact_phi required, as in the
 *      reference).  d_rates [B][2][Nr] = photoionization, recombination (1/s); NaN for failed models. */
int pwb_h_rates_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_mdot, const double* d_hfrac,
                      const double* d_mu0, const double* d_r, int Nr, const pwb_h_opts* opts, double* d_rates,
                      void* stream);

/* ---- helium.population_fraction, batched (helium.py:413-785) --------------------
 * inputs : d_T, d_hfrac, d_vs, d_rs, d_rhos [B] (d_vs.. may alias columns of d_scal
 *          with stride `scal_stride` doubles; pass 1 for dense arrays);
 *          d_r [Nr]; d_v, d_rho, d_f_h [B][Nr]
 * outputs: d_f1, d_f3 [B][Nr]; d_scal [B][4] = n_relax, nfe, nst, nje;
 *          d_rates [B][11][Nr] or NULL (return_rates); d_status [B] is read first
 *          (models already failed are skipped) and updated. */
int pwb_he_population_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_hfrac,
                            const double* d_vs, const double* d_rs, const double* d_rhos,
                            int scal_stride, const double* d_r, int Nr, const double* d_v,
                            const double* d_rho, const double* d_f_h, const pwb_he_opts* opts,
                            double* d_f1, double* d_f3, double* d_scal, double* d_rates,
                            int* d_status, void* stream);

/* ---- oxygen.ion_fraction, batched (oxygen.py:186-491) -----------
 * Same per-model inputs as the helium stage plus d_f_he_ion [B][Nr] = 1 - f_1 - f_3.
 * outputs: d_f_o [B][Nr]; d_scal [B][4] = n_relax, nfev, n_solves, -; d_rates (may be NULL)
 *          [B][5][Nr] = photoionization, recombination, e_impact_ionization,
 *          charge_exchange_HII, charge_exchange_HI (1/s); d_status in/out as for helium
 *          (status 2 = the RuntimeError of oxygen.py:419-433). */
int pwb_o_ion_fraction_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_hfrac,
                             const double* d_vs, const double* d_rs, const double* d_rhos, int sstride,
                             const double* d_r, int Nr, const double* d_v, const double* d_rho,
                             const double* d_f_h, const double* d_f_he_ion, const pwb_o_opts* opts,
                             double* d_f_o, double* d_scal, double* d_rates, int* d_status, void* stream);

/* ---- helium.ion_fraction, batched (helium.py:789-1032) -------------
 * outputs: d_f_he_ion [B][Nr]; d_scal [B][4] = n_relax, nfev, n_solves, - */
int pwb_he_ion_fraction_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_hfrac,
                              const double* d_vs, const double* d_rs, const double* d_rhos, int sstride,
                              const double* d_r, int Nr, const double* d_v, const double* d_rho,
                              const double* d_f_h, const pwb_heion_opts* opts, double* d_f_he_ion,
                              double* d_scal, int* d_status, void* stream);

/* ---- carbon.ion_fraction, batched (carbon.py:256-639) -------------
 * outputs: d_f_cii, d_f_ciii [B][Nr]; d_scal [B][4] = n_relax, nfev, n_solves, -; d_rates (may be
 *          NULL) [B][11][Nr] in the order of carbon.py:531-533; status 2 = the RuntimeError
 *          of carbon.py:541-557 (the default 'odeint' gives up on dense winds). */
int pwb_c_ion_fraction_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_hfrac,
                             const double* d_vs, const double* d_rs, const double* d_rhos, int sstride,
                             const double* d_r, int Nr, const double* d_v, const double* d_rho,
                             const double* d_f_h, const double* d_f_he_ion, const pwb_c_opts* opts,
                             double* d_f_cii, double* d_f_ciii, double* d_scal, double* d_rates,
                             int* d_status, void* stream);

/* ---- transit geometry (output of transit.draw_transit, transit.py:113-116) plus the
 *      radius grid of the RT call; builds the per-pixel interpolation table used by
 *      interp1d(radius_profile, tau, axis=0, fill_value=0)(r_from_planet)
 *      (transit.py:220-222).  d_i0, d_r_map [npix]; d_r_si [Nr]. */
int pwb_set_geometry(pwb_ctx* ctx, const double* d_i0, const double* d_r_map, int npix,
                     const double* d_r_si, int Nr, void* stream);

/* ---- phase-averaged transit (docs/source/advanced_tutorial.ipynb cell 11,
 *      `transmission_model`): one geometry per sampled phase, slot 0..15, with the geometric
 *      transit depth draw_transit returned for it.  After pwb_set_phase_count(n) the pixel
 *      sum of pwb_rt2d_batch / pwb_forward_batch is
 *          mean_{phase < n} ( sum_pixels I0 exp(-tau) + transit_depth_phase ),
 *      i.e. numpy.mean(spectra, axis=0) of the notebook.  pwb_set_geometry = slot 0, depth 0,
 *      one phase (transit.radiative_transfer_2d itself). */
int pwb_set_geometry_phase(pwb_ctx* ctx, int slot, const double* d_i0, const double* d_r_map, int npix,
                           const double* d_r_si, int Nr, double transit_depth, void* stream);
int pwb_set_phase_count(pwb_ctx* ctx, int n_phases);

/* ---- transit.radiative_transfer_2d, batched (transit.py:120-230) ----------------
 * inputs : d_n [B][Nr] number density (m^-3), d_v [B][Nr] velocity (m/s),
 *          d_T [B] (or [B][Nr] if temperature_is_profile), d_wl [Nw] (m);
 *          d_status may be NULL; failed models get NaN spectra.
 * output : d_spectrum [B][Nw]; d_tau [B][Nr][Nw] optical depth table or NULL
 *          (transit.optical_depth_2d, transit.py:312-536) */
int pwb_rt2d_batch(pwb_ctx* ctx, int B, const double* d_n, const double* d_v, const double* d_T,
                   const double* d_wl, int Nw, const pwb_rt_opts* opts, const int* d_status,
                   double* d_spectrum, double* d_tau, void* stream);

/* ---- whole forward model, batched: the docs' cascading_model
 *      (docs/source/quickstart.ipynb cells 4-12; advanced_tutorial.ipynb:226-345):
 *      ion_fraction -> structure -> population_fraction -> n(He 2^3S) ->
 *      radiative_transfer_2d.  species: 0 = He 2^3S (lines of the caller), 1 = neutral H
 *      (Ly-alpha, BASELINE config 2: no helium stage).
 * outputs as for the stage calls; any output pointer except d_spectrum and d_status
 * may be NULL (internal scratch is used). */
int pwb_forward_batch(pwb_ctx* ctx, int B, const double* d_T, const double* d_mdot,
                      const double* d_hfrac, const double* d_r, int Nr, const pwb_h_opts* hopts,
                      const pwb_he_opts* heopts, const pwb_rt_opts* rtopts, int species,
                      const double* d_wl, int Nw, double* d_spectrum, double* d_f_r, double* d_f1,
                      double* d_f3, double* d_v, double* d_rho, double* d_hscal, double* d_hescal,
                      int* d_status, void* stream);

/* ---- how chi^2 / log-likelihood score a model of status 3 (an `odeint` of the helium relax
 *      loop warned, helium.py:736).  The reference does not check that call and carries on
 *      with rows scipy left uninitialised, so its own result for such a model is not
 *      reproducible; the engine stops at the warning and returns the last completed pass.
 *      accept = 0 (default): status 3 is a failed model (+inf / -inf);
 *      accept = 1: the spectrum of the last completed pass is scored like any other. */
int pwb_set_relax_warning_policy(pwb_ctx* ctx, int accept);

/* ---- chi^2 of each model spectrum against data (advanced_tutorial.ipynb:432-438):
 *      chi2[b] = sum_k ((data_k - model_bk / norm) / sigma_k)^2 ; failed models -> +inf */
int pwb_chi2_batch(pwb_ctx* ctx, int B, const double* d_spectrum, int Nw, const double* d_data,
                   const double* d_sigma, double norm, const int* d_status, double* d_chi2,
                   void* stream);

/* ---- instrumental broadening + likelihood of the retrieval loop (advanced_tutorial.ipynb
 *      cells 13 and 15) ------------------------------------------------------------------
 * pwb_convolve_extend_batch: astropy.convolution.convolve(spec_b, kernel, boundary='extend')
 *      for every model; d_kernel [Nk], Nk odd, already normalised to unit sum
 *      (normalize_kernel=True); d_out [B][Nw] may not alias d_spectrum.
 * pwb_loglike_batch: out[b] = -0.5 * sum_k ((y_k - model_bk)^2 / yerr_k^2 + log(yerr_k^2));
 *      failed models -> -inf (the notebook's `except RuntimeError: return -np.inf`). */
int pwb_convolve_extend_batch(pwb_ctx* ctx, int B, const double* d_spectrum, int Nw, const double* d_kernel,
                              int Nk, double* d_out, void* stream);
int pwb_loglike_batch(pwb_ctx* ctx, int B, const double* d_model, int Nw, const double* d_y,
                      const double* d_yerr, const int* d_status, double* d_out, void* stream);

/* ---- transit.draw_transit (transit.py:20-116) with flatstar's rasteriser ---------
 * ld_law: 0 none, 1 linear, 2 quadratic, 3 square-root, 4 log, 5 exp, 6 s3, 7 c4
 * outputs: d_i0, d_r_map [grid_size^2]; *transit_depth (host) */
int pwb_draw_transit(pwb_ctx* ctx, double planet_to_star_ratio, double planet_physical_radius,
                     double impact_parameter, double phase, int grid_size, int supersampling,
                     int ld_law, const double* ld_coefficient, double* d_i0, double* d_r_map,
                     double* transit_depth, void* stream);

/* ---- FP64 FMA peak of the device (for the roofline denominator; MEASURED_PEAKS.json
 *      has no FP64 figure): runs a dependent-chain DFMA kernel and returns TFLOP/s. */
int pwb_fp64_peak(pwb_ctx* ctx, double* tflops, void* stream);

/* ---- Self-test of the step arithmetic (test hook, no reference counterpart): quot[i] =
 *      a[i] / b[i] and root[i] = a[i] ** (1 / k[i]) through the routines the device build of the
 *      LSODA replica uses in place of IEEE division and pow (csrc/pw_common.cuh: PW_DIV,
 *      pw_root), so that the GPU tests can compare them with numpy. */
int pwb_selftest_arith(pwb_ctx* ctx, int n, const double* d_a, const double* d_b, const int* d_k,
                       double* d_quot, double* d_root, void* stream);
/* ---- Self-test of the two e^x routines of the pixel sum (test hook): wide[i] = pw_exp_neg_wide(x[i])
 *      (NaN outside its range -11 < x <= 0), general[i] = pw_exp_neg_sum_sh(x[i]) (any x <= 0). */
int pwb_selftest_exp(pwb_ctx* ctx, int n, const double* d_x, double* d_wide, double* d_general, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PWINDS_B200_H */
