"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference is imported from /root/reference with the stand-ins under
oracle/standins (astropy, flatstar are absent from the image).  Every array
stored here is an output of the reference's own functions, never of the oracle or
of the CUDA engine.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refloader  # noqa: E402
from refchain import (spectrum_dict, one_model, R, R_PL, M_PL, M_STAR, A_AU, M_HE,  # noqa: E402,F401
                      TIGHT_H, TIGHT_HE, DATA)
import astropy.units as u  # noqa: E402,F401  (stand-in)
from p_winds import parker, hydrogen, helium, transit, tools, lines  # noqa: E402,F401

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def stack(models):
    keys = models[0].keys()
    return {k: np.array([m[k] for m in models]) for k in keys}


def main():
    os.makedirs(OUT, exist_ok=True)
    sp = spectrum_dict()

    # ---- data fixtures (the SED and the He 2^3S profile the reference tests use)
    np.savez_compressed(os.path.join(OUT, 'solar_sed.npz'),
                        wavelength=sp['wavelength'], flux_lambda=sp['flux_lambda'])
    prof = np.loadtxt(os.path.join(DATA, 'he_3_profile.dat'))
    np.savez_compressed(os.path.join(OUT, 'he_3_profile.npz'), r=prof[:, 0],
                        v=prof[:, 1], n=prof[:, 2])

    # ---- SED scalars
    phi, a0 = hydrogen.radiative_processes(sp)
    he = helium.radiative_processes(sp)
    np.savez(os.path.join(OUT, 'sed_scalars.npz'), h=np.array([phi, a0]),
             he=np.array(he), he_collision_9100=np.array(helium.collision(9100.)),
             he_recomb_9100=np.array(helium.recombination(9100.)),
             h_recomb_9100=hydrogen.recombination(9100.))

    # ---- geometry (quickstart: k=0.12086, b=0.499, G=100, ss=10) + the test geometry
    i0, depth, r_map = transit.draw_transit(0.12086, R_PL * 71492000, 0.499, 0.0,
                                            supersampling=10, grid_size=100)
    t0, td, tr = transit.draw_transit(0.12086, R_PL * 71492000, 0.0, 0.0,
                                      supersampling=10, grid_size=101)
    ld = {}
    for law, c in (('linear', 0.5), ('quadratic', [0.3, 0.2]),
                   ('square-root', [0.3, 0.2]), ('log', [0.3, 0.2]),
                   ('exp', [0.3, 0.01]), ('s3', [0.3, 0.2, 0.1]),
                   ('c4', [0.3, 0.2, 0.1, 0.05])):
        m, d, rm = transit.draw_transit(0.1, 1.0, 0.3, -0.2, supersampling=4,
                                        grid_size=40, limb_darkening_law=law,
                                        ld_coefficient=c)
        ld['ld_%s_map' % law] = m
        ld['ld_%s_depth' % law] = d
        ld['ld_%s_rmap' % law] = rm
    np.savez_compressed(os.path.join(OUT, 'geometry.npz'), i0=i0, depth=depth,
                        r_map=r_map, test_i0=t0, test_depth=td, test_r_map=tr, **ld)

    # ---- RT known-answer fixture (tests/test_transit.py:42-48 inputs)
    w0, w1, w2, f0, f1, f2, a_ij = lines.he_3_properties()
    wl = np.linspace(1.0827, 1.0832, 150) * 1E-6
    args = (prof[:, 0], prof[:, 2], prof[:, 1], np.array([w0, w1, w2]),
            np.array([f0, f1, f2]), np.array([a_ij] * 3))
    s_avg = transit.radiative_transfer_2d(t0, tr, *args, wl, 9E3, M_HE,
                                          bulk_los_velocity=-2E3)
    tau_avg = transit.optical_depth_2d(*args, wl, 9E3, M_HE, 200, -2E3)
    s_turb = transit.radiative_transfer_2d(t0, tr, *args, wl, 9E3, M_HE,
                                           bulk_los_velocity=-2E3,
                                           planet_radial_velocity=1.5E3,
                                           turbulence_broadening=True)
    s_formal = transit.radiative_transfer_2d(t0, tr, *args, wl[::6], 9E3, M_HE,
                                             bulk_los_velocity=-2E3,
                                             wind_broadening_method='formal')
    tprof = np.linspace(9E3, 7E3, len(prof))
    s_tprof = transit.radiative_transfer_2d(t0, tr, *args, wl, tprof, M_HE,
                                            bulk_los_velocity=-2E3)
    s_formal_t = transit.radiative_transfer_2d(t0, tr, *args, wl[::6], tprof, M_HE,
                                               bulk_los_velocity=-2E3,
                                               wind_broadening_method='formal')
    s_single = transit.radiative_transfer_2d(1.0, tr, *args[:3], w1, f1, a_ij, wl,
                                             9E3, M_HE)
    np.savez_compressed(os.path.join(OUT, 'rt_known_answer.npz'), wl=wl,
                        spectrum_average=s_avg, tau_average=tau_avg,
                        spectrum_turbulence_rv=s_turb, spectrum_formal=s_formal,
                        spectrum_tprofile=s_tprof, spectrum_formal_tprofile=s_formal_t,
                        tprofile=tprof, spectrum_single_line_scalar_i0=s_single)

    # ---- Parker structure
    rr = np.concatenate((np.logspace(-1.2, 0, 40), np.logspace(0, 1.2, 40)[1:]))
    v, rho = parker.structure(rr)
    cs = parker.sound_speed(9100., 0.8)
    rs = parker.radius_sonic_point_tidal(M_PL, cs, M_STAR, A_AU)
    vt, rhot = parker.structure_tidal(rr, cs, rs, M_PL, M_STAR, A_AU)
    np.savez(os.path.join(OUT, 'parker.npz'), r=rr, v=v, rho=rho, cs=cs, rs_tidal=rs,
             rs_plain=parker.radius_sonic_point(M_PL, cs),
             rhos=parker.density_sonic_point(1e11, rs, cs), v_tidal=vt, rho_tidal=rhot,
             v_at_1=np.array(parker.structure(1.0)),
             cs_1e4=parker.sound_speed(1e4, 1.0))

    # ---- hydrogen-stage goldens incl. the reference's own test configuration
    r500 = np.linspace(1, 15, 500)
    f_apx = hydrogen.ion_fraction(r500, R_PL, 9E3, 0.9, 8E10, M_PL, 0.7,
                                  spectrum_at_planet=sp, relax_solution=True)
    f_ex = hydrogen.ion_fraction(r500, R_PL, 9E3, 0.9, 8E10, M_PL, 0.7,
                                 spectrum_at_planet=sp, relax_solution=True,
                                 exact_phi=True)
    f_mono = hydrogen.ion_fraction(np.linspace(1.0, 20, 500), R_PL, 9100., 0.9,
                                   10 ** 10.27, M_PL, 0.7613, flux_euv=1000.,
                                   initial_f_ion=0.0, relax_solution=True)
    f_norelax, mu_norelax = hydrogen.ion_fraction(
        R, R_PL, 9100., 0.9, 10 ** 10.27, M_PL, 1.3, spectrum_at_planet=sp,
        exact_phi=True, return_mu=True)
    np.savez_compressed(os.path.join(OUT, 'hydrogen.npz'), r500=r500, f_approx=f_apx,
                        f_exact=f_ex, f_mono=f_mono, f_norelax=f_norelax,
                        mu_norelax=mu_norelax)

    # ---- anchor models through the whole chain
    geom = (i0, r_map)
    sets = {
        # defaults (reference tolerances): includes failing models (T0 <= 8000)
        'anchors_default': [(T, 10 ** lm, 0.9, False, False)
                            for T in (6000., 8000., 9100., 10000., 11500., 13000.)
                            for lm in (9.5, 10.27, 11.0, 11.6, 12.0)],
        'anchors_default_tidal': [(T, 10 ** lm, h, True, False)
                                  for T in (7000., 9000., 10500., 12500.)
                                  for lm in (9.7, 10.6, 11.5) for h in (0.85, 0.95)],
        # tight tolerances: the tier where 1e-6 / 1e-5 is enforceable per model
        'anchors_tight': [(T, 10 ** lm, 0.9, False, True)
                          for T in (9100., 10500., 12500.)
                          for lm in (9.7, 10.27, 11.0, 11.7)],
        'anchors_tight_tidal': [(T, 10 ** lm, h, True, True)
                                for T in (9100., 11500.) for lm in (10.0, 11.3)
                                for h in (0.85, 0.95)],
    }
    for name, plist in sets.items():
        models = []
        for (T, md, h, tidal, tight) in plist:
            m = one_model(sp, geom, T, md, h, tidal, tight)
            m.update(T0=T, mdot=md, h_fraction=h)
            models.append(m)
            print(name, T, '%.3e' % md, h, 'status', m['status'], flush=True)
        d = stack(models)
        d['tidal'] = plist[0][3]
        d['tight'] = plist[0][4]
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)


def oxygen_fixture(sp):
    """tests/golden/oxygen.npz: oxygen.radiative_processes and oxygen.ion_fraction of the
    unmodified reference on profiles from its own H / He chain."""
    from p_winds import oxygen
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        scal = np.array(oxygen.radiative_processes(sp))
    cases, keep = [], {}
    for (T0, md, h) in [(9100., 10 ** 10.27, 0.9), (9100., 1e11, 0.9), (11500., 10 ** 10.5, 0.85),
                        (13000., 10 ** 11.6, 0.9)]:
        he_h = (1 - h) / h
        mu_0 = (1 + 4 * he_h) / (1 + he_h)
        f_r, mu = hydrogen.ion_fraction(R, R_PL, T0, h, md, M_PL, mu_0, spectrum_at_planet=sp,
                                        initial_f_ion=0.0, relax_solution=True, exact_phi=True, return_mu=True)
        vs = parker.sound_speed(T0, mu)
        rs = parker.radius_sonic_point(M_PL, vs)
        rhos = parker.density_sonic_point(md, rs, vs)
        v, rho = parker.structure(R * R_PL / rs)
        f1, f3 = helium.population_fraction(R, v, rho, f_r, R_PL, T0, h, vs, rs, rhos, spectrum_at_planet=sp,
                                            initial_state=np.array([1.0, 0.0]), relax_solution=True)
        f_he_ion = 1 - f1 - f3
        for kw in (dict(), dict(relax_solution=True), dict(relax_solution=True, rtol=1e-8, atol=1e-11),
                   dict(relax_solution=True, method='odeint')):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                try:
                    f_o, rates = oxygen.ion_fraction(R, v, rho, f_r, f_he_ion, R_PL, T0, h, vs, rs, rhos, sp,
                                                     return_rates=True, **kw)
                    status = 0
                except RuntimeError:
                    f_o, rates, status = np.full(len(R), np.nan), None, 2
            # the reference's own round-off noise floor (SURVEY.md 8(c) protocol): re-run it with
            # 1-ulp noise on the density profile; the finite-difference Jacobian of scipy's Radau
            # turns that into a different Newton / step sequence
            noise = 0.0
            if status == 0:
                rng = np.random.default_rng(17)
                for _ in range(6):
                    rho_p = rho * (1 + 2.2e-16 * rng.standard_normal(len(rho)))
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        f_p = oxygen.ion_fraction(R, v, rho_p, f_r, f_he_ion, R_PL, T0, h, vs, rs, rhos, sp, **kw)
                    noise = max(noise, float(np.max(np.abs(f_p[1:] / f_o[1:] - 1))))
            cases.append(dict(T0=T0, mdot=md, h_fraction=h, vs=vs, rs=rs, rhos=rhos, v=v, rho=rho, f_h=f_r, noise=noise,
                              f_he_ion=f_he_ion, relax=int(kw.get('relax_solution', False)),
                              method=kw.get('method', 'Radau'), rtol=kw.get('rtol', 1e-3), atol=kw.get('atol', 1e-6),
                              f_o=f_o, status=status,
                              rates=np.array([rates[k] for k in ('photoionization', 'recombination',
                                                                 'e_impact_ionization', 'charge_exchange_HII',
                                                                 'charge_exchange_HI')])
                              if rates is not None else np.full((5, len(R)), np.nan)))
            print('oxygen', T0, '%.3e' % md, h, kw, 'status', status, 'f_o[-1] %.6f' % f_o[-1],
                  'self-noise %.1e' % noise, flush=True)
    for k in cases[0]:
        keep[k] = np.array([c[k] for c in cases])
    np.savez_compressed(os.path.join(OUT, 'oxygen.npz'), scalars=scal, r=R, **keep)


def hydrogen_rates_fixture(sp):
    """tests/golden/hydrogen_rates.npz: hydrogen.ion_fraction(..., return_rates=True)."""
    cases = []
    for (T0, md, relax, tidal) in [(9100., 10 ** 10.27, True, False), (9100., 1e11, False, False),
                                   (11500., 10 ** 10.8, True, True)]:
        kw = dict(star_mass=M_STAR, semimajor_axis=A_AU) if tidal else {}
        f_r, mu, rates = hydrogen.ion_fraction(R, R_PL, T0, 0.9, md, M_PL, 1.2, spectrum_at_planet=sp,
                                               initial_f_ion=0.0, relax_solution=relax, exact_phi=True,
                                               return_mu=True, return_rates=True, **kw)
        cases.append(dict(T0=T0, mdot=md, relax=int(relax), tidal=int(tidal), f_r=f_r, mu_bar=mu,
                          photoionization=rates['photoionization'], recombination=rates['recombination']))
        print('hydrogen rates', T0, '%.3e' % md, relax, tidal, flush=True)
    np.savez_compressed(os.path.join(OUT, 'hydrogen_rates.npz'), r=R,
                        **{k: np.array([c[k] for c in cases]) for k in cases[0]})


def heion_fixture(sp):
    """tests/golden/helium_ion.npz: helium.radiative_processes(combined_ionization=True) and
    helium.ion_fraction of the unmodified reference (default 'Radau', and 'odeint')."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        scal = np.array(helium.radiative_processes(sp, combined_ionization=True))
    cases, keep = [], {}
    for (T0, md, h) in [(9100., 10 ** 10.27, 0.9), (9100., 1e11, 0.9), (11500., 10 ** 10.5, 0.85),
                        (13000., 10 ** 11.6, 0.9)]:
        he_h = (1 - h) / h
        mu_0 = (1 + 4 * he_h) / (1 + he_h)
        f_r, mu = hydrogen.ion_fraction(R, R_PL, T0, h, md, M_PL, mu_0, spectrum_at_planet=sp,
                                        initial_f_ion=0.0, relax_solution=True, exact_phi=True, return_mu=True)
        vs = parker.sound_speed(T0, mu)
        rs = parker.radius_sonic_point(M_PL, vs)
        rhos = parker.density_sonic_point(md, rs, vs)
        v, rho = parker.structure(R * R_PL / rs)
        for kw in (dict(), dict(relax_solution=True), dict(relax_solution=True, rtol=1e-8, atol=1e-11),
                   dict(relax_solution=True, method='odeint')):
            def run(rho_in):
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    return helium.ion_fraction(R, v, rho_in, f_r, R_PL, T0, h, vs, rs, rhos, sp, **kw)
            try:
                f, status = run(rho), 0
            except RuntimeError:
                f, status = np.full(len(R), np.nan), 2
            noise = 0.0
            if status == 0:
                rng = np.random.default_rng(17)
                for _ in range(6):
                    g = run(rho * (1 + 2.2e-16 * rng.standard_normal(len(rho))))
                    noise = max(noise, float(np.max(np.abs(g[1:] / f[1:] - 1))))
            cases.append(dict(T0=T0, mdot=md, h_fraction=h, vs=vs, rs=rs, rhos=rhos, v=v, rho=rho, f_h=f_r,
                              relax=int(kw.get('relax_solution', False)), method=kw.get('method', 'Radau'),
                              rtol=kw.get('rtol', 1e-3), atol=kw.get('atol', 1e-6), f_he_ion=f, status=status,
                              noise=noise))
            print('he ion', T0, '%.3e' % md, h, kw, 'status', status, 'f[-1] %.6f' % f[-1],
                  'self-noise %.1e' % noise, flush=True)
    for k in cases[0]:
        keep[k] = np.array([c[k] for c in cases])
    np.savez_compressed(os.path.join(OUT, 'helium_ion.npz'), scalars=scal, r=R, **keep)


def carbon_fixture(sp):
    """tests/golden/carbon.npz: carbon.radiative_processes and carbon.ion_fraction of the
    unmodified reference (default 'odeint' -- which gives up on some of these models, as SURVEY.md
    8(f-3) notes -- and 'Radau', the method docs/source/exospheric_metals.ipynb uses)."""
    from p_winds import carbon
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        scal = np.array(carbon.radiative_processes(sp))
    cases, keep = [], {}
    keys = ('ionization_CI', 'ionization_CII', 'recombination_CII', 'recombination_CIII', 'e_impact_ion_CI',
            'e_impact_ion_CII', 'charge_exchange_CI_HII', 'charge_exchange_CI_HeII', 'charge_exchange_CII_HI',
            'charge_exchange_CIII_HI', 'charge_exchange_CIII_HeI')
    for (T0, md, h) in [(9100., 10 ** 10.27, 0.9), (9100., 1e11, 0.9), (11500., 10 ** 10.5, 0.85)]:
        he_h = (1 - h) / h
        mu_0 = (1 + 4 * he_h) / (1 + he_h)
        f_r, mu = hydrogen.ion_fraction(R, R_PL, T0, h, md, M_PL, mu_0, spectrum_at_planet=sp,
                                        initial_f_ion=0.0, relax_solution=True, exact_phi=True, return_mu=True)
        vs = parker.sound_speed(T0, mu)
        rs = parker.radius_sonic_point(M_PL, vs)
        rhos = parker.density_sonic_point(md, rs, vs)
        v, rho = parker.structure(R * R_PL / rs)
        f1, f3 = helium.population_fraction(R, v, rho, f_r, R_PL, T0, h, vs, rs, rhos, spectrum_at_planet=sp,
                                            initial_state=np.array([1.0, 0.0]), relax_solution=True)
        f_he_ion = 1 - f1 - f3
        for kw in (dict(), dict(relax_solution=True), dict(method='Radau'), dict(method='Radau', relax_solution=True),
                   dict(method='Radau', relax_solution=True, rtol=1e-8, atol=1e-11)):
            def run(rho_in):
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    return carbon.ion_fraction(R, v, rho_in, f_r, f_he_ion, R_PL, T0, h, vs, rs, rhos, sp,
                                               return_rates=True, **kw)
            try:
                f2, f3c, rates = run(rho)
                status = 0
            except RuntimeError:
                f2, f3c, rates, status = np.full(len(R), np.nan), np.full(len(R), np.nan), None, 2
            # noise floor of the reference under 1-ulp input noise: C II everywhere; C III where it
            # is within a factor 100 of its peak (below that it sits at the solver's atol)
            noise, noise3 = 0.0, 0.0
            if status == 0:
                rng = np.random.default_rng(17)
                for _ in range(6):
                    try:
                        g2, g3, _r = run(rho * (1 + 2.2e-16 * rng.standard_normal(len(rho))))
                    except RuntimeError:
                        noise = noise3 = np.inf
                        continue
                    big = f3c > 1e-2 * f3c.max()
                    noise = max(noise, float(np.max(np.abs(g2[1:] / f2[1:] - 1))))
                    noise3 = max(noise3, float(np.max(np.abs(g3[big] / f3c[big] - 1))))
            cases.append(dict(T0=T0, mdot=md, h_fraction=h, vs=vs, rs=rs, rhos=rhos, v=v, rho=rho, f_h=f_r,
                              f_he_ion=f_he_ion, relax=int(kw.get('relax_solution', False)),
                              method=kw.get('method', 'odeint'), rtol=kw.get('rtol', 1e-3), atol=kw.get('atol', 1e-6),
                              f_cii=f2, f_ciii=f3c, status=status, noise=noise, noise3=noise3,
                              rates=np.array([rates[k] for k in keys]) if rates is not None
                              else np.full((11, len(R)), np.nan)))
            print('carbon', T0, '%.3e' % md, h, kw, 'status', status, 'f_cii[-1] %.6f' % f2[-1],
                  'self-noise %.1e / %.1e' % (noise, noise3), 'max f_ciii %.2e' % np.nanmax(f3c), flush=True)
    for k in cases[0]:
        keep[k] = np.array([c[k] for c in cases])
    np.savez_compressed(os.path.join(OUT, 'carbon.npz'), scalars=scal, r=R, **keep)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'hrates':
        hydrogen_rates_fixture(spectrum_dict())
    elif len(sys.argv) > 1 and sys.argv[1] == 'heion':
        heion_fixture(spectrum_dict())
    elif len(sys.argv) > 1 and sys.argv[1] == 'carbon':
        carbon_fixture(spectrum_dict())
    elif len(sys.argv) > 1 and sys.argv[1] == 'oxygen':
        oxygen_fixture(spectrum_dict())
    else:
        main()
        oxygen_fixture(spectrum_dict())
        carbon_fixture(spectrum_dict())
        heion_fixture(spectrum_dict())
        hydrogen_rates_fixture(spectrum_dict())
