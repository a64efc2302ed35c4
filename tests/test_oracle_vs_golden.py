"""The oracle (oracle/pwinds_oracle.py) is pinned against (1) fixtures produced by
the unmodified reference (oracle/make_golden.py), (2) the reference's own
known-answer numbers (reference tests/test_*.py) and, when the reference tree is
mounted, (3) the live reference."""
import os

import numpy as np
import pytest

from conftest import golden

R = np.logspace(0, np.log10(20), 100)


def test_sed_scalars_bit_exact(oracle, sed):
    g = golden('sed_scalars')
    assert np.array_equal(np.array(oracle.h_radiative_processes(sed)), g['h'])
    assert np.array_equal(np.array(oracle.he_radiative_processes(sed)), g['he'])
    assert np.array_equal(np.array(oracle.he_collision(9100.)), g['he_collision_9100'])
    assert np.array_equal(np.array(oracle.he_recombination(9100.)), g['he_recomb_9100'])
    assert oracle.h_recombination(9100.) == g['h_recomb_9100']


def test_survey_appendix_b_anchor_values(oracle, sed):
    # SURVEY.md Appendix B (reference + scipy 1.18.1)
    phi, a0 = oracle.h_radiative_processes(sed)
    assert phi == pytest.approx(5.549833563862282e-05, rel=1e-14)
    assert a0 == pytest.approx(1.0951066697423974e-18, rel=1e-14)
    assert oracle.nearest_index(sed.wl, 504) == 504
    assert oracle.nearest_index(sed.wl, 911.65) == 911


def test_reference_parker_known_answers(oracle):
    # reference tests/test_parker.py:12-23
    v, rho = oracle.structure(1.0)
    assert abs(v - 1.0) < 1e-6 and abs(rho - 1.0) < 1e-6
    assert abs(oracle.sound_speed(1e4, 1.0) - 9.08537273) / 9.08537273 < 1e-6
    g = golden('parker')
    v, rho = oracle.structure(g['r'])
    assert np.array_equal(v, g['v']) and np.array_equal(rho, g['rho'])
    vt, rt = oracle.structure_tidal(g['r'], float(g['cs']), float(g['rs_tidal']), 0.73, 1.07, 0.04634)
    assert np.array_equal(vt, g['v_tidal']) and np.array_equal(rt, g['rho_tidal'])


def test_reference_hydrogen_known_answers(oracle, sed):
    # reference tests/test_hydrogen.py:31-42 (mu_0 = 0.7 passed positionally)
    g = golden('hydrogen')
    o = oracle.h_ion_fraction(g['r500'], 1.39, 9E3, 0.9, 8E10, 0.73, 0.7, sed=sed, relax_solution=True)
    assert abs((o['f_r'][-1] - 0.998737) / o['f_r'][-1]) < 1e-4
    assert np.array_equal(o['f_r'], g['f_approx'])
    o = oracle.h_ion_fraction(g['r500'], 1.39, 9E3, 0.9, 8E10, 0.73, 0.7, sed=sed, relax_solution=True,
                              exact_phi=True)
    assert abs((o['f_r'][-1] - 0.998913) / o['f_r'][-1]) < 1e-4
    assert np.array_equal(o['f_r'], g['f_exact'])
    o = oracle.h_ion_fraction(np.linspace(1.0, 20, 500), 1.39, 9100., 0.9, 10 ** 10.27, 0.73, 0.7613,
                              flux_euv=1000., relax_solution=True)
    assert np.array_equal(o['f_r'], g['f_mono'])


def test_reference_transit_known_answers(oracle):
    # reference tests/test_transit.py:28-48
    geo = golden('geometry')
    i0, depth, r_map = oracle.draw_transit(0.12086, 1.39 * 71492000, 0.0, 0.0, grid_size=101, supersampling=10)
    assert abs((i0[30, 30] - 1.249219E-4) / i0[30, 30]) < 5e-3
    assert abs((depth - 0.015012130070238605) / depth) < 1e-13      # reference asks 5e-3
    assert np.array_equal(i0, geo['test_i0']) and np.array_equal(r_map, geo['test_r_map'])
    prof, ka = golden('he_3_profile'), golden('rt_known_answer')
    w0, f, a = oracle.he3_lines()
    s = oracle.radiative_transfer_2d(i0, r_map, prof['r'], prof['n'], prof['v'], w0, f, a, ka['wl'], 9E3,
                                     4 * 1.67262192369e-27, v_bulk=-2E3)
    assert abs((s[60] - 0.97946) / s[60]) < 5e-3
    assert np.array_equal(s, ka['spectrum_average'])
    s = oracle.radiative_transfer_2d(i0, r_map, prof['r'], prof['n'], prof['v'], w0, f, a, ka['wl'][::6], 9E3,
                                     4 * 1.67262192369e-27, v_bulk=-2E3, method='formal')
    assert np.array_equal(s, ka['spectrum_formal'])


def test_draw_transit_limb_darkening_fixtures(oracle):
    geo = golden('geometry')
    for law, c in (('linear', 0.5), ('quadratic', [0.3, 0.2]), ('square-root', [0.3, 0.2]),
                   ('log', [0.3, 0.2]), ('s3', [0.3, 0.2, 0.1]), ('c4', [0.3, 0.2, 0.1, 0.05])):
        m, d, rm = oracle.draw_transit(0.1, 1.0, 0.3, -0.2, grid_size=40, supersampling=4,
                                       limb_darkening_law=law, ld_coefficient=c)
        assert np.array_equal(m, geo['ld_%s_map' % law]), law
        assert d == geo['ld_%s_depth' % law]


@pytest.mark.parametrize('name,idx', [('anchors_default', [5, 11, 12, 17, 29]),
                                      ('anchors_default_tidal', [0, 2, 9, 23]),
                                      ('anchors_tight', [2]), ('anchors_tight_tidal', [1])])
def test_forward_model_anchors(oracle, sed, name, idx):
    """Whole chain vs reference outputs, including the models the reference fails on."""
    d = golden(name)
    geo = golden('geometry')
    tight, tidal = bool(d['tight']), bool(d['tidal'])
    s = oracle.Setup(sed, r=R, tidal=tidal, i0=geo['i0'], r_map=geo['r_map'],
                     h_tol=dict(rtol=1e-10, atol=1e-13) if tight else None,
                     he_method='LSODA' if tight else 'odeint',
                     he_tol=dict(rtol=1e-10, atol=1e-16) if tight else None)
    for i in idx:
        o = oracle.forward_model(s, float(d['T0'][i]), float(d['mdot'][i]), float(d['h_fraction'][i]))
        assert o['status'] == int(d['status'][i])
        if o['status'] != 0:
            continue
        for key in ('f_r', 'v', 'rho', 'f_1', 'f_3', 'spectrum'):
            assert np.array_equal(o[key], d[key][i]), (name, i, key)
        assert o['mu_bar'] == d['mu_bar'][i]


def test_p3_anchor_fixture_and_oracle(oracle, sed):
    """tests/golden/anchors_p3.npz (oracle/make_anchors.py: 272 triples through the unmodified reference,
    each with K = 4 one-ulp-perturbed re-runs): shape of the fixture, and the oracle bit for bit on a
    sample of it incl. failing models."""
    d = golden('anchors_p3')
    geo = golden('geometry')
    n = len(d['T0'])
    assert n >= 256 and (d['tidal']).sum() >= 100 and (~d['tidal']).sum() >= 100
    assert (d['status'] == 1).sum() >= 60 and (d['status'] == 0).sum() >= 150
    for k in ('f_r', 'f_1', 'f_3_peak', 'spectrum'):
        runs = d['noise_%s_runs' % k]
        assert runs.shape == (n, int(d['k_noise'])) and np.nanmax(runs) > 0
    rng = np.random.default_rng(3)
    ok = np.flatnonzero(d['status'] == 0)
    idx = np.concatenate([rng.choice(ok, 5, replace=False), np.flatnonzero(d['status'] == 1)[:3]])
    for i in idx:
        s = oracle.Setup(sed, r=R, tidal=bool(d['tidal'][i]), i0=geo['i0'], r_map=geo['r_map'])
        o = oracle.forward_model(s, float(d['T0'][i]), float(d['mdot'][i]), float(d['h_fraction'][i]))
        assert o['status'] == int(d['status'][i])
        if o['status'] == 0:
            for key in ('f_r', 'f_1', 'f_3', 'spectrum'):
                assert np.array_equal(o[key], d[key][i]), (i, key)


def test_tight_grid_fixture_and_oracle(oracle, sed):
    """tests/golden/anchors_tight_grid.npz: shape, and the oracle bit for bit on a sample."""
    d = golden('anchors_tight_grid')
    geo = golden('geometry')
    assert len(d['T0']) == 100 and (d['status'] == 0).all()
    for i in (3, 40, 70, 99):
        s = oracle.Setup(sed, r=R, tidal=bool(d['tidal'][i]), i0=geo['i0'], r_map=geo['r_map'],
                         h_tol=dict(rtol=1e-12, atol=1e-15), he_method='LSODA', he_tol=dict(rtol=1e-10, atol=1e-16))
        o = oracle.forward_model(s, float(d['T0'][i]), float(d['mdot'][i]), float(d['h_fraction'][i]))
        assert o['status'] == 0
        for key in ('f_r', 'f_1', 'f_3', 'spectrum'):
            assert np.array_equal(o[key], d[key][i]), (i, key)


def test_lyman_alpha_fixture_and_oracle(oracle, sed):
    """tests/golden/anchors_lya.npz (BASELINE configs[1], unmodified reference): the oracle's hydrogen-only
    chain + Ly-alpha radiative transfer bit for bit on a sample."""
    d = golden('anchors_lya')
    geo = golden('geometry')
    assert len(d['T0']) == 25 and (d['status'] == 0).all() and len(d['grid_status']) == 1024
    wl = np.linspace(1214.0, 1217.4, 200) * 1e-10
    s = oracle.Setup(sed, r=R, i0=geo['i0'], r_map=geo['r_map'], wl=wl,
                     lines=(np.array([1215.67e-10]), np.array([0.4164]), np.array([6.265e8])),
                     mass=1.67262192369e-27, h_tol=dict(rtol=1e-10, atol=1e-13), species='h1')
    for i in (0, 7, 24):
        o = oracle.forward_model(s, float(d['T0'][i]), float(d['mdot'][i]), 0.9)
        assert o['status'] == 0
        assert np.array_equal(o['f_r'], d['f_r'][i])
        assert np.max(np.abs(o['spectrum'] - d['spectrum'][i])) <= 1e-15 * np.max(d['spectrum'][i])


def test_helium_failure_class_fixture_and_oracle(oracle, sed):
    """tests/golden/he_failure_classes.npz: >= 20 models of each helium failure class from the unmodified
    reference; class-3 results are not reproducible run to run (recorded), the class is.  The oracle
    returns the same class."""
    d = golden('he_failure_classes')
    geo = golden('geometry')
    st = d['status']
    assert (st == 2).sum() >= 20 and (st == 3).sum() >= 20 and (st == 0).sum() >= 20
    assert (d['f_1_run_spread'][st == 3] > 1e-3).sum() >= 10
    assert (d['f_1_run_spread'][st != 3] == 0).all()
    s = oracle.Setup(sed, r=R, i0=geo['i0'], r_map=geo['r_map'])
    for c in (0, 2, 3):
        for i in np.flatnonzero(st == c)[:4]:
            o = oracle.forward_model(s, float(d['T0'][i]), float(d['mdot'][i]), 0.9)
            assert o['status'] == c, (i, c, o['status'])
            if c == 3:       # the documented convention: the last completed pass, finite and clamped
                assert np.isfinite(o['f_1']).all() and (o['f_1'] <= 1).all() and np.isfinite(o['spectrum']).all()


def test_tidal_grid_status_fixture():
    d = golden('tidal_grid_status')
    assert len(d['status']) == 2048 and (d['status'] == 0).sum() > 2000
    assert (d['status'] == 2).sum() >= 5 and (d['status'] == 3).sum() >= 5


def test_bench_grid_status_fixture():
    d = golden('bench_grid_status')
    assert len(d['status']) == 4096 and (d['status'] == 0).sum() > 4000
    assert (d['status'] == 2).sum() >= 5 and (d['status'] == 3).sum() >= 5


def test_failure_map_in_fixtures():
    """The reference fails on a large part of a wide grid; the fixtures
    must contain both outcomes so that status parity is actually exercised."""
    d = golden('anchors_default')
    assert (d['status'] == 1).sum() >= 5 and (d['status'] == 0).sum() >= 15


@pytest.mark.skipif(not os.path.isdir('/root/reference/p_winds'), reason='reference tree not mounted')
def test_oracle_matches_live_reference(oracle, sed):
    import refloader
    refloader.load()
    import astropy.units as u
    from p_winds import hydrogen, helium, parker, tools
    units = {'wavelength': u.angstrom, 'flux': u.erg / u.s / u.cm ** 2 / u.angstrom}
    sp = tools.make_spectrum_from_file('/root/reference/data/solar_spectrum_scaled_lambda.dat', units)
    T0, md, h = 10400., 10 ** 10.9, 0.88
    mu0 = (1 + 4 * (1 - h) / h) / (1 + (1 - h) / h)
    f_r, mu = hydrogen.ion_fraction(R, 1.39, T0, h, md, 0.73, mu0, spectrum_at_planet=sp, exact_phi=True,
                                    relax_solution=True, return_mu=True)
    o = oracle.h_ion_fraction(R, 1.39, T0, h, md, 0.73, mu0, sed=sed, exact_phi=True, relax_solution=True)
    assert np.array_equal(o['f_r'], f_r) and o['mu_bar'] == mu
    vs = parker.sound_speed(T0, mu)
    rs = parker.radius_sonic_point(0.73, vs)
    rhos = parker.density_sonic_point(md, rs, vs)
    v, rho = parker.structure(R * 1.39 / rs)
    f1, f3 = helium.population_fraction(R, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, sp,
                                        initial_state=np.array([1.0, 0.0]), relax_solution=True)
    oh = oracle.he_population(R, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, sed=sed,
                              initial_state=np.array([1.0, 0.0]), relax_solution=True)
    assert np.array_equal(oh['f_1'], f1) and np.array_equal(oh['f_3'], f3)


def test_convolve_extend_definition(oracle):
    """The oracle's restatement of astropy.convolution.convolve(boundary='extend') equals the
    textbook convolution of the edge-padded array with the unit-sum kernel (numpy.convolve),
    for symmetric and asymmetric odd kernels."""
    rng = np.random.default_rng(11)
    x = 1.0 - 0.02 * np.exp(-0.5 * ((np.arange(169) - 70) / 6.0) ** 2) + 1e-4 * rng.standard_normal(169)
    for k in (np.exp(-0.5 * ((np.arange(169) - 84) / 2.5) ** 2), rng.uniform(0.1, 1.0, 31), np.array([1.0])):
        ref = np.convolve(np.pad(x, k.size // 2, mode='edge'), k / k.sum(), mode='valid')
        got = oracle.convolve_extend(x, k)
        assert got.shape == x.shape
        assert np.max(np.abs(got - ref)) < 2e-15
    with pytest.raises(ValueError):
        oracle.convolve_extend(x, np.ones(4))
    assert abs(oracle.log_likelihood(x, x, np.full_like(x, 0.1)) + 0.5 * 169 * np.log(0.01)) < 1e-10


def test_oxygen_oracle_matches_reference(oracle, sed):
    """SURVEY 8(f-3), oxygen: scalars and ion fractions of the unmodified reference
    (tests/golden/oxygen.npz, made by oracle/make_golden.py) -- bit for bit."""
    d = golden('oxygen')
    assert np.array_equal(np.array(oracle.o_radiative_processes(sed)), d['scalars'])
    for i in range(len(d['T0'])):
        kw = {} if d['method'][i] == 'odeint' else dict(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        o = oracle.o_ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], d['f_he_ion'][i], 1.39,
                                  float(d['T0'][i]), float(d['h_fraction'][i]), float(d['vs'][i]), float(d['rs'][i]),
                                  float(d['rhos'][i]), sed=sed, relax_solution=bool(d['relax'][i]),
                                  method=str(d['method'][i]), **kw)
        assert o['status'] == d['status'][i]
        assert np.array_equal(o['f_o'], d['f_o'][i]), i


def test_carbon_oracle_matches_reference(oracle, sed):
    """SURVEY 8(f-3), carbon: scalars and ion fractions of the unmodified reference
    (tests/golden/carbon.npz) -- bit for bit, including the models its default `odeint` gives up on."""
    d = golden('carbon')
    assert np.array_equal(np.array(oracle.c_radiative_processes(sed)), d['scalars'])
    for T in (7000., 9100., 12000.):
        assert all(np.isfinite(x) for x in oracle.c_charge_transfer(T))
    for i in range(len(d['T0'])):
        kw = {} if d['method'][i] == 'odeint' else dict(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        o = oracle.c_ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], d['f_he_ion'][i], 1.39,
                                  float(d['T0'][i]), float(d['h_fraction'][i]), float(d['vs'][i]), float(d['rs'][i]),
                                  float(d['rhos'][i]), sed=sed, relax_solution=bool(d['relax'][i]),
                                  method=str(d['method'][i]), **kw)
        assert o['status'] == d['status'][i], i
        assert np.array_equal(o['f_cii'], d['f_cii'][i], equal_nan=True), i
        assert np.array_equal(o['f_ciii'], d['f_ciii'][i], equal_nan=True), i


def test_helium_ion_fraction_oracle_matches_reference(oracle, sed):
    """SURVEY 8(f-3), helium.ion_fraction (helium.py:789-1032): bit for bit."""
    d = golden('helium_ion')
    assert np.array_equal(np.array(oracle.he_radiative_processes_combined(sed)), d['scalars'])
    for i in range(len(d['T0'])):
        kw = {} if d['method'][i] == 'odeint' else dict(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        o = oracle.he_ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], 1.39, float(d['T0'][i]),
                                   float(d['h_fraction'][i]), float(d['vs'][i]), float(d['rs'][i]),
                                   float(d['rhos'][i]), sed=sed, relax_solution=bool(d['relax'][i]),
                                   method=str(d['method'][i]), **kw)
        assert o['status'] == d['status'][i] == 0
        assert np.array_equal(o['f_he_ion'], d['f_he_ion'][i]), i


def test_hydrogen_rates_oracle_matches_reference(oracle, sed):
    """hydrogen.ion_fraction(return_rates=True) (hydrogen.py:538-561): bit for bit."""
    d = golden('hydrogen_rates')
    for i in range(len(d['T0'])):
        star = (1.07, 0.04634) if d['tidal'][i] else (1.0, 1.0)
        o = oracle.h_ion_fraction(d['r'], 1.39, float(d['T0'][i]), 0.9, float(d['mdot'][i]), 0.73, 1.2, *star,
                                  sed=sed, exact_phi=True, relax_solution=bool(d['relax'][i]))
        assert np.array_equal(o['f_r'], d['f_r'][i]) and o['mu_bar'] == d['mu_bar'][i]
        assert np.array_equal(o['rates']['photoionization'], d['photoionization'][i])
        assert np.array_equal(o['rates']['recombination'], d['recombination'][i])
