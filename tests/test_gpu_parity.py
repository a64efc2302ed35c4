"""GPU parity tests (run on the B200 box: `python -m pytest tests -m gpu`).

Everything goes through the C ABI (ctypes -> libpwinds_b200.so).  Checkers: the
committed reference fixtures (tests/golden, produced by the unmodified reference)
and the CPU oracle.  Tolerances follow BASELINE.json / SURVEY.md 8(c):

* P2 -- solver tolerances tightened on both sides (H rtol=1e-10/atol=1e-13, He
  LSODA rtol=1e-10/atol=1e-16): fractions <= 1e-6, spectra <= 1e-5 relative, per model.
* P3 -- reference defaults (RK45 1e-3/1e-6, odeint 1.49e-8): the reference's own
  output is only reproducible to its round-off noise floor (1-ulp input noise moves
  9 % of its f_r by > 1e-6, its spectra by 7.5e-7 median / 3.5e-5 max), so the
  comparison is statistical, plus status (ok / failed) agreement.
* deterministic maps (Parker, SED integrals, Voigt, RT): <= 1e-10 in all tiers.
"""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

R = np.logspace(0, np.log10(20), 100)
WL = np.linspace(1.0827, 1.0832, 200) * 1E-6
M_HE = 4 * 1.67262192369e-27


@pytest.fixture(scope='module')
def eng(built):
    from p_winds_b200.engine import Engine
    g = golden('solar_sed')
    geo = golden('geometry')
    e = Engine(0)
    e.set_sed(g['wavelength'], g['flux_lambda'])
    e.set_geometry(geo['i0'], geo['r_map'], R * (1.39 * 71492000))
    return e


def _opts(eng, tight, tidal, h_tol=None, he_tol=None):
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    sc = eng.sed_scalars()
    rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
    hk = (h_tol or dict(rtol=1e-10, atol=1e-13)) if tight else {}
    ho = Engine.h_opts(1.39, 0.73, 1.07 if tidal else 1.0, 0.04634 if tidal else 1.0, relax_solution=True,
                       exact_phi=True, structure_tidal=tidal, **hk)
    hek = dict(method='LSODA', **(he_tol or dict(rtol=1e-10, atol=1e-16))) if tight else {}
    heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True, **hek)
    w = lines.he_3_properties()
    rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, M_HE)
    return ho, heo, rto


def _forward(eng, T, md, hf, tight=False, tidal=False, h_tol=None, he_tol=None):
    ho, heo, rto = _opts(eng, tight, tidal, h_tol, he_tol)
    out = eng.forward_batch(T, md, hf, R, ho, heo, rto, WL)
    return {k: v.cpu().numpy() for k, v in out.items()}


def _rel(a, b):
    return np.abs(a / b - 1)


# --------------------------------------------------------------------------- SED / Parker
def test_sed_scalars(eng):
    sc = eng.sed_scalars()
    g = golden('sed_scalars')
    got = np.array([sc[k] for k in ('h_phi', 'h_a0', 'he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1',
                                    'he_a_h_3')])
    assert np.max(_rel(got, np.concatenate((g['h'], g['he'])))) < 1e-13
    assert (sc['n_cut_h'], sc['i1'], sc['i0']) == (912, 504, 911)


def test_parker_structure_reference_signature(eng):
    from p_winds_b200 import parker
    g = golden('parker')
    v, rho = parker.structure(g['r'])
    m = np.abs(g['r'] - 1.0) > 1e-6
    assert np.max(_rel(v[m], g['v'][m])) < 1e-12 and np.max(_rel(rho[m], g['rho'][m])) < 1e-12
    v1, rho1 = parker.structure(1.0)                       # reference tests/test_parker.py:12-15
    assert abs(v1 - 1.0) < 1e-6 and abs(rho1 - 1.0) < 1e-6
    vt, rt = parker.structure_tidal(g['r'], float(g['cs']), float(g['rs_tidal']), 0.73, 1.07, 0.04634)
    assert np.max(_rel(vt[m], g['v_tidal'][m])) < 1e-12 and np.max(_rel(rt[m], g['rho_tidal'][m])) < 1e-12
    assert abs((parker.sound_speed(1e4, 1.0) - 9.08537273) / 9.08537273) < 1e-6


# --------------------------------------------------------------------------- P2
@pytest.mark.parametrize('name', ['anchors_tight', 'anchors_tight_tidal'])
def test_forward_tight_tolerances_per_model(eng, name):
    d = golden(name)
    o = _forward(eng, d['T0'], d['mdot'], d['h_fraction'], tight=True, tidal=bool(d['tidal']))
    assert np.array_equal(o['status'], d['status']) and (o['status'] == 0).all()
    assert np.max(_rel(o['f_r'][:, 1:], d['f_r'][:, 1:])) < 1e-6
    assert np.max(_rel(o['hscal'][:, 0], d['mu_bar'])) < 1e-6
    assert np.max(_rel(o['v'], d['v'])) < 1e-6 and np.max(_rel(o['rho'], d['rho'])) < 1e-6
    assert np.max(_rel(o['f_1'], d['f_1'])) < 1e-6
    for b in range(len(d['T0'])):
        pk = d['f_3'][b].max()
        m = d['f_3'][b] > 1e-3 * pk
        assert np.max(_rel(o['f_3'][b][m], d['f_3'][b][m])) < 1e-6, b
    assert np.max(_rel(o['spectrum'], d['spectrum'])) < 1e-5      # He 10830 transit depth contract
    # in practice the spectra agree ~1e-9; keep a tripwire well inside the contract
    assert np.max(_rel(o['spectrum'], d['spectrum'])) < 1e-7


def test_forward_tight_grid_100_models(eng):
    """BASELINE.json's tolerances, model by model, over the parameter box of the bench grids: 64
    non-tidal + 36 tidal models through the unmodified reference at tight solver tolerances (H rtol 1e-12 /
    atol 1e-15, He LSODA rtol 1e-10 / atol 1e-16: tests/golden/anchors_tight_grid.npz, oracle/make_anchors.py
    tight).  Ion and population fractions within 1e-6, He 10830 transit spectrum within 1e-5."""
    d = golden('anchors_tight_grid')
    worst = {'f_r': 0.0, 'f_1': 0.0, 'f_3': 0.0, 'spectrum': 0.0, 'mu': 0.0}
    for tidal in (False, True):
        m = d['tidal'] == tidal
        o = _forward(eng, d['T0'][m], d['mdot'][m], d['h_fraction'][m], tight=True, tidal=tidal,
                     h_tol=dict(rtol=1e-12, atol=1e-15))
        assert (o['status'] == 0).all() and (d['status'][m] == 0).all()
        worst['f_r'] = max(worst['f_r'], np.max(_rel(o['f_r'][:, 1:], d['f_r'][m][:, 1:])))
        worst['mu'] = max(worst['mu'], np.max(_rel(o['hscal'][:, 0], d['mu_bar'][m])))
        worst['f_1'] = max(worst['f_1'], np.max(_rel(o['f_1'], d['f_1'][m])))
        f3 = d['f_3'][m]
        pk = f3.max(axis=1, keepdims=True)
        big, fringe = f3 > 1e-2 * pk, (f3 > 1e-3 * pk) & (f3 <= 1e-2 * pk)
        worst['f_3'] = max(worst['f_3'], np.max(_rel(o['f_3'][big], f3[big])))
        worst['f_3_fringe'] = max(worst.get('f_3_fringe', 0.0), np.max(_rel(o['f_3'][fringe], f3[fringe])))
        worst['spectrum'] = max(worst['spectrum'], np.max(_rel(o['spectrum'], d['spectrum'][m])))
    print('tight grid, worst relative deviations over 100 models:', {k: float('%.2e' % v) for k, v in worst.items()})
    assert worst['f_r'] < 1e-6 and worst['mu'] < 1e-6 and worst['f_1'] < 1e-6 and worst['f_3'] < 1e-6
    # below 1 % of its peak f_3 is a small difference of large rates: LSODA at rtol 1e-10 determines it to a few
    # 1e-7 only (the reference's own tight LSODA and tight Radau differ by up to 4e-7 there)
    assert worst['f_3_fringe'] < 5e-6
    assert worst['spectrum'] < 1e-5


# --------------------------------------------------------------------------- P3
def _p3_forward(eng, d):
    """The 272 anchor triples of tests/golden/anchors_p3.npz through forward_batch (reference
    default tolerances), non-tidal and tidal halves, reassembled in fixture order."""
    keys = ('status', 'f_r', 'f_1', 'f_3', 'spectrum', 'hscal')
    o = {}
    for tidal in (False, True):
        m = d['tidal'] == tidal
        part = _forward(eng, d['T0'][m], d['mdot'][m], d['h_fraction'][m], tight=False, tidal=tidal)
        for k in keys:
            if k not in o:
                o[k] = np.zeros((len(d['T0']),) + part[k].shape[1:], dtype=part[k].dtype)
            o[k][m] = part[k]
    return o


def test_p3_anchor_protocol(eng):
    """SURVEY.md 8(c) protocol P3 as written, thresholds taken from the fixture.  At the reference's
    default tolerances its own output moves under 1-ulp input noise (tests/golden/anchors_p3.npz:
    K = 4 perturbed runs of the unmodified reference per model).  Pass =
      * status agrees wherever the reference's own status is stable under that noise;
      * per quantity: median engine-vs-reference deviation <= 2 x the median self-noise (floored at the
        1e-10 of the deterministic maps); the fraction of models above the stated tolerance (1e-6
        fractions, 1e-5 flux) <= the reference's own fraction + 5 points;
      * per model: deviation <= max(stated, 3 x that model's own noise).  K = 4 runs do not sample a
        chaotic model's noise fully, so the reference is held to the same rule (its 4th run against
        the noise of the other three) and the engine may miss it on <= that share of models + 5 points."""
    d = golden('anchors_p3')
    o = _p3_forward(eng, d)
    ref_st, st = d['status'], o['status']
    # status 9: the reference did not finish within 90 s (RK45 crawling); the engine gives up (5) or fails (1)
    crawl = ref_st == 9
    assert np.isin(st[crawl], (1, 5)).all()
    stable = (d['status_flips'] == 0) & ~crawl
    mism = (st != ref_st) & stable
    assert mism.sum() <= 2, np.flatnonzero(mism)
    assert (st[ref_st == 1] == 1).mean() > 0.95 and (ref_st == 1).sum() >= 60
    both = (ref_st == 0) & (st == 0)
    assert both.sum() >= 150
    ipk = np.argmax(d['f_3'], axis=1)
    rows = np.arange(len(ipk))
    dev = {'f_r': _rel(o['f_r'][:, 1:], d['f_r'][:, 1:]).max(axis=1),
           'f_1': _rel(o['f_1'], d['f_1']).max(axis=1),
           'f_3_peak': _rel(o['f_3'][rows, ipk], d['f_3'][rows, ipk]),
           'spectrum': _rel(o['spectrum'], d['spectrum']).max(axis=1)}
    stated = {'f_r': 1e-6, 'f_1': 1e-6, 'f_3_peak': 1e-6, 'spectrum': 1e-5}
    report = {}
    for k, e in dev.items():
        e = e[both]
        runs = d['noise_%s_runs' % k][both]                 # [n, K]
        noise = np.nanmax(runs, axis=1)
        med_ref = np.nanmedian(runs)
        exc_ref = np.nanmean(runs > stated[k])
        assert np.median(e) <= max(2 * med_ref, 1e-10), (k, np.median(e), med_ref)
        assert (e > stated[k]).mean() <= exc_ref + 0.05, (k, (e > stated[k]).mean(), exc_ref)
        miss = (e > np.maximum(stated[k], 3 * noise)).mean()
        # the reference against the same rule: run K against the noise of runs 1..K-1
        own = np.nanmax(runs[:, :-1], axis=1)
        miss_ref = np.nanmean(runs[:, -1] > np.maximum(stated[k], 3 * own))
        assert miss <= miss_ref + 0.05, (k, miss, miss_ref)
        report[k] = (float(np.median(e)), float(med_ref), float(miss), float(miss_ref))
    print('P3 (median engine, median self-noise, per-model miss engine, reference):', report)
    assert np.isfinite(o['spectrum'][both]).all()
    assert np.isnan(o['spectrum'][(st == 1) | (st == 5)]).all()          # failed models carry no spectrum


def test_helium_failure_classes_match_reference(eng):
    """Status classes 2 / 3 of helium.population_fraction (helium.py:691-697, :736) on models of the bench
    grid and of the strip around them, against the unmodified reference (tests/golden/
    he_failure_classes.npz: 34 class-2, 32 class-3 models and 40 neighbours that succeed; each run
    three times, plus four times under 1-ulp SED noise).  The class is compared wherever the
    reference's own class is stable under that noise.  For class 3 the reference's *profiles* differ
    from run to run (key f_1_run_spread: scipy leaves the rows after the failing point
    uninitialised) -- nothing to compare; the engine returns the last completed pass, finite and
    clamped, and chi^2 scores such a model +inf unless the caller opts in."""
    d = golden('he_failure_classes')
    o = _forward(eng, d['T0'], d['mdot'], d['h_fraction'])
    st, ref = o['status'], d['status']
    stable = (d['status_perturbed'] == ref[:, None]).all(axis=1)
    assert (ref == 2).sum() >= 20 and (ref == 3).sum() >= 20
    assert (d['f_1_run_spread'][ref == 3] > 0).sum() >= 10       # the reference is not reproducible there
    assert (d['f_1_run_spread'][ref != 3] == 0).all()
    flips = (st != ref) & stable
    print('helium failure classes: %d of %d stable models differ; unstable models: %d, of which %d differ'
          % (flips.sum(), stable.sum(), (~stable).sum(), ((st != ref) & ~stable).sum()))
    assert flips.sum() <= 2, (np.flatnonzero(flips), st[flips], ref[flips])
    w = st == 3
    assert np.isfinite(o['f_1'][w]).all() and ((o['f_1'][w] >= 0) & (o['f_1'][w] <= 1)).all()
    assert np.isfinite(o['spectrum'][w]).all() and np.isnan(o['spectrum'][st == 2]).all()
    n_relax = o['hescal'][:, 0]
    assert (n_relax[w] >= 1).all() and (n_relax[w] <= 10).all()
    # chi^2 policy (pwb_set_relax_warning_policy): status 3 is a failed model by default
    data, sigma = np.ones(200), np.full(200, 1e-3)
    c0 = eng.chi2_batch(o['spectrum'], data, sigma, 1.0, o['status']).cpu().numpy()
    assert np.isinf(c0[w]).all() and np.isinf(c0[st == 2]).all() and np.isfinite(c0[st == 0]).all()
    eng.set_relax_warning_policy(True)
    try:
        c1 = eng.chi2_batch(o['spectrum'], data, sigma, 1.0, o['status']).cpu().numpy()
    finally:
        eng.set_relax_warning_policy(False)
    assert np.isfinite(c1[w]).all() and np.isinf(c1[st == 2]).all()


def test_bench_grid_status_map_matches_reference(eng):
    """Every model of the 64 x 64 bench grid: the engine's status against the status the unmodified
    reference ends with (tests/golden/bench_grid_status.npz, oracle/make_anchors.py scan: 4076 ok,
    11 class 2, 9 class 3).  Models on a class boundary can flip under round-off (the reference's do
    under 1-ulp noise, see he_failure_classes.npz), so a handful of differences is allowed -- and they
    have to lie next to a model the reference itself fails."""
    d = golden('bench_grid_status')
    o = _forward(eng, d['T0'], d['mdot'], np.full(len(d['T0']), 0.9))
    st, ref = o['status'].reshape(64, 64), d['status'].reshape(64, 64).astype(int)
    diff = st != ref
    print('bench grid: engine', {int(k): int((st == k).sum()) for k in np.unique(st)},
          'reference', {int(k): int((ref == k).sum()) for k in np.unique(ref)}, 'differ', int(diff.sum()))
    assert diff.sum() <= 12
    assert abs(int((st != 0).sum()) - int((ref != 0).sum())) <= 8
    strip = (d['mdot'].reshape(64, 64) > 10 ** 11.2)      # all helium failures of the reference lie here
    assert not diff[~strip].any()


def test_north_star_grid_sample_status_and_depth_match_reference(eng):
    """A 32 x 32 x 2 sample of the north-star grid (BASELINE configs[3]: tidal Parker profile, T0 9000-13000 K,
    h = 0.85 / 0.95) through the unmodified reference (tests/golden/tidal_grid_status.npz: 2025 ok, 12 class 2,
    11 class 3): the engine ends every model in the same class, and the depth of the He 10830 line agrees
    inside the reference's own default-tolerance noise (P3: median ~1e-6)."""
    d = golden('tidal_grid_status')
    o = _forward(eng, d['T0'], d['mdot'], d['h_fraction'], tidal=True)
    st, ref = o['status'], d['status'].astype(int)
    diff = st != ref
    print('north-star sample: engine', {int(k): int((st == k).sum()) for k in np.unique(st)},
          'reference', {int(k): int((ref == k).sum()) for k in np.unique(ref)}, 'differ', int(diff.sum()))
    assert diff.sum() <= 4
    ok = (st == 0) & (ref == 0)
    dev = np.abs(o['spectrum'][ok].min(axis=1) / d['spectrum_min'][ok] - 1)
    assert np.median(dev) < 3e-6 and np.percentile(dev, 99) < 2e-4


def test_hydrogen_work_limit(eng):
    """A model on which the reference's RK45 crawls (1.9e8 right-hand sides, hours in Python, before it
    raises): the engine gives it up after max_nfev evaluations with status 5 and says so."""
    from p_winds_b200.engine import Engine
    from p_winds_b200 import hydrogen
    ho = Engine.h_opts(1.39, 0.73, 1.07, 0.04634, relax_solution=True, exact_phi=True, structure_tidal=True,
                       max_nfev=200000)
    he_h = 0.15 / 0.85
    o = eng.h_ion_fraction_batch(np.array([5000.0, 9000.0]), np.array([7.196857e+09, 1e10]), np.array([0.85, 0.85]),
                                 np.full(2, (1 + 4 * he_h) / (1 + he_h)), R, ho)
    st = o['status'].cpu().numpy()
    assert st[0] == 5 and st[1] == 0
    assert o['scal'][0, 5].item() >= 200000 and np.isnan(o['f_r'][0].cpu().numpy()).all()
    with pytest.raises(RuntimeError, match='max_nfev'):
        hydrogen.ion_fraction(R, 1.39, 5000., 0.85, 7.196857e+09, 0.73, (1 + 4 * he_h) / (1 + he_h), 1.07, 0.04634,
                              spectrum_at_planet=_spectrum_dict(), exact_phi=True, relax_solution=True)


def test_status_map_on_a_wide_grid_matches_oracle(eng, oracle, sed):
    """8x8 grid over T0 5000-13000 K: the reference fails below ~7700 K (RK45 step
    underflow on the first step); the engine must fail the same models."""
    T0 = np.linspace(5000, 13000, 8)
    md = np.logspace(9.5, 12, 8)
    TT, MM = np.meshgrid(T0, md, indexing='ij')
    T, M = TT.ravel(), MM.ravel()
    o = _forward(eng, T, M, np.full(T.size, 0.9))
    mu0 = (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9)
    ref = np.array([oracle.h_ion_fraction(R, 1.39, t, 0.9, m, 0.73, mu0, sed=sed, exact_phi=True,
                                          relax_solution=True)['status'] for t, m in zip(T, M)])
    gpu = (o['status'] == 1).astype(int)
    assert ref.sum() >= 16 and (ref == 0).sum() >= 16
    assert (ref != gpu).sum() <= 2


# --------------------------------------------------------------------------- stage APIs
def _spectrum_dict():
    g = golden('solar_sed')
    return {'wavelength': g['wavelength'], 'flux_lambda': g['flux_lambda'], 'wavelength_unit': 'angstrom',
            'flux_unit': 'erg / (s cm2 angstrom)'}


def test_hydrogen_reference_signature(eng):
    """reference tests/test_hydrogen.py:31-42 through p_winds_b200.hydrogen.ion_fraction."""
    from p_winds_b200 import hydrogen
    g = golden('hydrogen')
    sp = _spectrum_dict()
    f = hydrogen.ion_fraction(g['r500'], 1.39, 9E3, 0.9, 8E10, 0.73, 0.7, spectrum_at_planet=sp,
                              relax_solution=True)
    assert abs((f[-1] - 0.998737) / f[-1]) < 1e-4
    assert np.max(_rel(f[1:], g['f_approx'][1:])) < 1e-5
    f, mu = hydrogen.ion_fraction(g['r500'], 1.39, 9E3, 0.9, 8E10, 0.73, 0.7, spectrum_at_planet=sp,
                                  relax_solution=True, exact_phi=True, return_mu=True)
    assert abs((f[-1] - 0.998913) / f[-1]) < 1e-4
    assert np.max(_rel(f[1:], g['f_exact'][1:])) < 1e-5 and 0.5 < mu < 1.3
    f = hydrogen.ion_fraction(np.linspace(1.0, 20, 500), 1.39, 9100., 0.9, 10 ** 10.27, 0.73, 0.7613,
                              flux_euv=1000., initial_f_ion=0.0, relax_solution=True)
    assert np.max(_rel(f[1:], g['f_mono'][1:])) < 1e-5
    f, mu = hydrogen.ion_fraction(R, 1.39, 9100., 0.9, 10 ** 10.27, 0.73, 1.3, spectrum_at_planet=sp,
                                  exact_phi=True, return_mu=True)
    assert np.max(_rel(f[1:], g['f_norelax'][1:])) < 1e-5 and abs(mu / g['mu_norelax'] - 1) < 1e-6
    phi, a0 = hydrogen.radiative_processes(sp)
    assert abs(phi / 5.549833563862282e-05 - 1) < 1e-13
    with pytest.raises(ValueError, match='Either `spectrum_at_planet` or `flux_euv`'):
        hydrogen.ion_fraction(R, 1.39, 9100., 0.9, 1e10, 0.73)
    with pytest.raises(RuntimeError, match='solve_ivp'):      # the reference fails here too
        hydrogen.ion_fraction(R, 1.39, 6000., 0.9, 1e11, 0.73, 1.3, spectrum_at_planet=sp, exact_phi=True,
                              relax_solution=True)
    with pytest.raises(ValueError, match='no CPU fallback'):
        hydrogen.ion_fraction(R, 1.39, 9100., 0.9, 1e10, 0.73, flux_euv=1e3, method='Radau')


def test_hydrogen_return_rates(eng):
    """hydrogen.ion_fraction(..., return_rates=True) (hydrogen.py:538-573): both rate profiles and
    every return arity, against the unmodified reference (tests/golden/hydrogen_rates.npz)."""
    from p_winds_b200 import hydrogen
    d = golden('hydrogen_rates')
    g = golden('solar_sed')
    sp = {'wavelength': g['wavelength'], 'flux_lambda': g['flux_lambda'], 'wavelength_unit': 'angstrom',
          'flux_unit': 'erg / (s cm2 angstrom)'}
    for i in range(len(d['T0'])):
        kw = dict(star_mass=1.07, semimajor_axis=0.04634) if d['tidal'][i] else {}
        f_r, mu, rates = hydrogen.ion_fraction(d['r'], 1.39, float(d['T0'][i]), 0.9, float(d['mdot'][i]), 0.73, 1.2,
                                               spectrum_at_planet=sp, relax_solution=bool(d['relax'][i]),
                                               exact_phi=True, return_mu=True, return_rates=True, **kw)
        # reference tolerances (RK45 1e-3): same step sequence as the reference -> ~1e-9
        assert np.max(_rel(f_r[1:], d['f_r'][i][1:])) < 1e-6 and abs(mu / d['mu_bar'][i] - 1) < 1e-8
        assert np.max(_rel(rates['recombination'][1:], d['recombination'][i][1:])) < 1e-6
        m = d['f_r'][i] < 0.999
        assert np.max(_rel(rates['photoionization'][m], d['photoionization'][i][m])) < 1e-5
    f2, r2 = hydrogen.ion_fraction(d['r'], 1.39, 9100., 0.9, 10 ** 10.27, 0.73, 1.2, spectrum_at_planet=sp,
                                   relax_solution=True, exact_phi=True, return_rates=True)
    assert np.array_equal(f2, f_r) is False and set(r2) == {'photoionization', 'recombination'}
    with pytest.raises(ValueError, match='exact_phi'):
        hydrogen.ion_fraction(d['r'], 1.39, 9100., 0.9, 10 ** 10.27, 0.73, 1.2, spectrum_at_planet=sp,
                              return_rates=True)


def test_helium_reference_signature(eng, oracle, sed):
    """reference tests/test_helium.py:28-79 (range checks) + oracle comparison at tight
    tolerance + return_rates + the monochromatic channel."""
    from p_winds_b200 import helium, parker
    sp = _spectrum_dict()
    r = np.logspace(0, np.log10(20), 100)
    T0, md, h = 9000., 5e10, 0.9
    d = golden('anchors_tight')
    f_r = d['f_r'][1]
    T0, md = float(d['T0'][1]), float(d['mdot'][1])
    vs, rs, rhos = float(d['vs'][1]), float(d['rs'][1]), float(d['rhos'][1])
    v, rho = d['v'][1], d['rho'][1]
    f1, f3 = helium.population_fraction(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, sp,
                                        initial_state=np.array([1.0, 0.0]), relax_solution=True)
    assert ((f1 >= 0) & (f1 <= 1) & (f3 >= 0) & (f3 <= 1)).all()
    f1t, f3t, rates = helium.population_fraction(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, sp,
                                                 initial_state=np.array([1.0, 0.0]), relax_solution=True,
                                                 method='LSODA', rtol=1e-10, atol=1e-16, return_rates=True)
    assert np.max(_rel(f1t, d['f_1'][1])) < 1e-6
    m = d['f_3'][1] > 1e-3 * d['f_3'][1].max()
    assert np.max(_rel(f3t[m], d['f_3'][1][m])) < 1e-6
    assert set(rates) == {'ionization_1', 'ionization_3', 'recombination_1', 'recombination_3',
                          'radiative_transition', 'transition_1_to_3', 'transition_3_to_21s',
                          'transition_3_to_21p', 'other_ionization', 'charge_exchange_1',
                          'charge_exchange_he_ion'}
    assert all(np.isfinite(x).all() and x.shape == (100,) for x in rates.values())
    # RK45 through solve_ivp (helium.py:702) and the monochromatic fluxes (helium.py:181-266)
    oh = oracle.he_population(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos,
                              rates=helium.radiative_processes_mono(500., 5e6),
                              initial_state=np.array([1.0, 0.0]), relax_solution=True, method='LSODA',
                              rtol=1e-10, atol=1e-16)
    f1m, f3m = helium.population_fraction(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, flux_euv=500.,
                                          flux_fuv=5e6, initial_state=np.array([1.0, 0.0]),
                                          relax_solution=True, method='LSODA', rtol=1e-10, atol=1e-16)
    assert np.max(_rel(f1m, oh['f_1'])) < 1e-6
    assert helium.radiative_processes_mono(500., 5e6) == pytest.approx(
        oracle_mono(oracle, 500., 5e6), rel=1e-14)
    with pytest.raises(ValueError, match='no CPU fallback'):
        helium.population_fraction(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos, sp, method='Radau')
    # status 3 through the reference signature: a RuntimeWarning (the reference's user sees scipy's ODEintWarning)
    # and the populations of the last completed pass
    c = golden('he_failure_classes')
    i3 = int(np.flatnonzero(c['status'] == 3)[0])
    o3 = _forward(eng, c['T0'][i3:i3 + 1], c['mdot'][i3:i3 + 1], c['h_fraction'][i3:i3 + 1])
    assert o3['status'][0] == 3
    with pytest.warns(RuntimeWarning, match='relax loop'):
        g1, g3 = helium.population_fraction(r, o3['v'][0], o3['rho'][0], o3['f_r'][0], 1.39, float(c['T0'][i3]), 0.9,
                                            float(o3['hscal'][0, 1]), float(o3['hscal'][0, 2]), float(o3['hscal'][0, 3]),
                                            sp, initial_state=np.array([1.0, 0.0]), relax_solution=True)
    assert np.array_equal(g1, o3['f_1'][0]) and np.array_equal(g3, o3['f_3'][0])
    with pytest.raises(ValueError, match='Either `spectrum_at_planet` must be provided'):
        helium.population_fraction(r, v, rho, f_r, 1.39, T0, h, vs, rs, rhos)
    assert helium.collision(9100.) == pytest.approx(tuple(oracle.he_collision(9100.)), rel=1e-15)


def oracle_mono(oracle, flux_euv, flux_fuv):
    """helium.radiative_processes_mono restated with the oracle's microphysics."""
    a_1 = float(oracle.sigma_he_singlet(np.array([242.0]))[0])
    wl3, a3 = oracle.sigma_he_triplet_table()
    a_3 = np.interp(2348.0, wl3, a3)
    a_h_1 = 6.3E-18 * (242.0 / 13.6) ** (-3)
    e1, e3 = 12398.419843320025 / 242.0, 12398.419843320025 / 2348.0
    return (flux_euv * 6.24150907e+11 * a_1 / e1, flux_fuv * 6.24150907e+11 * a_3 / e3, a_1, a_3, a_h_1, 0.0)


def test_radiative_transfer_reference_signature(eng):
    """reference tests/test_transit.py:42-48 inputs through transit.radiative_transfer_2d;
    deterministic map -> 1e-10."""
    from p_winds_b200 import transit, lines
    geo, prof, ka = golden('geometry'), golden('he_3_profile'), golden('rt_known_answer')
    w = lines.he_3_properties()
    w0, f, a = np.array(w[:3]), np.array(w[3:6]), np.array([w[6]] * 3)
    args = (prof['r'], prof['n'], prof['v'], w0, f, a)
    s = transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'], 9E3, M_HE,
                                      bulk_los_velocity=-2E3, wind_broadening_method='average')
    assert abs((s[60] - 0.97946) / s[60]) < 5e-3                  # the reference's own golden number
    assert np.max(_rel(s, ka['spectrum_average'])) < 1e-10
    s = transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'], 9E3, M_HE,
                                      bulk_los_velocity=-2E3, planet_radial_velocity=1.5E3,
                                      turbulence_broadening=True)
    assert np.max(_rel(s, ka['spectrum_turbulence_rv'])) < 1e-10
    s = transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'], ka['tprofile'],
                                      M_HE, bulk_los_velocity=-2E3)
    assert np.max(_rel(s, ka['spectrum_tprofile'])) < 1e-10
    # scalar intensity, single line given as floats (transit.py:418-421, :225)
    s = transit.radiative_transfer_2d(1.0, geo['test_r_map'], prof['r'], prof['n'], prof['v'], w[1], w[4],
                                      w[6], ka['wl'], 9E3, M_HE)
    assert np.max(_rel(s, ka['spectrum_single_line_scalar_i0'])) < 1e-10
    # 'formal' wind broadening (per-cell Doppler width and wind shift, transit.py:463-473)
    s = transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'][::6], 9E3, M_HE,
                                      bulk_los_velocity=-2E3, wind_broadening_method='formal')
    assert np.max(_rel(s, ka['spectrum_formal'])) < 1e-10
    s = transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'][::6], ka['tprofile'],
                                      M_HE, bulk_los_velocity=-2E3, wind_broadening_method='formal')
    assert np.max(_rel(s, ka['spectrum_formal_tprofile'])) < 1e-10
    tau = transit.optical_depth_2d(*args, ka['wl'], 9E3, M_HE, 200, -2E3)
    ref = ka['tau_average']
    m = ref > 0
    assert tau.shape == ref.shape and np.array_equal(tau[~m], ref[~m])
    assert np.max(_rel(tau[m], ref[m])) < 1e-10
    with pytest.raises(ValueError, match='`gas_temperature` has to be either'):
        transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'], '9000', M_HE)
    with pytest.raises(ValueError, match='wind_broadening_method'):
        transit.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], *args, ka['wl'], 9E3, M_HE,
                                      wind_broadening_method='nope')


# --------------------------------------------------------------------------- properties
@pytest.fixture(scope='module')
def grid64(eng):
    T0 = np.linspace(10000, 13000, 64)
    md = np.logspace(9.5, 12, 64)
    TT, MM = np.meshgrid(T0, md, indexing='ij')
    T, M, H = TT.ravel(), MM.ravel(), np.full(4096, 0.9)
    return T, M, H, _forward(eng, T, M, H)


def test_full_grid_is_deterministic_and_batch_independent(eng, grid64):
    """BASELINE config 3 size (64x64).  Same inputs -> bit-identical outputs; a model's
    result does not depend on what else is in the batch or on its position."""
    T, M, H, o = grid64
    o2 = _forward(eng, T, M, H)
    for k in ('spectrum', 'f_r', 'f_1', 'f_3', 'status'):
        assert np.array_equal(o[k], o2[k], equal_nan=True), k
    rng = np.random.default_rng(0)
    pick = rng.choice(4096, 37, replace=False)
    o3 = _forward(eng, T[pick], M[pick], H[pick])
    for k in ('spectrum', 'f_r', 'f_1', 'f_3', 'status'):
        assert np.array_equal(o[k][pick], o3[k], equal_nan=True), k
    ok = o['status'] == 0
    assert ok.mean() > 0.98                                 # throughput grid: >= 99 % succeed


def test_full_grid_physical_invariants(eng, grid64):
    T, M, H, o = grid64
    ok = o['status'] == 0
    geo = golden('geometry')
    cont = geo['i0'].sum()
    sp = o['spectrum'][ok]
    assert (sp > 0).all() and (sp <= cont * (1 + 1e-12)).all()
    assert np.max(np.abs(sp[:, 0] / cont - 1)) < 2e-2       # far blue wing ~ continuum
    f_r, f1, f3 = o['f_r'][ok], o['f_1'][ok], o['f_3'][ok]
    assert (f_r >= 0).all() and (f_r <= 1 + 1e-9).all() and (f_r[:, 0] == 0).all()
    assert (f1 >= 0).all() and (f1 <= 1).all() and (f3 >= 0).all() and (f3 <= 1).all()
    assert (f1[:, 0] == 1).all() and (f3[:, 0] == 0).all()   # initial_state = [1, 0]
    assert (np.diff(o['v'][ok], axis=1) > 0).all()           # transonic Parker wind accelerates


def test_rt_linear_in_intensity_and_chi2(eng, grid64):
    T, M, H, o = grid64
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    geo = golden('geometry')
    w = lines.he_3_properties()
    rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, M_HE)
    ok = np.where(o['status'] == 0)[0][:64]
    he = 0.1
    n = o['f_3'][ok] * o['rho'][ok] * o['hscal'][ok, 3:4] * he / (0.9 + 4 * he) / 1.67262192369e-24 * 1e6
    v = o['v'][ok] * o['hscal'][ok, 1:2] * 1000
    s1 = eng.rt2d_batch(n, v, T[ok], WL, rto).cpu().numpy()
    assert np.max(_rel(s1, o['spectrum'][ok])) < 1e-13          # same stage called on its own
    eng.set_geometry(2.0 * geo['i0'], geo['r_map'], R * (1.39 * 71492000))
    s2 = eng.rt2d_batch(n, v, T[ok], WL, rto).cpu().numpy()
    eng.set_geometry(geo['i0'], geo['r_map'], R * (1.39 * 71492000))
    assert np.array_equal(s2, 2.0 * s1)                          # exact: scaling by a power of two
    data = s1[5] / s1[5, 0]
    sig = np.full(200, 1e-3)
    chi2 = eng.chi2_batch(s1, data, sig, norm=float(s1[5, 0])).cpu().numpy()
    ref = (((data[None, :] - s1 / s1[5, 0]) / sig) ** 2).sum(axis=1)
    assert np.allclose(chi2, ref, rtol=1e-12, atol=1e-9) and chi2[5] == 0.0


def test_edge_cases(eng):
    from p_winds_b200.engine import Engine
    e = np.zeros(0)
    o = _forward(eng, e, e, e)
    assert o['spectrum'].shape == (0, 200)
    # a batch in which every model fails
    o = _forward(eng, np.full(3, 5500.), np.full(3, 1e11), np.full(3, 0.9))
    assert (o['status'] == 1).all() and np.isnan(o['spectrum']).all() and np.isnan(o['f_r']).all()
    # one model == the same model inside a batch (B = 1 path uses pixel splitting)
    a = _forward(eng, np.array([9100.]), np.array([1e11]), np.array([0.9]))
    b = _forward(eng, np.array([12000., 9100.]), np.array([1e10, 1e11]), np.array([0.9, 0.9]))
    assert np.max(_rel(a['spectrum'][0], b['spectrum'][1])) < 1e-13
    assert np.array_equal(a['f_3'][0], b['f_3'][1])
    with pytest.raises(ValueError, match='no CPU fallback'):
        Engine.he_opts(1.39, [0] * 6, method='BDF')


# --------------------------------------------------------------------------- draw_transit
def test_draw_transit_reference_signature(eng, oracle):
    """reference tests/test_transit.py:28-39 through transit.draw_transit, plus the
    quickstart geometry and the limb-darkening laws (fixtures from the reference with the
    recovered flatstar rasteriser)."""
    from p_winds_b200 import transit
    geo = golden('geometry')
    i0, depth, rm = transit.draw_transit(planet_to_star_ratio=0.12086, planet_physical_radius=1.39 * 71492000,
                                         impact_parameter=0.0, phase=0.0, supersampling=10, grid_size=101)
    assert abs((i0[30, 30] - 1.249219E-4) / i0[30, 30]) < 5e-3
    assert abs((depth - 0.015012130070238605) / depth) < 5e-3
    assert abs(depth / geo['test_depth'] - 1) < 1e-12
    m = geo['test_i0'] != 0
    assert np.array_equal(i0 != 0, m) and np.max(_rel(i0[m], geo['test_i0'][m])) < 1e-12
    assert np.max(np.abs(rm - geo['test_r_map'])) < 1e-12 * geo['test_r_map'].max()
    i0, depth, rm = transit.draw_transit(0.12086, 1.39 * 71492000, 0.499, 0.0, supersampling=10, grid_size=100)
    m = geo['i0'] != 0
    assert np.array_equal(i0 != 0, m) and np.max(_rel(i0[m], geo['i0'][m])) < 1e-12
    assert abs(depth / geo['depth'] - 1) < 1e-12 and abs(i0.sum() - (1 - depth)) < 1e-12
    assert np.max(np.abs(rm - geo['r_map'])) < 1e-12 * geo['r_map'].max()
    for law, c in (('linear', 0.5), ('quadratic', [0.3, 0.2]), ('square-root', [0.3, 0.2]), ('log', [0.3, 0.2]),
                   ('s3', [0.3, 0.2, 0.1]), ('c4', [0.3, 0.2, 0.1, 0.05])):
        i0, depth, rm = transit.draw_transit(0.1, 1.0, 0.3, -0.2, supersampling=4, grid_size=40,
                                             limb_darkening_law=law, ld_coefficient=c)
        ref = geo['ld_%s_map' % law]
        m = ref != 0
        assert np.array_equal(i0 != 0, m), law
        assert np.max(_rel(i0[m], ref[m])) < 1e-11, law
        assert abs(depth / geo['ld_%s_depth' % law] - 1) < 1e-11, law
    # no supersampling, planet partly off the disk, vs the oracle run live
    ref = oracle.draw_transit(0.2, 2.5, 0.7, 0.33, grid_size=64)
    i0, depth, rm = transit.draw_transit(0.2, 2.5, 0.7, 0.33, grid_size=64)
    assert 0 < ref[1] < 0.04            # the planet straddles the limb
    m = ref[0] != 0
    assert np.array_equal(i0 != 0, m) and np.max(_rel(i0[m], ref[0][m])) < 1e-12
    assert abs(depth / ref[1] - 1) < 1e-12 and np.max(np.abs(rm - ref[2])) < 1e-12 * ref[2].max()
    with pytest.raises(ValueError, match='box'):
        transit.draw_transit(0.1, 1.0, resample_method='lanczos', supersampling=2)


# --------------------------------------------------------------------------- BASELINE configs 2 and 5
def test_config2_lyman_alpha_hydrogen_only(eng, oracle, sed):
    """BASELINE config 2: hydrogen.ion_fraction + Ly-alpha radiative_transfer_2d on a 32x32
    (T0, Mdot) grid, no helium stage (species = neutral H; line data supplied by the caller
    because the reference's lines.py has no Ly-alpha entry)."""
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    geo = golden('geometry')
    w0, f, a = lines.lyman_alpha_properties()
    m_p = 1.67262192369e-27
    wl = np.linspace(1214.0, 1217.4, 200) * 1e-10
    sc = eng.sed_scalars()
    rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
    heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
    rto = Engine.rt_opts(w0, f, a, m_p)
    # tight tolerances, per model vs the oracle
    T = np.array([9100., 11000., 12500.])
    md = np.array([10 ** 10.27, 1e11, 10 ** 9.8])
    hf = np.full(3, 0.9)
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True, rtol=1e-10, atol=1e-13)
    o = eng.forward_batch(T, md, hf, R, ho, heo, rto, wl, species='h1')
    s = oracle.Setup(sed, r=R, i0=geo['i0'], r_map=geo['r_map'], wl=wl, lines=(np.array([w0]), np.array([f]),
                     np.array([a])), mass=m_p, h_tol=dict(rtol=1e-10, atol=1e-13), species='h1')
    spec = o['spectrum'].cpu().numpy()
    for b in range(3):
        ref = oracle.forward_model(s, T[b], md[b], hf[b])
        assert ref['status'] == 0 and int(o['status'][b]) == 0
        assert np.max(_rel(spec[b], ref['spectrum'])) < 1e-5        # Ly-alpha transit depth contract
        assert np.max(_rel(o['f_r'][b].cpu().numpy()[1:], ref['f_r'][1:])) < 1e-6
        assert spec[b].min() < 0.9 * spec[b].max()                   # saturated core (tau up to ~1e7)
    # full 32x32 grid at the reference's tolerances: invariants + determinism
    T0 = np.linspace(10000, 13000, 32)
    Md = np.logspace(9.5, 12, 32)
    TT, MM = np.meshgrid(T0, Md, indexing='ij')
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
    o1 = eng.forward_batch(TT.ravel(), MM.ravel(), np.full(1024, 0.9), R, ho, heo, rto, wl, species='h1')
    sp1, st1 = o1['spectrum'].cpu().numpy(), o1['status'].cpu().numpy()
    o2 = eng.forward_batch(TT.ravel(), MM.ravel(), np.full(1024, 0.9), R, ho, heo, rto, wl, species='h1')
    assert np.array_equal(sp1, o2['spectrum'].cpu().numpy(), equal_nan=True)
    assert (st1 == 0).mean() > 0.98
    ok = st1 == 0
    cont = geo['i0'].sum()
    assert (sp1[ok] >= 0).all() and (sp1[ok] <= cont * (1 + 1e-12)).all()
    assert (sp1[ok].min(axis=1) < 0.95 * cont).all()


def test_config2_lyman_alpha_against_reference_fixture(eng):
    """BASELINE configs[1] against the unmodified reference (tests/golden/anchors_lya.npz, made by
    oracle/make_anchors.py lya: hydrogen.ion_fraction -> parker.structure -> n(H I) ->
    transit.radiative_transfer_2d with the caller-supplied Ly-alpha line).  25 models at tight solver
    tolerances: ion fractions within 1e-6, Ly-alpha transit spectrum within 1e-5, model by model; the 32 x 32
    grid at the reference's defaults: same status for every model, line cores as deep."""
    from p_winds_b200.engine import Engine
    d = golden('anchors_lya')
    m_p = 1.67262192369e-27
    wl = np.linspace(1214.0, 1217.4, 200) * 1e-10
    rto = Engine.rt_opts([1215.67e-10], [0.4164], [6.265e8], m_p)
    sc = eng.sed_scalars()
    heo = Engine.he_opts(1.39, [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')],
                         (1.0, 0.0), True)
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True, rtol=1e-10, atol=1e-13)
    n = len(d['T0'])
    o = eng.forward_batch(d['T0'], d['mdot'], np.full(n, 0.9), R, ho, heo, rto, wl, species='h1')
    st = o['status'].cpu().numpy()
    assert (st == 0).all() and (d['status'] == 0).all()
    f_r, spec = o['f_r'].cpu().numpy(), o['spectrum'].cpu().numpy()
    e_f = np.max(_rel(f_r[:, 1:], d['f_r'][:, 1:]))
    live = d['spectrum'] > 1e-12                   # saturated cores: compare absolutely there
    e_s = np.max(_rel(spec[live], d['spectrum'][live]))
    print('Ly-alpha, 25 models at tight tolerances: f_r %.2e, spectrum %.2e' % (e_f, e_s))
    assert e_f < 1e-6 and e_s < 1e-5
    assert np.max(np.abs(spec[~live] - d['spectrum'][~live])) < 1e-12 if (~live).any() else True
    assert d['spectrum'].min() < 0.5 * d['spectrum'].max()          # the line is there
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
    og = eng.forward_batch(d['grid_T0'], d['grid_mdot'], np.full(1024, 0.9), R, ho, heo, rto, wl, species='h1')
    stg = og['status'].cpu().numpy()
    assert np.array_equal(stg, d['grid_status'].astype(stg.dtype))
    smin = og['spectrum'].cpu().numpy().min(axis=1)
    ok = stg == 0
    # default tolerances: the H stage is only reproducible to its RK45 noise (P3); the core depth follows f_r
    assert np.median(np.abs(smin[ok] - d['grid_spectrum_min'][ok])) < 1e-6
    assert np.max(np.abs(smin[ok] - d['grid_spectrum_min'][ok])) < 5e-3


def test_config5_size_radiative_transfer(eng, oracle):
    """BASELINE config 5 sizes: 512 x 512 transit cells, 4096 wavelength bins, He 10830 +
    C II + O I line sets.  The reference needs 11.9 s and 34 GB per model and line set
    here; the oracle is run on a wavelength subset (the 'average' method is independent per
    wavelength) and the full-size result is checked through its invariants."""
    from p_winds_b200 import transit, lines
    prof = golden('he_3_profile')
    r, n, v = prof['r'][::5].copy(), prof['n'][::5].copy(), prof['v'][::5].copy()
    i0, depth, rm = transit.draw_transit(0.12086, 1.39 * 71492000, 0.499, 0.0, grid_size=512, supersampling=2)
    assert i0.shape == (512, 512) and abs(i0.sum() - (1 - depth)) < 1e-12
    ref_geo = oracle.draw_transit(0.12086, 1.39 * 71492000, 0.499, 0.0, grid_size=512, supersampling=2)
    assert np.max(np.abs(i0 - ref_geo[0])) < 1e-15 and abs(depth - ref_geo[1]) < 1e-14
    m_p = 1.67262192369e-27
    he, c2, o1 = lines.he_3_properties(), lines.c_ii_properties(), lines.o_i_properties()
    sets = [((np.array(he[:3]), np.array(he[3:6]), np.array([he[6]] * 3)), 4 * m_p,
             np.linspace(1.0827, 1.0832, 4096) * 1e-6),
            ((np.array(c2[:3]), np.array(c2[3:6]), np.array(c2[6:9])), 12 * m_p,
             np.linspace(1332.9, 1337.1, 4096) * 1e-10),
            ((np.array(o1[:3]), np.array(o1[3:6]), np.array(o1[6:9])), 16 * m_p,
             np.linspace(1300.9, 1307.1, 4096) * 1e-10)]
    for (w0, f, a), mass, wl in sets:
        s = transit.radiative_transfer_2d(i0, rm, r, n, v, w0, f, a, wl, 9E3, mass, bulk_los_velocity=-2E3)
        assert s.shape == (4096,) and np.isfinite(s).all()
        assert (s > 0).all() and (s <= i0.sum() * (1 + 1e-12)).all()
        sub = np.arange(0, 4096, 128)
        ref = oracle.radiative_transfer_2d(i0, rm, r, n, v, w0, f, a, wl[sub], 9E3, mass, v_bulk=-2E3)
        assert np.max(_rel(s[sub], ref)) < 1e-10
        # linear in the intensity map, exactly (power-of-two scaling)
        s2 = transit.radiative_transfer_2d(2.0 * i0, rm, r, n, v, w0, f, a, wl, 9E3, mass, bulk_los_velocity=-2E3)
        assert np.array_equal(s2, 2.0 * s)


# --------------------------------------------------------------------------- SURVEY 8(f-1)
def test_retrieval_loop_phase_mean_convolution_likelihood(built, oracle):
    """The step around the path in every retrieval (docs/source/advanced_tutorial.ipynb cells
    11, 13, 15): spectrum averaged over sampled transit phases with the geometric depth added,
    Gaussian instrumental profile by astropy-style `convolve(boundary='extend')`, Gaussian
    log-likelihood.  Own engine instance: the phase slots are per-context state."""
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    prof = golden('he_3_profile')
    w = lines.he_3_properties()
    w0, f, a = np.array(w[:3]), np.array(w[3:6]), np.array([w[6]] * 3)
    wl = np.linspace(1.0826, 1.0834, 169) * 1E-6
    r_si, n_si, v_si = prof['r'], prof['n'], prof['v']
    r_phys = float(r_si[0])
    phases = np.linspace(-0.5, 0.5, 5)
    maps = [oracle.draw_transit(0.12086, r_phys, 0.499, ph, 60, 3) for ph in phases]
    maps = [(m[0], m[2], m[1]) for m in maps]                      # (I0, r_map, depth)
    eng = Engine(0)
    eng.set_geometry_phases(maps, r_si)
    rto = Engine.rt_opts(w0, f, a, M_HE, bulk_los_velocity=-2E3)
    scale = np.array([1.0, 0.3, 3.0])
    T = np.array([9000., 7000., 11000.])
    spec = eng.rt2d_batch(n_si[None, :] * scale[:, None], np.tile(v_si, (3, 1)), T, wl, rto).cpu().numpy()
    for b in range(3):
        ref = oracle.transmission_model(maps, r_si, n_si * scale[b], v_si, w0, f, a, wl, T[b], M_HE, v_wind=-2E3)
        assert np.max(_rel(spec[b], ref)) < 1e-10, b
    assert np.all(spec > 0.9) and np.all(spec <= 1.0 + 1e-12)      # depth added back: continuum at 1
    # one phase with its depth = radiative_transfer_2d + depth
    eng.set_geometry_phases(maps[2:3], r_si)
    one = eng.rt2d_batch(n_si[None, :], v_si[None, :], T[:1], wl, rto).cpu().numpy()[0]
    ref = oracle.radiative_transfer_2d(maps[2][0], maps[2][1], r_si, n_si, v_si, w0, f, a, wl, 9000., M_HE,
                                       v_bulk=-2E3) + maps[2][2]
    assert np.max(_rel(one, ref)) < 1e-10
    # instrumental profile (cell 3): Gaussian sampled on the wavelength grid, odd length
    wl_a = wl * 1e10
    sigma_wl = 3.7 / (2 * (2 * np.log(2)) ** 0.5) / 299792.458 * np.mean(wl_a)
    kern = 1 / sigma_wl / (2 * np.pi) ** 0.5 * np.exp(-0.5 * (wl_a - np.mean(wl_a)) ** 2 / sigma_wl ** 2)
    conv = eng.convolve_extend_batch(spec, kern).cpu().numpy()
    for b in range(3):
        assert np.max(np.abs(conv[b] - oracle.convolve_extend(spec[b], kern))) < 1e-14, b
    asym = np.linspace(0.2, 1.0, 31)                               # kernel flip matters here
    got = eng.convolve_extend_batch(spec[0], asym).cpu().numpy()[0]
    assert np.max(np.abs(got - oracle.convolve_extend(spec[0], asym))) < 1e-14
    with pytest.raises(ValueError, match='Kernel size must be odd'):
        eng.convolve_extend_batch(spec, np.ones(4))
    # log-likelihood (cell 15), failed models -> -inf
    rng = np.random.default_rng(2)
    y = conv[0] + 2e-4 * rng.standard_normal(169)
    yerr = np.full(169, 2e-4) * rng.uniform(0.8, 1.2, 169)
    import torch
    status = torch.tensor([0, 2, 3], dtype=torch.int32, device='cuda')
    ll = eng.loglike_batch(conv, y, yerr, status).cpu().numpy()
    assert abs(ll[0] / oracle.log_likelihood(conv[0], y, yerr) - 1) < 1e-12
    assert ll[1] == -np.inf
    assert ll[2] == -np.inf          # status 3: the reference's result is undefined -> failed by default
    eng.set_relax_warning_policy(True)
    ll = eng.loglike_batch(conv, y, yerr, status).cpu().numpy()
    eng.set_relax_warning_policy(False)
    assert ll[1] == -np.inf
    assert abs(ll[2] / oracle.log_likelihood(conv[2], y, yerr) - 1) < 1e-12
    assert ll[0] > ll[2]                                            # the true model wins


def test_pixel_sum_kernels_agree_bit_for_bit(built, oracle):
    """The two launch shapes of the pixel sum (k_rt_sum2): one CTA per model and wavelength tile summing
    all pixel segments and phases with the segment sums combined in shared memory (large batches), and
    one CTA per pixel segment and phase + k_rt_reduce (few models) -- same segments, same order ->
    np.array_equal, for the quickstart sizes, a ragged tile (Nw = 150), five phases with their depths,
    Nw = 4096, and an optically thick case that takes the general exp path (tau >> 11)."""
    import os
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    prof = golden('he_3_profile')
    geo = golden('geometry')
    w = lines.he_3_properties()
    rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, M_HE, bulk_los_velocity=-2E3)
    r_si, n_si, v_si = prof['r'][::5].copy(), prof['n'][::5].copy(), prof['v'][::5].copy()
    rng = np.random.default_rng(11)
    B = 160
    scale = 10 ** rng.uniform(-1, 1, B)
    n = n_si[None, :] * scale[:, None]
    v = np.tile(v_si, (B, 1))
    T = rng.uniform(7000, 12000, B)
    phases = np.linspace(-0.4, 0.4, 5)
    maps = [oracle.draw_transit(0.12086, float(r_si[0]), 0.499, ph, 60, 3) for ph in phases]
    maps = [(m[0], m[2], m[1]) for m in maps]
    cases = [('quickstart', lambda e: e.set_geometry(geo['i0'], geo['r_map'], r_si), np.linspace(1.0827, 1.0832, 200) * 1e-6),
             ('ragged', lambda e: e.set_geometry(geo['test_i0'], geo['test_r_map'], r_si), np.linspace(1.0827, 1.0832, 150) * 1e-6),
             ('phases', lambda e: e.set_geometry_phases(maps, r_si), np.linspace(1.0826, 1.0834, 169) * 1e-6),
             ('wide', lambda e: e.set_geometry(geo['i0'], geo['r_map'], r_si), np.linspace(1.0827, 1.0832, 4096) * 1e-6),
             ('thick', lambda e: e.set_geometry(geo['i0'], geo['r_map'], r_si), np.linspace(1.0827, 1.0832, 200) * 1e-6)]
    for name, setup, wl in cases:
        if name == 'thick':
            n = n * 300.0
        eng = Engine(0)
        setup(eng)
        out = {}
        for key in ('PWB_RT_SPLIT', 'PWB_RT_NOSPLIT'):
            os.environ[key] = '1'
            try:
                out[key] = eng.rt2d_batch(n, v, T, wl, rto, return_tau=(name == 'ragged'))
            finally:
                del os.environ[key]
        a, b2 = out['PWB_RT_SPLIT'], out['PWB_RT_NOSPLIT']
        if name == 'ragged':
            assert np.array_equal(a[1].cpu().numpy(), b2[1].cpu().numpy()), name
            a, b2 = a[0], b2[0]
        a, b2 = a.cpu().numpy(), b2.cpu().numpy()
        assert np.isfinite(a).all() and (a.min() < 0.999 * a.max())
        assert np.array_equal(a, b2), (name, np.max(np.abs(a - b2)))


@pytest.mark.parametrize('nw', [1, 2, 3, 63, 511, 513, 1025])
def test_pixel_sum_odd_shapes(built, oracle, nw):
    """Tile and pair edge cases of k_rt_sum2 (one bin, an odd bin count, one bin past a tile boundary,
    fewer than 64 threads per CTA), a single model, both launch shapes, against the oracle."""
    import os
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    prof = golden('he_3_profile')
    geo = golden('geometry')
    w = lines.he_3_properties()
    w0, f, a = np.array(w[:3]), np.array(w[3:6]), np.array([w[6]] * 3)
    rto = Engine.rt_opts(w0, f, a, M_HE, bulk_los_velocity=-2E3)
    r_si, n_si, v_si = prof['r'][::5].copy(), prof['n'][::5].copy(), prof['v'][::5].copy()
    wl = np.linspace(1.0829, 1.0831, nw) * 1e-6 if nw > 1 else np.array([1.08303e-6])
    eng = Engine(0)
    eng.set_geometry(geo['test_i0'], geo['test_r_map'], r_si)
    ref = oracle.radiative_transfer_2d(geo['test_i0'], geo['test_r_map'], r_si, n_si, v_si, w0, f, a, wl[:40], 9000., M_HE,
                                       v_bulk=-2E3)
    out = {}
    for key in ('PWB_RT_SPLIT', 'PWB_RT_NOSPLIT'):
        os.environ[key] = '1'
        try:
            out[key] = eng.rt2d_batch(n_si[None, :], v_si[None, :], np.array([9000.]), wl, rto).cpu().numpy()[0]
        finally:
            del os.environ[key]
    assert np.array_equal(out['PWB_RT_SPLIT'], out['PWB_RT_NOSPLIT'])
    assert out['PWB_RT_SPLIT'].shape == (nw,)
    assert np.max(_rel(out['PWB_RT_SPLIT'][:40], ref)) < 1e-10
    # a geometry without a single lit pixel inside the radius grid: the spectrum is the unocculted flux
    far = np.full_like(geo['test_r_map'], 10 * r_si[-1])
    eng.set_geometry(geo['test_i0'], far, r_si)
    flat = eng.rt2d_batch(n_si[None, :], v_si[None, :], np.array([9000.]), wl, rto).cpu().numpy()[0]
    assert np.max(np.abs(flat - geo['test_i0'].sum())) < 1e-12


def test_config3_share_tidal_grid_properties(eng):
    """BASELINE configs[3] at full size for one GPU: every 8th model (one rank's interleaved share)
    of the 256 x 256 (T0, Mdot) x 3 h_fraction grid with the tidal Parker profile -- 24 576 models
    in one call.  No oracle can run this in seconds; check what must hold at any size: the
    result of a model does not depend on the batch it is in (cost-ordered helium schedule, packed
    lanes, phase slots), statuses are the reference's classes, physical bounds."""
    T0 = np.linspace(9000, 13000, 256)
    md = np.logspace(9.5, 12, 256)
    hf = np.array([0.85, 0.90, 0.95])
    HH, TT, MM = np.meshgrid(hf, T0, md, indexing='ij')
    T, M, H = TT.ravel()[0::8].copy(), MM.ravel()[0::8].copy(), HH.ravel()[0::8].copy()
    assert T.size == 24576
    big = _forward(eng, T, M, H, tidal=True)
    st = big['status']
    assert set(np.unique(st)) <= {0, 1, 2, 3}
    ok = (st == 0) | (st == 3)
    assert ok.mean() > 0.98
    assert np.isfinite(big['spectrum'][ok]).all() and np.isnan(big['spectrum'][~ok]).all()
    sp = big['spectrum'][ok] / float(golden('geometry')['i0'].sum())
    # (status 3 = odeint warned inside the relax loop: the reference returns whatever it has, e.g.
    # populations clamped to 1 -> a nearly opaque disc; only the trivial bounds hold for those)
    assert sp.max() <= 1.0 + 1e-12 and sp.min() >= 0.0
    good = st == 0
    assert (big['spectrum'][good] / float(golden('geometry')['i0'].sum())).min() > 0.2
    for k in ('f_r', 'f_1', 'f_3'):
        assert (big[k][good] >= 0).all() and (big[k][good] <= 1.0).all(), k
    # a scattered sub-batch (different batch size, order and lane packing) gives the same bits
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(T.size, 700, replace=False))[::-1].copy()
    small = _forward(eng, T[pick], M[pick], H[pick], tidal=True)
    assert np.array_equal(small['status'], st[pick])
    for k in ('f_r', 'f_1', 'f_3', 'spectrum'):
        assert np.array_equal(small[k], big[k][pick], equal_nan=True), k


# --------------------------------------------------------------------------- SURVEY 8(f-3)
def test_oxygen_reference_signature_and_batch(eng, oracle, sed):
    """oxygen.ion_fraction on the GPU (k_oxygen: Radau / LSODA replicas) against the unmodified
    reference (tests/golden/oxygen.npz): scalars of oxygen.radiative_processes to 1e-10, ion
    fractions inside the reference's own 1-ulp noise envelope (fixture key `noise`; see
    tests/test_hostsim_stages.py::_oxygen_tolerance), exceptions and return arity."""
    from p_winds_b200 import oxygen
    from p_winds_b200.engine import Engine
    d = golden('oxygen')
    g = golden('solar_sed')
    sp = {'wavelength': g['wavelength'], 'flux_lambda': g['flux_lambda'], 'wavelength_unit': 'angstrom',
          'flux_unit': 'erg / (s cm2 angstrom)'}
    scal = np.array(oxygen.radiative_processes(sp))
    assert np.max(_rel(scal, d['scalars'])) < 1e-10
    for T in (7000., 9100., 12000.):
        assert oxygen.electron_impact_ionization(T) == oracle.o_electron_impact_ionization(T)
        assert oxygen.recombination(T) == oracle.o_recombination(T)
        assert oxygen.charge_transfer(T) == oracle.o_charge_transfer(T)
    for i in range(len(d['T0'])):
        kw = dict(relax_solution=bool(d['relax'][i]))
        if str(d['method'][i]) == 'odeint':
            kw['method'] = 'odeint'
        else:
            kw.update(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        f, rates = oxygen.ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], d['f_he_ion'][i], 1.39,
                                       float(d['T0'][i]), float(d['h_fraction'][i]), float(d['vs'][i]),
                                       float(d['rs'][i]), float(d['rhos'][i]), sp, return_rates=True, **kw)
        tol = 4.0 * float(d['noise'][i]) + 1e-6
        err = np.max(_rel(f[1:], d['f_o'][i][1:]))
        assert err < tol, (i, err, tol)
        assert set(rates) == {'photoionization', 'recombination', 'e_impact_ionization', 'charge_exchange_HII',
                              'charge_exchange_HI'}
    # the whole fixture as one batch, plus a model that starts from a failed upstream stage
    sel = [i for i in range(len(d['T0'])) if str(d['method'][i]) == 'Radau' and d['relax'][i] and d['rtol'][i] > 1e-6]
    import torch
    status = torch.zeros(len(sel), dtype=torch.int32, device='cuda')
    status[1] = 1
    opts = Engine.o_opts(1.39, scal, relax_solution=True)
    o = eng.o_ion_fraction_batch(d['T0'][sel], d['h_fraction'][sel], d['vs'][sel], d['rs'][sel], d['rhos'][sel],
                                 d['r'], d['v'][sel], d['rho'][sel], d['f_h'][sel], d['f_he_ion'][sel], opts,
                                 status=status)
    fo = o['f_o'].cpu().numpy()
    assert np.isnan(fo[1]).all() and o['status'].cpu().numpy().tolist() == [0, 1, 0, 0]
    for k, i in enumerate(sel):
        if k != 1:
            assert np.max(_rel(fo[k][1:], d['f_o'][i][1:])) < 4.0 * float(d['noise'][i]) + 1e-6
    with pytest.raises(ValueError, match='not available on the GPU path'):
        oxygen.ion_fraction(d['r'], d['v'][0], d['rho'][0], d['f_h'][0], d['f_he_ion'][0], 1.39, 9100., 0.9,
                            float(d['vs'][0]), float(d['rs'][0]), float(d['rhos'][0]), sp, method='BDF')


def test_carbon_reference_signature(eng, oracle):
    """carbon.ion_fraction on the GPU (k_carbon: LSODA / Radau replicas) against the unmodified
    reference (tests/golden/carbon.npz): scalars to 1e-10, C II / C III fractions inside the
    reference's own 1-ulp noise envelope, the RuntimeError of the default `odeint` on the dense
    model, and the reference's rates dict (including its 'ionization_CII' = electron-impact slip)."""
    from p_winds_b200 import carbon
    d = golden('carbon')
    g = golden('solar_sed')
    sp = {'wavelength': g['wavelength'], 'flux_lambda': g['flux_lambda'], 'wavelength_unit': 'angstrom',
          'flux_unit': 'erg / (s cm2 angstrom)'}
    scal = np.array(carbon.radiative_processes(sp))
    assert np.max(_rel(scal, d['scalars'])) < 1e-10
    for T in (7000., 9100., 12000.):
        assert carbon.electron_impact_ionization(T) == oracle.c_electron_impact_ionization(T)
        assert carbon.recombination(T) == oracle.c_recombination(T)
        assert carbon.charge_transfer(T) == oracle.c_charge_transfer(T)
    keys = ('ionization_CI', 'ionization_CII', 'recombination_CII', 'recombination_CIII', 'e_impact_ion_CI',
            'e_impact_ion_CII', 'charge_exchange_CI_HII', 'charge_exchange_CI_HeII', 'charge_exchange_CII_HI',
            'charge_exchange_CIII_HI', 'charge_exchange_CIII_HeI')
    n_fail = 0
    for i in range(len(d['T0'])):
        kw = dict(relax_solution=bool(d['relax'][i]))
        if str(d['method'][i]) != 'odeint':
            kw.update(method=str(d['method'][i]), rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        args = (d['r'], d['v'][i], d['rho'][i], d['f_h'][i], d['f_he_ion'][i], 1.39, float(d['T0'][i]),
                float(d['h_fraction'][i]), float(d['vs'][i]), float(d['rs'][i]), float(d['rhos'][i]), sp)
        if d['status'][i] == 2:
            n_fail += 1
            with pytest.raises(RuntimeError, match='odeint'):
                carbon.ion_fraction(*args, **kw)
            continue
        f2, f3, rates = carbon.ion_fraction(*args, return_rates=True, **kw)
        tol2, tol3 = 4.0 * float(d['noise'][i]) + 1e-5, 4.0 * float(d['noise3'][i]) + 1e-3
        big = d['f_ciii'][i] > 1e-2 * d['f_ciii'][i].max()
        e2, e3 = np.max(_rel(f2[1:], d['f_cii'][i][1:])), np.max(_rel(f3[big], d['f_ciii'][i][big]))
        assert e2 < tol2 and e3 < tol3, (i, e2, tol2, e3, tol3)
        assert tuple(rates) == keys
        assert np.array_equal(rates['ionization_CII'], rates['e_impact_ion_CII'])
        m = (d['f_cii'][i] > 1e-3) & (d['f_cii'][i] < 0.99)
        m[0] = False
        if m.any():
            assert np.max(_rel(rates['recombination_CII'][m], d['rates'][i][2][m])) < 100 * tol2
    assert n_fail == 2


def test_helium_ion_fraction_reference_signature(eng):
    """helium.ion_fraction on the GPU (k_he_ion) against the unmodified reference
    (tests/golden/helium_ion.npz): scalars of radiative_processes(combined_ionization=True) to
    1e-10, fractions inside the reference's noise envelope + solver tolerance."""
    from p_winds_b200 import helium
    d = golden('helium_ion')
    g = golden('solar_sed')
    sp = {'wavelength': g['wavelength'], 'flux_lambda': g['flux_lambda'], 'wavelength_unit': 'angstrom',
          'flux_unit': 'erg / (s cm2 angstrom)'}
    scal = np.array(helium.radiative_processes(sp, combined_ionization=True))
    assert np.max(_rel(scal, d['scalars'])) < 1e-10
    for i in range(len(d['T0'])):
        kw = dict(relax_solution=bool(d['relax'][i]))
        if str(d['method'][i]) == 'odeint':
            kw['method'] = 'odeint'
        else:
            kw.update(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
        f = helium.ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], 1.39, float(d['T0'][i]),
                                float(d['h_fraction'][i]), float(d['vs'][i]), float(d['rs'][i]),
                                float(d['rhos'][i]), sp, **kw)
        floor = 1e-5 if str(d['method'][i]) == 'odeint' else max(1e-6, float(d['rtol'][i]))
        err = np.max(_rel(f[1:], d['f_he_ion'][i][1:]))
        assert err < 4.0 * float(d['noise'][i]) + floor, (i, err)


def test_metals_chain_config5_lines_on_gpu(eng, oracle, sed):
    """BASELINE configs[4] names C II and O I lines next to He 10830.  Here their number
    densities come from the GPU as well: H -> Parker -> He populations -> O II and C II / C III
    fractions (Radau) -> n(O I), n(C II) -> radiative transfer at the O I and C II lines, for a
    small batch, against the oracle running the same chain.  Tight solver tolerances on both
    sides, so that the comparison is meaningful per model (tier P2): spectra <= 1e-5 relative."""
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines
    T = np.array([9100., 11000.])
    md = np.array([10 ** 10.27, 10 ** 10.8])
    hf = np.array([0.9, 0.9])
    ho, heo, _ = _opts(eng, True, False)
    mu0 = (1 + 4 * (1 - hf) / hf) / (1 + (1 - hf) / hf)
    h = eng.h_ion_fraction_batch(T, md, hf, mu0, R, ho)
    vs, rs, rhos = (h['scal'][:, k].contiguous() for k in (1, 2, 3))
    he = eng.he_population_batch(T, hf, vs, rs, rhos, R, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
    f_he_ion = 1 - he['f_1'] - he['f_3']
    sc = eng.sed_scalars()
    tight = dict(relax_solution=True, method='Radau', rtol=1e-10, atol=1e-13)
    oo = Engine.o_opts(1.39, [sc[k] for k in ('o_phi', 'o_a', 'o_a_h', 'o_a_he')], **tight)
    co = Engine.c_opts(1.39, [sc[k] for k in ('c_phi_1', 'c_phi_2', 'c_a_1', 'c_a_2', 'c_a_h_1', 'c_a_h_2', 'c_a_he')],
                       **tight)
    ox = eng.o_ion_fraction_batch(T, hf, vs, rs, rhos, R, h['v'], h['rho'], h['f_r'], f_he_ion, oo,
                                  status=he['status'].clone())
    ca = eng.c_ion_fraction_batch(T, hf, vs, rs, rhos, R, h['v'], h['rho'], h['f_r'], f_he_ion, co,
                                  status=he['status'].clone())
    assert (ox['status'].cpu().numpy() == 0).all() and (ca['status'].cpu().numpy() == 0).all()
    o_frac, c_frac = 10 ** (8.69 - 12.00), 10 ** (8.43 - 12.00)
    m_p = 1.67262192369e-27
    r_si = R * 1.39 * 71492000
    o1, c2 = lines.o_i_properties(), lines.c_ii_properties()
    wl_o = np.linspace(1300.9, 1307.1, 120) * 1e-10
    wl_c = np.linspace(1332.9, 1337.1, 120) * 1e-10
    geo = golden('geometry')
    hfd, rho_d, rhos_d = eng.dev(hf)[:, None], h['rho'], rhos[:, None]
    v_si = h['v'] * vs[:, None] * 1000
    n_o = rho_d * rhos_d * o_frac / (hfd + 4 * (1 - hfd) + 16 * o_frac) / 1.67262192369e-24 * (1 - ox['f_o']) * 1e6
    n_c = rho_d * rhos_d * c_frac / (hfd + 4 * (1 - hfd) + 12 * c_frac) / 1.67262192369e-24 * ca['f_cii'] * 1e6
    s_o = eng.rt2d_batch(n_o, v_si, T, wl_o, Engine.rt_opts(o1[:3], o1[3:6], o1[6:9], 16 * m_p)).cpu().numpy()
    s_c = eng.rt2d_batch(n_c, v_si, T, wl_c, Engine.rt_opts(c2[:3], c2[3:6], c2[6:9], 12 * m_p)).cpu().numpy()
    he_rates = oracle.he_radiative_processes(sed)
    for b in range(2):
        oh = oracle.h_ion_fraction(R, 1.39, T[b], hf[b], md[b], 0.73, mu0[b], sed=sed, exact_phi=True,
                                   relax_solution=True, rtol=1e-10, atol=1e-13)
        ovs = oracle.sound_speed(T[b], oh['mu_bar'])
        ors = oracle.radius_sonic_point(0.73, ovs)
        orhos = oracle.density_sonic_point(md[b], ors, ovs)
        ov, orho = oracle.structure(R * 1.39 / ors)
        ohe = oracle.he_population(R, ov, orho, oh['f_r'], 1.39, T[b], hf[b], ovs, ors, orhos, rates=he_rates,
                                   initial_state=np.array([1., 0.]), relax_solution=True, method='LSODA',
                                   rtol=1e-10, atol=1e-16)
        fhe = 1 - ohe['f_1'] - ohe['f_3']
        kw = dict(relax_solution=True, method='Radau', rtol=1e-10, atol=1e-13)
        oo_ = oracle.o_ion_fraction(R, ov, orho, oh['f_r'], fhe, 1.39, T[b], hf[b], ovs, ors, orhos, sed=sed, **kw)
        oc_ = oracle.c_ion_fraction(R, ov, orho, oh['f_r'], fhe, 1.39, T[b], hf[b], ovs, ors, orhos, sed=sed, **kw)
        assert np.max(_rel(ox['f_o'][b].cpu().numpy()[1:], oo_['f_o'][1:])) < 1e-6
        assert np.max(_rel(ca['f_cii'][b].cpu().numpy()[1:], oc_['f_cii'][1:])) < 1e-6
        on_o = orho * orhos * o_frac / (hf[b] + 4 * (1 - hf[b]) + 16 * o_frac) / 1.67262192369e-24 * (1 - oo_['f_o']) * 1e6
        on_c = orho * orhos * c_frac / (hf[b] + 4 * (1 - hf[b]) + 12 * c_frac) / 1.67262192369e-24 * oc_['f_cii'] * 1e6
        ref_o = oracle.radiative_transfer_2d(geo['i0'], geo['r_map'], r_si, on_o, ov * ovs * 1000, np.array(o1[:3]),
                                             np.array(o1[3:6]), np.array(o1[6:9]), wl_o, T[b], 16 * m_p)
        ref_c = oracle.radiative_transfer_2d(geo['i0'], geo['r_map'], r_si, on_c, ov * ovs * 1000, np.array(c2[:3]),
                                             np.array(c2[3:6]), np.array(c2[6:9]), wl_c, T[b], 12 * m_p)
        assert np.max(_rel(s_o[b], ref_o)) < 1e-5 and np.max(_rel(s_c[b], ref_c)) < 1e-5, b
        assert ref_c.min() < 0.999 * ref_c.max()      # the C II lines are actually there


def test_pixel_sum_exponentials_on_gpu(eng):
    """The two e^x of k_rt_sum2 against numpy: pw_exp_neg_wide (table of 4097 powers of two, series to
    r^4; tau < 11) and pw_exp_neg_sum_sh (64-entry table, exponent arithmetic, exact zero below e^-708).
    Relative error <= 3e-16 scaled by max(1, |x|) (the one-constant argument reduction), i.e. an absolute
    error below one ulp of an O(1) sum."""
    rng = np.random.default_rng(8)
    x = -np.concatenate([10 ** rng.uniform(-14, np.log10(10.99), 300000), rng.uniform(0, 10.99, 300000),
                         [0.0, 1e-300, 10.99, np.log(2) / 256, np.log(2) / 512, 5.5]])
    w, g = (t.cpu().numpy() for t in eng.selftest_exp(x))
    ref = np.exp(x)
    assert np.max(np.abs(w / ref - 1) / np.maximum(1.0, -x)) < 3e-16
    assert np.max(np.abs(w - ref)) < 1.2e-16
    assert np.max(np.abs(g / ref - 1) / np.maximum(1.0, -x)) < 3e-16
    assert w[600000] == 1.0 and g[600000] == 1.0
    far = -np.array([11.5, 50.0, 700.0, 707.9, 708.5, 1e4, 1e7, 1e12])
    w2, g2 = (t.cpu().numpy() for t in eng.selftest_exp(far))
    assert np.isnan(w2).all()
    assert np.max(np.abs(g2[:4] / np.exp(far[:4]) - 1) / 708) < 3e-16 and (g2[4:] == 0).all()


def test_step_arithmetic_on_gpu(eng):
    """The device build of the LSODA step replaces `a / b` by the fast path of the IEEE division
    algorithm (csrc/pw_common.cuh PW_DIV) and pow(a, 1/k) by a Newton root (pw_root).  Over the
    operand ranges of the step the quotient must be the correctly rounded one, bit for bit, the
    special cases (0, inf, nan, denormal divisors) must keep their IEEE results, and the root
    must sit within 2e-15 of the true one."""
    rng = np.random.default_rng(11)
    n = 200000
    a = np.concatenate([rng.standard_normal(n) * 10.0 ** rng.uniform(-40, 40, n),
                        rng.uniform(-2, 2, n), [0.0, 1.0, -1.0, np.inf, np.nan, 1e300, 1.0, 1.0, 5e-324, 1.0]])
    b = np.concatenate([rng.standard_normal(n) * 10.0 ** rng.uniform(-40, 40, n),
                        1.49012e-8 * (1 + rng.uniform(0, 1, n)), [1.0, 0.0, np.inf, 2.0, 1.0, 1e-300, 5e-324, np.nan, 3.0, -0.0]])
    k = np.concatenate([rng.integers(1, 14, 2 * n), np.arange(1, 11)]).astype(np.int32)
    q, r = eng.selftest_arith(a, b, k)
    q = q.cpu().numpy(); r = r.cpu().numpy()
    with np.errstate(all='ignore'):
        want = a / b
    assert np.array_equal(q, want, equal_nan=True)
    # roots: a > 0 in the range the step-size selection produces (error ratios)
    ar = np.abs(a[:2 * n]) + 1e-300
    _, r2 = eng.selftest_arith(ar, np.ones_like(ar), k[:2 * n])
    r2 = r2.cpu().numpy()
    ref = ar ** (1.0 / k[:2 * n])
    m = (ar > 1e-30) & (ar < 1e30)
    assert np.max(np.abs(r2[m] / ref[m] - 1)) < 2e-15
    assert np.max(np.abs(r2[~m] / ref[~m] - 1)) < 1e-13
