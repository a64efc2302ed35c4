"""Parity tier P1 for the physics stages: the host build of the device code against
the oracle (== the reference, see test_oracle_vs_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest
from scipy.special import wofz

from conftest import P, golden

R = np.logspace(0, np.log10(20), 100)


@pytest.fixture(scope='module')
def tables(hostsim, sed):
    n = len(sed.wl)
    gw, sh, she, sc = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(32)
    hostsim.hs_sed_setup(P(sed.wl), P(sed.flux), n, P(gw), P(sh), P(she), P(sc))
    return gw, sh, she, sc


def test_sed_tables_and_scalars(tables, oracle, sed):
    gw, sh, she, sc = tables
    ref = np.array(list(oracle.h_radiative_processes(sed)) + list(oracle.he_radiative_processes(sed)))
    assert np.max(np.abs(sc[:8] / ref - 1)) < 5e-15
    assert (int(sc[8]), int(sc[9]), int(sc[10])) == (912, 504, 911)
    nh = int(sc[8])
    assert np.max(np.abs(sh[:nh] / oracle.sigma_h(sed.wl[:nh]) - 1)) < 5e-15
    ref_he = oracle.sigma_he_total(sed.wl[:nh])
    m = ref_he > 0
    assert np.array_equal(she[:nh][~m], ref_he[~m])
    assert np.max(np.abs(she[:nh][m] / ref_he[m] - 1)) < 1e-12


def test_parker_structure(hostsim, oracle):
    g = golden('parker')
    r = np.ascontiguousarray(g['r'])
    v, rho = np.zeros_like(r), np.zeros_like(r)
    hostsim.hs_parker(P(r), len(r), P(v), P(rho))
    m = np.abs(r - 1.0) > 1e-6     # at r == 1 the reference's secant is only sqrt(eps) accurate
    assert np.max(np.abs(v[m] / g['v'][m] - 1)) < 2e-13
    assert np.max(np.abs(rho[m] / g['rho'][m] - 1)) < 2e-13
    assert abs(v[~m][0] - 1.0) < 1e-15 and abs(g['v'][~m][0] - 1.0) < 1e-6   # reference tests/test_parker.py
    hostsim.hs_parker_tidal.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 5 + [C.c_void_p] * 2
    hostsim.hs_parker_tidal(P(r), len(r), float(g['cs']), float(g['rs_tidal']), 0.73, 1.07, 0.04634, P(v), P(rho))
    assert np.max(np.abs(v[m] / g['v_tidal'][m] - 1)) < 2e-13
    assert np.max(np.abs(rho[m] / g['rho_tidal'][m] - 1)) < 2e-13


def _h_host(hostsim, tables, T, md, h, tidal, r, exact=1, relax=1, mu0=None, rtol=1e-3, atol=1e-6,
            phi_abs=0., a0=0.):
    gw, sh, she, sc = tables
    he_h = (1 - h) / h
    if mu0 is None:
        mu0 = (1 + 4 * he_h) / (1 + he_h)
    model = np.array([T, md, h, mu0])
    opts = np.array([1.39, 0.73, 1.07 if tidal else 1.0, 0.04634 if tidal else 1.0, 0.0, 0.01, 10, relax, exact,
                     phi_abs, a0, rtol, atol, 0., 0., 1 if tidal else 0], dtype=float)
    nr = len(r)
    f, v, rho, scal = np.zeros(nr), np.zeros(nr), np.zeros(nr), np.zeros(8)
    st = hostsim.hs_h_ion_fraction(P(model), P(opts), P(np.ascontiguousarray(r)), nr, P(gw), P(sh), P(she),
                                   int(sc[8]), P(f), P(v), P(rho), P(scal))
    return st, f, v, rho, scal


@pytest.mark.parametrize('tidal', [False, True])
@pytest.mark.parametrize('T0,md', [(9100., 10 ** 10.27), (9100., 1e11), (12000., 1e10), (7000., 1e11),
                                   (13000., 10 ** 9.5)])
def test_hydrogen_stage_trace(hostsim, tables, oracle, sed, tidal, T0, md):
    """Same solver outcome, same number of relax passes, same total nfev as
    scipy's RK45 inside the reference algorithm; values to rounding."""
    mu0 = (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9)
    o = oracle.h_ion_fraction(R, 1.39, T0, 0.9, md, 0.73, mu0, *((1.07, 0.04634) if tidal else (1.0, 1.0)),
                              sed=sed, exact_phi=True, relax_solution=True)
    st, f, v, rho, scal = _h_host(hostsim, tables, T0, md, 0.9, tidal, R)
    assert st == o['status']
    assert int(scal[5]) == sum(o['nfev'])
    if st == 0:
        assert int(scal[4]) == o['n_relax']
        assert np.max(np.abs(f[1:] / o['f_r'][1:] - 1)) < 1e-8
        assert abs(scal[0] / o['mu_bar'] - 1) < 1e-9


def test_hydrogen_work_limit(hostsim, tables):
    """The reference's RK45 has no step limit: at T0 = 5000 K, Mdot = 7.2e9 g/s, h = 0.85 (tidal) it
    crawls for 1.9e8 right-hand sides before it fails (oracle/make_anchors.py gave the unmodified
    reference up after 90 s).  With Rk45Opts::max_nfev the engine stops early with status 5."""
    lim = C.c_longlong.in_dll(hostsim, 'hs_h_max_nfev')
    try:
        lim.value = 200000
        st, f, v, rho, scal = _h_host(hostsim, tables, 5000.0, 7.196857e+09, 0.85, True, R)
        assert st == 5 and 200000 <= scal[5] < 200100
        st, *_ = _h_host(hostsim, tables, 9100.0, 10 ** 10.27, 0.9, False, R)
        assert st == 0
    finally:
        lim.value = 0


def test_hydrogen_reference_test_configuration(hostsim, tables, oracle, sed):
    """reference tests/test_hydrogen.py (500 radii, mu_0 = 0.7): approximate and exact phi."""
    g = golden('hydrogen')
    sc = tables[3]
    st, f, *_ = _h_host(hostsim, tables, 9E3, 8E10, 0.9, False, g['r500'], exact=0, mu0=0.7,
                        phi_abs=sc[0], a0=sc[1])
    assert st == 0 and abs((f[-1] - 0.998737) / f[-1]) < 1e-4
    assert np.max(np.abs(f[1:] / g['f_approx'][1:] - 1)) < 1e-8
    st, f, *_ = _h_host(hostsim, tables, 9E3, 8E10, 0.9, False, g['r500'], exact=1, mu0=0.7)
    assert st == 0 and abs((f[-1] - 0.998913) / f[-1]) < 1e-4
    assert np.max(np.abs(f[1:] / g['f_exact'][1:] - 1)) < 1e-8
    # monochromatic EUV channel (quickstart.ipynb last cells)
    phi, a0 = oracle.h_radiative_processes_mono(1000.)
    st, f, *_ = _h_host(hostsim, tables, 9100., 10 ** 10.27, 0.9, False, np.linspace(1.0, 20, 500), exact=0,
                        mu0=0.7613, phi_abs=phi, a0=a0)
    assert st == 0 and np.max(np.abs(f[1:] / g['f_mono'][1:] - 1)) < 1e-8


def _he_host(hostsim, rates, T, h, vs, rs, rhos, r, v, rho, fh, method=0, rtol=0., atol=0., relax=1):
    model = np.array([T, h, vs, rs, rhos])
    opts = np.array([1.39, 1.0, 0.0, relax, 10, 0.01, method, rtol, atol, 0., 0.], dtype=float)
    nr = len(r)
    f1, f3, scal = np.zeros(nr), np.zeros(nr), np.zeros(4)
    st = hostsim.hs_he_population(P(model), P(opts), P(rates), P(np.ascontiguousarray(r)),
                                  P(np.ascontiguousarray(v)), P(np.ascontiguousarray(rho)),
                                  P(np.ascontiguousarray(fh)), nr, P(f1), P(f3), P(scal), None)
    return st, f1, f3, scal


def _he_inputs(oracle, sed, T0, md, tight):
    mu0 = (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9)
    kw = dict(rtol=1e-10, atol=1e-13) if tight else {}
    o = oracle.h_ion_fraction(R, 1.39, T0, 0.9, md, 0.73, mu0, sed=sed, exact_phi=True, relax_solution=True, **kw)
    vs = oracle.sound_speed(T0, o['mu_bar'])
    rs = oracle.radius_sonic_point(0.73, vs)
    rhos = oracle.density_sonic_point(md, rs, vs)
    v, rho = oracle.structure(R * 1.39 / rs)
    return o['f_r'], vs, rs, rhos, v, rho


@pytest.mark.parametrize('T0,md', [(9100., 10 ** 10.27), (11000., 1e11)])
def test_helium_stage_tight_tolerance(hostsim, oracle, sed, T0, md):
    """Tier P2 on the host: at rtol=1e-10 / atol=1e-16 the LSODA replica and scipy's
    LSODA agree far inside the 1e-6 contract; relax counts identical."""
    rates = np.array(oracle.he_radiative_processes(sed))
    f_r, vs, rs, rhos, v, rho = _he_inputs(oracle, sed, T0, md, True)
    oh = oracle.he_population(R, v, rho, f_r, 1.39, T0, 0.9, vs, rs, rhos, rates=tuple(rates),
                              initial_state=np.array([1., 0.]), relax_solution=True, method='LSODA',
                              rtol=1e-10, atol=1e-16)
    st, f1, f3, scal = _he_host(hostsim, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r, method=1, rtol=1e-10,
                                atol=1e-16)
    assert st == oh['status'] == 0 and int(scal[0]) == oh['n_relax']
    assert np.max(np.abs(f1 / oh['f_1'] - 1)) < 2e-8
    ipk = np.argmax(oh['f_3'])
    assert abs(f3[ipk] / oh['f_3'][ipk] - 1) < 2e-8
    big = oh['f_3'] > 1e-3 * oh['f_3'][ipk]
    assert np.max(np.abs(f3[big] / oh['f_3'][big] - 1)) < 1e-6


@pytest.mark.parametrize('T0,md', [(9100., 10 ** 10.27), (9100., 1e11), (10000., 1e12)])
def test_helium_stage_default_tolerance_envelope(hostsim, oracle, sed, T0, md):
    """Tier P3 on the host: at odeint's default tolerance the output is only defined
    to the integrator's own noise (f1 ~5e-6 median / 4e-5 max,
    f3 at its peak ~1e-5 median / 1.6e-4 max for *any* valid integrator, including
    the reference re-run with 1-ulp input noise).  Work counters within 10 %."""
    rates = np.array(oracle.he_radiative_processes(sed))
    f_r, vs, rs, rhos, v, rho = _he_inputs(oracle, sed, T0, md, False)
    oh = oracle.he_population(R, v, rho, f_r, 1.39, T0, 0.9, vs, rs, rhos, rates=tuple(rates),
                              initial_state=np.array([1., 0.]), relax_solution=True)
    st, f1, f3, scal = _he_host(hostsim, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r)
    assert st == oh['status'] == 0 and int(scal[0]) == oh['n_relax']
    assert np.max(np.abs(f1 / oh['f_1'] - 1)) < 1e-4
    ipk = np.argmax(oh['f_3'])
    assert abs(f3[ipk] / oh['f_3'][ipk] - 1) < 1e-3
    for k, key in ((1, 'nfe'), (2, 'nst'), (3, 'nje')):
        assert abs(scal[k] / sum(oh[key]) - 1) < 0.10, key


def test_helium_rhs_literal_logexp_build(built, hostsim, oracle, sed):
    """-DPW_HE_LOGEXP: the right-hand side with the reference's literal exp(log k1 + log rho + log coef)
    (helium.py:637-652) instead of the product k1 * rho * coef the engine uses.  Both
    host builds on the same model: same status and relax count, populations equal far inside odeint's
    own noise -- the detour through logarithms only adds ~1e-14 of rounding."""
    import subprocess
    lib = os.path.join(os.path.dirname(built.HOSTSIM), 'libhostsim_logexp.so')
    src = os.path.join(os.path.dirname(built.HOSTSIM), '..', 'hostsim', 'hostsim.cu')
    want = built.source_hash([src])
    if built._lib_hash(lib, b'hostsim src:') != want:
        subprocess.check_call([built.NVCC, '-Wno-deprecated-gpu-targets', '-O2', '-std=c++17', '--expt-relaxed-constexpr',
                               '--expt-extended-lambda', '-Xcompiler', '-fPIC,-ffp-contract=off', '-shared',
                               '-DPW_HE_LOGEXP', '-DPW_SRC_HASH="%s"' % want, '-o', lib, src])
    alt = C.CDLL(lib)
    rates = np.array(oracle.he_radiative_processes(sed))
    T0, md = 9100., 10 ** 10.27
    f_r, vs, rs, rhos, v, rho = _he_inputs(oracle, sed, T0, md, False)
    st0, a1, a3, sc0 = _he_host(hostsim, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r)
    st1, b1, b3, sc1 = _he_host(alt, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r)
    assert st0 == st1 == 0 and sc0[0] == sc1[0]
    assert not np.array_equal(a1, b1)                       # it is a different build
    # default tolerance: ~1e-14 of input noise is amplified to odeint's own noise floor (f_1 ~5e-6 median,
    # 5e-5 at the 90th percentile over tests/golden/anchors_p3.npz) -- the envelope of the P3 host test
    assert np.max(np.abs(a1 / b1 - 1)) < 1e-4
    ipk = np.argmax(a3)
    assert abs(a3[ipk] / b3[ipk] - 1) < 1e-3
    # tight tolerances: the two forms agree to rounding
    st0, a1, a3, _ = _he_host(hostsim, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r, method=1, rtol=1e-10, atol=1e-16)
    st1, b1, b3, _ = _he_host(alt, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r, method=1, rtol=1e-10, atol=1e-16)
    assert st0 == st1 == 0 and np.max(np.abs(a1 / b1 - 1)) < 1e-9


def test_wofz_against_scipy(hostsim):
    rng = np.random.default_rng(0)
    sets = [(np.linspace(-15, 15, 2001), np.full(2001, 1.4e-4)),        # He 10830
            (np.linspace(-35, 35, 2001), np.full(2001, 4.9e-4)),        # Ly-alpha
            (rng.uniform(-60, 60, 20000), 10 ** rng.uniform(-8, 1.5, 20000)),
            (10 ** rng.uniform(-8, -2, 5000), 10 ** rng.uniform(-6, 0, 5000)),
            (10 ** rng.uniform(1, 9, 5000), 10 ** rng.uniform(-12, 3, 5000))]
    for x, y in sets:
        x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        out = np.zeros_like(x)
        hostsim.hs_wofz_re(P(x), P(y), len(x), P(out))
        assert np.max(np.abs(out / wofz(x + 1j * y).real - 1)) < 2e-14


def test_optical_depth_average_against_reference(hostsim):
    """tau[Nr, Nw] of transit.optical_depth_2d ('average'), reference fixture."""
    d, ka = golden('he_3_profile'), golden('rt_known_answer')
    w0 = np.array([1.082909114, 1.083025010, 1.083033977]) * 1e-6
    f = np.array([5.9902e-02, 1.7974e-01, 2.9958e-01])
    a = np.array([1.0216e+07] * 3)
    wl = np.ascontiguousarray(ka['wl'])
    tau = np.zeros((500, 150))
    hostsim.hs_tau_average.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + \
        [C.c_double] * 3 + [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    hostsim.hs_tau_average(P(np.ascontiguousarray(d['r'])), P(np.ascontiguousarray(d['n'])),
                           P(np.ascontiguousarray(d['v'])), None, 500, P(w0), P(f), P(a), 3,
                           4 * 1.67262192369e-27, -2e3, 0.0, 200, 0, 9e3, P(wl), 150, P(tau), None)
    ref = ka['tau_average']
    m = ref > 0
    assert np.array_equal(tau[~m], ref[~m])
    assert np.max(np.abs(tau[m] / ref[m] - 1)) < 1e-12


def test_pil_ellipse_spans_against_pillow(hostsim, oracle):
    """The span table behind pwb_draw_transit vs Pillow itself (flatstar's rasteriser)
    and vs the oracle's restatement, incl. clipping, odd/even diameters, off-grid planets."""
    from PIL import Image, ImageDraw
    hostsim.hs_pil_ellipse_rows.argtypes = [C.c_double] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(1)
    cases = [(505.0, 505.0, 505.0, 1010), (505.0, 505.0, 505 * 0.12086, 1010), (50.5, 50.5, 50.5, 101)]
    for _ in range(60):
        n = int(rng.integers(20, 400))
        cases.append((rng.uniform(-0.2, 1.2) * n, rng.uniform(-0.2, 1.2) * n, rng.uniform(0.5, 0.7 * n), n))
    for cx, cy, rad, n in cases:
        lo, hi = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        hostsim.hs_pil_ellipse_rows(cx - rad, cy - rad, cx + rad, cy + rad, n, P(lo), P(hi))
        mask = np.zeros((n, n))
        for r in range(n):
            if hi[r] >= lo[r]:
                mask[r, lo[r]:hi[r] + 1] = 1.0
        im = Image.new('1', (n, n))
        ImageDraw.Draw(im).ellipse([(cx - rad, cy - rad), (cx + rad, cy + rad)], outline=1, fill=1)
        assert np.array_equal(mask, np.array(im).astype(float)), (cx, cy, rad, n)
        assert np.array_equal(mask, oracle._disk((cx, cy), rad, n))


@pytest.mark.parametrize('T0,md', [(10428.571428571428, 10 ** 11.365079365079366),
                                   (12714.285714285714, 10 ** 11.642857142857142),
                                   (10619.047619047618, 10 ** 11.365079365079366)])
def test_helium_relax_warning_stops_the_loop(hostsim, oracle, sed, T0, md):
    """Status 3 (helium.py:736, the unchecked odeint of the relax loop).  The reference's output
    is not reproducible for these models (scipy leaves the rows after the failing point
    uninitialised); what is defined is the class and the pass that warned.  Engine (host build)
    and oracle both stop there and return the last completed pass."""
    rates = np.array(oracle.he_radiative_processes(sed))
    f_r, vs, rs, rhos, v, rho = _he_inputs(oracle, sed, T0, md, False)
    oh = oracle.he_population(R, v, rho, f_r, 1.39, T0, 0.9, vs, rs, rhos, rates=tuple(rates),
                              initial_state=np.array([1., 0.]), relax_solution=True)
    st, f1, f3, scal = _he_host(hostsim, rates, T0, 0.9, vs, rs, rhos, R, v, rho, f_r)
    assert st == oh['status'] == 3
    assert int(scal[0]) == oh['n_relax']
    assert np.all(np.isfinite(f1)) and np.all((f1 >= 0) & (f1 <= 1)) and np.all((f3 >= 0) & (f3 <= 1))
    # the last completed pass is an ordinary default-tolerance solve: same envelope as the status-0 test
    assert np.max(np.abs(f1 / oh['f_1'] - 1)) < 1e-4
    ipk = np.argmax(oh['f_3'])
    assert abs(f3[ipk] / oh['f_3'][ipk] - 1) < 1e-3


def test_exp_neg(hostsim):
    """pw_exp_neg (table + degree-6 series, the e^{-tau} of the pixel sum) against numpy."""
    rng = np.random.default_rng(5)
    x = -np.concatenate([10 ** rng.uniform(-12, 2.85, 200000), rng.uniform(0, 40, 200000),
                         [0.0, 1e-300, 707.9, 708.0, 708.5, 745.0, 800.0, np.inf]])
    out = np.zeros_like(x)
    hostsim.hs_exp_neg.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    hostsim.hs_exp_neg(P(x), x.size, P(out))
    ref = np.exp(x)
    live = x > -708.0
    assert np.max(np.abs(out[live] / ref[live] - 1)) < 3e-16
    assert np.all(out[~live] == 0.0)
    assert out[0 + 400000] == 1.0


def test_exp_neg_sum(hostsim):
    """pw_exp_neg_sum (64-entry table, degree-5 series, one-constant reduction: the e^{-tau} of the
    pixel sum): absolute error below half an ulp of the O(1) sum everywhere, relative error at
    rounding level where the term matters."""
    rng = np.random.default_rng(6)
    x = -np.concatenate([10 ** rng.uniform(-12, 2.85, 200000), rng.uniform(0, 40, 200000),
                         [0.0, 1e-300, 707.9, 708.0, 708.5, 745.0, 800.0, np.inf]])
    out = np.zeros_like(x)
    hostsim.hs_exp_neg_sum.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    hostsim.hs_exp_neg_sum(P(x), x.size, P(out))
    ref = np.exp(x)
    live = x > -708.0
    assert np.max(np.abs(out[live] - ref[live])) < 1.2e-16          # one ulp below 1
    near = x > -2.0
    assert np.max(np.abs(out[near] / ref[near] - 1)) < 3e-16
    assert np.max(np.abs(out[live] / ref[live] - 1) / np.maximum(1.0, -x[live])) < 3e-16
    assert np.all(out[~live] == 0.0)
    assert out[0 + 400000] == 1.0


def _oxygen_tolerance(d, i):
    """The reference's oxygen.ion_fraction is reproducible only to its round-off noise floor:
    scipy's Radau builds its Jacobian by finite differences, which turns 1-ulp input noise into
    another Newton / step sequence (fixture key `noise`: largest relative change of f_o over
    six 1-ulp-perturbed runs of the unmodified reference, 2e-4 .. 3e-3 for Radau at any rtol,
    <= 1e-5 for odeint).  Tolerance = 4 x that floor + 1e-6 (SURVEY.md 8(c) P3 protocol)."""
    return 4.0 * float(d['noise'][i]) + 1e-6


def test_oxygen_stage_against_reference(hostsim, oracle, sed):
    """o_ion_fraction_model (Radau / odeint replicas) on the host build against the
    unmodified reference's oxygen.ion_fraction (tests/golden/oxygen.npz) -- same relax-loop
    outcome, right-hand-side count within the reference's own scatter, values inside its
    noise envelope.  (Step-for-step identity of the Radau replica itself is shown on
    pure-arithmetic problems in tests/test_hostsim_solvers.py.)"""
    d = golden('oxygen')
    rates = np.ascontiguousarray(d['scalars'])
    meth = {'Radau': 0, 'odeint': 1}
    for i in range(len(d['T0'])):
        model = np.array([d['T0'][i], d['h_fraction'][i], d['vs'][i], d['rs'][i], d['rhos'][i]], dtype=float)
        opts = np.array([1.39, 10 ** (8.69 - 12.00), 0.0, d['relax'][i], 10, 0.01, meth[str(d['method'][i])],
                         d['rtol'][i], d['atol'][i], 0., 0.], dtype=float)
        f, scal, rr = np.zeros(100), np.zeros(4), np.zeros((5, 100))
        st = hostsim.hs_o_ion_fraction(P(model), P(opts), P(rates), P(np.ascontiguousarray(d['r'])),
                                       P(np.ascontiguousarray(d['v'][i])), P(np.ascontiguousarray(d['rho'][i])),
                                       P(np.ascontiguousarray(d['f_h'][i])), P(np.ascontiguousarray(d['f_he_ion'][i])),
                                       100, P(f), P(scal), P(rr))
        assert st == d['status'][i] == 0
        tol = _oxygen_tolerance(d, i)
        err = np.max(np.abs(f[1:] / d['f_o'][i][1:] - 1))
        assert err < tol, (i, err, tol)
        # the rates multiply (1 - f) or f: compare where neither factor is lost to cancellation
        m = (d['f_o'][i] > 1e-3) & (d['f_o'][i] < 0.99)
        m[0] = False
        if m.any():
            assert np.max(np.abs(rr[:, m] / d['rates'][i][:, m] - 1)) < 100 * tol
        if str(d['method'][i]) == 'Radau':
            kw = dict(rtol=float(d['rtol'][i]), atol=float(d['atol'][i]))
            o = oracle.o_ion_fraction(d['r'], d['v'][i], d['rho'][i], d['f_h'][i], d['f_he_ion'][i], 1.39,
                                      float(d['T0'][i]), float(d['h_fraction'][i]), float(d['vs'][i]),
                                      float(d['rs'][i]), float(d['rhos'][i]), sed=sed,
                                      relax_solution=bool(d['relax'][i]), **kw)
            assert int(scal[0]) == o['n_relax']
            assert abs(scal[1] / sum(o['nfev']) - 1) < 0.12     # the reference's own nfev scatters by +-5 %


def _carbon_tolerances(d, i):
    """Same protocol as _oxygen_tolerance; C III is compared where it is within a factor 100 of
    its peak (fixture keys `noise`, `noise3`).  The default odeint path is far less noise
    sensitive than Radau's finite-difference Jacobian: floor 1e-5 / 1e-3."""
    return 4.0 * float(d['noise'][i]) + 1e-5, 4.0 * float(d['noise3'][i]) + 1e-3


def test_carbon_stage_against_reference(hostsim):
    """c_ion_fraction_model (LSODA / Radau replicas) on the host build against the
    unmodified reference's carbon.ion_fraction (tests/golden/carbon.npz), including the dense
    model on which the default `odeint` gives up (status 2 = RuntimeError, carbon.py:541-546)."""
    d = golden('carbon')
    rates = np.ascontiguousarray(d['scalars'])
    meth = {'odeint': 0, 'Radau': 1}
    n_fail = 0
    for i in range(len(d['T0'])):
        model = np.array([d['T0'][i], d['h_fraction'][i], d['vs'][i], d['rs'][i], d['rhos'][i]], dtype=float)
        opts = np.array([1.39, 10 ** (8.43 - 12.00), 0.0, 0.0, d['relax'][i], 10, 0.01, meth[str(d['method'][i])],
                         d['rtol'][i], d['atol'][i], 0., 0.], dtype=float)
        f2, f3, scal, rr = np.zeros(100), np.zeros(100), np.zeros(4), np.zeros((11, 100))
        st = hostsim.hs_c_ion_fraction(P(model), P(opts), P(rates), P(np.ascontiguousarray(d['r'])),
                                       P(np.ascontiguousarray(d['v'][i])), P(np.ascontiguousarray(d['rho'][i])),
                                       P(np.ascontiguousarray(d['f_h'][i])), P(np.ascontiguousarray(d['f_he_ion'][i])),
                                       100, P(f2), P(f3), P(scal), P(rr))
        assert st == d['status'][i], i
        if st != 0:
            n_fail += 1
            assert np.isnan(f2).all() and np.isnan(f3).all()
            continue
        tol2, tol3 = _carbon_tolerances(d, i)
        e2 = np.max(np.abs(f2[1:] / d['f_cii'][i][1:] - 1))
        big = d['f_ciii'][i] > 1e-2 * d['f_ciii'][i].max()
        e3 = np.max(np.abs(f3[big] / d['f_ciii'][i][big] - 1))
        assert e2 < tol2 and e3 < tol3, (i, e2, tol2, e3, tol3)
    assert n_fail == 2


def _heion_tolerance(d, i):
    """4 x the reference's own 1-ulp noise (fixture key `noise`) + the solver's relative tolerance
    (Radau: the step sequence either survives last-bit differences, ~1e-9, or it does not, ~rtol)."""
    floor = 1e-5 if str(d['method'][i]) == 'odeint' else max(1e-6, float(d['rtol'][i]))
    return 4.0 * float(d['noise'][i]) + floor


def test_helium_ion_fraction_stage_against_reference(hostsim):
    """he_ion_fraction_model on the host build against the unmodified reference's
    helium.ion_fraction (tests/golden/helium_ion.npz)."""
    d = golden('helium_ion')
    rates = np.ascontiguousarray(d['scalars'])
    meth = {'Radau': 0, 'odeint': 1}
    errs = []
    for i in range(len(d['T0'])):
        model = np.array([d['T0'][i], d['h_fraction'][i], d['vs'][i], d['rs'][i], d['rhos'][i]], dtype=float)
        opts = np.array([1.39, 0.0, d['relax'][i], 10, 0.01, meth[str(d['method'][i])], d['rtol'][i], d['atol'][i],
                         0., 0.], dtype=float)
        f, scal = np.zeros(100), np.zeros(4)
        st = hostsim.hs_he_ion_fraction(P(model), P(opts), P(rates), P(np.ascontiguousarray(d['r'])),
                                        P(np.ascontiguousarray(d['v'][i])), P(np.ascontiguousarray(d['rho'][i])),
                                        P(np.ascontiguousarray(d['f_h'][i])), 100, P(f), P(scal))
        assert st == 0
        err = np.max(np.abs(f[1:] / d['f_he_ion'][i][1:] - 1))
        errs.append(err)
        assert err < _heion_tolerance(d, i), (i, err, _heion_tolerance(d, i))
    # where the reference's step sequence is stable under 1-ulp noise the replica reproduces it
    # to rounding -- the realistic-RHS counterpart of the trace tests in test_hostsim_solvers.py
    stable = [e for e, nz, m in zip(errs, d['noise'], d['method']) if nz < 1e-8 and str(m) == 'Radau']
    assert len(stable) >= 2 and min(stable) < 1e-7
