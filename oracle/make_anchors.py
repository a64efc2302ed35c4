"""Anchor fixtures of parity tier P3 and of the helium failure classes, made by running
the UNMODIFIED reference (SURVEY.md 8(c) "Golden vectors to carry", "Parity protocol").

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/make_anchors.py scan      # status map of the 64x64 bench grid -> oracle/_build/
    python oracle/make_anchors.py ridge     # off-grid models of the strip where helium fails
    python oracle/make_anchors.py p3        # tests/golden/anchors_p3.npz
    python oracle/make_anchors.py classes   # tests/golden/he_failure_classes.npz
    python oracle/make_anchors.py tight     # tests/golden/anchors_tight_grid.npz
    python oracle/make_anchors.py lya       # tests/golden/anchors_lya.npz (BASELINE configs[1])
    python oracle/make_anchors.py tidal     # tests/golden/tidal_grid_status.npz (sample of BASELINE configs[3])

``anchors_p3.npz``: 272 (T0, Mdot, h) triples (144 non-tidal, 128 tidal, T0 from 5000 K so
that the models the reference fails on are in it).  For each triple the reference's own
outputs at its default tolerances, plus that model's *self-noise*: the change of each output
in each of K = 4 re-runs of the reference with the SED flux multiplied by 1 + 2.2e-16 N(0, 1)
(one ulp) -- keys ``noise_*_runs`` [n, K], their maxima ``noise_*`` -- and how often its
status flipped (``status_flips``).  f_3 is compared at its peak (elsewhere it sits below
odeint's atol = 1.49e-8 and is not determined in relative terms).  Status 9 = the
reference did not finish the model within 90 s (explicit RK45 crawling on a stiff start).

``he_failure_classes.npz``: models of the bench grid on which the reference's helium stage
fails -- class 2 (the first ``odeint`` warns -> RuntimeError, helium.py:691-697) and class 3
(an ``odeint`` of the relax loop warns, helium.py:736, which the reference does not check).
Class-3 models are run three times: scipy's odeint leaves the rows after the failing output
point uninitialised, so the reference's result for them is not reproducible; the fixture
records the spread as evidence and the tests compare the status class only.
"""
import os
import sys
import warnings
import multiprocessing as mp

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
BUILD = os.path.join(HERE, '_build')
K_NOISE = 4
LIMIT_S = 90      # per model; the reference crawls (explicit RK45 on a stiff start) on a few cold models
_G = {}


class _Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise _Timeout()


def _init():
    os.environ['OMP_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = '1'
    warnings.simplefilter('ignore')
    import make_golden as mg
    _G['mg'] = mg
    _G['sp'] = mg.spectrum_dict()
    geo = np.load(os.path.join(OUT, 'geometry.npz'))
    _G['geom'] = (geo['i0'], geo['r_map'])


def _run_tight(args):
    """One model at the solver tolerances of the 100-model P2 grid: H rtol 1e-12 / atol 1e-15, He
    method='LSODA' rtol 1e-10 / atol 1e-16.  (At H rtol 1e-10 the reference's own round-off noise in f_r
    reaches 8e-7, SURVEY.md 8(c): too close to the 1e-6 it is held to; at 1e-12 it is two orders lower.  The He
    tolerances stay: scipy's LSODA at rtol 1e-12 / atol 1e-18 does not come back within an hour here.)"""
    T, md, h, tidal = args
    import refchain
    saved = dict(refchain.TIGHT_H)
    refchain.TIGHT_H.update(rtol=1e-12, atol=1e-15)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return _G['mg'].one_model(_G['sp'], _G['geom'], float(T), float(md), float(h), bool(tidal), True)
    finally:
        refchain.TIGHT_H.clear()
        refchain.TIGHT_H.update(saved)


def _run(args):
    """One model; seed < 0: unperturbed, else 1-ulp noise on the SED flux."""
    T, md, h, tidal, seed = args
    mg, sp = _G['mg'], _G['sp']
    if seed >= 0:
        rng = np.random.default_rng(seed)
        sp = dict(sp)
        sp['flux_lambda'] = sp['flux_lambda'] * (1 + 2.2e-16 * rng.standard_normal(len(sp['flux_lambda'])))
    import signal
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(LIMIT_S)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            o = mg.one_model(sp, _G['geom'], float(T), float(md), float(h), bool(tidal), False)
    except _Timeout:
        # status 9: the unmodified reference did not finish this model within LIMIT_S seconds
        o = mg.one_model.__globals__['np'] and dict(
            status=9, f_r=np.full(100, np.nan), mu_bar=np.nan, v=np.full(100, np.nan), rho=np.full(100, np.nan),
            f_1=np.full(100, np.nan), f_3=np.full(100, np.nan), spectrum=np.full(200, np.nan), vs=np.nan,
            rs=np.nan, rhos=np.nan)
        print('timeout (> %d s): T0 %.4f Mdot %.6e h %.2f tidal %d seed %d' % (LIMIT_S, T, md, h, tidal, seed),
              flush=True)
    finally:
        signal.alarm(0)
    return o


def _rel(a, b, sel=slice(None)):
    a, b = np.asarray(a)[sel], np.asarray(b)[sel]
    with np.errstate(all='ignore'):
        return float(np.max(np.abs(a / b - 1)))


def bench_grid():
    T0 = np.linspace(10000, 13000, 64)
    md = np.logspace(9.5, 12, 64)
    TT, MM = np.meshgrid(T0, md, indexing='ij')
    return TT.ravel(), MM.ravel()


def scan(pool):
    T, M = bench_grid()
    res = pool.map(_run, [(t, m, 0.9, False, -1) for t, m in zip(T, M)], chunksize=8)
    st = np.array([o['status'] for o in res])
    os.makedirs(BUILD, exist_ok=True)
    np.savez(os.path.join(BUILD, 'bench_grid_status.npz'), T0=T, mdot=M, status=st)
    np.savez_compressed(os.path.join(OUT, 'bench_grid_status.npz'), T0=T, mdot=M, status=st.astype(np.int8))
    print('bench grid status histogram', {int(k): int((st == k).sum()) for k in np.unique(st)})
    return T, M, st


def scan_tidal(pool):
    """tests/golden/tidal_grid_status.npz: the status the reference ends with on a 32 x 32 x 2 sample of the
    north-star grid (BASELINE configs[3]: T0 9000-13000 K, Mdot 10^9.5-10^12 g/s, h = 0.85 / 0.95, tidal term)."""
    TT, MM, HH = np.meshgrid(np.linspace(9000, 13000, 32), np.logspace(9.5, 12, 32), np.array([0.85, 0.95]), indexing='ij')
    T, M, H = TT.ravel(), MM.ravel(), HH.ravel()
    res = pool.map(_run, [(t, m, h, True, -1) for t, m, h in zip(T, M, H)], chunksize=8)
    st = np.array([o['status'] for o in res], dtype=np.int8)
    smin = np.array([np.nanmin(o['spectrum']) if o['status'] in (0, 3) else np.nan for o in res])
    np.savez_compressed(os.path.join(OUT, 'tidal_grid_status.npz'), T0=T, mdot=M, h_fraction=H, status=st,
                        spectrum_min=smin)
    print('tidal grid status histogram', {int(k): int((st == k).sum()) for k in np.unique(st)})


def ridge_grid():
    """Off-grid models in the strip of the bench grid where the helium stage fails
    (10^11.2 .. 10^12 g/s): T0 between the bench grid's rows, 100 mass-loss rates."""
    T0 = np.linspace(10000, 13000, 64)
    T0 = (0.5 * (T0[:-1] + T0[1:]))[::2]
    md = np.logspace(11.2, 12.0, 100)
    TT, MM = np.meshgrid(T0, md, indexing='ij')
    return TT.ravel(), MM.ravel()


def scan_ridge(pool):
    T, M = ridge_grid()
    res = pool.map(_run, [(t, m, 0.9, False, -1) for t, m in zip(T, M)], chunksize=8)
    st = np.array([o['status'] for o in res])
    os.makedirs(BUILD, exist_ok=True)
    np.savez(os.path.join(BUILD, 'ridge_status.npz'), T0=T, mdot=M, status=st)
    print('ridge status histogram', {int(k): int((st == k).sum()) for k in np.unique(st)})
    return T, M, st


def p3(pool):
    trip = [(T, md, 0.9, False) for T in np.linspace(5000, 13000, 12) for md in np.logspace(9.5, 12, 12)]
    trip += [(T, md, h, True) for T in np.linspace(5000, 13000, 8) for md in np.logspace(9.5, 12, 8)
             for h in (0.85, 0.95)]
    jobs = [(T, md, h, td, s) for (T, md, h, td) in trip for s in [-1] + [1000 + k for k in range(K_NOISE)]]
    res = pool.map(_run, jobs, chunksize=4)
    n = len(trip)
    per = 1 + K_NOISE
    keys = ('status', 'f_r', 'mu_bar', 'v', 'rho', 'f_1', 'f_3', 'spectrum', 'vs', 'rs', 'rhos')
    d = {k: np.array([res[i * per][k] for i in range(n)]) for k in keys}
    d['T0'] = np.array([t[0] for t in trip])
    d['mdot'] = np.array([t[1] for t in trip])
    d['h_fraction'] = np.array([t[2] for t in trip])
    d['tidal'] = np.array([t[3] for t in trip])
    names = ('noise_f_r', 'noise_mu', 'noise_f_1', 'noise_f_3_peak', 'noise_spectrum')
    runs = {k: np.full((n, K_NOISE), np.nan) for k in names}   # deviation of each perturbed run
    flips = np.zeros(n, dtype=np.int32)
    for i in range(n):
        base = res[i * per]
        for k in range(1, per):
            o = res[i * per + k]
            if o['status'] != base['status']:
                flips[i] += 1
                continue
            if base['status'] in (1, 2, 9):
                continue
            runs['noise_f_r'][i, k - 1] = _rel(o['f_r'], base['f_r'], slice(1, None))
            runs['noise_mu'][i, k - 1] = abs(o['mu_bar'] / base['mu_bar'] - 1)
            if base['status'] != 0:
                continue
            pk = slice(int(np.argmax(base['f_3'])), int(np.argmax(base['f_3'])) + 1)
            runs['noise_f_1'][i, k - 1] = _rel(o['f_1'], base['f_1'])
            runs['noise_f_3_peak'][i, k - 1] = _rel(o['f_3'], base['f_3'], pk)
            runs['noise_spectrum'][i, k - 1] = _rel(o['spectrum'], base['spectrum'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        nz = {k: np.nan_to_num(np.nanmax(runs[k], axis=1), nan=0.0) for k in names}
    for k in names:
        d[k + '_runs'] = runs[k]
    d.update(nz)
    d['status_flips'] = flips
    d['k_noise'] = K_NOISE
    np.savez_compressed(os.path.join(OUT, 'anchors_p3.npz'), **d)
    st = d['status']
    print('anchors_p3: %d models, status histogram %s, %d with a status flip under 1-ulp noise'
          % (n, {int(k): int((st == k).sum()) for k in np.unique(st)}, int((flips > 0).sum())))
    ok = st == 0
    for k in nz:
        print('  %-16s median %.2e  p90 %.2e  max %.2e' % (k, np.median(nz[k][ok]), np.percentile(nz[k][ok], 90),
                                                          nz[k][ok].max()))


def tight(pool):
    """tests/golden/anchors_tight_grid.npz: 100 models over the parameter box of the bench grids at the
    tight tolerances, where BASELINE's 1e-6 (fractions) / 1e-5 (transit depth) hold model by model."""
    trip = [(T, md, 0.9, False) for T in np.linspace(9000, 13000, 8) for md in np.logspace(9.5, 12, 8)]
    trip += [(T, md, h, True) for T in np.linspace(9000, 13000, 6) for md in np.logspace(9.5, 11.8, 3)
             for h in (0.85, 0.95)]
    res = pool.map(_run_tight, trip, chunksize=1)
    keys = ('status', 'f_r', 'mu_bar', 'v', 'rho', 'f_1', 'f_3', 'spectrum', 'vs', 'rs', 'rhos')
    d = {k: np.array([o[k] for o in res]) for k in keys}
    d['T0'] = np.array([t[0] for t in trip])
    d['mdot'] = np.array([t[1] for t in trip])
    d['h_fraction'] = np.array([t[2] for t in trip])
    d['tidal'] = np.array([t[3] for t in trip])
    np.savez_compressed(os.path.join(OUT, 'anchors_tight_grid.npz'), **d)
    st = d['status']
    print('anchors_tight_grid: %d models, status histogram %s' % (len(trip), {int(k): int((st == k).sum()) for k in np.unique(st)}))


LYA = dict(w0=1215.67e-10, f=0.4164, a=6.265e8, m=1.67262192369e-27)   # SURVEY.md 8(d), cfg 2 (caller-supplied line)


def _run_lya(args):
    """BASELINE configs[1]: hydrogen-only chain + Ly-alpha radiative transfer through the unmodified reference
    (hydrogen.ion_fraction -> parker.structure -> n(H I) -> transit.radiative_transfer_2d)."""
    T, md, h, tight = args
    mg = _G['mg']
    from p_winds import hydrogen, parker, transit
    he_h = (1 - h) / h
    mu_0 = (1 + 4 * he_h) / (1 + he_h)
    o = dict(status=0, f_r=np.full(100, np.nan), mu_bar=np.nan, spectrum=np.full(200, np.nan))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            f_r, mu = hydrogen.ion_fraction(mg.R, mg.R_PL, float(T), float(h), float(md), mg.M_PL, mu_0,
                                            spectrum_at_planet=_G['sp'], exact_phi=True, initial_f_ion=0.0,
                                            relax_solution=True, return_mu=True, **(mg.TIGHT_H if tight else {}))
        except RuntimeError:
            o['status'] = 1
            return o
        vs = parker.sound_speed(float(T), mu)
        rs = parker.radius_sonic_point(mg.M_PL, vs)
        rhos = parker.density_sonic_point(float(md), rs, vs)
        v, rho = parker.structure(mg.R * mg.R_PL / rs)
        n_h1 = (1 - f_r) * rho * rhos * h / (h + 4 * (1 - h)) / 1.67262192369e-24 * 1E6
        wl = np.linspace(1214.0, 1217.4, 200) * 1e-10
        i0, r_map = _G['geom']
        spec = transit.radiative_transfer_2d(i0, r_map, mg.R * (mg.R_PL * 71492000), n_h1, v * vs * 1000, LYA['w0'],
                                             LYA['f'], LYA['a'], wl, float(T), LYA['m'])
    o.update(f_r=f_r, mu_bar=mu, spectrum=spec)
    return o


def lya(pool):
    """tests/golden/anchors_lya.npz: 5 x 5 models at tight tolerances (values) and the status map of the 32 x 32
    grid of BASELINE configs[1] at the reference's defaults."""
    trip = [(T, md, 0.9, True) for T in np.linspace(9000, 13000, 5) for md in np.logspace(9.5, 12, 5)]
    res = pool.map(_run_lya, trip, chunksize=1)
    d = {k: np.array([o[k] for o in res]) for k in ('status', 'f_r', 'mu_bar', 'spectrum')}
    d['T0'] = np.array([t[0] for t in trip]); d['mdot'] = np.array([t[1] for t in trip])
    TT, MM = np.meshgrid(np.linspace(10000, 13000, 32), np.logspace(9.5, 12, 32), indexing='ij')
    grid = pool.map(_run_lya, [(t, m, 0.9, False) for t, m in zip(TT.ravel(), MM.ravel())], chunksize=8)
    d['grid_T0'] = TT.ravel(); d['grid_mdot'] = MM.ravel()
    d['grid_status'] = np.array([o['status'] for o in grid], dtype=np.int8)
    d['grid_spectrum_min'] = np.array([np.nanmin(o['spectrum']) if o['status'] == 0 else np.nan for o in grid])
    np.savez_compressed(os.path.join(OUT, 'anchors_lya.npz'), **d)
    print('anchors_lya: tight status', np.bincount(d['status']), 'grid status', np.bincount(d['grid_status']),
          'deepest core %.3f' % np.nanmin(d['grid_spectrum_min']))


def classes(pool):
    parts = []
    for name, fn in (('bench_grid_status.npz', scan), ('ridge_status.npz', scan_ridge)):
        path = os.path.join(BUILD, name)
        if os.path.exists(path):
            z = np.load(path)
            parts.append((z['T0'], z['mdot'], z['status']))
        else:
            parts.append(fn(pool))
    n_bench = len(parts[0][0])
    T, M, st = (np.concatenate([q[k] for q in parts]) for k in range(3))
    rng = np.random.default_rng(0)
    pick = {}
    for c in (2, 3):
        idx = np.flatnonzero(st == c)
        on_grid = idx[idx < n_bench]          # every failing model of the bench grid itself
        off = idx[idx >= n_bench]
        extra = off if len(off) <= 32 else np.sort(rng.choice(off, 32, replace=False))
        pick[c] = np.concatenate([on_grid, extra])
    # neighbours that succeed, so that the class boundary is in the fixture from both sides
    ok = np.flatnonzero(st == 0)
    near = [i for i in ok if any(abs(i - j) == 1 for c in (2, 3) for j in pick[c])]
    sel = np.concatenate([pick[2], pick[3], np.array(near[::max(1, len(near) // 40)][:40], dtype=int)])
    runs = 3
    res = pool.map(_run, [(T[i], M[i], 0.9, False, -1) for i in sel for _ in range(runs)], chunksize=2)
    status = np.array([[res[a * runs + k]['status'] for k in range(runs)] for a in range(len(sel))])
    assert (status == status[:, :1]).all(), 'the status class itself must be reproducible'
    # is the class stable under 1-ulp noise on the SED?  (a model on a class boundary is not)
    pert = pool.map(_run, [(T[i], M[i], 0.9, False, 2000 + k) for i in sel for k in range(K_NOISE)], chunksize=2)
    status_pert = np.array([[pert[a * K_NOISE + k]['status'] for k in range(K_NOISE)] for a in range(len(sel))])
    spread = np.zeros(len(sel))
    f1 = np.array([[res[a * runs + k]['f_1'] for k in range(runs)] for a in range(len(sel))])
    for a in range(len(sel)):
        if status[a, 0] in (0, 3):
            with np.errstate(all='ignore'):
                spread[a] = max(float(np.nanmax(np.abs(f1[a, k] / f1[a, 0] - 1))) for k in range(1, runs))
    np.savez_compressed(os.path.join(OUT, 'he_failure_classes.npz'), T0=T[sel], mdot=M[sel],
                        h_fraction=np.full(len(sel), 0.9), status=status[:, 0], f_1_runs=f1,
                        f_1_run_spread=spread, on_bench_grid=(sel < n_bench),
                        status_perturbed=status_pert)
    print('%d models change class under 1-ulp noise on the SED' % int((status_pert != status[:, :1]).any(axis=1).sum()))
    for c in (0, 2, 3):
        m = status[:, 0] == c
        print('class %d: %d models, run-to-run spread of f_1: max %.2e, %d of them not reproducible'
              % (c, m.sum(), spread[m].max() if m.any() else 0.0, int((spread[m] > 0).sum())))


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    with mp.get_context('fork').Pool(len(os.sched_getaffinity(0)), initializer=_init) as pool:
        if what in ('scan', 'all'):
            scan(pool)
        if what in ('ridge', 'all'):
            scan_ridge(pool)
        if what in ('p3', 'all'):
            p3(pool)
        if what in ('classes', 'all'):
            classes(pool)
        if what in ('tight', 'all'):
            tight(pool)
        if what in ('lya', 'all'):
            lya(pool)
        if what in ('tidal', 'all'):
            scan_tidal(pool)
