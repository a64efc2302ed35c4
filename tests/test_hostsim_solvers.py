"""Parity tier P1: the engine's `__host__ __device__` numerical
core, compiled for the host, stepped against scipy itself.

* RK45 replica: identical nfev and bit-level-close output vs scipy.solve_ivp.
* LSODA replica: identical step / f-evaluation / Jacobian counters, method-switch
  time and last step size vs scipy.odeint(full_output=True) on problems whose
  right-hand side is pure arithmetic (so both sides see identical doubles).
"""
import ctypes as C

import numpy as np
import pytest
from scipy.integrate import odeint, solve_ivp

from conftest import P


def _vdp(mu):
    return lambda t, y: [y[1], mu * ((1.0 - y[0] * y[0]) * y[1]) - y[0]]


@pytest.mark.parametrize('mu,tend,rtol,atol', [(1.0, 20., 1e-3, 1e-6), (5.0, 30., 1e-3, 1e-6),
                                               (1.0, 20., 1e-8, 1e-10), (50., 100., 1e-3, 1e-6)])
def test_rk45_replica_trace(hostsim, mu, tend, rtol, atol):
    hostsim.hs_rk45_vdp.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                    C.c_void_p, C.c_void_p]
    t = np.linspace(0, tend, 60)
    y0 = np.array([2.0, 0.0])
    s = solve_ivp(_vdp(mu), (t[0], t[-1]), y0, t_eval=t, rtol=rtol, atol=atol)
    yo = np.zeros((60, 2))
    cnt = np.zeros(3, dtype=np.int32)
    got = hostsim.hs_rk45_vdp(mu, P(y0), P(t), 60, rtol, atol, P(yo), P(cnt))
    assert got == len(s.t) == 60
    assert cnt[0] == s.nfev                       # same accept/reject sequence
    # nfev identical; values differ only by summation order inside numpy.dot (BLAS), which
    # the rtol=1e-3 step controller amplifies to <= ~1e-8
    assert np.max(np.abs(yo - s.y.T)) < 1e-6


@pytest.mark.parametrize('mu,tend,rtol,atol,exact', [(1.0, 20., 1.49012e-8, 1.49012e-8, True),
                                                     (5.0, 30., 1e-10, 1e-13, True),
                                                     (50., 100., 1.49012e-8, 1.49012e-8, False)])
def test_lsoda_replica_trace(hostsim, mu, tend, rtol, atol, exact):
    hostsim.hs_lsoda_vdp.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_void_p, C.c_void_p]
    t = np.linspace(0, tend, 60)
    y0 = np.array([2.0, 0.0])
    sol, info = odeint(_vdp(mu), y0, t, tfirst=True, full_output=True, rtol=rtol, atol=atol)
    yout = np.zeros((60, 2))
    stats = np.zeros((60, 8))
    assert hostsim.hs_lsoda_vdp(mu, P(y0), P(t), 60, rtol, atol, 500, P(yout), P(stats)) == 2
    for col, key in ((0, 'nst'), (1, 'nfe'), (2, 'nje'), (3, 'nqu'), (4, 'mused')):
        assert np.array_equal(stats[1:, col], info[key]), key
    if exact:   # Adams phase / few Jacobians: bit-identical trajectory
        assert np.array_equal(stats[1:, 5], info['hu'])
        assert np.array_equal(stats[1:, 6], info['tcur'])
        assert np.array_equal(stats[1:, 7], info['tsw'])
        assert np.max(np.abs(yout - sol)) <= 4e-16 * np.max(np.abs(sol))
    else:       # stiff phase: last-bit differences in the 2x2 linear algebra only
        assert np.allclose(stats[1:, 5], info['hu'], rtol=1e-5)
        assert np.max(np.abs(yout - sol)) < 1e-10


def test_lsoda_excess_work_is_reported(hostsim):
    """mxstep = 500 per output interval (odeint default) -> istate -1, which the
    reference turns into RuntimeError (helium.py:691-697)."""
    hostsim.hs_lsoda_vdp.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_void_p, C.c_void_p]
    t = np.array([0.0, 3000.0])
    y0 = np.array([2.0, 0.0])
    yout = np.zeros((2, 2))
    stats = np.zeros((2, 8))
    assert hostsim.hs_lsoda_vdp(1000., P(y0), P(t), 2, 1e-8, 1e-10, 500, P(yout), P(stats)) == -1
    assert stats[1, 0] == 500


def test_lsoda_coefficient_tables(hostsim):
    """cfode: BDF l-polynomials and the Adams error constants are textbook values."""
    elco = np.zeros((2, 13, 14))
    tesco = np.zeros((2, 13, 4))
    hostsim.hs_lsoda_tables(P(elco), P(tesco))
    # BDF order 1..3 leading coefficients el(1) = 1, 2/3, 6/11
    assert np.allclose(elco[1, 1:4, 1], [1.0, 2 / 3, 6 / 11], rtol=1e-15)
    assert np.allclose(elco[1, 2, 1:4], [2 / 3, 1.0, 1 / 3], rtol=1e-15)
    # Adams-Moulton order 2 (trapezoid): el = [1/2, 1, 1/2]
    assert np.allclose(elco[0, 2, 1:4], [0.5, 1.0, 0.5], rtol=1e-15)
    assert tesco[0, 1, 2] == 2.0 and tesco[1, 1, 2] == 2.0


def test_step_selection_root(hostsim):
    """pw_root_newton (the device's a**(1/k): float seed + two Newton steps) against pow,
    over the range LSODA's step / order selection feeds it (error ratios, k = 2..13)."""
    rng = np.random.default_rng(3)
    a = np.concatenate([10 ** rng.uniform(-29, 29, 20000), 10 ** rng.uniform(-12, 2, 20000),
                        [1.0, 1e-30, 1e30, 0.0, 1e-200, 3e200]])
    k = rng.integers(1, 14, a.size).astype(np.int32)
    out = np.zeros_like(a)
    hostsim.hs_root_newton.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    hostsim.hs_root_newton(P(a), P(k), a.size, P(out))
    ref = a ** (1.0 / k)
    m = ref > 0
    assert np.array_equal(out[~m], ref[~m])
    newton = m & (a > 1e-30) & (a < 1e30)          # the range the Newton path serves
    # (pow(a, 1.0 / k) itself carries |ln a| * eps from the rounded exponent)
    assert np.max(np.abs(out[newton] / ref[newton] - 1)) < 2e-15
    assert np.max(np.abs(out[m] / ref[m] - 1)) < 1e-13   # exp(log(a)/k) outside it


@pytest.mark.parametrize('mu,tend,rtol,atol', [(1.0, 20., 1e-3, 1e-6), (50.0, 100., 1e-3, 1e-6),
                                               (1000.0, 1500., 1e-3, 1e-6), (5.0, 30., 1e-8, 1e-10)])
def test_radau_replica_trace_two_equations(hostsim, mu, tend, rtol, atol):
    """scipy's Radau (radau.py) restated: identical nfev / njev / nlu / step count on a stiff
    problem whose right-hand side is pure arithmetic; values to the rounding the step
    controller amplifies."""
    hostsim.hs_radau_vdp.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p]
    t = np.linspace(0, tend, 80)
    y0 = np.array([2.0, 0.0])
    s = solve_ivp(_vdp(mu), (t[0], t[-1]), y0, t_eval=t, method='Radau', rtol=rtol, atol=atol)
    yo = np.zeros((80, 2))
    cnt = np.zeros(5, dtype=np.int32)
    got = hostsim.hs_radau_vdp(mu, P(y0), P(t), 80, rtol, atol, P(yo), P(cnt))
    assert got == len(s.t) == 80
    assert (cnt[0], cnt[1], cnt[2]) == (s.nfev, s.njev, s.nlu)
    assert np.max(np.abs(yo - s.y.T)) < 1e-6 * max(1.0, np.max(np.abs(s.y)))


@pytest.mark.parametrize('lam,rtol,atol', [(1e3, 1e-3, 1e-6), (1e6, 1e-3, 1e-6), (50.0, 1e-6, 1e-9)])
def test_radau_replica_trace_one_equation(hostsim, lam, rtol, atol):
    hostsim.hs_radau_cos.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p]
    hostsim.hs_radau_cos.restype = C.c_int
    t = np.linspace(0, 10.0, 50)
    s = solve_ivp(lambda tt, y: [-lam * (y[0] - np.cos(tt)) - np.sin(tt)], (0, 10.0), [0.0], t_eval=t,
                  method='Radau', rtol=rtol, atol=atol)
    yo = np.zeros(50)
    cnt = np.zeros(5, dtype=np.int32)
    got = hostsim.hs_radau_cos(lam, 0.0, P(t), 50, rtol, atol, P(yo), P(cnt))
    assert got == len(s.t) == 50
    assert (cnt[0], cnt[1], cnt[2]) == (s.nfev, s.njev, s.nlu)
    assert np.max(np.abs(yo - s.y[0])) < 1e-6
