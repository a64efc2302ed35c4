"""Developer script: a small pass over every kernel of the hot path, for compute-sanitizer
(memcheck / racecheck):  compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
from p_winds_b200 import lines
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz')); geo = np.load(os.path.join(ROOT, 'tests/golden/geometry.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0)
eng.set_sed(g['wavelength'], g['flux_lambda'])
sc = eng.sed_scalars()
eng.set_geometry(geo['i0'], geo['r_map'], r * 1.39 * 71492000)
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
w = lines.he_3_properties()
T = np.array([9100., 11000., 12500., 6000., 10428.571428571428, 10048.0])
md = np.array([10 ** 10.27, 1e11, 10 ** 9.8, 1e11, 10 ** 11.365079365079366, 10 ** 11.444])
hf = np.full(T.size, 0.9)
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
for nw in (200, 150, 700):
    wl = np.linspace(1.0827, 1.0832, nw) * 1E-6
    rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27)
    for env in ('PWB_RT_SPLIT', 'PWB_RT_NOSPLIT'):
        os.environ[env] = '1'
        out = eng.forward_batch(T, md, hf, r, ho, heo, rto, wl)
        del os.environ[env]
    torch.cuda.synchronize()
    print(nw, out['status'].cpu().numpy(), float(out['spectrum'][0].min()))
rto_f = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27, wind_broadening_method='formal')
n = out['rho'][:2] * 1e20
spec = eng.rt2d_batch(torch.nan_to_num(n), torch.nan_to_num(out['v'][:2]) * 1e4, T[:2], np.linspace(1.0827, 1.0832, 40) * 1E-6, rto_f)
chi = eng.chi2_batch(out['spectrum'], np.ones(700), np.full(700, 1e-3), 1.0, out['status'])
torch.cuda.synchronize()
print('done', chi.cpu().numpy())
