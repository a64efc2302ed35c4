"""Developer script: does a warp run faster when its lanes hold parameter-space neighbours?
Timing build needed (see he_timing.py).  Compares, for L models per warp, the cost-sorted
schedule (PWB_HE_SCHED=1.0:L) with the caller's order (PWB_HE_LANES=L: consecutive models =
adjacent Mdot at the same T0), in microseconds per right-hand side and model."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['PWB_LIB_PATH'] = os.path.join(ROOT, 'build', 'lib_timing.so')
from p_winds_b200.engine import Engine
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0); eng.set_sed(g['wavelength'], g['flux_lambda']); sc = eng.sed_scalars()
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
N = 64
T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True); heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
mu0 = torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)


def run(env):
    for k in ('PWB_HE_SCHED', 'PWB_HE_LANES'):
        os.environ.pop(k, None)
    os.environ.update(env)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        he = eng.he_population_batch(*args, status=h['status'].clone())
        e1.record()
        torch.cuda.synchronize()
    s = he['scal'].cpu().numpy()
    dur, nfe = (s[:, 3] - s[:, 0]) * 1e-3, s[:, 1]
    ok = nfe > 0
    us = dur[ok] / nfe[ok]
    hi = nfe > np.percentile(nfe[ok], 80)
    print('%-28s kernel %.1f ms | us per RHS: median %.2f, costly fifth %.2f | sum(dur) %.0f ms'
          % (env, e0.elapsed_time(e1), np.median(us), np.median(dur[hi] / nfe[hi]), dur[ok].sum() * 1e-3), flush=True)


for L in (2, 4, 8, 16, 32):
    run({'PWB_HE_SCHED': '1.0:%d' % L})
    run({'PWB_HE_LANES': '%d' % L})
