"""Developer script: sweep PWB_HE_SCHED patterns for the helium stage of an N x N grid."""
import os, sys, itertools
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0)
eng.set_sed(g['wavelength'], g['flux_lambda'])
sc = eng.sed_scalars()
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
mu0 = torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)
def run(sched):
    if sched: os.environ['PWB_HE_SCHED'] = sched
    else: os.environ.pop('PWB_HE_SCHED', None)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.he_population_batch(*args, status=h['status'].clone())
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = []
for sched in sys.argv[2:]:
    res.append((run(sched), sched))
    print('%8.2f ms  %s' % res[-1], flush=True)
print('best:', min(res))
