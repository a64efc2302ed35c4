"""Developer script: per-model start / end times of k_helium on one GPU's share of the 256x256x3 tidal grid,
by schedule segment (timing build, see he_timing.py)."""
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
T0 = np.linspace(9000, 13000, 256); Md = np.logspace(9.5, 12, 256)
HH, TT, MM = np.meshgrid(np.array([0.85, 0.9, 0.95]), T0, Md, indexing='ij')
T, md, hf = TT.ravel()[::8].copy(), MM.ravel()[::8].copy(), HH.ravel()[::8].copy()
ho = Engine.h_opts(1.39, 0.73, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True, exact_phi=True, structure_tidal=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)
for sched in sys.argv[1:] or ['']:
    os.environ.pop('PWB_HE_SCHED', None)
    if sched:
        os.environ['PWB_HE_SCHED'] = sched
    for rep in range(2):
        he = eng.he_population_batch(*args, status=h['status'].clone())
        torch.cuda.synchronize()
    s = he['scal'].cpu().numpy()
    t0, t1, nfe = s[:, 0], s[:, 3], s[:, 1]
    ok = nfe > 0
    base = t0[ok].min()
    end = (t1 - base) / 1e6
    dur = (t1 - t0) / 1e6
    print('schedule %s: kernel span %.1f ms, %d models, sum nfe %.1fM' % (sched or 'default', end[ok].max(), ok.sum(), nfe.sum() / 1e6))
    # utilisation over time: number of models running
    grid = np.linspace(0, end[ok].max(), 11)
    run = [int(((t0[ok] - base) / 1e6 <= x).sum() - (end[ok] <= x).sum()) for x in grid]
    print('  models running at t =', ', '.join('%.0f:%d' % (x, n) for x, n in zip(grid, run)))
    q = np.argsort(-nfe)
    for lo, hi in ((0, 0.02), (0.02, 0.1), (0.1, 0.3), (0.3, 0.6), (0.6, 1.0)):
        idx = q[int(lo * len(q)):int(hi * len(q))]
        idx = idx[nfe[idx] > 0]
        print('  nfe rank %3.0f-%3.0f%%: nfe median %.0f max %.0f | end median %.1f max %.1f ms | us per RHS median %.2f'
              % (100 * lo, 100 * hi, np.median(nfe[idx]), nfe[idx].max(), np.median(end[idx]), end[idx].max(), np.median(dur[idx] * 1e3 / nfe[idx])))
