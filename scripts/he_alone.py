"""Developer script: k_helium on the costliest n models of the 64x64 grid alone, one model per
warp -- the per-RHS latency of a model when its SM is (nearly) its own."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0); eng.set_sed(g['wavelength'], g['flux_lambda']); sc = eng.sed_scalars()
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
N = 64
T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel()
o = np.argsort(-md * np.sqrt(T), kind='stable')
for n in [int(a) for a in sys.argv[1:]] or [148, 205, 592]:
    Ts, ms = T[o[:n]], md[o[:n]]; hf = np.full(n, 0.9)
    Td, mdd, hfd = eng.dev(Ts), eng.dev(ms), eng.dev(hf)
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True); heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
    mu0 = torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
    h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
    args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)
    os.environ['PWB_HE_LANES'] = '1'
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        he = eng.he_population_batch(*args, status=h['status'].clone())
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    nfe = he['scal'][:, 1].cpu().numpy()
    print('n=%d alone, L=1: %.1f ms | nfe max %.0f median %.0f' % (n, best, nfe.max(), np.median(nfe)), flush=True)
