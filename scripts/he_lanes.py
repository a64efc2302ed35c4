"""Developer script: per-RHS latency of k_helium against models per warp (L) and warps per SM,
on cost-sorted samples of the 64x64 grid (needs the timing build, see he_timing.py)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['PWB_LIB_PATH'] = os.path.join(ROOT, 'build', os.environ.get('PWB_TIMING_LIB', 'lib_timing.so'))
from p_winds_b200.engine import Engine
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0); eng.set_sed(g['wavelength'], g['flux_lambda']); sc = eng.sed_scalars()
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
N = 64
T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel()
o = np.argsort(-md * np.sqrt(T), kind='stable')
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True); heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
for warps in (592, 1184):
    for L in (1, 2, 4, 8, 16, 32):
        n = min(4096, warps * L)
        # a contiguous run of the cost-sorted list starting at the 10 % mark: neighbours in a warp
        # have similar cost, as in the production schedule
        start = min(410, 4096 - n)
        idx = o[start:start + n]
        Ts, ms = T[idx], md[idx]; hf = np.full(n, 0.9)
        Td, mdd, hfd = eng.dev(Ts), eng.dev(ms), eng.dev(hf)
        mu0 = torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
        h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
        args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)
        os.environ['PWB_HE_LANES'] = str(L)
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            he = eng.he_population_batch(*args, status=h['status'].clone())
            e1.record(); torch.cuda.synchronize()
        sc4 = he['scal'].cpu().numpy(); t0, t1, nfe = sc4[:, 0], sc4[:, 3], sc4[:, 1]
        ok = nfe > 0
        us = (t1 - t0)[ok] / 1e3 / nfe[ok]
        print('warps %4d L=%2d n=%4d kernel %.1f ms | us/RHS median %.2f p10 %.2f p90 %.2f | nfe median %.0f max %.0f' % (
            (n + L - 1) // L, L, n, e0.elapsed_time(e1), np.median(us), np.percentile(us, 10), np.percentile(us, 90), np.median(nfe), nfe.max()), flush=True)
        if n == 4096: break
