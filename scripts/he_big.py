"""Developer script: helium stage on one GPU's share of the 256x256x3 tidal grid (24 576 models) and on the
64x64 grid, for library variants (PS_LIB=build/<lib>) and warp capacities (PWB_HE_CAP)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get('PS_LIB'):
    os.environ['PWB_LIB_PATH'] = os.path.join(ROOT, 'build', os.environ['PS_LIB'])
from p_winds_b200.engine import Engine
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz'))
r = np.logspace(0, np.log10(20), 100)
eng = Engine(0); eng.set_sed(g['wavelength'], g['flux_lambda']); sc = eng.sed_scalars()
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)


def prep(big):
    if big:
        T0 = np.linspace(9000, 13000, 256); Md = np.logspace(9.5, 12, 256)
        HH, TT, MM = np.meshgrid(np.array([0.85, 0.9, 0.95]), T0, Md, indexing='ij')
        T, md, hf = TT.ravel()[::8].copy(), MM.ravel()[::8].copy(), HH.ravel()[::8].copy()
        ho = Engine.h_opts(1.39, 0.73, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True, exact_phi=True, structure_tidal=True)
    else:
        T0 = np.linspace(10000, 13000, 64); Md = np.logspace(9.5, 12, 64)
        TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
        ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
    Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
    mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
    h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
    return (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo), h['status']


for big in ((False,) if os.environ.get("SMALL_ONLY") else (False, True)):
    args, st = prep(big)
    for cap in os.environ.get('CAPS', '8').split(','):
        os.environ.pop('PWB_HE_SCHED', None); os.environ.pop('PWB_HE_CAP', None)
        if ':' in cap:
            os.environ['PWB_HE_SCHED'] = cap.replace(';', ',')
        else:
            os.environ['PWB_HE_CAP'] = cap
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            eng.he_population_batch(*args, status=st.clone())
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('%s  %s  cap %s : %.2f ms' % (os.environ.get('PS_LIB', 'default'), 'big (24576)' if big else '64x64', cap, best), flush=True)
