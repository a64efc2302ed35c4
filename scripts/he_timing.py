"""Developer script: per-model start / end times of k_helium by schedule segment.

Needs the timing build of the library:
    PW_DEFS="-DPW_HE_TIMING" PW_OUT=build/lib_timing.so ./build.sh
(the per-model slots n_relax and the spare one then carry globaltimer values)."""
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
for rep in range(2):
    he = eng.he_population_batch(*args, status=h['status'].clone())
    torch.cuda.synchronize()
sc4 = he['scal'].cpu().numpy()
t0, t1, nfe = sc4[:, 0], sc4[:, 3], sc4[:, 1]
base = t0.min()
dr = np.diff(r); dr = np.concatenate((dr, dr[-1:]))
key = (h['scal'][:, 3] * (h['rho'] * eng.dev(dr)[None, :]).sum(1)).cpu().numpy()   # column-density proxy of k_he_key
order = np.argsort(-key, kind='stable')
B = T.size
frac = [0.05, 0.25, 0.40, 0.30]; lanes = [1, 2, 4, 8]
pos = 0
print('kernel span %.1f ms' % ((t1.max() - base) / 1e6))
for f, L in zip(frac, lanes):
    n = int(round(f * B)) if L != 8 else B - pos
    idx = order[pos:pos + n]; pos += n
    st, en = (t0[idx] - base) / 1e6, (t1[idx] - base) / 1e6
    dur = en - st
    print('L=%d models %4d  start %.1f..%.1f  end median %.1f max %.1f ms | nfe median %.0f max %.0f | us per RHS median %.2f p90 %.2f' % (
        L, n, st.min(), st.max(), np.median(en), en.max(), np.median(nfe[idx]), nfe[idx].max(), np.median(dur * 1e3 / nfe[idx]), np.percentile(dur * 1e3 / nfe[idx], 90)))
