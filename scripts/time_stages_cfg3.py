"""Developer script: per-stage device time of the whole 256x256x3 tidal grid (BASELINE configs[3]) on one GPU,
or of a 1/8 share of it (argument `share`)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
from p_winds_b200 import lines
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz')); geo = np.load(os.path.join(ROOT, 'tests/golden/geometry.npz'))
r = np.logspace(0, np.log10(20), 100); wl = np.linspace(1.0827, 1.0832, 200) * 1E-6
eng = Engine(0)
eng.set_sed(g['wavelength'], g['flux_lambda'])
sc = eng.sed_scalars()
eng.set_geometry(geo['i0'], geo['r_map'], r * 1.39 * 71492000)
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
w = lines.he_3_properties()
rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27)
T0 = np.linspace(9000, 13000, 256); Md = np.logspace(9.5, 12, 256)
HH, TT, MM = np.meshgrid(np.array([0.85, 0.9, 0.95]), T0, Md, indexing='ij')
T, md, hf = TT.ravel(), MM.ravel(), HH.ravel()
if len(sys.argv) > 1 and sys.argv[1] == 'share':
    T, md, hf = T[::8].copy(), md[::8].copy(), hf[::8].copy()
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
ho = Engine.h_opts(1.39, 0.73, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True, exact_phi=True, structure_tidal=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
    ev[1].record()
    he = eng.he_population_batch(Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
    hh = hfd[:, None]
    nhe = h['rho'] * h['scal'][:, 3:4] * (1 - hh) / (hh + 4 * (1 - hh)) / 1.67262192369e-24 * he['f_3'] * 1e6
    vsi = h['v'] * h['scal'][:, 1:2] * 1000
    ev[2].record()
    sp = eng.rt2d_batch(nhe, vsi, Td, wl, rto, status=he['status'])
    ev[3].record(); torch.cuda.synchronize()
    print('%d models rep %d: H %.1f  He %.1f  RT %.1f ms' % (T.size, rep, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), flush=True)
