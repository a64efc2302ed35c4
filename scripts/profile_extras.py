"""Developer script for the ncu record (profiles/): exercises the kernels the headline step does not --
'formal' broadening (k_rt_formal_tau), the metals (k_he_ion, k_oxygen, k_carbon), the pixel sum at the
configs[4] sizes (512 x 512 cells, 4096 bins) and the FP64 peak probe."""
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
r_si = r * 1.39 * 71492000
eng.set_geometry(geo['i0'], geo['r_map'], r_si)
N = 16
TT, MM = np.meshgrid(np.linspace(10000, 13000, N), np.logspace(9.5, 12, N), indexing='ij')
T, md, hf = eng.dev(TT.ravel()), eng.dev(MM.ravel()), eng.dev(np.full(N * N, 0.9))
mu0 = torch.full_like(T, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
h = eng.h_ion_fraction_batch(T, md, hf, mu0, r, ho)
vs, rs, rhos = (h['scal'][:, k].contiguous() for k in (1, 2, 3))
he = eng.he_population_batch(T, hf, vs, rs, rhos, r, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
f_he_ion = 1 - he['f_1'] - he['f_3']
# metals (reference defaults: Radau for He+ / O, Radau as in the docs for C)
hio = Engine.o_opts(1.39, [sc['hei_phi'], sc['hei_a_he'], sc['hei_a_h'], 0.0], relax_solution=True)
eng.he_ion_fraction_batch(T, hf, vs, rs, rhos, r, h['v'], h['rho'], h['f_r'], hio, status=h['status'].clone())
oo = Engine.o_opts(1.39, [sc[k] for k in ('o_phi', 'o_a', 'o_a_h', 'o_a_he')], relax_solution=True)
eng.o_ion_fraction_batch(T, hf, vs, rs, rhos, r, h['v'], h['rho'], h['f_r'], f_he_ion, oo, status=he['status'].clone())
co = Engine.c_opts(1.39, [sc[k] for k in ('c_phi_1', 'c_phi_2', 'c_a_1', 'c_a_2', 'c_a_h_1', 'c_a_h_2', 'c_a_he')],
                   relax_solution=True, method='Radau')
eng.c_ion_fraction_batch(T, hf, vs, rs, rhos, r, h['v'], h['rho'], h['f_r'], f_he_ion, co, status=he['status'].clone())
# 'formal' broadening at the quickstart sizes
w = lines.he_3_properties()
hh = hf[:, None]
n_he = h['rho'] * rhos[:, None] * (1 - hh) / (hh + 4 * (1 - hh)) / 1.67262192369e-24 * he['f_3'] * 1e6
v_si = h['v'] * vs[:, None] * 1000
wl = np.linspace(1.0827, 1.0832, 200) * 1e-6
rto_f = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27, wind_broadening_method='formal')
eng.rt2d_batch(n_he[:64], v_si[:64], T[:64], wl, rto_f, status=he['status'][:64].contiguous())
# configs[4] pixel sum: 512 x 512 cells, 4096 bins
eng2 = Engine(0)
eng2.set_sed(g['wavelength'], g['flux_lambda'])
i0, depth, rm = eng2.draw_transit(0.12086, 1.39 * 71492000, 0.499, 0.0, grid_size=512, supersampling=2,
                                  limb_darkening_law='quadratic', ld_coefficient=(0.3, 0.2))
eng2.set_geometry(i0, rm, r_si)
rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27)
eng2.rt2d_batch(n_he, v_si, T, np.linspace(1.0827, 1.0832, 4096) * 1e-6, rto, status=he['status'])
print('fp64 peak %.2f TFLOP/s' % eng.fp64_peak_tflops())
torch.cuda.synchronize()
print('done')
