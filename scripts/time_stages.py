"""Per-stage device timing of the 64x64 He-triplet grid (developer script)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
from p_winds_b200 import lines
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz')); geo = np.load(os.path.join(ROOT, 'tests/golden/geometry.npz'))
r = np.logspace(0, np.log10(20), 100); wl = np.linspace(1.0827, 1.0832, 200) * 1E-6
eng = Engine(0)
eng.set_sed(g['wavelength'], g['flux_lambda'])
sc = eng.sed_scalars()
eng.set_geometry(geo['i0'], geo['r_map'], r * 1.39 * 71492000)
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
w = lines.he_3_properties()
rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27)
T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
if os.environ.get('SORTED'):
    o = np.argsort(-md * np.sqrt(T), kind='stable'); T = T[o]; md = md[o]
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
mu0 = torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9))
for rep in range(int(os.environ.get('REPS', '3'))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    ev[0].record()
    h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
    ev[1].record()
    he = eng.he_population_batch(Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
    nhe = h['rho'] * h['scal'][:, 3:4] * 0.1 / (0.9 + 0.4) / 1.67262192369e-24 * he['f_3'] * 1e6
    vsi = h['v'] * h['scal'][:, 1:2] * 1000
    ev[2].record()
    sp = eng.rt2d_batch(nhe, vsi, Td, wl, rto, status=he['status'])
    ev[3].record(); torch.cuda.synchronize()
    tot = ev[0].elapsed_time(ev[3])
    print('N=%d lanes=%s rep %d total %.1f ms -> %.0f models/s | H %.1f He %.1f RT %.1f ms' % (N, os.environ.get('PWB_HE_LANES', 'auto'), rep, tot, T.size / tot * 1e3, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
nst = he['scal'][:, 2].cpu().numpy(); nfe = he['scal'][:, 1].cpu().numpy(); st = he['status'].cpu().numpy()
print('  He nst mean %.0f max %.0f p99 %.0f | nfe mean %.0f max %.0f | status counts' % (nst.mean(), nst.max(), np.percentile(nst, 99), nfe.mean(), nfe.max()), np.bincount(st, minlength=5).tolist(), 'relax max', he['scal'][:, 0].max().item())
if os.environ.get('DUMP'):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    np.savez(os.path.join(ROOT, 'gpurun_out', os.environ['DUMP']), T=T, md=md, nst=nst, nfe=nfe, status=st,
             n_relax=he['scal'][:, 0].cpu().numpy(), nje=he['scal'][:, 3].cpu().numpy())
