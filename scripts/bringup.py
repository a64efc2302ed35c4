"""GPU bring-up diagnostics (developer script, not part of the test-suite)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import pwinds_oracle as O
from p_winds_b200.engine import Engine

g = np.load(os.path.join(ROOT, 'tests/golden/solar_sed.npz')); geo = np.load(os.path.join(ROOT, 'tests/golden/geometry.npz'))
sed = O.Sed(g['wavelength'], g['flux_lambda'])
r = np.logspace(0, np.log10(20), 100); wl = np.linspace(1.0827, 1.0832, 200) * 1E-6
eng = Engine(0)
print(torch.cuda.get_device_name(0))
eng.set_sed(sed.wl, sed.flux)
sc = eng.sed_scalars(); print('sed scalars', sc)
ref = list(O.h_radiative_processes(sed)) + list(O.he_radiative_processes(sed))
print('sed scalar rel err', [abs(a / b - 1) for a, b in zip(list(sc.values())[:8], ref)])
eng.set_geometry(geo['i0'], geo['r_map'], r * 1.39 * 71492000)
rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
w0, f, a = O.he3_lines()
rto = Engine.rt_opts(w0, f, a, 4 * 1.67262192369e-27)
print('fp64 peak TFLOP/s', eng.fp64_peak_tflops())

def run(T, md, hf, tight, tidal=False):
    kw = dict(rtol=1e-10, atol=1e-13) if tight else {}
    ho = Engine.h_opts(1.39, 0.73, 1.07 if tidal else 1.0, 0.04634 if tidal else 1.0, relax_solution=True, exact_phi=True, structure_tidal=tidal, **kw)
    heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True, **(dict(method='LSODA', rtol=1e-10, atol=1e-16) if tight else {}))
    return eng.forward_batch(T, md, hf, r, ho, heo, rto, wl)

for name in ('anchors_tight', 'anchors_default', 'anchors_tight_tidal', 'anchors_default_tidal'):
    d = np.load(os.path.join(ROOT, 'tests/golden/%s.npz' % name))
    tight = bool(d['tight']); tidal = bool(d['tidal'])
    t = time.time(); out = run(d['T0'], d['mdot'], d['h_fraction'], tight, tidal); torch.cuda.synchronize(); dt = time.time() - t
    st = out['status'].cpu().numpy()
    print(name, 'B', len(st), '%.1f ms' % (dt * 1e3), 'status gpu', st.tolist()); print('   status ref', d['status'].tolist())
    ok = (st == 0) & (d['status'] == 0)
    for key, gk in (('f_r', 'f_r'), ('f_1', 'f_1'), ('f_3', 'f_3'), ('spectrum', 'spectrum'), ('v','v'), ('rho','rho')):
        a_ = out[gk].cpu().numpy()[ok]; b_ = d[key][ok]
        if key == 'f_r': a_, b_ = a_[:, 1:], b_[:, 1:]
        rel = np.abs(a_ / b_ - 1)
        if key == 'f_3':
            ipk = np.argmax(b_, axis=1); relpk = rel[np.arange(len(ipk)), ipk]
            print('   %-9s rel err at peak: median %.2e max %.2e | anywhere max %.2e' % (key, np.median(relpk), relpk.max(), rel.max()))
        else:
            m = rel.max(axis=1); print('   %-9s max rel err per model: median %.2e max %.2e' % (key, np.median(m), m.max()))
    mu = out['hscal'][:, 0].cpu().numpy()[ok]; print('   mu_bar rel', np.max(np.abs(mu / d['mu_bar'][ok] - 1)), 'n_relax', out['hscal'][:, 4].cpu().numpy()[ok].astype(int).tolist(), 'he n_relax', out['hescal'][:, 0].cpu().numpy()[ok].astype(int).tolist())

# throughput: 64x64 grid, default tolerances
for N in (32, 64):
    T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
    TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
    Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
        heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
        ev[0].record()
        h = eng.h_ion_fraction_batch(Td, mdd, hfd, torch.full_like(Td, (1 + 4 * 0.1 / 0.9) / (1 + 0.1 / 0.9)), r, ho)
        ev[1].record()
        he = eng.he_population_batch(Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
        ev[2].record()
        nhe = h['rho'] * h['scal'][:, 3:4] * 0.1 / (0.9 + 0.4) / 1.67262192369e-24 * he['f_3'] * 1e6
        vsi = h['v'] * h['scal'][:, 1:2] * 1000
        ev[3].record()
        sp = eng.rt2d_batch(nhe, vsi, Td, wl, rto, status=he['status'])
        ev[4].record(); torch.cuda.synchronize(); dt = time.time() - t
        print('N=%d rep %d total %.1f ms -> %.0f models/s | H %.1f He %.1f prep %.1f RT %.1f ms | fail H %d He %d' % (N, rep, dt * 1e3, T.size / dt, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]), int((h['status'] != 0).sum()), int((he['status'] == 2).sum())))
    print('  nfev H mean', float(h['scal'][:, 5].mean()), 'He nfe mean', float(he['scal'][:, 1].mean()), 'nst', float(he['scal'][:, 2].mean()))
