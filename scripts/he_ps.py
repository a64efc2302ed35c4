"""Developer script: thread-per-model helium kernel (k_helium) against the warp-scheduled experiment
(k_helium_ps, PWB_HE_PS=1): bit-identical outputs, time per schedule.  Needs the developer builds

    PW_DEFS="-DPW_HE_PS_KERNEL" PW_OUT=build/lib_ps.so ./build.sh
    PW_DEFS="-DPW_HE_PS_KERNEL -DPW_PS_FREE" PW_OUT=build/lib_ps_free.so ./build.sh     # control: no vote

and PS_LIB=lib_ps.so (or lib_ps_free.so) in the environment.  Results of round 2: profiles/r2_he_ps.log."""
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
big = len(sys.argv) > 1 and sys.argv[1] == 'big'
if big:
    T0 = np.linspace(9000, 13000, 256); Md = np.logspace(9.5, 12, 256)
    HH, TT, MM = np.meshgrid(np.array([0.85, 0.9, 0.95]), T0, Md, indexing='ij')
    T, md, hf = TT.ravel()[::8].copy(), MM.ravel()[::8].copy(), HH.ravel()[::8].copy()
    ho = Engine.h_opts(1.39, 0.73, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True, exact_phi=True, structure_tidal=True)
else:
    N = 64
    T0 = np.linspace(10000, 13000, N); Md = np.logspace(9.5, 12, N)
    TT, MM = np.meshgrid(T0, Md, indexing='ij'); T = TT.ravel(); md = MM.ravel(); hf = np.full(T.size, 0.9)
    ho = Engine.h_opts(1.39, 0.73, relax_solution=True, exact_phi=True)
Td, mdd, hfd = eng.dev(T), eng.dev(md), eng.dev(hf)
heo = Engine.he_opts(1.39, rates, (1.0, 0.0), True)
mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, r, ho)
args = (Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(), h['scal'][:, 3].contiguous(), r, h['v'], h['rho'], h['f_r'], heo)


def run(env):
    for k in ('PWB_HE_SCHED', 'PWB_HE_LANES', 'PWB_HE_PS'):
        os.environ.pop(k, None)
    os.environ.update(env)
    best, he = 1e9, None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        he = eng.he_population_batch(*args, status=h['status'].clone())
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, he


t0, ref = run({})
print('%d models; k_helium default schedule %.2f ms' % (T.size, t0), flush=True)
scheds = ['', '1.0:1', '1.0:2', '1.0:4', '1.0:8', '1.0:16', '1.0:32']
for sch in scheds:
    env = {'PWB_HE_PS': '1'}
    if sch:
        env['PWB_HE_SCHED'] = sch
    t1, he = run(env)
    same = all(torch.equal(torch.nan_to_num(he[k], nan=-1.0), torch.nan_to_num(ref[k], nan=-1.0)) for k in ('f_1', 'f_3', 'scal', 'status'))
    env.pop('PWB_HE_PS')
    t2, _ = run(env)
    print('  sched %-34s  k_helium %.2f ms   k_helium_ps %.2f ms   identical: %s' % (sch or 'default', t2, t1, same), flush=True)
