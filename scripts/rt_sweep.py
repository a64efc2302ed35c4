"""Developer script: pixel-sum kernel variants (PWB_RT_SPLIT = one CTA per segment; PWB_RT_G / PWB_RT_CHUNK)
on the quickstart sizes (B = 4096, 100 x 100 map, 200 bins) and on the configs[4] sizes
(B = 64, 512 x 512 map, 4096 bins)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from p_winds_b200.engine import Engine
from p_winds_b200 import lines
prof = np.load(os.path.join(ROOT, 'tests/golden/he_3_profile.npz'))
geo = np.load(os.path.join(ROOT, 'tests/golden/geometry.npz'))
w = lines.he_3_properties()
rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, 4 * 1.67262192369e-27, bulk_los_velocity=-2E3)
r_si, n_si, v_si = prof['r'][::5].copy(), prof['n'][::5].copy(), prof['v'][::5].copy()
rng = np.random.default_rng(1)


def case(name, B, setup, nw):
    eng = Engine(0)
    setup(eng)
    n = eng.dev(n_si[None, :] * 10 ** rng.uniform(-1, 1, B)[:, None])
    v = eng.dev(np.tile(v_si, (B, 1)))
    T = eng.dev(rng.uniform(7000, 12000, B))
    wl = eng.dev(np.linspace(1.0827, 1.0832, nw) * 1e-6)
    print(name, 'active pixels', eng.n_active_pixels(), flush=True)
    for env in ([{'PWB_RT_SPLIT': '1'}] + [{'PWB_RT_G': str(g), 'PWB_RT_CHUNK': str(c)} for g in (1, 2, 4) for c in (256, 1024, 2048)]):
        for k in ('PWB_RT_SPLIT', 'PWB_RT_G', 'PWB_RT_CHUNK'):
            os.environ.pop(k, None)
        os.environ.update(env)
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            eng.rt2d_batch(n, v, T, wl, rto)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('  %-44s %8.3f ms' % (env, best), flush=True)


case('quickstart', 4096, lambda e: e.set_geometry(geo['i0'], geo['r_map'], r_si), 200)


def hi(e):
    i0, depth, rm = e.draw_transit(0.12086, 1.39 * 71492000, 0.499, 0.0, grid_size=512, supersampling=2,
                                   limb_darkening_law='quadratic', ld_coefficient=(0.3, 0.2))
    e.set_geometry(i0, rm, r_si)


case('highres', 64, hi, 4096)
