#!/usr/bin/env python
"""Headline benchmark: He 10830 forward models per second (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the largest single-GPU configuration the metric
is quoted on): a 64 x 64 (T0, Mdot) He-triplet retrieval grid per GPU -- HD 209458 b,
solar SED, 100 radii, exact photoionisation, relax loops on, reference solver
tolerances (RK45 1e-3/1e-6, odeint 1.49e-8), G = 100 transit map (supersampling 10),
200 wavelength bins, 3 lines.  One step = one pass of the whole forward-model path
over the rank's 4096 models + chi^2.  With N GPUs the Mdot axis is sampled N times finer
(64 x 64N models, h_fraction = 0.90 throughout) and the flattened grid is interleaved over
the ranks, so every rank integrates a 64 x 64 grid over the same parameter box -- the
per-rank work is statistically identical at every N (weak scaling); the only exchange is
the final all_gather of chi^2 over NCCL.  The line also carries, measured on all ranks at
every N: `cfg3` (BASELINE configs[3], the north-star grid: 256 x 256 x 3 tidal models cut
over the ranks, strong scaling) and `highres_batch` (BASELINE configs[4]: 512 x 512 cells,
4096 bins, He + C II + O I; 256 models cut over the ranks).

The reference arm (`--impl reference`) times the UNMODIFIED reference
(hydrogen.ion_fraction -> parker.structure -> helium.population_fraction ->
transit.radiative_transfer_2d through its stock code path, imported from oracle/_ref --
the byte-for-byte copy made by oracle/build_ref.py -- with the astropy / flatstar stand-ins)
on all host cores on a bounded sample of the same grid; under torchrun only rank 0 runs
it.  If oracle/_ref is absent it falls back to the oracle port and says so (`kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

N_GRID = 64
R_PL, M_PL = 1.39, 0.73
NR, NW, NLINES = 100, 200, 3
M_HE = 4 * 1.67262192369e-27
METRIC = 'He10830 forward models/sec'
UNIT = 'models/s'


def grid(n_gpus):
    """Global (T0, Mdot, h_fraction) arrays, flattened; SURVEY.md 8(d) throughput grid.
    The Mdot axis is the fastest one and has 64 * n_gpus samples: interleaved over the ranks
    (rank g takes g, g + N, ...) every rank gets a 64 x 64 grid whose Mdot samples are shifted
    by a fraction of a cell -- the same work per rank at every N."""
    T0 = np.linspace(10000, 13000, N_GRID)
    md = np.logspace(9.5, 12, N_GRID * n_gpus)
    TT, MM = np.meshgrid(T0, md, indexing='ij')
    return TT.ravel(), MM.ravel(), np.full(TT.size, 0.90)


def grid_cfg3():
    """BASELINE configs[3]: 256 x 256 (T0, Mdot) x 3 h_fractions, tidal Parker profile
    (M* = 1.07, a = 0.04634 au); 196 608 models."""
    T0 = np.linspace(9000, 13000, 256)
    md = np.logspace(9.5, 12, 256)
    hf = np.array([0.85, 0.90, 0.95])
    HH, TT, MM = np.meshgrid(hf, T0, md, indexing='ij')
    return TT.ravel(), MM.ravel(), HH.ravel()


def grid_cfg3_share():
    """One GPU's share (1/8, interleaved) of the configs[3] grid."""
    T, md, hf = grid_cfg3()
    return T[0::8].copy(), md[0::8].copy(), hf[0::8].copy()


def config(n_gpus, workload='grid64'):
    if workload == 'cfg3':
        n = 256 * 256 * 3
        return {'workload': 'He-triplet retrieval grid 256x256 (T0, Mdot) x 3 h_fractions, tidal Parker profile '
                            '(BASELINE configs[3], the north-star target); HD 209458 b, solar SED, Nr=100, G=100 '
                            '(ss=10), Nw=200, 3 lines, exact_phi, relax_solution, reference solver tolerances',
                'models_per_gpu': -(-n // n_gpus), 'global_models': n,
                'sharding': 'interleaved over ranks, final all_gather of chi2+status' if n_gpus > 1 else 'single GPU',
                'l2_policy': 'L2 flushed (512 MiB write) before every timed step'}
    return {'workload': 'He-triplet retrieval grid 64x64 (T0, Mdot) per GPU (BASELINE configs[2]; with N GPUs '
                        '64 x 64N, Mdot axis N times finer, h_fraction 0.90, interleaved over ranks); '
                        'HD 209458 b, solar SED, Nr=100, G=100 (ss=10), Nw=200, 3 lines, exact_phi, '
                        'relax_solution, reference solver tolerances',
            'models_per_gpu': N_GRID * N_GRID, 'global_models': N_GRID * N_GRID * n_gpus,
            'sharding': 'interleaved over ranks, final all_gather of chi2+status' if n_gpus > 1 else 'single GPU',
            'l2_policy': 'L2 flushed (512 MiB write) before every timed step'}


# ----------------------------------------------------------------------------- CPU side
def _cpu_init():
    """Worker initialiser: the unmodified reference when it is there (oracle/_ref on the GPU box,
    /root/reference in the build container), else the oracle port."""
    global _S, _O, _RC, _SP, _GEOM, _KIND
    os.environ['OMP_NUM_THREADS'] = os.environ['OPENBLAS_NUM_THREADS'] = '1'
    import warnings
    warnings.simplefilter('ignore')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import refloader
    if refloader.available() and not os.environ.get('PWB_BENCH_FORCE_PORT'):
        import refchain as RC
        _KIND, _RC = 'reference', RC
        _SP = RC.spectrum_dict()
        _GEOM = RC.reference_geometry()   # transit.draw_transit hoisted out of the loop, as the docs do
        return
    import pwinds_oracle as O
    g = np.load(os.path.join(GOLDEN, 'solar_sed.npz'))
    geo = np.load(os.path.join(GOLDEN, 'geometry.npz'))
    _KIND, _O = 'port', O
    _S = O.Setup(O.Sed(g['wavelength'], g['flux_lambda']), i0=geo['i0'], r_map=geo['r_map'])


def _cpu_model(args):
    import warnings
    T, md, h = args
    t = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if _KIND == 'reference':
            o = _RC.one_model(_SP, _GEOM, float(T), float(md), float(h), False, False)
        else:
            o = _O.forward_model(_S, float(T), float(md), float(h))
    return o['status'], time.perf_counter() - t, _KIND


def cpu_loop(T, md, hf, n_sample, pool, cores):
    """Reference per-model CPU loop on a deterministic subsample (every k-th model)."""
    k = max(1, len(T) // n_sample)
    idx = np.arange(0, len(T), k)[:n_sample]
    t0 = time.perf_counter()
    res = pool.map(_cpu_model, [(T[i], md[i], hf[i]) for i in idx], chunksize=1)
    dt = time.perf_counter() - t0
    fails = sum(1 for s, _, _ in res if s in (1, 2))
    kind = res[0][2]
    what = ('the unmodified reference (p_winds %s from oracle/_ref or /root/reference: hydrogen.ion_fraction -> '
            'parker.structure -> helium.population_fraction -> transit.radiative_transfer_2d, draw_transit '
            'hoisted; astropy / flatstar stand-ins)' % _ref_version()) if kind == 'reference' else \
        'oracle/pwinds_oracle.py (numpy + scipy solve_ivp/odeint; oracle/_ref absent)'
    return {'value': len(idx) / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': 'every %d-th model of the grid (%d models), multiprocessing.Pool(%d), OMP_NUM_THREADS=1; '
                      '%s; %d failed, mean %.3f s/model/core' % (k, len(idx), cores, what, fails,
                                                              float(np.mean([x for _, x, _ in res])))}, dt


def _ref_version():
    try:
        with open(os.path.join(ROOT, 'oracle', '_ref', 'MANIFEST.json')) as fh:
            return json.load(fh).get('version', '')
    except Exception:
        return ''


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    T, md, hf = grid(args.gpus)
    n_sample = max(cores * 8, 64)
    with mp.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(args.warmup):
            cpu_loop(T, md, hf, cores, pool, cores)
        tot_models, tot_t, base = 0, 0.0, None
        for _ in range(args.steps):
            base, dt = cpu_loop(T, md, hf, n_sample, pool, cores)
            tot_models += min(n_sample, len(T))
            tot_t += dt
    value = tot_models / tot_t
    base['value'] = value
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic grid over the reference solar SED', 'config': config(args.gpus),
            'cpu_baseline': base,
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU side
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(self.rows[0][1]),
                'reasons': reasons, 'samples': len(self.rows)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines, grid as pgrid

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    g = np.load(os.path.join(GOLDEN, 'solar_sed.npz'))
    geo = np.load(os.path.join(GOLDEN, 'geometry.npz'))
    r = np.logspace(0, np.log10(20), NR)
    wl = np.linspace(1.0827, 1.0832, NW) * 1E-6
    eng = Engine(local_rank)
    eng.set_sed(g['wavelength'], g['flux_lambda'])
    eng.set_geometry(geo['i0'], geo['r_map'], r * (R_PL * 71492000))
    sc = eng.sed_scalars()
    rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
    cfg3 = (args.workload == 'cfg3')
    if cfg3:   # the whole 256 x 256 x 3 tidal grid, cut over the ranks (total work fixed: strong scaling)
        ho = Engine.h_opts(R_PL, M_PL, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True,
                           exact_phi=True, structure_tidal=True)
        args.no_large_batch = True
        args.no_cpu_baseline = True
    else:
        ho = Engine.h_opts(R_PL, M_PL, relax_solution=True, exact_phi=True)
    heo = Engine.he_opts(R_PL, rates, (1.0, 0.0), True)
    w = lines.he_3_properties()
    rto = Engine.rt_opts(w[:3], w[3:6], [w[6]] * 3, M_HE)
    fp64_peak = eng.fp64_peak_tflops()

    T, md, hf = grid_cfg3() if cfg3 else grid(world)
    n_global = len(T)
    idx = pgrid.shard_indices(n_global, rank, world)
    B = len(idx)
    # host side of the public API: pinned host buffers in, pinned host buffers out
    h_in = torch.from_numpy(np.stack([T[idx], md[idx], hf[idx]])).pin_memory()
    d_in = torch.empty_like(h_in, device=dev)
    out = eng.alloc_forward(B, NR, NW)
    h_spec = torch.empty((B, NW), dtype=torch.float64).pin_memory()
    h_chi2 = torch.empty(n_global, dtype=torch.float64).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    rd, wd = eng.dev(r), eng.dev(wl)
    # synthetic "observation": the model at the grid centre + fixed noise level
    data = torch.ones(NW, dtype=torch.float64, device=dev)
    sigma = torch.full((NW,), 1e-3, dtype=torch.float64, device=dev)
    norm = float(geo['i0'].sum())
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e, events=None):
        if e2e:
            d_in.copy_(h_in, non_blocking=True)
        Td, mdd, hfd = d_in[0], d_in[1], d_in[2]
        if events:
            events[0].record()
        eng.forward_batch(Td, mdd, hfd, rd, ho, heo, rto, wd, out=out)
        chi2 = eng.chi2_batch(out['spectrum'], data, sigma, norm, out['status'])
        if events:
            events[1].record()
        allchi = pgrid.gather_interleaved(chi2, n_global, rank, world)
        if events:
            events[2].record()
        if e2e:
            h_spec.copy_(out['spectrum'], non_blocking=True)
            h_chi2.copy_(allchi, non_blocking=True)
            h_status.copy_(out['status'], non_blocking=True)
        return allchi

    split = {'compute_ms': 0.0, 'gather_ms': 0.0}

    def timed(e2e, steps):
        """K steps timed with CUDA events on the launching stream; L2 flushed before each.  The
        device-timed pass also splits each step into this rank's compute (forward + chi^2) and
        the all_gather (which includes waiting for the slowest rank)."""
        total_ms = 0.0
        for _ in range(steps):
            flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if not e2e else None
            e0.record()
            step(e2e, ev)
            e1.record()
            if e2e:
                torch.cuda.current_stream().synchronize()
            barrier()
            total_ms += e0.elapsed_time(e1)
            if ev:
                split['compute_ms'] += ev[0].elapsed_time(ev[1]) / steps
                split['gather_ms'] += ev[1].elapsed_time(ev[2]) / steps
        return total_ms

    d_in.copy_(h_in)
    for _ in range(max(args.warmup, 3)):
        step(True)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    t_wall = time.perf_counter()
    ms_dev = timed(False, args.steps)
    wall_dev = time.perf_counter() - t_wall
    launches = eng.launch_count() - launches0
    ms_e2e = timed(True, args.steps)
    sampler.stop_flag = True
    sampler.join(2)

    # per-kernel device time of one step (CUDA events around each stage call)
    Td, mdd, hfd = d_in[0], d_in[1], d_in[2]
    mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
    for _rep in range(2):  # first round warms the stage-level entry points (allocations), second is timed
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        flush.fill_(1.0)
        barrier()
        ev[0].record()
        h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, rd, ho)
        ev[1].record()
        he = eng.he_population_batch(Td, hfd, h['scal'][:, 1].contiguous(), h['scal'][:, 2].contiguous(),
                                     h['scal'][:, 3].contiguous(), rd, h['v'], h['rho'], h['f_r'], heo,
                                     status=h['status'].clone())
        ev[2].record()
        nhe = h['rho'] * h['scal'][:, 3:4] * (1 - hfd[:, None]) / (hfd[:, None] + 4 * (1 - hfd[:, None])) \
            / 1.67262192369e-24 * he['f_3'] * 1e6
        vsi = h['v'] * h['scal'][:, 1:2] * 1000
        ev[3].record()
        eng.rt2d_batch(nhe, vsi, Td, wd, rto, status=he['status'])
        ev[4].record()
        barrier()
    t_h, t_he, t_rt = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[3].elapsed_time(ev[4])
    status = out['status'].cpu().numpy()
    nfe_he = float(he['scal'][:, 1].sum())
    nfev_h = float(h['scal'][:, 5].sum())
    n_relax_h = float(h['scal'][:, 4].sum())
    n_ok = int((status == 0).sum())          # models with a defined result
    n_warn = int((status == 3).sum())        # relax-loop odeint warning: spectrum of the last completed pass

    # algorithmic FP64 work, SURVEY.md 8(d) convention (add/mul 1, FMA 2, div/sqrt 8, exp/log 20,
    # linear interp 20, Lambert-W 200, wofz 150); measured counters instead of estimates
    px = int((geo['i0'] != 0).sum())
    w_he = nfe_he * 540.0
    w_h = (n_relax_h + B) * NR * sc['n_cut_h'] * 28.0 + nfev_h * 85.0 + (2 * (n_relax_h + B) + B) * NR * 200.0
    w_rt = (n_ok + n_warn) * (px * (35.0 + NW * 26.0) + 200 * NR * 60.0 + NLINES * NW * 150.0)
    tms = torch.tensor([ms_dev, ms_e2e, t_h, t_he, t_rt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, t_h, t_he, t_rt = tms.tolist()
    value = n_global * args.steps / (ms_dev * 1e-3)
    e2e = n_global * args.steps / (ms_e2e * 1e-3)

    kernels = {
        'k_helium': {'ms': t_he, 'share_of_step': t_he / (t_h + t_he + t_rt), 'flop': w_he,
                     'achieved_tflops': w_he / (t_he * 1e-3) / 1e12,
                     'note': 'latency bound: %.0f dependent RHS evaluations per model (max %.0f), one thread per '
                             'model, cost-ordered schedule' % (nfe_he / B, float(he['scal'][:, 1].max()))},
        'k_h_fields+k_h_solve+k_h_finish': {'ms': t_h, 'share_of_step': t_h / (t_h + t_he + t_rt), 'flop': w_h,
                       'achieved_tflops': w_h / (t_h * 1e-3) / 1e12},
        'k_rt_los+k_rt_sum2': {'ms': t_rt, 'share_of_step': t_rt / (t_h + t_he + t_rt),
                                           'flop': w_rt, 'achieved_tflops': w_rt / (t_rt * 1e-3) / 1e12,
                                           'note': 'flop = the reference algorithm (every lit pixel, 20-flop exp); '
                                                   'the kernel merges equal-radius pixels (%d -> %d terms) and '
                                                   'needs 10 FP64 instructions per term (ncu: FP64 pipe 72 %% busy)'
                                                   % (px, eng.n_active_pixels())},
    }
    dom = max(kernels, key=lambda k: kernels[k]['ms'])
    # DRAM bytes of the dominant kernel per launch, from the committed `ncu --set full` capture
    # (profiles/r2_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum); null if absent
    traffic = None
    try:
        import csv
        with open(os.path.join(ROOT, 'profiles', 'r2_ncu_full_summary.csv')) as fh:
            rows = list(csv.reader(fh))
        hdr = rows[0]
        for row in rows[1:]:
            if 'k_helium' in row[0]:
                def _bytes(cell):
                    val, unit = cell.split()[:2]
                    return float(val) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
                traffic = _bytes(row[hdr.index('dram__bytes_read.sum')]) + _bytes(row[hdr.index('dram__bytes_write.sum')])
                break
    except Exception:
        traffic = None
    roofline = {'bound': 'fp64', 'kernel': dom, 'achieved': kernels[dom]['achieved_tflops'],
                'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': kernels[dom]['achieved_tflops'] / fp64_peak,
                'traffic': traffic,
                'traffic_source': 'profiles/r2_ncu_full_summary.csv (ncu --set full capture of k_helium, bytes per launch); '
                                  'not an HBM-bound kernel' if traffic is not None else None,
                'peak_source': 'FP64 DFMA microbenchmark run in this process (pwb_fp64_peak); '
                               'MEASURED_PEAKS.json has no FP64 figure; no tensor cores on this path',
                'whole_step': {'flop': w_h + w_he + w_rt,
                               'achieved': (w_h + w_he + w_rt) / ((t_h + t_he + t_rt) * 1e-3) / 1e12,
                               'frac': (w_h + w_he + w_rt) / ((t_h + t_he + t_rt) * 1e-3) / 1e12 / fp64_peak},
                'hbm': {'bytes_per_model': 8 * (4 * NR + NW) + 12 + 24,
                        'achieved_gbs': n_global * args.steps * (8 * (4 * NR + NW) + 36) / (ms_dev * 1e-3) / 1e9,
                        'note': 'profiles + spectrum written per model; inputs (SED tables 22 KB, pixel table '
                                '%d KB) stay L2 resident: HBM is not the bound' % (px * 28 // 1024)},
                'kernels': kernels}

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def all_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(q[0]) for q in parts]

    per_rank = {'compute_ms': all_ranks(split['compute_ms']), 'gather_ms': all_ranks(split['gather_ms']),
                'he_ms': all_ranks(ev[1].elapsed_time(ev[2])), 'h_ms': all_ranks(ev[0].elapsed_time(ev[1])),
                'rt_ms': all_ranks(ev[3].elapsed_time(ev[4])),
                'note': 'per rank, mean over the timed steps: compute = forward_batch + chi2 on that rank; gather = '
                        'the NCCL all_gather of chi2 including the wait for the slowest rank'}

    hol = Engine.h_opts(R_PL, M_PL, star_mass=1.07, semimajor_axis=0.04634, relax_solution=True,
                        exact_phi=True, structure_tidal=True)

    def tidal_pass(Tl, ml, hl, n_all, reps):
        """forward + chi2 + gather over this rank's models of a tidal grid; max over ranks, best of reps."""
        outl = eng.alloc_forward(Tl.numel(), NR, NW)
        best = None
        for _ in range(reps + 1):   # first pass warms allocations
            flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.forward_batch(Tl, ml, hl, rd, hol, heo, rto, wd, out=outl)
            c2 = eng.chi2_batch(outl['spectrum'], data, sigma, norm, outl['status'])
            pgrid.gather_interleaved(c2, n_all, rank, world)
            e1.record()
            barrier()
            t = max_over_ranks(e0.elapsed_time(e1))
            best = t if (best is None or _ == 1) else min(best, t)
        stl = outl['status']
        cnt = torch.stack([(stl == k).sum() for k in range(6)]).to(torch.float64)
        if world > 1:
            dist.all_reduce(cnt)
        return best, [int(x) for x in cnt.tolist()]

    # the production-size batch: one GPU's share of the 256 x 256 x 3 tidal grid (24 576 models)
    large = None
    if world == 1 and not args.no_large_batch and not cfg3:
        Tl, ml, hl = (eng.dev(x) for x in grid_cfg3_share())
        best, hist = tidal_pass(Tl, ml, hl, Tl.numel(), 3)
        large = {'workload': 'one GPU share (1/8, interleaved) of the 256x256x3 tidal He-triplet grid '
                             '(BASELINE configs[3]): %d models' % Tl.numel(),
                 'value': Tl.numel() / (best * 1e-3), 'unit': UNIT, 'ms_per_step': best,
                 'status_histogram': hist, 'models_ok_frac': hist[0] / Tl.numel(),
                 'note': 'device-timed, inputs resident, best of 3 after L2 flush; not the headline config'}

    # BASELINE configs[3], the north-star target, on all ranks at every N: the whole 256 x 256 x 3 tidal
    # grid (196 608 models) cut over the ranks -- total work fixed (strong scaling)
    cfg3_line = None
    if not args.no_large_batch and not cfg3:
        Tg, mg_, hg = grid_cfg3()
        ig = pgrid.shard_indices(len(Tg), rank, world)
        best, hist = tidal_pass(eng.dev(Tg[ig]), eng.dev(mg_[ig]), eng.dev(hg[ig]), len(Tg), 2)
        cfg3_line = {'workload': config(world, 'cfg3')['workload'], 'global_models': len(Tg),
                     'models_per_gpu': len(ig), 'n_gpus': world, 'scaling': 'strong',
                     'value': len(Tg) / (best * 1e-3), 'unit': UNIT, 'ms_per_step': best,
                     'status_histogram': hist, 'models_ok_frac': hist[0] / len(Tg),
                     'note': 'forward_batch + chi2 + all_gather of chi2; device-timed, max over ranks, best of 2 '
                             'after L2 flush'}

    # BASELINE configs[4]: high-resolution transit grid (512 x 512 cells, 4096 wavelength bins),
    # He 10830 + C II + O I line sets, densities of all three species made on the GPU; 256 models cut
    # over the ranks
    highres = None
    if not args.no_large_batch and not cfg3:
        highres = highres_batch(local_rank, g, sc, r, flush, rank, world, barrier, max_over_ranks)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = len(os.sched_getaffinity(0))
        with mp.get_context('fork').Pool(cores, initializer=_cpu_init) as pool:
            cpu_base, _ = cpu_loop(T, md, hf, max(64, 2 * cores), pool, cores)

    hist = torch.stack([(out['status'] == k).sum() for k in range(6)]).to(torch.float64)
    if world > 1:
        dist.all_reduce(hist)
    hist = [int(x) for x in hist.tolist()]
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
                'scaling': 'strong' if cfg3 else 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic (T0, Mdot, h) grid; reference solar SED and HD 209458 b geometry',
                'config': config(world, args.workload),
                'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                        'h2d_bytes_per_step': int(h_in.numel() * 8),
                        'd2h_bytes_per_step': int(h_spec.numel() * 8 + h_chi2.numel() * 8 + h_status.numel() * 4),
                        'api': 'p_winds_b200.engine.Engine.forward_batch + chi2_batch (C ABI pwb_forward_batch), '
                               'pinned host buffers in and out'},
                'gpu_launches': int(launches), 'clocks': sampler.summary(), 'roofline': roofline,
                'cpu_baseline': cpu_base, 'per_rank': per_rank, 'cfg3': cfg3_line, 'large_batch': large,
                'highres_batch': highres,
                'status_histogram': {'ok': hist[0], 'h_solver_failed': hist[1], 'he_solver_failed': hist[2],
                                     'he_relax_warning': hist[3], 'nan': hist[4], 'h_work_limit': hist[5],
                                     'note': 'status 3 = an odeint of the helium relax loop warned: the reference\'s own '
                                             'result is undefined there; chi2 = +inf by default'},
                'models_ok_frac': hist[0] / n_global, 'wall_s_timed_region': wall_dev}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def highres_batch(device, sed, sc, r, flush, rank, world, barrier, max_over_ranks, n_models=256):
    """One pass of the configs[4] workload over `n_models` models (a 16 x 16 (T0, Mdot) grid, cut over
    the ranks): H -> He populations -> O II and C II / C III fractions -> three radiative-transfer calls
    (He 10830, C II 1335, O I 1302-1306; 4096 bins each) on a 512 x 512 limb-darkened transit map drawn
    on the GPU.  Own engine instance: the pixel table is per-context state."""
    import torch
    from p_winds_b200.engine import Engine
    from p_winds_b200 import lines, grid as pgrid
    eng = Engine(device)
    eng.set_sed(sed['wavelength'], sed['flux_lambda'])
    i0, depth, rm = eng.draw_transit(0.12086, R_PL * 71492000, 0.499, 0.0, grid_size=512, supersampling=2,
                                     limb_darkening_law='quadratic', ld_coefficient=(0.3, 0.2))
    r_si = r * R_PL * 71492000
    eng.set_geometry(i0, rm, r_si)
    n_side = int(round(n_models ** 0.5))
    TT, MM = np.meshgrid(np.linspace(10000, 13000, n_side), np.logspace(9.5, 12, n_side), indexing='ij')
    T, md = TT.ravel(), MM.ravel()
    n_models = T.size
    idx = pgrid.shard_indices(n_models, rank, world)
    Td, mdd, hfd, rd = eng.dev(T[idx]), eng.dev(md[idx]), eng.dev(np.full(len(idx), 0.9)), eng.dev(r)
    mu0 = (1 + 4 * (1 - hfd) / hfd) / (1 + (1 - hfd) / hfd)
    rates = [sc[k] for k in ('he_phi_1', 'he_phi_3', 'he_a_1', 'he_a_3', 'he_a_h_1', 'he_a_h_3')]
    ho = Engine.h_opts(R_PL, M_PL, relax_solution=True, exact_phi=True)
    heo = Engine.he_opts(R_PL, rates, (1.0, 0.0), True)
    oo = Engine.o_opts(R_PL, [sc[k] for k in ('o_phi', 'o_a', 'o_a_h', 'o_a_he')], relax_solution=True)
    co = Engine.c_opts(R_PL, [sc[k] for k in ('c_phi_1', 'c_phi_2', 'c_a_1', 'c_a_2', 'c_a_h_1', 'c_a_h_2', 'c_a_he')],
                       relax_solution=True, method='Radau')
    m_p = 1.67262192369e-27
    he3, c2, o1 = lines.he_3_properties(), lines.c_ii_properties(), lines.o_i_properties()
    sets = [(Engine.rt_opts(he3[:3], he3[3:6], [he3[6]] * 3, 4 * m_p), eng.dev(np.linspace(1.0827, 1.0832, 4096) * 1e-6)),
            (Engine.rt_opts(c2[:3], c2[3:6], c2[6:9], 12 * m_p), eng.dev(np.linspace(1332.9, 1337.1, 4096) * 1e-10)),
            (Engine.rt_opts(o1[:3], o1[3:6], o1[6:9], 16 * m_p), eng.dev(np.linspace(1300.9, 1307.1, 4096) * 1e-10))]
    o_frac, c_frac = 10 ** (8.69 - 12.00), 10 ** (8.43 - 12.00)
    stage = {}

    def run(ev=None):
        def mark(k):
            if ev is not None:
                ev[k] = torch.cuda.Event(enable_timing=True)
                ev[k].record()
        mark('t0')
        h = eng.h_ion_fraction_batch(Td, mdd, hfd, mu0, rd, ho)
        vs, rs, rhos = (h['scal'][:, k].contiguous() for k in (1, 2, 3))
        he = eng.he_population_batch(Td, hfd, vs, rs, rhos, rd, h['v'], h['rho'], h['f_r'], heo, status=h['status'].clone())
        f_he_ion = 1 - he['f_1'] - he['f_3']
        mark('t1')
        ox = eng.o_ion_fraction_batch(Td, hfd, vs, rs, rhos, rd, h['v'], h['rho'], h['f_r'], f_he_ion, oo,
                                      status=he['status'].clone())
        ca = eng.c_ion_fraction_batch(Td, hfd, vs, rs, rhos, rd, h['v'], h['rho'], h['f_r'], f_he_ion, co,
                                      status=he['status'].clone())
        mark('t2')
        hh, rho_d, rhos_d = hfd[:, None], h['rho'], rhos[:, None]
        v_si = h['v'] * vs[:, None] * 1000
        n_he = rho_d * rhos_d * (1 - hh) / (hh + 4 * (1 - hh)) / 1.67262192369e-24 * he['f_3'] * 1e6
        n_c = rho_d * rhos_d * c_frac / (hh + 4 * (1 - hh) + 12 * c_frac) / 1.67262192369e-24 * ca['f_cii'] * 1e6
        n_o = rho_d * rhos_d * o_frac / (hh + 4 * (1 - hh) + 16 * o_frac) / 1.67262192369e-24 * (1 - ox['f_o']) * 1e6
        spec = [eng.rt2d_batch(n, v_si, Td, wl, opt, status=st)
                for (opt, wl), n, st in zip(sets, (n_he, n_c, n_o), (he['status'], ca['status'], ox['status']))]
        mark('t3')
        return spec, he['status']

    run()
    best = None
    for _ in range(2):
        flush.fill_(1.0)
        barrier()
        ev = {}
        spec, st = run(ev)
        barrier()
        t = max_over_ranks(ev['t0'].elapsed_time(ev['t3']))
        if best is None or t < best:
            best = t
            stage = {'h_he_ms': ev['t0'].elapsed_time(ev['t1']), 'metals_ms': ev['t1'].elapsed_time(ev['t2']),
                     'rt_ms': ev['t2'].elapsed_time(ev['t3'])}
    ok = all(bool(torch.isfinite(s_[(st == 0) | (st == 3)]).all()) for s_ in spec)
    px = eng.n_active_pixels()
    # algorithmic work of the three pixel sums, SURVEY 8(d) convention: every lit cell, 26 flop per (cell, bin)
    lit = int((i0 != 0).sum())
    flop_rt = 3.0 * len(idx) * lit * (35.0 + 4096 * 26.0)
    return {'workload': 'BASELINE configs[4]: %d models x (He 10830 + C II + O I), 512x512 transit cells '
                        '(%d merged lit pixels), 4096 wavelength bins per line set; models cut over the ranks'
                        % (n_models, px),
            'n_gpus': world, 'models_per_gpu': len(idx), 'scaling': 'strong',
            'value': n_models / (best * 1e-3), 'unit': UNIT, 'ms_per_step': best, 'finite': ok,
            'rank0_stage_ms': stage,
            'rank0_rt_tflops_algorithmic': flop_rt / (stage['rt_ms'] * 1e-3) / 1e12,
            'note': 'device-timed through the stage-level C ABI calls, max over ranks, best of 2 after L2 flush; '
                    'the pixel sum (3 x 4096 bins x lit pixels per model) dominates; not the headline config'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-large-batch', action='store_true')
    ap.add_argument('--workload', default='grid64', choices=['grid64', 'cfg3'],
                    help="grid64: 64x64 grid per GPU (default, the headline line); cfg3: the whole 256x256x3 tidal "
                         "grid of BASELINE configs[3] cut over the ranks (ours only)")
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
