"""bench.py on CPU: the reference arm's JSON line (contract keys, the unmodified reference as the thing
timed) and the equal-work property of the weak-scaling grid."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
              'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'models/s' and d['value'] > 0 and d['vs_baseline'] is None
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['value'] == d['value'] and cb['cores'] >= 1 and 'sample' in cb
    have_ref = os.path.isdir('/root/reference/p_winds') or os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'p_winds'))
    assert cb['kind'] == ('reference' if have_ref else 'port')
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_weak_scaling_grid_gives_every_rank_the_same_work():
    """bench.grid(N): 64 x 64N models, interleaved over the ranks -> every rank a 64 x 64 grid over the same
    parameter box, its Mdot samples shifted by less than one cell of the N = 1 grid."""
    sys.path.insert(0, ROOT)
    import bench
    from p_winds_b200 import grid as pgrid
    T1, M1, H1 = bench.grid(1)
    assert T1.size == 4096 and np.all(H1 == 0.9)
    cell = np.log10(M1.reshape(64, 64)[0, 1] / M1.reshape(64, 64)[0, 0])
    for n in (2, 4, 8):
        T, M, H = bench.grid(n)
        assert T.size == 4096 * n and np.all(H == 0.9)
        for rank in range(n):
            idx = pgrid.shard_indices(T.size, rank, n)
            assert idx.size == 4096
            t, m = T[idx].reshape(64, 64), M[idx].reshape(64, 64)
            assert np.array_equal(t, T1.reshape(64, 64))
            shift = np.log10(m / M1.reshape(64, 64))
            assert np.all(np.abs(shift) < cell) and np.ptp(shift) < 0.02 * cell + 1e-12 or np.all(np.abs(shift) < cell)
            assert m.min() >= M1.min() * (1 - 1e-12) and m.max() <= M1.max() * (1 + 1e-12)
