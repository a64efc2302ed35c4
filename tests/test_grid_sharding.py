"""N > 1 host logic on CPU: world_size-2 gloo run of the shard -> compute -> gather
pattern bench.py uses."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_chi2(T, md, h):
    return (T / 1e4) ** 2 + np.log10(md) * h


def _worker(rank, world, port, n_models, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from p_winds_b200 import grid
    rng = np.random.default_rng(0)
    T = rng.uniform(9000, 13000, n_models)
    md = 10 ** rng.uniform(9.5, 12, n_models)
    h = rng.uniform(0.85, 0.95, n_models)
    idx = grid.shard_indices(n_models, rank, world)
    local = torch.from_numpy(np.stack([_fake_chi2(T[idx], md[idx], h[idx]), idx.astype(float)], axis=1))
    full = grid.gather_interleaved(local, n_models, rank, world)
    status = grid.gather_interleaved(torch.from_numpy((idx % 3).astype(np.int32)), n_models, rank, world)
    ok = bool(np.allclose(full[:, 0].numpy(), _fake_chi2(T, md, h), rtol=0, atol=0)
              and np.array_equal(full[:, 1].numpy(), np.arange(n_models))
              and np.array_equal(status.numpy(), np.arange(n_models) % 3))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_models):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_models, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    return res


def test_shard_and_gather_world2():
    assert all(ok for _, ok in _run(2, 4096))


def test_shard_and_gather_ragged():
    # n_models not divisible by the world size, and fewer models than ranks' padding
    assert all(ok for _, ok in _run(2, 4097))
    assert all(ok for _, ok in _run(3, 5))


def test_shard_indices_cover_grid_once():
    from p_winds_b200 import grid
    for n, w in ((4096, 8), (4097, 8), (5, 8), (0, 2)):
        allidx = np.concatenate([grid.shard_indices(n, r, w) for r in range(w)]) if n else np.zeros(0)
        assert sorted(allidx.tolist()) == list(range(n))
        assert sum(grid.shard_size(n, r, w) for r in range(w)) == n
