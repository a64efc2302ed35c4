"""Sharding of a retrieval grid over ranks and the final gather.

Models are independent, so the grid is split with no data-path collective: rank g
of W takes the flattened indices g, g+W, g+2W, ... (interleaved, because model cost
varies smoothly with Mdot and failed models exit early); the only exchange is one
all_gather of chi^2 / status (8 + 4 bytes per model) at the end."""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_models, rank, world_size):
    return np.arange(rank, n_models, world_size)


def shard_size(n_models, rank, world_size):
    return len(range(rank, n_models, world_size))


def gather_interleaved(local, n_models, rank, world_size, group=None):
    """all_gather the per-rank result rows (first axis = local models) back into grid
    order.  `local` is a torch tensor on the backend's device (cuda for NCCL, cpu for
    gloo).  Returns a tensor with n_models rows on every rank."""
    if world_size == 1:
        return local
    per = (n_models + world_size - 1) // world_size
    pad_shape = (per,) + tuple(local.shape[1:])
    padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    out = torch.empty((n_models,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for g in range(world_size):
        n_g = shard_size(n_models, g, world_size)
        out[g::world_size] = parts[g][:n_g]
    return out
