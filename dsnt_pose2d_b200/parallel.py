"""Batch sharding of the head across ranks.

Every heatmap is independent; ranks own contiguous slices of the batch dimension and the only exchange
is the 3-float all-reduce inside `dsnt_head(..., group=...)`.  One process per GPU, NCCL over NVLink.
"""

import os

import torch


def shard_range(batch, rank, world):
    """Samples [lo, hi) of a global batch owned by `rank`; the first `batch % world` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad rank/world: %r/%r' % (rank, world))
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t, rank, world, dim=0):
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, local_rank, world)."""
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world
