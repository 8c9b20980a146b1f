"""Batch sharding of the head across ranks.

Every heatmap is independent; ranks own contiguous slices of the batch dimension and the only exchange
is that of three partial sums inside `dsnt_head(..., group=...)`.  One process per GPU.  On one node the sums
travel through peer-mapped memory from inside the finishing kernel (`PeerExchange`, include/dsnt_b200.h:
dsnt_finish_loss_peer); otherwise, or with DSNT_PEER_EXCHANGE=0, through a 3-float NCCL all-reduce.
"""

import ctypes
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


class PeerExchange:
    """Exchange buffers of one process group on one node for the fused finishing reductions.

    Every rank allocates a small symmetric-memory buffer that all ranks map (torch.distributed._symmetric_memory:
    CUDA VMM handles shared between the processes, P2P over NVLink); the kernels then write their partial sums into
    every rank's buffer and read the others' from their own -- no collective launch, CUDA-graph capturable.
    `PeerExchange.get(group, device)` returns None where that is not possible (single rank, no symmetric memory,
    DSNT_PEER_EXCHANGE=0); the caller then all-reduces with NCCL."""

    _cache = {}        # (ranks of the group, its name, device index) -> (weakref to the group, PeerExchange or None)

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError('peer exchange serves at most 16 ranks')
        nfloats = _lib.LIB.dsnt_peer_exchange_bytes() // 4
        buf = symm.empty(nfloats, dtype=torch.float32, device=device)
        buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm.rendezvous(buf, group)
        self.buf = buf
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError('symmetric memory returned %r peer pointers for %d ranks' % (ptrs, self.world))
        self.peers = (ctypes.c_void_p * self.world)(*ptrs)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)     # [epoch, error]
        torch.cuda.synchronize(device)
        self._args = (ctypes.cast(self.peers, ctypes.c_void_p), self.rank, self.world, self.state[0:1].data_ptr(),
                      self.state[1:2].data_ptr())

    def args(self):
        """(peers, rank, world, epoch*, error*) as the *_peer entry points take them."""
        return self._args

    def check(self):
        """Host read of the error flag (a peer that never arrived); not called on the hot path."""
        if int(self.state[1].item()) != 0:
            raise RuntimeError('peer exchange timed out: a rank of the group did not take part')

    @classmethod
    def check_all(cls):
        """`check()` for every exchange of this process: call where the host synchronises anyway (after a bench run, when
        the coordinates are copied out, when a loss is not finite)."""
        for _, px in list(cls._cache.values()):
            if px is not None:
                px.check()

    @staticmethod
    def _key(group, device):
        import torch.distributed as dist
        try:
            ranks = tuple(dist.get_process_group_ranks(group))
        except Exception:          # noqa: BLE001
            ranks = (dist.get_world_size(group),)
        return ranks, getattr(group, 'group_name', None), torch.device(device).index

    @classmethod
    def get(cls, group, device):
        import weakref
        import torch.distributed as dist
        if group is None or not dist.is_available() or not dist.is_initialized():
            return None
        if os.environ.get('DSNT_PEER_EXCHANGE', '1') == '0' or dist.get_world_size(group) < 2:
            return None
        if dist.get_backend(group) != 'nccl' or torch.device(device).type != 'cuda':
            return None
        key = cls._key(group, device)
        hit = cls._cache.get(key)
        if hit is not None and hit[0]() is group:      # a NEW group with the ranks and name of a destroyed one rebuilds
            return hit[1]
        px, why = None, None
        try:
            px = cls(group, torch.device(device))
        except Exception as e:          # noqa: BLE001  (no symmetric memory on this system: NCCL all-reduce instead)
            why = e
        # Every rank must end up on the SAME transport: a rank that fell back to NCCL while the others spin in the peer
        # exchange would hang them until the 20 s timeout.  Agree on the minimum, which is also the barrier that makes sure
        # every buffer is zero before anybody's first exchange can land in it.
        ok = torch.tensor([1 if px is not None else 0], dtype=torch.int32, device=torch.device(device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if px is None or why is not None:
                import warnings
                warnings.warn('dsnt_pose2d_b200: peer exchange unavailable (%s); using NCCL all-reduce' % (why,))
            px = None
        try:
            ref = weakref.ref(group)
        except TypeError:
            ref = (lambda g: (lambda: g))(group)
        cls._cache[key] = (ref, px)
        return px
