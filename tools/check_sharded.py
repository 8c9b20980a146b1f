#!/usr/bin/env python
"""Multi-GPU check of the sharded head (developer tool; run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py

Every rank evaluates its batch shard with `group=`; rank 0 also evaluates the WHOLE batch in one process.  Loss and
gradients of the sharded run (both the two-kernel and the one-pass path) must equal the single-process ones."""

import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dsnt_pose2d_b200 as dp  # noqa: E402
from dsnt_pose2d_b200 import head as _head  # noqa: E402
from dsnt_pose2d_b200.parallel import init_from_env, shard  # noqa: E402


def main():
    rank, local, world = init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    ok = True
    default_min = _head.STEP_MIN_BYTES
    # (global batch, STEP_MIN_BYTES): uneven shards; a batch whose shards straddle the default threshold (128 | 127 samples
    # of 256 KiB on two ranks: every rank must still take the same path); fewer samples than ranks (EMPTY shards)
    cases = [(8 * world + 3, 0), (128 * world - 1, default_min), (world - 1, 0)]
    cases = [c for c in cases if c[0] > 0]
    for b, min_bytes in cases:
        gen = torch.Generator().manual_seed(7 + b)
        z = torch.randn(b, 16, 64, 64, generator=gen)
        target = torch.rand(b, 16, 2, generator=gen) * 1.6 - 0.8
        mask = (torch.rand(b, 16, generator=gen) > 0.3).float()
        mask[:5] = 0.0                          # rank 0 sees very few visible joints
        if b > 5:
            mask[5, 0] = 1.0
        _head.STEP_MIN_BYTES = min_bytes
        for one_pass in (False, True):
            for reg in ('js', 'var'):
                zs = shard(z, rank, world).contiguous().to(dev).requires_grad_(True)
                out = dp.dsnt_head(zs, shard(target, rank, world).contiguous().to(dev),
                                   shard(mask, rank, world).contiguous().to(dev),
                                   reg=reg, group=dist.group.WORLD if world > 1 else None, one_pass=one_pass)
                out.loss.backward()
                if zs.grad is None:                 # an empty shard has no gradient to speak of
                    zs.grad = torch.zeros_like(zs)
                if world > 1:
                    # shards are uneven: gather through a padded buffer
                    pad = torch.zeros((max(1, shard(z, 0, world).shape[0]),) + tuple(z.shape[1:]), device=dev)
                    pad[:zs.shape[0]] = zs.grad
                    bufs = [torch.empty_like(pad) for _ in range(world)]
                    dist.all_gather(bufs, pad)
                    grads = [bufs[r][:shard(z, r, world).shape[0]] for r in range(world)]
                    # every rank must hold the SAME loss (the totals are added in rank order on every rank)
                    losses = [torch.empty_like(out.loss) for _ in range(world)]
                    dist.all_gather(losses, out.loss.detach())
                    same = all(torch.equal(losses[0], x) for x in losses)
                    ok &= same
                    if rank == 0 and not same:
                        print('loss differs between ranks: %r' % ([x.item() for x in losses],), flush=True)
                else:
                    grads = [zs.grad]
                if rank == 0:
                    zf = z.to(dev).requires_grad_(True)
                    full = dp.dsnt_head(zf, target.to(dev), mask.to(dev), reg=reg, one_pass=False)
                    full.loss.backward()
                    g = torch.cat(grads, 0)
                    e_loss = abs(out.loss.item() - full.loss.item()) / max(abs(full.loss.item()), 1e-30)
                    e_dz = ((g - zf.grad).norm() / zf.grad.norm().clamp_min(1e-30)).item()
                    good = e_loss < 2e-6 and e_dz < 2e-6
                    ok &= good
                    print('world %d batch %4d one_pass %-5s reg %-3s loss rel.err %.1e dz rel.L2 %.1e %s' % (
                        world, b, one_pass, reg, e_loss, e_dz, 'ok' if good else 'MISMATCH'), flush=True)
    _head.STEP_MIN_BYTES = default_min
    if world > 1:
        from dsnt_pose2d_b200.parallel import PeerExchange
        peer = PeerExchange.get(dist.group.WORLD, dev)
        if peer is not None:
            peer.check()
        if rank == 0:
            print('exchange of the partial sums: %s' % ('peer memory, fused into the finishing kernels' if peer is not None
                                                         else 'NCCL all-reduce'), flush=True)
        dist.barrier()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    if rank == 0:
        print('check_sharded: %s' % ('PASS' if ok else 'FAIL'), flush=True)
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
