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
    gen = torch.Generator().manual_seed(7)
    b = 8 * world + 3                       # uneven shards
    z = torch.randn(b, 16, 64, 64, generator=gen)
    target = torch.rand(b, 16, 2, generator=gen) * 1.6 - 0.8
    mask = (torch.rand(b, 16, generator=gen) > 0.3).float()
    mask[:5] = 0.0                          # rank 0 sees very few visible joints
    ok = True
    _head.STEP_MIN_BYTES = 0          # the batch here is small: do not let the dispatch route the one-pass step to two kernels
    for one_pass in (False, True):
        for reg in ('js', 'var'):
            zs = shard(z, rank, world).contiguous().to(dev).requires_grad_(True)
            out = dp.dsnt_head(zs, shard(target, rank, world).contiguous().to(dev), shard(mask, rank, world).contiguous().to(dev),
                               reg=reg, group=dist.group.WORLD if world > 1 else None, one_pass=one_pass)
            out.loss.backward()
            grads = [torch.empty(shard(z, r, world).shape, device=dev) for r in range(world)]
            if world > 1:
                # shards are uneven: gather through a padded buffer
                pad = torch.zeros(shard(z, 0, world).shape, device=dev)
                pad[:zs.shape[0]] = zs.grad
                bufs = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(bufs, pad)
                grads = [bufs[r][:shard(z, r, world).shape[0]] for r in range(world)]
            else:
                grads = [zs.grad]
            if world > 1:
                # every rank must hold the SAME loss (the totals are added in rank order on every rank)
                losses = [torch.empty_like(out.loss) for _ in range(world)]
                dist.all_gather(losses, out.loss.detach())
                same = all(torch.equal(losses[0], x) for x in losses)
                ok &= same
                if rank == 0 and not same:
                    print('loss differs between ranks: %r' % ([x.item() for x in losses],), flush=True)
            if rank == 0:
                zf = z.to(dev).requires_grad_(True)
                full = dp.dsnt_head(zf, target.to(dev), mask.to(dev), reg=reg, one_pass=False)
                full.loss.backward()
                g = torch.cat(grads, 0)
                e_loss = abs(out.loss.item() - full.loss.item()) / abs(full.loss.item())
                e_dz = ((g - zf.grad).norm() / zf.grad.norm()).item()
                good = e_loss < 2e-6 and e_dz < 2e-6
                ok &= good
                print('world %d one_pass %-5s reg %-3s loss rel.err %.1e dz rel.L2 %.1e %s' % (
                    world, one_pass, reg, e_loss, e_dz, 'ok' if good else 'MISMATCH'), flush=True)
    if world > 1:
        from dsnt_pose2d_b200.parallel import PeerExchange
        peer = PeerExchange.get(dist.group.WORLD, dev)
        if peer is not None:
            peer.check()
        if rank == 0:
            print('exchange of the partial sums: %s' % ('peer memory, fused into the finishing kernels' if peer is not None
                                                         else 'NCCL all-reduce'), flush=True)
        dist.barrier()
    if rank == 0:
        print('check_sharded: %s' % ('PASS' if ok else 'FAIL'), flush=True)
    os._exit(0 if ok else 1)


if __name__ == '__main__':
    main()
