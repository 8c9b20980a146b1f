#!/usr/bin/env python
"""Small workload for compute-sanitizer (tools/sanitize.sh): every kernel family of the hot path once or twice, on few
heatmaps -- the sanitizer instruments every memory access, so sizes are what finishes in minutes.

    compute-sanitizer --tool racecheck python tools/sanitize_workload.py [case ...]

Cases: step (64x64 single-launch one-pass, fp32 + bf16, every regulariser), stacked (hourglass single launch),
generic (one-pass kernels for 28x28 / KL: dsnt_mask_count + dsnt_head_step + dsnt_finish_loss), pair (256x256 fp32 / bf16 on
clusters of two and four CTAs: DSMEM exchange, staged bulk stores), two (forward + streaming backward), level1 (nn.* operators), peer (2 ranks under
torchrun: the exchanges over peer memory)."""

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dsnt_pose2d_b200 as dp  # noqa: E402
from dsnt_pose2d_b200 import head as _head  # noqa: E402

DEV = 'cuda'


def inputs(b, c, h, w, dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(b, c, h, w, generator=g).to(DEV).to(dtype).requires_grad_(True)
    t = (torch.rand(b, c, 2, generator=g) * 1.6 - 0.8).to(DEV)
    m = (torch.rand(b, c, generator=g) > 0.2).float().to(DEV)
    return z, t, m


def run(z, t, m, reg, one_pass, group=None):
    out = dp.dsnt_head(z, t, m, reg=reg, hm_sigma=1.0, one_pass=one_pass, group=group)
    out.loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(out.loss).item() and torch.isfinite(z.grad.float()).all().item()
    return out.loss.item()


def case_step():
    for dtype in (torch.float32, torch.bfloat16):
        for reg in ('none', 'var', 'js', 'mse'):
            z, t, m = inputs(20, 16, 64, 64, dtype)           # 320 heatmaps: more than one per CTA, ring wraps in a few CTAs
            print('step', dtype, reg, run(z, t, m, reg, True))


def case_stacked():
    z, t, m = inputs(4, 16, 64, 64)
    zs = [z.detach().clone().requires_grad_(True) for _ in range(3)]
    coords, loss = dp.dsnt_head_stacked(zs, t, m, reg='js', hm_sigma=1.0, one_pass=True)
    loss.backward()
    torch.cuda.synchronize()
    print('stacked', loss.item())


def case_generic():
    _head.STEP_MIN_BYTES = 0
    for (h, w, reg) in ((28, 28, 'js'), (64, 64, 'kl'), (32, 32, 'var')):
        z, t, m = inputs(4, 16, h, w)
        print('generic', h, w, reg, run(z, t, m, reg, True))


def case_pair():
    # two clusters only, so that each walks three heatmaps: the staged buffers are handed on (bulk-store read -> next load)
    os.environ['DSNT_TUNE_STEP_PAIR_CLUSTERS'] = '2'
    _head.STEP_MIN_BYTES = 0
    _head.USE_PAIR_STEP_BF16 = True
    for cs in ('2', '4'):
        os.environ['DSNT_TUNE_STEP_PAIR_CS'] = cs
        for dtype in (torch.float32, torch.bfloat16):
            for reg in ('var', 'none', 'js', 'mse'):
                z, t, m = inputs(2, 3, 256, 256, dtype)
                print('pair', cs, dtype, reg, run(z, t, m, reg, True))
    del os.environ['DSNT_TUNE_STEP_PAIR_CS']


def case_two():
    for dtype in (torch.float32, torch.bfloat16):
        for (h, w, reg) in ((64, 64, 'js'), (64, 64, 'kl'), (28, 28, 'mse'), (7, 7, 'var'), (128, 128, 'js')):
            z, t, m = inputs(2, 16, h, w, dtype)
            print('two', dtype, h, w, reg, run(z, t, m, reg, False))


def case_level1():
    z, t, m = inputs(2, 16, 32, 32)
    p = dp.nn.flat_softmax(z)
    c = dp.nn.dsnt(p)
    loss = dp.nn.euclidean_loss(c, t, m) + dp.nn.js_reg_loss(p, t, 2.0 / 32, m) + dp.nn.variance_reg_loss(p, t, 2.0 / 32, m)
    loss.backward()
    g = dp.nn.make_gauss(t, 32, 32, 2.0 / 32)
    torch.cuda.synchronize()
    print('level1', loss.item(), g.sum().item())


def case_peer():
    import torch.distributed as dist
    from dsnt_pose2d_b200.parallel import init_from_env, shard, PeerExchange
    rank, local, world = init_from_env('nccl')
    torch.cuda.set_device(local)
    global DEV
    DEV = 'cuda:%d' % local
    _head.STEP_MIN_BYTES = 0
    z, t, m = inputs(4 * world + 1, 16, 64, 64, seed=5)
    for one_pass in (True, False):
        zs = shard(z.detach(), rank, world).contiguous().requires_grad_(True)
        loss = run(zs, shard(t, rank, world).contiguous(), shard(m, rank, world).contiguous(), 'js', one_pass, dist.group.WORLD)
        print('peer rank', rank, one_pass, loss, flush=True)
    PeerExchange.check_all()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == '__main__':
    cases = sys.argv[1:] or ['step', 'stacked', 'generic', 'pair', 'two', 'level1']
    for name in cases:
        globals()['case_' + name]()
    print('SANITIZE-WORKLOAD-DONE')
