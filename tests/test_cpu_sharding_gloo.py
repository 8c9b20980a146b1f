"""world_size-2 gloo test of the N>1 host logic: batch-sharded ranks exchange only the three
partial sums, and with the GLOBAL denominator the sharded loss and gradients equal the single-process ones.
The per-shard arithmetic here is the fp64 oracle (the CUDA kernels need a GPU); the exchange code under
test is the product's `all_reduce_sums` + `shard`."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form as cf


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, z, target, mask, reg, sigma, coeff, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dsnt_pose2d_b200.head import all_reduce_sums
        from dsnt_pose2d_b200.parallel import shard
        zs = shard(z, rank, world).numpy().astype(np.float64)
        ts = shard(target, rank, world).numpy().astype(np.float64)
        ms = None if mask is None else shard(mask, rank, world).numpy().astype(np.float64)
        b, c, h, w = zs.shape
        n = b * c
        local = cf.head(zs.reshape(n, h, w), ts.reshape(n, 2), None if ms is None else ms.reshape(n), reg=reg,
                        sigma=sigma, reg_coeff=coeff)
        wts = np.ones(n) if ms is None else ms.reshape(n)
        out8 = torch.zeros(8, dtype=torch.float64)
        out8[0] = float((wts * local['dist']).sum())
        out8[1] = float((wts * local['reg_terms']).sum())
        out8[2] = float(wts.sum())
        all_reduce_sums(out8, dist.group.WORLD)
        denom = max(out8[2].item(), 1.0)
        loss = (out8[0].item() + coeff * out8[1].item()) / denom
        # gradients with the global denominator: rescale the local-mean gradients
        local_denom = max(wts.sum(), 1.0)
        dz = local['dz'] * (local_denom / denom)
        ret[rank] = (loss, dz.reshape(b, c, h, w))
    finally:
        dist.destroy_process_group()


def _run(mask_kind, reg):
    torch.manual_seed(0)
    b, c, h, w = 6, 4, 12, 12
    z = torch.randn(b, c, h, w)
    target = torch.rand(b, c, 2) * 1.6 - 0.8
    if mask_kind == 'none':
        mask = None
    else:
        mask = (torch.rand(b, c) > 0.4).float()
        mask[:3] = 0.0                                  # rank 0 sees no visible joint at all
        mask[3, 0] = 1.0
    sigma, coeff = 2.0 / w, 0.5
    n = b * c
    full = cf.head(z.numpy().astype(np.float64).reshape(n, h, w), target.numpy().astype(np.float64).reshape(n, 2),
                   None if mask is None else mask.numpy().astype(np.float64).reshape(n), reg=reg, sigma=sigma,
                   reg_coeff=coeff)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), z, target, mask, reg, sigma, coeff, ret), nprocs=world, join=True)
    losses = [ret[r][0] for r in range(world)]
    assert abs(losses[0] - losses[1]) < 1e-15                       # every rank ends with the same global loss
    assert abs(losses[0] - full['loss']) < 1e-12 * max(1.0, abs(full['loss']))
    dz = np.concatenate([ret[r][1] for r in range(world)], axis=0)
    assert np.abs(dz - full['dz'].reshape(b, c, h, w)).max() < 1e-12


def test_sharded_equals_single_process_with_uneven_mask():
    _run('uneven', 'js')


def test_sharded_equals_single_process_without_mask():
    _run('none', 'var')


def _worker_one_pass(rank, world, port, z, target, mask, reg, sigma, coeff, ret):
    """The exchange order of the one-pass step (head._FusedHeadStep): the mask COUNT is all-reduced before the
    forward so every rank differentiates with the global denominator directly; the loss sums follow."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from dsnt_pose2d_b200.head import all_reduce_sums
        from dsnt_pose2d_b200.parallel import shard
        zs = shard(z, rank, world).numpy().astype(np.float64)
        ts = shard(target, rank, world).numpy().astype(np.float64)
        ms = shard(mask, rank, world).numpy().astype(np.float64)
        b, c, h, w = zs.shape
        n = b * c
        cnt8 = torch.zeros(8, dtype=torch.float64)
        cnt8[2] = float(ms.sum())                                  # dsnt_mask_count on the local shard
        all_reduce_sums(cnt8, dist.group.WORLD)
        denom = max(cnt8[2].item(), 1.0)                           # dsnt_combine_loss
        local = cf.head(zs.reshape(n, h, w), ts.reshape(n, 2), ms.reshape(n), reg=reg, sigma=sigma, reg_coeff=coeff)
        local_denom = max(ms.sum(), 1.0)
        dz = local['dz'] * (local_denom / denom)                   # = the step kernel's weights mask/denom_global
        out8 = torch.zeros(8, dtype=torch.float64)
        out8[0] = float((ms.reshape(n) * local['dist']).sum())
        out8[1] = float((ms.reshape(n) * local['reg_terms']).sum())
        out8[2] = float(ms.sum())
        all_reduce_sums(out8, dist.group.WORLD)
        assert out8[2].item() == cnt8[2].item()
        ret[rank] = ((out8[0].item() + coeff * out8[1].item()) / denom, dz.reshape(b, c, h, w))
    finally:
        dist.destroy_process_group()


def test_one_pass_exchange_order_equals_single_process():
    torch.manual_seed(1)
    b, c, h, w = 6, 4, 12, 12
    z = torch.randn(b, c, h, w)
    target = torch.rand(b, c, 2) * 1.6 - 0.8
    mask = (torch.rand(b, c) > 0.4).float()
    mask[:3] = 0.0
    mask[3, 0] = 1.0
    sigma, coeff, n = 2.0 / w, 0.5, b * c
    full = cf.head(z.numpy().astype(np.float64).reshape(n, h, w), target.numpy().astype(np.float64).reshape(n, 2),
                   mask.numpy().astype(np.float64).reshape(n), reg='js', sigma=sigma, reg_coeff=coeff)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_one_pass, args=(2, _free_port(), z, target, mask, 'js', sigma, coeff, ret), nprocs=2, join=True)
    assert abs(ret[0][0] - ret[1][0]) < 1e-15 and abs(ret[0][0] - full['loss']) < 1e-12 * max(1.0, abs(full['loss']))
    dz = np.concatenate([ret[r][1] for r in range(2)], axis=0)
    assert np.abs(dz - full['dz'].reshape(b, c, h, w)).max() < 1e-12
