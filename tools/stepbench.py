#!/usr/bin/env python
"""Head step time of the launch-bound BASELINE configs (developer tool, not the contract bench).

    python tools/stepbench.py [--iters 200]

cfg 1 (32x16 heatmaps of 64x64), cfg 2 head (64x16 of 28x28) and cfg 3 head (8 hourglass stacks of 32x16 of
64x64) move 3-64 MiB per step: they are bound by launch count and latency, not by HBM, and
their logits are L2-resident by nature (the backbone's last conv has just written them).  Reported per
config: microseconds per head step (fused forward + finishing reduction + backward through the public
autograd API), eager and replayed from a CUDA graph, and the number of kernels of ours per step.  For cfg 3
the one-launch-for-all-stacks path (dsnt_head_stacked) is shown next to eight per-stack calls.
"""

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dsnt_pose2d_b200 as dp  # noqa: E402
from dsnt_pose2d_b200 import _lib  # noqa: E402


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def graphed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--reg', default='js')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    cases = [('cfg1 32x16x64x64', 1, 32, 16, 64, 64), ('cfg2 64x16x28x28', 1, 64, 16, 28, 28),
             ('cfg3 8 stacks 32x16x64x64', 8, 32, 16, 64, 64)]
    print('%-28s %-30s %10s %10s %9s %12s' % ('config', 'path', 'eager us', 'graph us', 'launches', 'Mhm/s graph'))
    for name, stacks, b, c, h, w in cases:
        zs = [torch.randn(b, c, h, w, device=dev, requires_grad=True) for _ in range(stacks)]
        target = torch.rand(b, c, 2, device=dev) * 1.6 - 0.8
        mask = (torch.rand(b, c, device=dev) > 0.1).float()

        def per_stack():
            for z in zs:
                z.grad = None
            total = None
            for z in zs:
                out = dp.dsnt_head(z, target, mask, reg=args.reg, hm_sigma=1.0, one_pass=False)
                total = out.loss if total is None else total + out.loss
            total.backward()

        def per_stack_step():
            for z in zs:
                z.grad = None
            total = None
            for z in zs:
                out = dp.dsnt_head(z, target, mask, reg=args.reg, hm_sigma=1.0, one_pass=True)
                total = out.loss if total is None else total + out.loss
            total.backward()

        def one_launch_step():
            for z in zs:
                z.grad = None
            _, total = dp.dsnt_head_stacked(zs, target, mask, reg=args.reg, hm_sigma=1.0, one_pass=True)
            total.backward()

        def one_launch():
            for z in zs:
                z.grad = None
            _, total = dp.dsnt_head_stacked(zs, target, mask, reg=args.reg, hm_sigma=1.0)
            total.backward()

        paths = [('two-kernel, per stack', per_stack), ('single-launch step, per stack', per_stack_step)]
        if stacks > 1:
            paths += [('two-kernel, stacked launches', one_launch), ('single-launch step, stacked', one_launch_step)]
        for pname, fn in paths:
            before = _lib.launch_count
            fn()
            launches = _lib.launch_count - before
            eager = timed(fn, args.iters)
            replay = graphed(fn)
            graph = timed(replay, args.iters)
            print('%-28s %-30s %10.1f %10.1f %9d %12.1f' % (name, pname, eager, graph, launches,
                                                           stacks * b * c / graph))


if __name__ == '__main__':
    main()
