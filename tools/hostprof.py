#!/usr/bin/env python
"""Where the HOST time of one eager head step goes (developer tool): cProfile over the cfg 1 step through the public API.

    python tools/hostprof.py [--iters 2000] [--stacks 1|8]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dsnt_pose2d_b200 as dp  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--stacks', type=int, default=1)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    zs = [torch.randn(32, 16, 64, 64, device=dev, requires_grad=True) for _ in range(args.stacks)]
    target = torch.rand(32, 16, 2, device=dev) * 1.6 - 0.8
    mask = (torch.rand(32, 16, device=dev) > 0.1).float()

    def step():
        for z in zs:
            z.grad = None
        if args.stacks == 1:
            dp.dsnt_head(zs[0], target, mask, reg='js', hm_sigma=1.0, one_pass=True).loss.backward()
        else:
            dp.dsnt_head_stacked(zs, target, mask, reg='js', hm_sigma=1.0, one_pass=True)[1].backward()

    for _ in range(50):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        step()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print('%d stacks: host %.1f us/step to enqueue, %.1f us/step until the GPU is done' % (
        args.stacks, t_host / args.iters * 1e6, t_all / args.iters * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.iters):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(22)


if __name__ == '__main__':
    main()
