#!/usr/bin/env python
"""Kernel micro-benchmark (developer tool, not the contract bench): times dsnt_head_fwd / dsnt_head_bwd
through the C ABI with CUDA events and prints achieved GB/s against the measured HBM peak.

    python tools/kbench.py [--configs cfg4,cfg5,...] [--regs js,var] [--dtypes f32,bf16] [--variants 0]
"""

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dsnt_pose2d_b200 import _lib  # noqa: E402

CONFIGS = {
    'cfg1': (32 * 16, 64, 64),
    'cfg2': (64 * 16, 28, 28),
    'cfg4': (4096 * 16, 64, 64),
    'cfg4s': (1024 * 16, 64, 64),
    'cfg5': (512 * 16, 256, 256),
    'cfg5s': (128 * 16, 256, 256),
    'c128': (2048 * 16, 128, 128),
    'c32': (16384 * 16, 32, 32),
}


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


def time_calls(fn, iters, warmup=3):
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='cfg4,cfg5,cfg1,cfg2')
    ap.add_argument('--regs', default='js,var,none,kl,mse')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--variants', default='0')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--step-only', action='store_true', help='time only the one-pass step kernel')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    peak, how = peak_gbs()
    print('HBM peak %.0f GB/s (%s)' % (peak, how))
    print('%-6s %-5s %-5s %3s | %9s %8s %6s | %9s %8s %6s | %8s %6s' % (
        'cfg', 'dtype', 'reg', 'var', 'fwd us', 'GB/s', 'frac', 'bwd us', 'GB/s', 'frac', 'Mhm/s', 'frac'))
    for cfg in args.configs.split(','):
        n, h, w = CONFIGS[cfg]
        for dt in args.dtypes.split(','):
            tdt = torch.float32 if dt == 'f32' else torch.bfloat16
            es = 4 if dt == 'f32' else 2
            nbytes = n * h * w * es
            nbuf = max(1, min(8, (400 << 20) // nbytes + 1))      # rotate buffers when one fits in L2
            zs = [torch.randn(n, h, w, device=dev).to(tdt) for _ in range(nbuf)]
            dzs = [torch.empty_like(zs[0]) for _ in range(min(nbuf, 2) if nbytes > (200 << 20) else nbuf)]
            target = torch.rand(n, 2, device=dev) * 1.6 - 0.8
            mask = (torch.rand(n, device=dev) > 0.1).float()
            coords = torch.empty(n, 2, device=dev)
            stats = torch.empty(n, 8, device=dev)
            terms = torch.empty(n, 2, device=dev)
            out8 = torch.empty(8, device=dev)
            gl = torch.ones((), device=dev)
            ws = _lib.finish_workspace(dev)
            stream = torch.cuda.current_stream().cuda_stream
            for reg in args.regs.split(','):
                rid = _lib.REG_IDS[reg]
                sigma = 2.0 / w
                for variant in [int(v) for v in args.variants.split(',')]:
                    def fwd(i):
                        _lib.call('dsnt_head_fwd', zs[i % nbuf].data_ptr(), _lib.dtype_id(zs[0]), 1, n, h, w,
                                  target.data_ptr(), rid, sigma, coords.data_ptr(), stats.data_ptr(),
                                  terms.data_ptr(), variant, stream)

                    def bwd(i):
                        _lib.call('dsnt_head_bwd', zs[i % nbuf].data_ptr(), _lib.dtype_id(zs[0]), 1, n, h, w,
                                  target.data_ptr(), mask.data_ptr(), stats.data_ptr(), None, None, gl.data_ptr(),
                                  out8[3:4].data_ptr(), 1.0, rid, sigma, 0, dzs[i % len(dzs)].data_ptr(), variant,
                                  stream)
                    try:
                        if args.step_only:
                            raise StopIteration
                        tf, tf_min = time_calls(fwd, args.iters)
                        _lib.call('dsnt_finish_loss', terms.data_ptr(), mask.data_ptr(), n, 1.0, out8.data_ptr(),
                                  ws.data_ptr(), stream)
                        tb, tb_min = time_calls(bwd, args.iters)
                    except StopIteration:
                        tf = tb = None
                    except RuntimeError as e:
                        print('%-6s %-5s %-5s %3d | FAILED %s' % (cfg, dt, reg, variant, e))
                        continue
                    ts = None
                    if _lib.LIB.dsnt_head_step_supported_reg(_lib.dtype_id(zs[0]), h, w, rid):
                        cnt8 = torch.empty(8, device=dev)
                        _lib.call('dsnt_mask_count', mask.data_ptr(), n, cnt8.data_ptr(), ws.data_ptr(), stream)

                        def step(i):
                            _lib.call('dsnt_head_step', zs[i % nbuf].data_ptr(), _lib.dtype_id(zs[0]), n, h, w,
                                      target.data_ptr(), mask.data_ptr(), cnt8[3:4].data_ptr(), None, 1.0, rid, sigma, 0,
                                      coords.data_ptr(), stats.data_ptr(), terms.data_ptr(), dzs[i % len(dzs)].data_ptr(),
                                      stream)
                        ts, _ = time_calls(step, args.iters)
                    fused_note = ''
                    if _lib.LIB.dsnt_head_step_fused_supported(_lib.dtype_id(zs[0]), h, w, rid, sigma):
                        def fused(i):
                            _lib.call('dsnt_head_step_fused', zs[i % nbuf].data_ptr(), _lib.dtype_id(zs[0]), n, h, w,
                                      target.data_ptr(), mask.data_ptr(), None, 1.0, rid, sigma, 0, coords.data_ptr(),
                                      stats.data_ptr(), dzs[i % len(dzs)].data_ptr(), out8.data_ptr(), ws.data_ptr(), stream)
                        tfu, _ = time_calls(fused, args.iters)
                        # back to back: one pair of events around all launches (no event between two kernels)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for i in range(args.iters):
                            fused(i)
                        e1.record()
                        torch.cuda.synchronize()
                        tb2b = e0.elapsed_time(e1) / args.iters
                        off = _lib.LIB.dsnt_finish_trace_offset_bytes() // 4
                        st = ws[off:off + 16].view(torch.int64).cpu()
                        fused_note = ' || single launch %7.1f us (%5.3f), back to back %7.1f us (%5.3f), inside the kernel %7.1f us' % (
                            tfu * 1e3, (2 * nbytes + 60 * n) / tfu / 1e6 / peak, tb2b * 1e3,
                            (2 * nbytes + 60 * n) / tb2b / 1e6 / peak, (st[4] - st[0]).item() / 1e3)
                    if tf is None:
                        print('%-6s %-5s %-5s || one-pass step %7.1f us %6.0f GB/s %5.3f %8.2f Mhm/s' % (
                            cfg, dt, reg, ts * 1e3, 2 * nbytes / ts / 1e6, 2 * nbytes / ts / 1e6 / peak, n / ts / 1e3) + fused_note)
                        continue
                    gf = nbytes / tf / 1e6
                    gb = 2 * nbytes / tb / 1e6
                    tot = (3 * nbytes + 96 * n) / (tf + tb) / 1e6
                    print('%-6s %-5s %-5s %3d | %9.1f %8.0f %6.3f | %9.1f %8.0f %6.3f | %8.2f %6.3f' % (
                        cfg, dt, reg, variant, tf * 1e3, gf, gf / peak, tb * 1e3, gb, gb / peak,
                        n / (tf + tb) / 1e3, tot / peak) + (
                        '' if ts is None else ' || one-pass step %7.1f us %6.0f GB/s %5.3f %8.2f Mhm/s' % (
                            ts * 1e3, 2 * nbytes / ts / 1e6, 2 * nbytes / ts / 1e6 / peak, n / ts / 1e3)) + fused_note)
            del zs, dzs
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
