#!/usr/bin/env python
"""Kernel timings of the SURVEY.md 8(f) rows (developer tool, not the contract bench): CUDA events around each C-ABI
call, achieved GB/s over the ALGORITHMIC bytes of the call against the HBM peak.

    python tools/nextbench.py [--iters 20]

  preact   dsnt_head_preact_fwd / _bwd at cfg 4 shape (65 536 heatmaps of 64x64): fwd reads z once (H*W*s bytes),
           bwd reads z and writes dz (2*H*W*s)
  flip     dsnt_flip_tta_fwd on 2*B images x 16 joints: reads both heatmap sets once (2*H*W*s per output heatmap)
  encode   dsnt_draw_gaussians: writes H*W*4 per heatmap;  decode  dsnt_decode_heatmaps: reads H*W*s per heatmap
Next to each, the same computation as the reference's own op sequence in eager torch ON THE SAME GPU (labelled
"torch eager"; not the CPU baseline), where that is a handful of ops.
"""

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dsnt_pose2d_b200 as dp  # noqa: E402
from dsnt_pose2d_b200 import _lib, util as du  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


def time_calls(fn, iters, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def line(name, ms, nbytes, n, peak, extra=''):
    gbs = nbytes / ms / 1e6
    print('%-44s %9.1f us %8.0f GB/s %6.3f of peak %9.2f M heatmaps/s %s' % (name, ms * 1e3, gbs, gbs / peak,
                                                                             n / ms / 1e3, extra))


def eager_flip(pair, perm):
    """src/dsnt/inference.py:43-47 as torch ops on the GPU."""
    hm1, hm2 = pair.chunk(2, 0)
    hm2 = hm2.flip(-1).index_select(1, perm)
    hm = (hm1 + hm2) / 2
    b, c, h, w = hm.shape
    p = torch.softmax(hm.reshape(-1, h * w), -1).view(b, c, h, w)
    xs = torch.linspace(-(w - 1) / w, (w - 1) / w, w, device=hm.device)
    ys = torch.linspace(-(h - 1) / h, (h - 1) / h, h, device=hm.device)
    return torch.stack(((p * xs.view(1, 1, 1, w)).sum((-1, -2)), (p * ys.view(1, 1, h, 1)).sum((-1, -2))), -1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--variant', type=int, default=0, help='1 = force the generic pre-activation kernels')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    peak, how = peak_gbs()
    print('HBM peak %.0f GB/s (%s)' % (peak, how))
    b, c, h, w = args.batch, 16, 64, 64
    n = b * c
    stream = torch.cuda.current_stream().cuda_stream
    target = torch.rand(n, 2, device=dev) * 1.6 - 0.8
    mask = (torch.rand(n, device=dev) > 0.1).float()
    coords = torch.empty(n, 2, device=dev)
    stats = torch.empty(n, 8, device=dev)
    terms = torch.empty(n, 2, device=dev)
    out8 = torch.empty(8, device=dev)
    gl = torch.ones((), device=dev)
    ws = _lib.finish_workspace(dev)

    # ------------------------------------------------------------------ pre-activations
    for dt, tdt, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        z = torch.randn(n, h, w, device=dev).to(tdt)
        dz = torch.empty_like(z)
        nbytes = n * h * w * es
        for preact in ('thresholded_softmax', 'relu', 'sigmoid', 'abs'):
            pid = _lib.PREACT_IDS[preact]
            thr, eps = dp.head.PREACT_DEFAULTS[preact]
            for reg in ('js', 'var'):
                rid = _lib.REG_IDS[reg]

                def fwd(i):
                    _lib.call('dsnt_head_preact_fwd', z.data_ptr(), _lib.dtype_id(z), pid, thr, eps, n, h, w,
                              target.data_ptr(), rid, 2.0 / w, coords.data_ptr(), stats.data_ptr(), terms.data_ptr(),
                              args.variant, stream)

                def bwd(i):
                    _lib.call('dsnt_head_preact_bwd', z.data_ptr(), _lib.dtype_id(z), pid, thr, n, h, w,
                              target.data_ptr(), mask.data_ptr(), stats.data_ptr(), None, None, gl.data_ptr(),
                              out8[3:4].data_ptr(), 1.0, rid, 2.0 / w, 0, dz.data_ptr(), args.variant, stream)
                tf = time_calls(fwd, args.iters)
                _lib.call('dsnt_finish_loss', terms.data_ptr(), mask.data_ptr(), n, 1.0, out8.data_ptr(), ws.data_ptr(),
                          stream)
                tb = time_calls(bwd, args.iters)
                line('preact %-20s %-4s %-3s fwd' % (preact, dt, reg), tf, nbytes, n, peak)
                line('preact %-20s %-4s %-3s bwd' % (preact, dt, reg), tb, 2 * nbytes, n, peak)
                line('preact %-20s %-4s %-3s fwd+bwd' % (preact, dt, reg), tf + tb, 3 * nbytes, n, peak)
        del z, dz
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ flip TTA
    for dt, tdt, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        fb = b // 2
        pair = torch.randn(2 * fb, c, h, w, device=dev).to(tdt)
        perm = torch.tensor(dp.MPII_HFLIP_INDICES, device=dev)
        nn_ = fb * c
        t = time_calls(lambda i: dp.flip_tta_coords(pair), args.iters)
        line('flip-TTA fused head %-4s (2x%d images)' % (dt, fb), t, 2 * nn_ * h * w * es, nn_, peak)
        t = time_calls(lambda i: dp.flip_tta_coords(pair, return_heatmaps=True), args.iters)
        line('flip-TTA fused head + averaged hm out %-4s' % dt, t, 3 * nn_ * h * w * es, nn_, peak)
        if dt == 'f32':
            t = time_calls(lambda i: eager_flip(pair, perm), max(3, args.iters // 4))
            line('flip-TTA torch eager on this GPU %-4s' % dt, t, 2 * nn_ * h * w * es, nn_, peak, '(reference op sequence)')
        if fb >= 1:
            one = pair[:2 * 1].clone() if fb == 1 else torch.cat([pair[:1], pair[fb:fb + 1]], 0).contiguous()
            t = time_calls(lambda i: dp.flip_tta_coords(one), 200)
            print('%-44s %9.1f us per call (batch 1 = the reference\'s own setting, launch-bound)' % (
                'flip-TTA fused head %-4s one image' % dt, t * 1e3))
            if dt == 'f32':
                t = time_calls(lambda i: eager_flip(one, perm), 100)
                print('%-44s %9.1f us per call' % ('flip-TTA torch eager on this GPU, one image', t * 1e3))
        del pair
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------ gauss helpers
    cn = (torch.rand(b, c, 2, device=dev) * 2.2 - 1.1)
    t = time_calls(lambda i: du.encode_heatmaps(cn, w, h, 1), args.iters)
    line('encode_heatmaps f32 (write only)', t, n * h * w * 4, n, peak)
    for dt, tdt, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        hm = torch.randn(b, c, h, w, device=dev).to(tdt)
        t = time_calls(lambda i: du.decode_heatmaps(hm), args.iters)
        line('decode_heatmaps %-4s (read only)' % dt, t, n * h * w * es, n, peak)
        if dt == 'f32':
            def eager_decode(i):
                flat = hm.view(b, c, -1)
                mv, idx = flat.max(2)
                return torch.stack((idx % w, idx // h), -1).float() * (mv > 0).unsqueeze(-1)
            t = time_calls(eager_decode, args.iters)
            line('decode torch eager arg-max only, on this GPU', t, n * h * w * es, n, peak, '(no neighbour offset)')
        del hm


if __name__ == '__main__':
    main()
