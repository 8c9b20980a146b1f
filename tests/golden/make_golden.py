#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

It imports `dsnt.nn` and `dsnt.model` from /root/reference/src (with a 5-name stub for the
absent third-party `torchdata.mpii`, which only the dataset code uses), evaluates the head in
float64 on CPU exactly the way `tests/common.py:18` makes the reference's own tests run, and
writes small .npz fixtures next to this file.  The fixtures -- not the reference -- travel to
the GPU box; `tests/test_oracle_golden.py` pins the oracle to them and the `-m gpu` tests pin
the CUDA kernels to them.
"""

import os
import sys
import types
import warnings

import numpy as np
import torch

REF_SRC = '/root/reference/src'
OUT_DIR = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    stub = types.ModuleType('torchdata.mpii')
    for name in ('MpiiData', 'MPII_Joint_Horizontal_Flips', 'MPII_Image_Mean', 'MPII_Image_Stddev',
                 'transform_keypoints'):
        setattr(stub, name, None)
    stub.MPII_Joint_Horizontal_Flips = list(range(16))
    pkg = types.ModuleType('torchdata')
    pkg.mpii = stub
    sys.modules['torchdata'] = pkg
    sys.modules['torchdata.mpii'] = stub
    sys.path.insert(0, REF_SRC)
    import dsnt.nn as ref_nn
    try:
        import dsnt.model as ref_model
    except Exception as exc:                                     # pragma: no cover
        print('dsnt.model not importable (%r); composing the head from dsnt.nn only' % (exc,))
        ref_model = None
    return ref_nn, ref_model


def ref_head(ref_nn, ref_model, z, target, mask, reg, hm_sigma, reg_coeff):
    """forward_part2 + forward_loss of ResNetHumanPoseModel (src/dsnt/model.py:138-145,176-183)."""
    if ref_model is not None:
        hpm = ref_model.HumanPoseModel
        p = hpm._hm_preact(None, z, 'softmax')
        coords = ref_nn.dsnt(p)
        euc = ref_nn.euclidean_loss(coords, target, mask)
        rv = hpm._calculate_reg_loss(None, target, mask, reg, p, hm_sigma)
    else:
        h, w = z.shape[-2:]
        p = torch.nn.functional.softmax(z.view(-1, h * w), dim=-1).view(-1, z.shape[-3], h, w)
        coords = ref_nn.dsnt(p)
        euc = ref_nn.euclidean_loss(coords, target, mask)
        sigma = 2.0 * hm_sigma / w
        fn = {'var': ref_nn.variance_reg_loss, 'kl': ref_nn.kl_reg_loss, 'js': ref_nn.js_reg_loss,
              'mse': ref_nn.mse_reg_loss}.get(reg)
        rv = fn(p, target, sigma, mask) if fn else 0
    return euc + reg_coeff * rv, coords, euc, rv, p


def trained_like_logits(ref_nn, b, c, h, w, gen):
    """Peaked 'trained network' logits: log of a 1-px Gaussian plus noise."""
    mu = torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8
    g = ref_nn.make_gauss(mu, w, h, 2.0 / w)
    return (g + 1e-6).log() + 0.1 * torch.randn(b, c, h, w, generator=gen)


def main():
    torch.set_default_dtype(torch.float64)          # tests/common.py:18
    warnings.simplefilter('ignore')
    ref_nn, ref_model = import_reference()
    gen = torch.Generator().manual_seed(0)          # tests/common.py:20
    regs = ['none', 'var', 'kl', 'js', 'mse']

    # ------------------------------------------------------------------ fused head on logits
    head_cases = [
        # name, (B, C, H, W), scale ('trained' = peaked), with mask?, hm_sigma, reg_coeff
        ('s5x5',       (2, 3, 5, 5),   1.0, True,  1.0, 1.0),
        ('s5x5_nomask', (2, 3, 5, 5),  1.0, False, 1.0, 1.0),
        ('s7x7_x5',    (2, 2, 7, 7),   5.0, True,  1.0, 0.5),
        ('s6x10',      (1, 3, 6, 10),  1.0, True,  1.5, 1.0),
        ('s14x14',     (2, 2, 14, 14), 1.0, False, 1.0, 2.0),
        ('s28x28',     (2, 2, 28, 28), 1.0, True,  1.0, 1.0),
        ('s56x56_tr',  (1, 2, 56, 56), 'trained', True, 1.0, 1.0),
        ('s64x64',     (2, 2, 64, 64), 1.0, True,  1.0, 1.0),
        ('s64x64_tr',  (1, 2, 64, 64), 'trained', False, 1.0, 1.0),
    ]
    out = {}
    meta = []
    for name, (b, c, h, w), scale, with_mask, hm_sigma, coeff in head_cases:
        if scale == 'trained':
            z32 = trained_like_logits(ref_nn, b, c, h, w, gen).float()
        else:
            z32 = (torch.randn(b, c, h, w, generator=gen) * scale).float()
        target = (torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8).float()
        mask = (torch.rand(b, c, generator=gen) > 0.25).float() if with_mask else None
        if mask is not None and mask.sum() == 0:
            mask[0, 0] = 1
        out[name + '/z'] = z32.numpy()
        out[name + '/target'] = target.numpy()
        if mask is not None:
            out[name + '/mask'] = mask.numpy()
        big = h * w >= 28 * 28
        for reg in regs:
            z = z32.double().clone().requires_grad_(True)
            loss, coords, euc, rv, _ = ref_head(ref_nn, ref_model, z, target.double(),
                                                None if mask is None else mask.double(),
                                                reg, hm_sigma, coeff)
            loss.backward()
            out['%s/%s/loss' % (name, reg)] = np.float64(loss.item())
            out['%s/%s/euclid' % (name, reg)] = np.float64(euc.item())
            out['%s/%s/reg' % (name, reg)] = np.float64(float(rv))
            dz = z.grad.numpy()
            out['%s/%s/dz' % (name, reg)] = dz.astype(np.float32) if big else dz
            if reg == 'none':
                out[name + '/coords'] = coords.detach().numpy()
        meta.append((name, b, c, h, w, hm_sigma, coeff, int(with_mask)))
    out['__cases__'] = np.array([m[0] for m in meta])
    out['__params__'] = np.array([m[1:] for m in meta], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT_DIR, 'head_logits.npz'), **out)

    # ------------------------------------------------------------------ level-1 API on heatmaps P
    out = {}
    p_cases = [('p5x5', (2, 2, 5, 5), True), ('p4x9', (3, 4, 9), True), ('p8x8_unnorm', (1, 3, 8, 8), False),
               ('p16x16', (2, 2, 16, 16), True)]
    names = []
    for name, shape, normalised in p_cases:
        h, w = shape[-2:]
        raw = torch.randn(*shape, generator=gen)
        if normalised:
            p0 = torch.softmax(raw.reshape(-1, h * w), dim=-1).view(shape)
        else:
            p0 = raw.abs() * 0.05
            p0[..., 0, 0] = 0.0                       # exact zeros exercise the epsilon handling
        p32 = p0.float()
        mu = (torch.rand(*shape[:-2], 2, generator=gen) * 1.6 - 0.8).float()
        mask = (torch.rand(*shape[:-2], generator=gen) > 0.3).float()
        if mask.sum() == 0:
            mask.view(-1)[0] = 1
        sigma = 0.3
        out[name + '/p'] = p32.numpy()
        out[name + '/mu'] = mu.numpy()
        out[name + '/mask'] = mask.numpy()
        out[name + '/sigma'] = np.float64(sigma)
        # dsnt forward + backward with a random upstream gradient
        p = p32.double().clone().requires_grad_(True)
        coords = ref_nn.dsnt(p)
        gc = torch.randn(coords.shape, generator=gen)
        coords.backward(gc)
        out[name + '/dsnt/coords'] = coords.detach().numpy()
        out[name + '/dsnt/g_coords'] = gc.numpy()
        out[name + '/dsnt/dp'] = p.grad.numpy()
        for reg, fn in (('var', ref_nn.variance_reg_loss), ('kl', ref_nn.kl_reg_loss),
                        ('js', ref_nn.js_reg_loss), ('mse', ref_nn.mse_reg_loss)):
            for mtag, mm in (('mask', mask.double()), ('nomask', None)):
                p = p32.double().clone().requires_grad_(True)
                val = fn(p, mu.double(), sigma, mm)
                val.backward()
                out['%s/%s/%s/loss' % (name, reg, mtag)] = np.float64(val.item())
                out['%s/%s/%s/dp' % (name, reg, mtag)] = p.grad.numpy()
        names.append(name)
    out['__cases__'] = np.array(names)

    # euclidean loss (any trailing dimension d) + gradient
    for tag, shape in (('e2', (4, 16, 2)), ('e3', (5, 3)), ('e_single', (2,))):
        actual32 = torch.randn(*shape, generator=gen).float()
        target32 = torch.randn(*shape, generator=gen).float()
        mask = (torch.rand(*shape[:-1], generator=gen) > 0.3).float() if len(shape) > 1 else None
        a = actual32.double().clone().requires_grad_(True)
        for mtag, mm in (('mask', None if mask is None else mask.double()), ('nomask', None)):
            a.grad = None
            val = ref_nn.euclidean_loss(a, target32.double(), mm)
            val.backward()
            out['euclid/%s/%s/loss' % (tag, mtag)] = np.float64(val.item())
            out['euclid/%s/%s/grad' % (tag, mtag)] = a.grad.numpy().copy()
        out['euclid/%s/actual' % tag] = actual32.numpy()
        out['euclid/%s/target' % tag] = target32.numpy()
        if mask is not None:
            out['euclid/%s/mask' % tag] = mask.numpy()

    # softmax_2d and thresholded softmax fwd/bwd
    z32 = (torch.randn(2, 3, 6, 7, generator=gen) * 2).float()
    z = z32.double().clone().requires_grad_(True)
    sm = ref_nn.softmax_2d(z)
    gs = torch.randn(sm.shape, generator=gen)
    sm.backward(gs)
    out['softmax2d/z'] = z32.numpy()
    out['softmax2d/out'] = sm.detach().numpy()
    out['softmax2d/g'] = gs.numpy()
    out['softmax2d/dz'] = z.grad.numpy()
    x32 = torch.randn(5, 37, generator=gen).float()
    for tag, thr in (('thr0', 0.0), ('thrm05', -0.5), ('thrinf', float('-inf'))):
        x = x32.double().clone().requires_grad_(True)
        ts = ref_nn.thresholded_softmax(x, thr)
        gt = torch.randn(ts.shape, generator=gen)
        ts.backward(gt)
        out['tsoftmax/%s/out' % tag] = ts.detach().numpy()
        out['tsoftmax/%s/g' % tag] = gt.numpy()
        out['tsoftmax/%s/dx' % tag] = x.grad.numpy()
        out['tsoftmax/%s/thr' % tag] = np.float64(thr)
    out['tsoftmax/x'] = x32.numpy()

    # make_gauss forward and gradient wrt coords
    mu32 = (torch.rand(2, 3, 2, generator=gen) * 1.6 - 0.8).float()
    mu = mu32.double().clone().requires_grad_(True)
    g = ref_nn.make_gauss(mu, 9, 6, 0.25)            # width 9, height 6 -> [2,3,6,9]
    gg = torch.randn(g.shape, generator=gen)
    g.backward(gg)
    out['gauss/mu'] = mu32.numpy()
    out['gauss/out'] = g.detach().numpy()
    out['gauss/g'] = gg.numpy()
    out['gauss/dmu'] = mu.grad.numpy()
    np.savez_compressed(os.path.join(OUT_DIR, 'level1_api.npz'), **out)

    # ------------------------------------------------------------------ hourglass-style stacked loss
    out = {}
    zs32 = [(torch.randn(2, 2, 16, 16, generator=gen)).float() for _ in range(3)]
    target = (torch.rand(2, 2, 2, generator=gen) * 1.6 - 0.8).float()
    mask = torch.tensor([[1.0, 0.0], [1.0, 1.0]])
    zs = [z.double().clone().requires_grad_(True) for z in zs32]
    total = 0
    for z in zs:                                       # src/dsnt/model.py:238-246
        loss, _, _, _, _ = ref_head(ref_nn, ref_model, z, target.double(), mask.double(), 'js', 1.0, 1.0)
        total = total + loss
    total.backward()
    for i, z in enumerate(zs):
        out['z%d' % i] = zs32[i].numpy()
        out['dz%d' % i] = z.grad.numpy()
    out['target'] = target.numpy()
    out['mask'] = mask.numpy()
    out['loss'] = np.float64(total.item())
    np.savez_compressed(os.path.join(OUT_DIR, 'stacked_js.npz'), **out)

    for f in ('head_logits.npz', 'level1_api.npz', 'stacked_js.npz'):
        print(f, os.path.getsize(os.path.join(OUT_DIR, f)), 'bytes')
    print('reference model layer used:', ref_model is not None)


if __name__ == '__main__':
    main()
