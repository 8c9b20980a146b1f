#!/usr/bin/env python
"""Golden vectors for the SURVEY.md 8(f) rows, produced by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_next.py

Writes, next to this file:
    preact_heads.npz   head fwd+bwd for every `--preact` choice (src/dsnt/model.py:24-45) x every regulariser
    flip_tta.npz       flip test-time augmentation of raw heatmaps (src/dsnt/inference.py:36-48)
    gauss_util.npz     the 'gauss' output strategy helpers (src/dsnt/util.py:70-198)

`make_golden.py` (the fixtures of the main path) is left alone so its random stream does not move.
"""

import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, trained_like_logits, OUT_DIR  # noqa: E402

PREACTS = ['softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid']
REGS = ['none', 'var', 'kl', 'js', 'mse']


def ref_head_preact(ref_nn, ref_model, z, target, mask, preact, reg, hm_sigma, reg_coeff):
    """forward_part2 + forward_loss of ResNetHumanPoseModel (src/dsnt/model.py:138-145,176-183)."""
    hpm = ref_model.HumanPoseModel
    p = hpm._hm_preact(None, z, preact)
    coords = ref_nn.dsnt(p)
    euc = ref_nn.euclidean_loss(coords, target, mask)
    rv = hpm._calculate_reg_loss(None, target, mask, reg, p, hm_sigma)
    return euc + reg_coeff * rv, coords, euc, rv


def make_preact(ref_nn, ref_model, gen):
    cases = [
        # name, (B, C, H, W), scale, mask?, hm_sigma, reg_coeff
        ('a5x5',      (2, 3, 5, 5),   1.0, True,  1.0, 1.0),
        ('a7x9',      (1, 3, 7, 9),   2.0, False, 1.5, 0.5),
        ('a28x28',    (2, 2, 28, 28), 1.0, True,  1.0, 1.0),
        ('a64x64',    (1, 2, 64, 64), 1.0, True,  1.0, 1.0),
        ('a64x64_tr', (1, 1, 64, 64), 'trained', False, 1.0, 2.0),
    ]
    out, meta = {}, []
    for name, (b, c, h, w), scale, with_mask, hm_sigma, coeff in cases:
        if scale == 'trained':
            z32 = (trained_like_logits(ref_nn, b, c, h, w, gen) + 8.0).float()   # positive peak, negative tails
        else:
            z32 = (torch.randn(b, c, h, w, generator=gen) * scale).float()
        z32[0, 0, 0, 0] = 0.0                                # abs'(0) = relu'(0) = 0 (torch's subgradient choice)
        target = (torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8).float()
        mask = (torch.rand(b, c, generator=gen) > 0.25).float() if with_mask else None
        if mask is not None and mask.sum() == 0:
            mask[0, 0] = 1
        out[name + '/z'] = z32.numpy()
        out[name + '/target'] = target.numpy()
        if mask is not None:
            out[name + '/mask'] = mask.numpy()
        big = h * w >= 28 * 28
        for preact in PREACTS:
            for reg in REGS:
                z = z32.double().clone().requires_grad_(True)
                loss, coords, euc, rv = ref_head_preact(ref_nn, ref_model, z, target.double(),
                                                        None if mask is None else mask.double(),
                                                        preact, reg, hm_sigma, coeff)
                loss.backward()
                key = '%s/%s/%s' % (name, preact, reg)
                out[key + '/loss'] = np.float64(loss.item())
                out[key + '/euclid'] = np.float64(euc.item())
                out[key + '/reg'] = np.float64(float(rv))
                dz = z.grad.numpy()
                out[key + '/dz'] = dz.astype(np.float32) if big else dz
                if reg == 'none':
                    out['%s/%s/coords' % (name, preact)] = coords.detach().numpy()
        meta.append((name, b, c, h, w, hm_sigma, coeff, int(with_mask)))
    # a heatmap that is entirely below the threshold / non-positive: P = 0 everywhere, loss finite, gradient 0
    z32 = -(torch.rand(1, 2, 6, 6, generator=gen) + 1.0).float()
    target = torch.tensor([[[0.3, -0.2], [-0.5, 0.1]]])   # away from coords = (0, 0): sqrt'(0) would give NaN
    out['dead/z'] = z32.numpy()
    out['dead/target'] = target.numpy()
    for preact in ('thresholded_softmax', 'relu'):
        for reg in REGS:
            z = z32.double().clone().requires_grad_(True)
            loss, coords, euc, rv = ref_head_preact(ref_nn, ref_model, z, target.double(), None, preact, reg, 1.0, 1.0)
            loss.backward()
            key = 'dead/%s/%s' % (preact, reg)
            out[key + '/loss'] = np.float64(loss.item())
            out[key + '/dz'] = z.grad.numpy()
            out[key + '/coords'] = coords.detach().numpy()
    out['__cases__'] = np.array([m[0] for m in meta])
    out['__params__'] = np.array([m[1:] for m in meta], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT_DIR, 'preact_heads.npz'), **out)
    print('preact_heads.npz: %d arrays' % len(out))


MPII_FLIPS = [5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10]   # torchdata.mpii.MPII_Joint_Horizontal_Flips


def make_flip(ref_nn, ref_model, gen):
    """src/dsnt/inference.py:36-48 with the reference's own reverse_tensor / type_as_index / _hm_preact / dsnt
    (inference.py itself needs progressbar + tele, absent here; these are its lines 43-48 verbatim in effect)."""
    import dsnt.util as ref_util
    hpm = ref_model.HumanPoseModel
    out = {}
    names = []
    for name, (c, h, w), flips in (('f16x32', (16, 32, 32), MPII_FLIPS), ('f16x28', (16, 28, 28), MPII_FLIPS), ('f2x64', (2, 64, 64), [1, 0]),
                                   ('f3x7x9', (3, 7, 9), [2, 1, 0]), ('f4x6x10', (4, 6, 10), [1, 0, 3, 2])):
        hm_var = torch.randn(2, c, h, w, generator=gen).float().double() * 2     # one image + its mirror image
        hm1, hm2 = hm_var.split(1)                                                 # :43
        hm2 = ref_util.reverse_tensor(hm2, -1)                                     # :44
        hm2 = hm2.index_select(-3, ref_util.type_as_index(torch.LongTensor(flips), hm2))   # :45
        hm = (hm1 + hm2) / 2                                                       # :46
        out[name + '/hm_pair'] = hm_var.float().numpy()
        out[name + '/flips'] = np.array(flips, dtype=np.int64)
        out[name + '/hm'] = hm.numpy()
        for preact in PREACTS:
            coords = ref_nn.dsnt(hpm._hm_preact(None, hm, preact))                 # :47 forward_part2
            out['%s/%s/coords' % (name, preact)] = coords.numpy()
        names.append(name)
    out['__cases__'] = np.array(names)
    np.savez_compressed(os.path.join(OUT_DIR, 'flip_tta.npz'), **out)
    print('flip_tta.npz: %d arrays' % len(out))


def main():
    torch.set_default_dtype(torch.float64)          # tests/common.py:18
    warnings.simplefilter('ignore')
    ref_nn, ref_model = import_reference()
    assert ref_model is not None, 'dsnt.model must import (torchdata.mpii stub)'
    which = sys.argv[1:] or ['preact', 'flip', 'gauss']
    if 'preact' in which:
        make_preact(ref_nn, ref_model, torch.Generator().manual_seed(1))
    if 'flip' in which:
        make_flip(ref_nn, ref_model, torch.Generator().manual_seed(2))
    if 'gauss' in which:
        make_gauss_util(torch.Generator().manual_seed(3))


if __name__ == '__main__':
    main()
