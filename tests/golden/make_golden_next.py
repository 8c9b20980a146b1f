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


def make_gauss_util(gen):
    """src/dsnt/util.py:70-198 through the reference's own functions."""
    import dsnt.util as ref_util
    torch.set_default_dtype(torch.float32)      # these helpers work on FloatTensors (util.py:140)
    # torch 0.3.1 (the reference's pin) returns a Python float when a FloatTensor is indexed down to one element, so
    # `round(coords[i, j, 0])` (util.py:142-143) is Python's round-half-to-even on a float.  torch 2.x returns a 0-dim
    # tensor, which has no __round__: give it the pinned version's meaning.  The reference source is untouched.
    torch.Tensor.__round__ = lambda self, ndigits=None: round(self.item())
    out = {}
    # encode_heatmaps: random coords incl. out-of-frame joints, half-pixel ties, corners
    for name, (b, c, h, w), sigma in (('e5x5', (1, 1, 5, 5), 1), ('e64', (2, 16, 64, 64), 1), ('e28', (3, 16, 28, 28), 2),
                                      ('e7x12', (2, 5, 7, 12), 1.5)):
        coords = (torch.rand(b, c, 2, generator=gen) * 2.6 - 1.3).float()
        coords[0, 0] = torch.tensor([-0.8, 0.8])                                  # tests/test_util.py:39
        if c > 4:
            coords[0, 1] = torch.tensor([-1.0 + 2.0 / w, 1.0 - 1.0 / h])        # lands exactly on x.5 / pixel centre
            coords[0, 2] = torch.tensor([-1.0, -1.0])                            # corner of the frame
            coords[0, 3] = torch.tensor([1.5, 0.0])                              # far outside: nothing drawn
            coords[0, 4] = torch.tensor([1.0 + 5.0 / w, 0.2])                    # outside but within the clip radius
        out[name + '/coords'] = coords.numpy().copy()
        out[name + '/sigma'] = np.float64(sigma)
        out[name + '/hm'] = ref_util.encode_heatmaps(coords.clone(), w, h, sigma).numpy()
    # draw_gaussian: unclipped + normalised + clipped
    for name, (h, w), (x, y), sigma, normalize, clip in (('d9', (9, 9), (4, 4), 1, False, None),
                                                         ('d5clip', (5, 5), (0, 4), 1, False, 7),
                                                         ('d12norm', (10, 12), (7.9, 2.2), 1.7, True, None),
                                                         ('d12clipnorm', (10, 12), (11, 9), 2.0, True, 5),
                                                         ('d_out', (6, 6), (-9, 2), 1, False, 7)):
        img = torch.zeros(1, h, w).float()
        ref_util.draw_gaussian(img, x, y, sigma, normalize=normalize, clip_size=clip)
        out[name + '/img'] = img.numpy()
        out[name + '/args'] = np.array([x, y, sigma, float(normalize), -1.0 if clip is None else clip], dtype=np.float64)
    # decode_heatmaps
    for name, (b, c, h, w) in (('g64', (2, 4, 64, 64)), ('g28', (2, 8, 28, 28)), ('g6x9', (2, 4, 6, 9)),
                               ('g9x6', (2, 4, 9, 6)), ('g2x2', (1, 3, 2, 2))):
        hm = torch.randn(b, c, h, w, generator=gen).float()
        hm[0, 0] = -hm[0, 0].abs() - 0.1                      # max <= 0 -> coords (0, 0) before normalisation
        hm[0, 1] = 0.0                                        # all equal: first index wins, max == 0 -> (0, 0)
        if h > 3 and w > 3:
            hm[0, 2, 2, 2] = 9.0                              # interior peak with EQUAL neighbours: sign(0) = 0
            hm[0, 2, 2, 1] = hm[0, 2, 2, 3] = 1.0
            hm[0, 3, 0, w - 1] = 9.0                          # peak on the border: no neighbour offset
            hm[1, 0, 1, 1] = hm[1, 0, h - 2, w - 2] = 7.0     # tie: the first maximum wins
        out[name + '/hm'] = hm.numpy()
        out[name + '/coords_nb'] = ref_util.decode_heatmaps(hm.clone(), use_neighbours=True).numpy()
        out[name + '/coords'] = ref_util.decode_heatmaps(hm.clone(), use_neighbours=False).numpy()
    # round trip of the two helpers
    coords = (torch.rand(2, 16, 2, generator=gen) * 1.8 - 0.9).float()
    out['rt/coords'] = coords.numpy().copy()
    out['rt/decoded'] = ref_util.decode_heatmaps(ref_util.encode_heatmaps(coords.clone(), 64, 64, 1)).numpy()
    torch.set_default_dtype(torch.float64)
    np.savez_compressed(os.path.join(OUT_DIR, 'gauss_util.npz'), **out)
    print('gauss_util.npz: %d arrays' % len(out))


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
