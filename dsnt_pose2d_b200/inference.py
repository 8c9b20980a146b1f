"""Inference path of the reference in front of the DSNT head.

`src/dsnt/inference.py:33-48` evaluates every image together with its mirror image, un-mirrors the second set of raw
heatmaps (reverse the last dim, permute left/right joints with `MPIIDataset.HFLIP_INDICES`, `src/dsnt/data.py:97`),
averages the two sets and only then applies `forward_part2` (softmax + dsnt).  Here those four full-size passes and the
head are ONE launch (`dsnt_flip_tta_fwd`, include/dsnt_b200.h): both heatmap sets are read once, nothing is written
but the coordinates (and, on request, the averaged heatmaps the reference calls `hm`).
"""

import torch

from . import _lib
from .head import PREACT_DEFAULTS

# MPII joint order: 0-5 right ankle, knee, hip, left hip, knee, ankle; 6 pelvis, 7 thorax, 8 upper neck, 9 head top;
# 10-15 right wrist, elbow, shoulder, left shoulder, elbow, wrist (the order src/dsnt/util.py:15-31 draws bones in).
# This is what torchdata.mpii.MPII_Joint_Horizontal_Flips holds (src/dsnt/data.py:15,97).
MPII_HFLIP_INDICES = (5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10)

_perm_cache = {}


def _perm_tensor(indices, n_chans, device):
    if indices is None:
        return None
    if torch.is_tensor(indices):
        indices = tuple(int(i) for i in indices.tolist())
    else:
        indices = tuple(int(i) for i in indices)
    if len(indices) != n_chans or sorted(indices) != list(range(n_chans)):
        raise ValueError('hflip_indices must be a permutation of range(%d), got %r' % (n_chans, indices))
    key = (indices, device)
    t = _perm_cache.get(key)
    if t is None:
        t = torch.tensor(indices, dtype=torch.int32, device=device)
        _perm_cache[key] = t
    return t


def flip_tta_coords(hm_pair, hflip_indices=MPII_HFLIP_INDICES, preact='softmax', return_heatmaps=False,
                    threshold=None, eps=None):
    """Coordinates from the raw heatmaps of [images, mirrored images] (src/dsnt/inference.py:36-48).

    Args:
        hm_pair: [2B, C, H, W] raw heatmaps (CUDA, float32/bfloat16), `model.forward_part1(cat([x, flip(x)]))`; for a
            stacked hourglass pass the LAST stack, as the reference does (inference.py:40-42).
        hflip_indices: joint permutation under a horizontal flip (length C), None = identity.
        preact: the model's heatmap pre-activation (src/dsnt/model.py:24-45).
        return_heatmaps: also return the averaged raw heatmaps `hm = (hm1 + unflipped hm2)/2`, [B, C, H, W].
    Returns:
        coords [B, C, 2] float32 on the device (`.cpu()` it for `compute_coords` semantics), or (coords, hm).
    """
    _lib.require_cuda(hm_pair, 'hm_pair')
    if hm_pair.dim() != 4 or hm_pair.size(0) % 2 != 0:
        raise ValueError('hm_pair must be [2B, C, H, W], got shape %s' % (tuple(hm_pair.shape),))
    if preact not in _lib.PREACT_IDS:
        raise Exception('unrecognised heatmap preactivation function: {}'.format(preact))
    z = hm_pair.detach().contiguous()
    b, c, h, w = z.size(0) // 2, z.size(1), z.size(2), z.size(3)
    d_thr, d_eps = PREACT_DEFAULTS[preact]
    dev = z.device
    with torch.cuda.device(dev):
        perm = _perm_tensor(hflip_indices, c, dev)
        coords = torch.empty(b, c, 2, dtype=torch.float32, device=dev)
        avg = torch.empty(b, c, h, w, dtype=z.dtype, device=dev) if return_heatmaps else None
        _lib.call('dsnt_flip_tta_fwd', z.data_ptr(), _lib.dtype_id(z), b, c, h, w, _lib.ptr(perm),
                  _lib.PREACT_IDS[preact], float(d_thr if threshold is None else threshold),
                  float(d_eps if eps is None else eps), coords.data_ptr(), _lib.ptr(avg), _lib.stream_of(z))
    return (coords, avg) if return_heatmaps else coords


def predict_flipped(model, images, hflip_indices=MPII_HFLIP_INDICES):
    """One batch of `generate_predictions(..., use_flipped=True)` (src/dsnt/inference.py:33-48) for a 'dsnt' model:
    backbone on [images, mirrored images] (stock cuDNN), then the fused un-mirror + average + head.
    Returns normalised coordinates [B, C, 2] as a float32 CPU tensor (`compute_coords`, src/dsnt/model.py:161-163);
    unlike the reference any batch size works, not just 1."""
    with torch.no_grad():
        pair = torch.cat([images, images.flip(-1)], 0)
        hm = model.forward_part1(pair)
        if isinstance(hm, (list, tuple)):
            hm = hm[-1]                       # just the last heatmap of a stacked hourglass (inference.py:40-42)
        coords = flip_tta_coords(hm, hflip_indices, preact=getattr(model, 'preact', 'softmax'))
    return coords.to('cpu', torch.float32)
