"""CPU restatement of the reference's 'gauss' output-strategy helpers (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows src/dsnt/util.py:70-198 line by line (same float32 op order, same Python rounding, same quirks); the loops
are pure Python like the reference's, so use it on small inputs.  Pinned by tests/test_oracle_gauss_util.py to the
known answers of the reference's tests/test_util.py and to golden vectors produced by the unmodified reference.
"""

import math

import numpy as np
import torch


def draw_gaussian(img_tensor, x, y, sigma, normalize=False, clip_size=None):
    """src/dsnt/util.py:70-126 -- draws into `img_tensor` ([H,W] or [1,H,W]) in place."""
    x = int(x)                                                        # :84-85 (truncation, not rounding)
    y = int(y)
    if img_tensor.dim() == 2:                                         # :87-94
        height, width = list(img_tensor.size())
    elif img_tensor.dim() == 3:
        n_chans, height, width = list(img_tensor.size())
        assert n_chans == 1, 'expected img_tensor to have one channel'
        img_tensor = img_tensor[0]
    else:
        raise Exception('expected img_tensor to have 2 or 3 dimensions')
    radius = max(width, height)                                       # :96-98
    if clip_size is not None:
        radius = clip_size / 2
    if radius < 0.5 or x <= -radius or y <= -radius or \
            x >= (width - 1) + radius or y >= (height - 1) + radius:  # :100-102
        return
    start_x = max(0, math.ceil(x - radius))                           # :104-109
    end_x = min(width, int(x + radius + 1))
    start_y = max(0, math.ceil(y - radius))
    end_y = min(height, int(y + radius + 1))
    w = end_x - start_x
    h = end_y - start_y
    subimg = img_tensor[start_y:end_y, start_x:end_x]                 # :111
    xs = torch.arange(start_x, end_x).type_as(img_tensor).view(1, w).expand_as(subimg)   # :113-114
    ys = torch.arange(start_y, end_y).type_as(img_tensor).view(h, 1).expand_as(subimg)
    k = -0.5 * (1 / sigma) ** 2                                       # :116
    subimg.copy_((xs - x) ** 2)                                       # :117-120
    subimg.add_((ys - y) ** 2)
    subimg.mul_(k)
    subimg.exp_()
    if normalize:                                                     # :122-125
        val_sum = subimg.sum()
        if val_sum > 0:
            subimg.div_(val_sum)


def encode_heatmaps(coords, width, height, sigma=1):
    """src/dsnt/util.py:129-148 -- normalised coords [B,C,2] -> float32 heatmaps [B,C,H,W] (7x7 clipped Gaussians).
    The reference converts `coords` to pixel units IN PLACE (:133-136); this restatement works on a copy."""
    coords = coords.clone().float()
    coords.add_(1)                                                    # :133-136
    coords[:, :, 0].mul_(width / 2)
    coords[:, :, 1].mul_(height / 2)
    coords.add_(-0.5)
    batch_size = coords.size(0)
    n_chans = coords.size(1)
    target = torch.zeros(batch_size, n_chans, height, width, dtype=torch.float32)   # :140
    for i in range(batch_size):
        for j in range(n_chans):
            x = round(coords[i, j, 0].item())                         # :143-144 (Python round: half to even)
            y = round(coords[i, j, 1].item())
            draw_gaussian(target[i, j], x, y, sigma, normalize=False, clip_size=7)   # :145
    return target


def get_preds(heatmaps):
    """src/dsnt/util.py:151-170 -- argmax pixel per heatmap; note y = idx / HEIGHT (:163), kept as is."""
    batch_size, n_chans, height, width = list(heatmaps.size())
    maxval, idx = torch.max(heatmaps.reshape(batch_size, n_chans, -1), 2)           # :154
    maxval = maxval.view(batch_size, n_chans, 1)
    idx = idx.view(batch_size, n_chans, 1)
    coords = idx.repeat(1, 1, 2)                                                     # :159
    coords[:, :, 0] = coords[:, :, 0] % width                                        # :161
    coords[:, :, 1] = torch.div(coords[:, :, 1], height, rounding_mode='floor')      # :162 (integer division in torch 0.3)
    coords = coords.float()
    pred_mask = maxval.gt(0).repeat(1, 1, 2).float()                                 # :166-168 (max <= 0 -> (0, 0))
    coords = coords * pred_mask
    return coords


def decode_heatmaps(heatmaps, use_neighbours=True):
    """src/dsnt/util.py:173-198 -- heatmaps [B,C,H,W] -> normalised coords [B,C,2] float32."""
    coords = get_preds(heatmaps)
    _, _, height, width = list(heatmaps.size())
    if use_neighbours:                                                               # :180-192
        for i, joint_coords in enumerate(coords):
            for j, (x, y) in enumerate(joint_coords):
                x = int(x)
                y = int(y)
                if x > 0 and x < width - 1 and y > 0 and y < height - 1:
                    hm = heatmaps[i, j]
                    joint_coords[j, 0] += (0.25 * np.sign(float(hm[y, x + 1] - hm[y, x - 1])))
                    joint_coords[j, 1] += (0.25 * np.sign(float(hm[y + 1, x] - hm[y - 1, x])))
    coords.add_(0.5)                                                                 # :195-198
    coords[:, :, 0].mul_(2 / width)
    coords[:, :, 1].mul_(2 / height)
    coords.add_(-1)
    return coords
