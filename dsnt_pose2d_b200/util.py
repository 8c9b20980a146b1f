"""The 'gauss' output-strategy helpers of the reference (`src/dsnt/util.py:70-198`) on the GPU.

Same names and argument meaning as `dsnt.util`:

    draw_gaussian(img_tensor, x, y, sigma, normalize=False, clip_size=None)     util.py:70-126
    encode_heatmaps(coords, width, height, sigma=1)                             util.py:129-148
    get_preds(heatmaps) / decode_heatmaps(heatmaps, use_neighbours=True)        util.py:151-198

The reference runs these on the CPU inside the training step (a Python double loop over (sample, joint), then an H2D
copy of the whole target tensor -- `model.py:148-154,247-256` -- and a D2H copy of the whole heatmap tensor before
`decode_heatmaps`, `model.py:165,269`).  Here each is one kernel launch on the device; CPU tensors are accepted for
drop-in use (the small coordinate tensor is uploaded, results of decode go back to where the input lived), but the
arithmetic always runs on the GPU -- there is no CPU implementation in this package.

Deliberate deviation: `encode_heatmaps` does not convert its `coords` argument to pixel units in place as
util.py:133-136 does (a side effect that corrupts the caller's targets when they already live on the CPU).
"""

import torch

from . import _lib


def _device_of(t):
    return t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())


def _draw(centres, centres_are_pixels, n, width, height, sigma, clip_size, normalize, dev):
    out = torch.empty(n, height, width, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call('dsnt_draw_gaussians', centres.data_ptr(), int(centres_are_pixels), n, width, height, float(sigma),
                  float(-1.0 if clip_size is None else clip_size), int(bool(normalize)), out.data_ptr(),
                  _lib.stream_of(out))
    return out


def encode_heatmaps(coords, width, height, sigma=1):
    """Convert normalised coordinates [B, C, 2] into float32 heatmaps [B, C, H, W] on the GPU (util.py:129-148):
    an unnormalised Gaussian clipped to 7x7 around the nearest pixel of every joint."""
    if coords.dim() != 3 or coords.size(-1) != 2:
        raise ValueError('coords must be [B, C, 2], got shape %s' % (tuple(coords.shape),))
    dev = _device_of(coords)
    c = coords.detach().to(dev, torch.float32).contiguous()
    b, n_chans = c.size(0), c.size(1)
    return _draw(c, False, b * n_chans, width, height, sigma, 7, False, dev).view(b, n_chans, height, width)


def draw_gaussian(img_tensor, x, y, sigma, normalize=False, clip_size=None):
    """Draw a Gaussian into a single-channel CUDA image in place (util.py:70-126)."""
    _lib.require_cuda(img_tensor, 'img_tensor')
    if img_tensor.dim() == 2:
        height, width = list(img_tensor.size())
        img = img_tensor
    elif img_tensor.dim() == 3:
        n_chans, height, width = list(img_tensor.size())
        assert n_chans == 1, 'expected img_tensor to have one channel'
        img = img_tensor[0]
    else:
        raise Exception('expected img_tensor to have 2 or 3 dimensions')
    x, y = int(x), int(y)                                        # util.py:84-85
    radius = max(width, height) if clip_size is None else clip_size / 2
    if radius < 0.5 or x <= -radius or y <= -radius or x >= (width - 1) + radius or y >= (height - 1) + radius:
        return                                                   # util.py:100-102: nothing is drawn, nothing is touched
    import math
    start_x, end_x = max(0, math.ceil(x - radius)), min(width, int(x + radius + 1))
    start_y, end_y = max(0, math.ceil(y - radius)), min(height, int(y + radius + 1))
    centre = torch.tensor([[float(x), float(y)]], dtype=torch.float32, device=img.device)
    drawn = _draw(centre, True, 1, width, height, sigma, clip_size, normalize, img.device)[0]
    # only the draw window is written, the rest of the image keeps its content (util.py:111)
    img[start_y:end_y, start_x:end_x] = drawn[start_y:end_y, start_x:end_x].to(img.dtype)


def decode_heatmaps(heatmaps, use_neighbours=True):
    """Convert heatmaps [B, C, H, W] into normalised coordinates [B, C, 2] float32 (util.py:173-198); the result lives
    where the input lived (the reference is handed `out_var.data.cpu()` and returns a CPU tensor, model.py:165)."""
    if heatmaps.dim() != 4:
        raise ValueError('heatmaps must be [B, C, H, W], got shape %s' % (tuple(heatmaps.shape),))
    dev = _device_of(heatmaps)
    hm = heatmaps.detach().to(dev).contiguous()
    b, c, h, w = hm.shape
    coords = torch.empty(b, c, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call('dsnt_decode_heatmaps', hm.data_ptr(), _lib.dtype_id(hm), b * c, h, w, int(bool(use_neighbours)),
                  coords.data_ptr(), _lib.stream_of(hm))
    return coords if heatmaps.is_cuda else coords.cpu()


def get_preds(heatmaps):
    """Arg-max pixel (x, y) per heatmap as float32 [B, C, 2], (0, 0) when the maximum is <= 0 (util.py:151-170)."""
    _, _, h, w = heatmaps.shape
    norm = decode_heatmaps(heatmaps, use_neighbours=False)
    px = (norm + 1) * norm.new_tensor([w / 2.0, h / 2.0]) - 0.5          # invert util.py:195-198
    return px.round()
