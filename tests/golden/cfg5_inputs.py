"""Deterministic 256x256 inputs (BASELINE config 5 shape) shared by tests/golden/make_golden_cfg5.py and the tests.

Two megabytes of random logits would not be a "small fixture", so they are not stored: both sides regenerate them from
numpy's frozen legacy generator (RandomState: same stream on every numpy version) and only the reference's OUTPUTS are kept
in head_256.npz -- loss terms, coordinates, and dL/dZ at 4096 sampled pixels per case plus its per-heatmap L2 norms."""

import numpy as np

CASES = {
    # name: (B, C, kind, hm_sigma, reg_coeff, seed)
    'c256_diffuse': (2, 2, 'diffuse', 1.0, 1.0, 2561),
    'c256_trained': (1, 3, 'trained', 1.0, 2.0, 2562),
}
H = W = 256
N_SAMPLES = 4096


def make_case(name):
    """-> z [B,C,H,W] float32, target [B,C,2] float32, mask [B,C] float32, flat sample indices into z (int64)."""
    b, c, kind, hm_sigma, coeff, seed = CASES[name]
    rs = np.random.RandomState(seed)
    target = (rs.rand(b, c, 2) * 1.2 - 0.6).astype(np.float32)
    target[0, 0, 1] = 0.002                     # window astride the two halves of the heatmap (cluster of two CTAs)
    if c > 2:
        target[0, 2] = np.array([0.994, -0.99], dtype=np.float32)      # window clipped by two borders
    if kind == 'diffuse':
        z = (rs.standard_normal((b, c, H, W)) * 2.0).astype(np.float32)
    else:
        # "trained network": log of a 1.5 px Gaussian about a point ~1 px from the target, plus noise
        centre = target.astype(np.float64) + rs.standard_normal((b, c, 2)) * (1.0 * 2.0 / W)
        xs = (2.0 * np.arange(W) + 1.0) / W - 1.0
        ys = (2.0 * np.arange(H) + 1.0) / H - 1.0
        s = 1.5 * 2.0 / W
        gx = np.exp(-0.5 * ((xs[None, None, :] - centre[..., 0:1]) / s) ** 2)
        gy = np.exp(-0.5 * ((ys[None, None, :] - centre[..., 1:2]) / s) ** 2)
        g = gy[..., :, None] * gx[..., None, :]
        z = (np.log(g / g.sum(axis=(-1, -2), keepdims=True) + 1e-9) + 0.05 * rs.standard_normal((b, c, H, W))).astype(np.float32)
    mask = np.ones((b, c), dtype=np.float32)
    mask[-1, -1] = 0.0                          # one joint not visible
    idx = np.sort(rs.choice(b * c * H * W, size=N_SAMPLES, replace=False)).astype(np.int64)
    return z, target, mask, idx
