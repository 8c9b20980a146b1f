"""CPU: oracle/util_port.py (src/dsnt/util.py:70-198 restated) against the known answers of the reference's
tests/test_util.py and golden vectors produced by the unmodified reference (gauss_util.npz)."""

import numpy as np
import torch

from oracle import util_port as up

GAUSS9 = [
    [0.00000, 0.00000, 0.00005, 0.00020, 0.00034, 0.00020, 0.00005, 0.00000, 0.00000],
    [0.00000, 0.00012, 0.00150, 0.00674, 0.01111, 0.00674, 0.00150, 0.00012, 0.00000],
    [0.00005, 0.00150, 0.01832, 0.08208, 0.13534, 0.08208, 0.01832, 0.00150, 0.00005],
    [0.00020, 0.00674, 0.08208, 0.36788, 0.60653, 0.36788, 0.08208, 0.00674, 0.00020],
    [0.00034, 0.01111, 0.13534, 0.60653, 1.00000, 0.60653, 0.13534, 0.01111, 0.00034],
    [0.00020, 0.00674, 0.08208, 0.36788, 0.60653, 0.36788, 0.08208, 0.00674, 0.00020],
    [0.00005, 0.00150, 0.01832, 0.08208, 0.13534, 0.08208, 0.01832, 0.00150, 0.00005],
    [0.00000, 0.00012, 0.00150, 0.00674, 0.01111, 0.00674, 0.00150, 0.00012, 0.00000],
    [0.00000, 0.00000, 0.00005, 0.00020, 0.00034, 0.00020, 0.00005, 0.00000, 0.00000],
]
CLIPPED5 = [
    [0.00000, 0.00000, 0.00000, 0.00000, 0.00000],
    [0.01111, 0.00674, 0.00150, 0.00012, 0.00000],
    [0.13534, 0.08208, 0.01832, 0.00150, 0.00000],
    [0.60653, 0.36788, 0.08208, 0.00674, 0.00000],
    [1.00000, 0.60653, 0.13534, 0.01111, 0.00000],
]


def test_draw_gaussian_known_answers():
    """tests/test_util.py:8-37."""
    actual = torch.zeros(1, 9, 9)
    up.draw_gaussian(actual, 4, 4, 1, normalize=False)
    assert (actual - torch.tensor([GAUSS9])).abs().max().item() < 1e-5
    actual = torch.zeros(1, 5, 5)
    up.draw_gaussian(actual, 0, 4, 1, normalize=False, clip_size=7)
    assert (actual - torch.tensor([CLIPPED5])).abs().max().item() < 1e-5


def test_encode_heatmaps_known_answer():
    """tests/test_util.py:39-52."""
    actual = up.encode_heatmaps(torch.tensor([[[-0.8, 0.8]]]), 5, 5)
    assert (actual - torch.tensor([[CLIPPED5]])).abs().max().item() < 1e-5


def test_decode_heatmaps_known_answers():
    """tests/test_util.py:54-77."""
    hm = torch.tensor([[[[0.0, 0.9], [0.0, 0.1]]]])
    assert (up.decode_heatmaps(hm) - torch.tensor([[[0.5, -0.5]]])).abs().max().item() < 1e-7
    hm = torch.tensor([[[[0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.9, 0.1, 0.0], [0.0, 0.1, 0.0, 0.0]]]])
    assert (up.decode_heatmaps(hm, use_neighbours=True) - torch.tensor([[[-0.125, 0.375]]])).abs().max().item() < 1e-7


def test_util_port_matches_reference_golden(golden_gauss):
    g = golden_gauss
    for name in ('e5x5', 'e64', 'e28', 'e7x12'):
        hm = g[name + '/hm']
        got = up.encode_heatmaps(torch.from_numpy(g[name + '/coords']), hm.shape[-1], hm.shape[-2], float(g[name + '/sigma']))
        assert np.array_equal(got.numpy(), hm), name
    for name in ('d9', 'd5clip', 'd12norm', 'd12clipnorm', 'd_out'):
        x, y, sigma, normalize, clip = g[name + '/args']
        img = torch.zeros(g[name + '/img'].shape)
        up.draw_gaussian(img, x, y, sigma, normalize=bool(normalize), clip_size=None if clip < 0 else clip)
        assert np.array_equal(img.numpy(), g[name + '/img']), name
    for name in ('g64', 'g28', 'g6x9', 'g9x6', 'g2x2'):
        hm = torch.from_numpy(g[name + '/hm'])
        assert np.array_equal(up.decode_heatmaps(hm, True).numpy(), g[name + '/coords_nb']), name
        assert np.array_equal(up.decode_heatmaps(hm, False).numpy(), g[name + '/coords']), name
    rt = up.decode_heatmaps(up.encode_heatmaps(torch.from_numpy(g['rt/coords']), 64, 64, 1))
    assert np.array_equal(rt.numpy(), g['rt/decoded'])
    # the round trip recovers every joint to within half a pixel
    assert np.abs(g['rt/decoded'] - g['rt/coords']).max() <= 1.0 / 64 + 1e-6
