"""CPU: the oracle's flip test-time augmentation (src/dsnt/inference.py:36-48) against reference golden vectors."""

import numpy as np
import torch

from oracle import torch_port as tp

PREACTS = ['softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid']


def test_oracle_flip_tta_matches_reference_golden(golden_flip):
    g = golden_flip
    for name in g.cases:
        pair = torch.from_numpy(g[name + '/hm_pair']).double()
        flips = [int(i) for i in g[name + '/flips']]
        for preact in PREACTS:
            coords, hm = tp.flip_tta_coords(pair, flips, preact)
            np.testing.assert_allclose(hm.numpy(), g[name + '/hm'], rtol=0, atol=1e-15)
            np.testing.assert_allclose(coords.numpy(), g['%s/%s/coords' % (name, preact)], rtol=0, atol=1e-13)


def test_mpii_flip_permutation_is_an_involution():
    p = tp.MPII_HFLIP_INDICES
    assert sorted(p) == list(range(16)) and all(p[p[i]] == i for i in range(16))
    # left/right pairs of the skeleton the reference draws (src/dsnt/util.py:15-31): ankles, knees, hips, wrists ...
    assert (p[0], p[1], p[2], p[10], p[11], p[12]) == (5, 4, 3, 15, 14, 13) and p[6:10] == (6, 7, 8, 9)


def test_flip_of_a_mirrored_pair_is_symmetric():
    """If the second half really is the mirror image of the first, the averaged heatmaps equal the first half."""
    gen = torch.Generator().manual_seed(0)
    hm1 = torch.randn(3, 16, 8, 10, generator=gen, dtype=torch.float64)
    idx = torch.tensor(tp.MPII_HFLIP_INDICES)
    hm2 = tp.reverse_tensor(hm1.index_select(1, idx), -1)
    hm = tp.flip_tta_heatmaps(torch.cat([hm1, hm2], 0))
    assert torch.equal(hm, hm1)
