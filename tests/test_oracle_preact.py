"""CPU: the oracle's restatement of every `--preact` choice (src/dsnt/model.py:24-45) is pinned to golden vectors
produced by the unmodified reference (tests/golden/make_golden_next.py -> preact_heads.npz)."""

import numpy as np
import pytest
import torch

from conftest import head_case_params, rel_l2
from oracle import torch_port as tp

PREACTS = ['softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid']
REGS = ['none', 'var', 'kl', 'js', 'mse']


@pytest.mark.parametrize('preact', PREACTS)
def test_oracle_preact_heads_match_reference_golden(golden_preact, preact):
    g = golden_preact
    for name in g.cases:
        b, c, h, w, hm_sigma, coeff, with_mask = head_case_params(g, name)
        z = torch.from_numpy(g[name + '/z'])
        target = torch.from_numpy(g[name + '/target'])
        mask = torch.from_numpy(g[name + '/mask']) if with_mask else None
        for reg in REGS:
            ref = tp.head_loss_and_grad(z, target, mask, reg, hm_sigma, coeff, dtype=torch.float64, preact=preact)
            key = '%s/%s/%s' % (name, preact, reg)
            assert abs(ref['loss'].item() - float(g[key + '/loss'])) <= 1e-12 * max(1.0, abs(float(g[key + '/loss'])))
            assert abs(ref['euclid'].item() - float(g[key + '/euclid'])) <= 1e-12
            assert rel_l2(ref['dz'].numpy(), g[key + '/dz']) < (1e-6 if g[key + '/dz'].dtype == np.float32 else 1e-12)
        np.testing.assert_allclose(ref['coords'].numpy(), g['%s/%s/coords' % (name, preact)], atol=1e-13)


@pytest.mark.parametrize('preact', ['thresholded_softmax', 'relu'])
def test_oracle_dead_heatmaps(golden_preact, preact):
    """Everything below the threshold / non-positive: P = 0 everywhere, finite loss, zero gradient."""
    g = golden_preact
    z = torch.from_numpy(g['dead/z'])
    target = torch.from_numpy(g['dead/target'])
    for reg in REGS:
        ref = tp.head_loss_and_grad(z, target, None, reg, 1.0, 1.0, dtype=torch.float64, preact=preact)
        assert abs(ref['loss'].item() - float(g['dead/%s/%s/loss' % (preact, reg)])) < 1e-12
        assert np.abs(ref['dz'].numpy() - g['dead/%s/%s/dz' % (preact, reg)]).max() < 1e-12


def test_unknown_preact_raises_like_the_reference():
    with pytest.raises(Exception, match='unrecognised heatmap preactivation function'):
        tp.hm_preact(torch.zeros(1, 1, 2, 2), 'tanh')
