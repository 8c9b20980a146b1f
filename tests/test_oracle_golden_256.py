"""The oracle at BASELINE config 5's heatmap size (256x256) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden_cfg5.py -> head_256.npz; the inputs are regenerated from a seed, tests/golden/cfg5_inputs.py)."""

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

from cfg5_inputs import CASES, make_case  # noqa: E402
from conftest import Golden, rel_l2  # noqa: E402
from oracle import torch_port as tp  # noqa: E402


@pytest.fixture(scope='module')
def golden():
    return Golden('head_256.npz')


def test_inputs_are_reproducible():
    """The fixture stores no logits: a numpy that changed the legacy RandomState stream would silently invalidate it."""
    z, target, mask, idx = make_case('c256_diffuse')
    assert z.shape == (2, 2, 256, 256) and z.dtype == np.float32
    assert abs(float(z[0, 0, 0, 0]) - 0.1752871424) < 1e-7 and abs(float(z.astype(np.float64).sum()) - (-234.448848)) < 1e-3
    assert idx.shape == (4096,) and int(idx[0]) >= 0 and np.all(np.diff(idx) > 0)


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('reg', ['none', 'var', 'kl', 'js', 'mse'])
def test_oracle_matches_the_reference_at_256(golden, case, reg):
    b, c, kind, hm_sigma, coeff, seed = CASES[case]
    z, target, mask, idx = make_case(case)
    ref = tp.head_loss_and_grad(torch.from_numpy(z), torch.from_numpy(target), torch.from_numpy(mask), reg, hm_sigma, coeff,
                                dtype=torch.float64)
    assert abs(ref['loss'].item() - float(golden['%s/%s/loss' % (case, reg)])) <= 1e-12 * abs(float(golden['%s/%s/loss' % (case, reg)]))
    assert abs(ref['euclid'].item() - float(golden['%s/%s/euclid' % (case, reg)])) <= 1e-12
    assert abs(ref['reg'].item() - float(golden['%s/%s/reg' % (case, reg)])) <= 1e-11 * max(abs(float(golden['%s/%s/reg' % (case, reg)])), 1e-30)
    assert np.abs(ref['coords'].numpy() - golden[case + '/coords']).max() <= 1e-14
    dz = ref['dz'].numpy()
    assert rel_l2(dz.reshape(-1)[idx], golden['%s/%s/dz_samples' % (case, reg)]) <= 1e-11
    assert rel_l2(np.sqrt((dz.reshape(b * c, -1) ** 2).sum(-1)), golden['%s/%s/dz_norms' % (case, reg)]) <= 1e-11
