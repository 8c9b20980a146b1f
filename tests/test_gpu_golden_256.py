"""The CUDA path at BASELINE config 5's heatmap size (256x256) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden_cfg5.py -> head_256.npz; inputs regenerated from a seed, tests/golden/cfg5_inputs.py): the one-pass
step on a cluster of CTAs (csrc/step_pair.cu: fp32, not KL), the two-kernel path, and bf16 logits through both.  Tolerances as in
test_gpu_parity.py: 1e-5 (coords max-abs, loss relative, dZ L2-relative); bf16: the reference evaluated on the rounded logits
is not in the fixture, so bf16 is compared at its own resolution (dZ 2e-2, loss 5e-3: rounding the LOGITS to 8 bits moves
every probability by |z| 2^-9)."""

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

from cfg5_inputs import CASES, make_case  # noqa: E402
from conftest import Golden, rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


@pytest.fixture(scope='module')
def golden():
    return Golden('head_256.npz')


@pytest.fixture(scope='module')
def dp():
    import dsnt_pose2d_b200
    return dsnt_pose2d_b200


@pytest.fixture(autouse=True)
def small_batches_take_the_step_kernels():
    from dsnt_pose2d_b200 import head
    old, head.STEP_MIN_BYTES = head.STEP_MIN_BYTES, 0
    yield
    head.STEP_MIN_BYTES = old


@pytest.mark.parametrize('case', list(CASES))
@pytest.mark.parametrize('reg', ['none', 'var', 'kl', 'js', 'mse'])
@pytest.mark.parametrize('one_pass', [True, False])
def test_head_matches_the_reference_at_256(dp, golden, case, reg, one_pass):
    from dsnt_pose2d_b200 import _lib
    b, c, kind, hm_sigma, coeff, seed = CASES[case]
    z, target, mask, idx = make_case(case)
    zz = torch.from_numpy(z).to(DEV).requires_grad_(True)
    before = _lib.launch_count
    out = dp.dsnt_head(zz, torch.from_numpy(target).to(DEV), torch.from_numpy(mask).to(DEV), reg=reg, hm_sigma=hm_sigma,
                       reg_coeff=coeff, one_pass=one_pass)
    out.loss.backward()
    launches = _lib.launch_count - before
    assert launches == (4 if one_pass and reg != 'kl' else 3)      # mask count, cluster step, finish, scale | fwd, finish, bwd
    g = lambda key: golden['%s/%s/%s' % (case, reg, key)]
    dz = zz.grad.detach().cpu().double().numpy()
    e_loss = abs(out.loss.item() - float(g('loss'))) / abs(float(g('loss')))
    e_coords = float(np.abs(out.coords.detach().cpu().double().numpy() - golden[case + '/coords']).max())
    e_dz = rel_l2(dz.reshape(-1)[idx], g('dz_samples'))
    e_norm = rel_l2(np.sqrt((dz.reshape(b * c, -1) ** 2).sum(-1)), g('dz_norms'))
    print('%s %s one_pass=%s: loss %.1e coords %.1e dz %.1e norms %.1e' % (case, reg, one_pass, e_loss, e_coords, e_dz, e_norm))
    tol = TOL
    if reg == 'kl' and kind == 'trained':
        # The stated exception: on peaked maps KL's ln(P + eps) of a float32 P loses digits in ANY fp32
        # evaluation -- the reference's own arithmetic run in float32 is 2.7e-5 from its float64 result on this very case.
        # The bar there: no further from float64 than the reference's own float32, within a factor 1.5.
        from oracle import torch_port as tp
        r32 = tp.head_loss_and_grad(torch.from_numpy(z), torch.from_numpy(target), torch.from_numpy(mask), reg, hm_sigma, coeff,
                                    dtype=torch.float32)
        ref32 = abs(r32['loss'].item() - float(g('loss'))) / abs(float(g('loss')))
        print('   the reference arithmetic in float32: loss %.1e' % ref32)
        assert ref32 > TOL
        tol = 1.5 * ref32
    assert e_loss < tol and e_coords < TOL and e_dz < tol and e_norm < tol
    if reg != 'none':
        assert abs(out.reg.item() - float(g('reg'))) < 2 * tol * abs(float(g('reg')))


@pytest.mark.parametrize('reg', ['var', 'js'])
@pytest.mark.parametrize('pair', [False, True])
def test_bf16_head_at_256_near_the_reference(dp, golden, reg, pair):
    """bf16 logits (both paths): against the reference on the fp32 logits, at bf16 resolution."""
    from dsnt_pose2d_b200 import head
    case = 'c256_diffuse'
    b, c, kind, hm_sigma, coeff, seed = CASES[case]
    z, target, mask, idx = make_case(case)
    old, head.USE_PAIR_STEP_BF16 = head.USE_PAIR_STEP_BF16, pair
    try:
        zz = torch.from_numpy(z).to(DEV).to(torch.bfloat16).requires_grad_(True)
        out = dp.dsnt_head(zz, torch.from_numpy(target).to(DEV), torch.from_numpy(mask).to(DEV), reg=reg, hm_sigma=hm_sigma,
                           reg_coeff=coeff, one_pass=True)
        out.loss.backward()
    finally:
        head.USE_PAIR_STEP_BF16 = old
    dz = zz.grad.detach().float().cpu().double().numpy()
    assert abs(out.loss.item() - float(golden['%s/%s/loss' % (case, reg)])) < 5e-3 * abs(float(golden['%s/%s/loss' % (case, reg)]))
    assert rel_l2(dz.reshape(-1)[idx], golden['%s/%s/dz_samples' % (case, reg)]) < 2e-2
