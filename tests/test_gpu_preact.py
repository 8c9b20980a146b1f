"""GPU parity of the fused head for the reference's other pre-activations (SURVEY.md 8f row 2;
src/dsnt/model.py:24-45): thresholded_softmax(-0.5), abs, relu, sigmoid -- and plain softmax through the same
epsilon-exact kernels -- against (a) golden vectors from the unmodified reference and (b) the fp64 oracle on
seeded inputs.  Tolerances as in test_gpu_parity.py: 1e-5 (coords max-abs, loss relative, dZ L2-relative)."""

import numpy as np
import pytest
import torch

from conftest import head_case_params, rel_l2, rel_max

pytestmark = pytest.mark.gpu

PREACTS = ['softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid']
REGS = ['none', 'var', 'kl', 'js', 'mse']
TOL = 1e-5
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def dp():
    import dsnt_pose2d_b200
    return dsnt_pose2d_b200


@pytest.fixture(scope='module')
def tp():
    from oracle import torch_port
    return torch_port


def run_head(dp, z, target, mask, preact, reg, hm_sigma=1.0, coeff=1.0, variant=0):
    zz = z.detach().clone().to(DEV).requires_grad_(True)
    tt = None if target is None else target.to(DEV)
    mm = None if mask is None else mask.to(DEV)
    # eps=... forces the epsilon-exact kernels for 'softmax' too (the default routes softmax to the tuned kernels)
    kw = {'eps': 0.0} if preact == 'softmax' else {}
    out = dp.dsnt_head(zz, tt, mm, reg=reg, hm_sigma=hm_sigma, reg_coeff=coeff, preact=preact, variant=variant, **kw)
    out.loss.backward()
    torch.cuda.synchronize()
    return {'loss': out.loss.item(), 'euclid': out.euclid.item(), 'reg': out.reg.item(),
            'coords': out.coords.detach().cpu().double().numpy(), 'dz': zz.grad.detach().cpu().double().numpy()}


def check(got, ref_loss, ref_coords, ref_dz, what, tol=TOL, dz_tol=None):
    dz_tol = tol if dz_tol is None else dz_tol
    e_loss = abs(got['loss'] - ref_loss) / max(abs(ref_loss), 1e-30)
    e_coords = float(np.abs(got['coords'] - ref_coords).max())
    e_l2 = rel_l2(got['dz'], ref_dz)
    e_max = rel_max(got['dz'], ref_dz)
    print('%-52s loss %.2e coords %.2e dz L2 %.2e max %.2e' % (what, e_loss, e_coords, e_l2, e_max))
    assert e_loss < tol, (what, 'loss', got['loss'], ref_loss)
    assert e_coords < tol, (what, 'coords', e_coords)
    assert e_l2 < dz_tol and e_max < dz_tol * 4, (what, 'dz', e_l2, e_max)


# variant 0 = the tuned streaming kernels where the layout qualifies (64x64 here), 1 = the generic epsilon-exact kernels
@pytest.mark.parametrize('variant', [0, 1])
@pytest.mark.parametrize('reg', REGS)
@pytest.mark.parametrize('preact', PREACTS)
def test_preact_head_matches_reference_golden(dp, golden_preact, preact, reg, variant):
    g = golden_preact
    for name in g.cases:
        b, c, h, w, hm_sigma, coeff, with_mask = head_case_params(g, name)
        z = torch.from_numpy(g[name + '/z'])
        target = torch.from_numpy(g[name + '/target'])
        mask = torch.from_numpy(g[name + '/mask']) if with_mask else None
        got = run_head(dp, z, target, mask, preact, reg, hm_sigma, coeff, variant)
        key = '%s/%s/%s' % (name, preact, reg)
        check(got, float(g[key + '/loss']), g['%s/%s/coords' % (name, preact)], g[key + '/dz'].astype(np.float64),
              'golden %s %s %s' % (name, preact, reg))
        assert abs(got['euclid'] - float(g[key + '/euclid'])) < TOL
        assert abs(got['reg'] - float(g[key + '/reg'])) < TOL * max(1.0, abs(got['reg']))


@pytest.mark.parametrize('shape', [(6, 6), (64, 64)])
@pytest.mark.parametrize('preact', ['thresholded_softmax', 'relu'])
def test_dead_heatmaps_give_zero_probability_and_zero_gradient(dp, tp, golden_preact, preact, shape):
    g = golden_preact
    z = torch.from_numpy(g['dead/z'])
    target = torch.from_numpy(g['dead/target'])
    if shape != (6, 6):      # the same situation at a size the tuned kernels take: checked against the fp64 oracle
        z = -(torch.rand(1, 2, *shape, generator=torch.Generator().manual_seed(5)) + 1.0)
        for reg in REGS:
            ref = tp.head_loss_and_grad(z, target, None, reg, 1.0, 1.0, dtype=torch.float64, preact=preact)
            got = run_head(dp, z, target, None, preact, reg)
            assert abs(got['loss'] - ref['loss'].item()) <= TOL * max(abs(ref['loss'].item()), 1e-3), (reg, got['loss'])
            assert np.abs(got['dz']).max() == 0.0 and np.abs(got['coords']).max() == 0.0
        return
    for reg in REGS:
        got = run_head(dp, z, target, None, preact, reg)
        ref = float(g['dead/%s/%s/loss' % (preact, reg)])
        assert abs(got['loss'] - ref) <= TOL * max(abs(ref), 1e-3), (reg, got['loss'], ref)
        assert np.abs(got['dz']).max() == 0.0
        assert np.abs(got['coords']).max() == 0.0


@pytest.mark.parametrize('variant', [0, 1])
@pytest.mark.parametrize('reg', REGS)
@pytest.mark.parametrize('preact', ['thresholded_softmax', 'abs', 'relu', 'sigmoid'])
@pytest.mark.parametrize('shape,scale', [((32, 16, 64, 64), 1.0), ((8, 16, 64, 64), 5.0), ((64, 16, 28, 28), 1.0),
                                         ((2, 3, 7, 7), 2.0), ((2, 2, 130, 132), 1.0), ((1, 2, 256, 256), 1.0),
                                         ((2, 2, 33, 31), 1.0), ((2, 4, 128, 128), 3.0), ((4, 4, 32, 32), 1.0)])
def test_preact_head_matches_fp64_oracle(dp, tp, preact, reg, shape, scale, variant):
    """BASELINE head shapes (cfg 1, cfg 2), an odd scalar-path size, a streaming-path size (> 128x128) and 256x256."""
    b, c, h, w = shape
    gen = torch.Generator().manual_seed(11)
    z = torch.randn(b, c, h, w, generator=gen) * scale
    target = torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8
    mask = (torch.rand(b, c, generator=gen) > 0.1).float()
    ref = tp.head_loss_and_grad(z, target, mask, reg, 1.0, 1.0, dtype=torch.float64, preact=preact)
    got = run_head(dp, z, target, mask, preact, reg, variant=variant)
    check(got, ref['loss'].item(), ref['coords'].numpy(), ref['dz'].numpy(),
          '%s %s %s v%d' % ('x'.join(map(str, shape)), preact, reg, variant))


@pytest.mark.parametrize('preact', ['thresholded_softmax', 'abs', 'relu', 'sigmoid'])
def test_preact_trained_like_heatmaps_and_any_sigma(dp, tp, preact):
    """Peaked 'trained network' maps (positive peak near the target, negative tails -> many exact zeros for relu /
    threshold) and Gaussian targets from sub-pixel to image-wide: the window optimisation must not show."""
    gen = torch.Generator().manual_seed(14)
    b, c, h, w = 4, 8, 64, 64
    target = torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8
    g = tp.make_gauss(target + 0.05 * torch.randn(b, c, 2, generator=gen), w, h, 2.0 / w)
    z = (g + 1e-6).log() + 8.0 + 0.1 * torch.randn(b, c, h, w, generator=gen)
    mask = (torch.rand(b, c, generator=gen) > 0.1).float()
    for reg in ('js', 'kl', 'mse'):
        for hm_sigma in (0.4, 1.0, 3.0, 20.0):
            ref = tp.head_loss_and_grad(z, target, mask, reg, hm_sigma, 1.0, dtype=torch.float64, preact=preact)
            got = run_head(dp, z, target, mask, preact, reg, hm_sigma=hm_sigma)
            check(got, ref['loss'].item(), ref['coords'].numpy(), ref['dz'].numpy(),
                  'trained %s %s sigma %.1f' % (preact, reg, hm_sigma))


@pytest.mark.parametrize('preact', ['thresholded_softmax', 'relu', 'sigmoid'])
def test_preact_head_bf16(dp, tp, preact):
    """bf16 raw heatmaps: the oracle runs on the bf16-rounded values; coords/loss 1e-5, dZ (bf16) 4e-3 L2-relative."""
    gen = torch.Generator().manual_seed(12)
    z = (torch.randn(8, 16, 64, 64, generator=gen)).to(torch.bfloat16)
    target = torch.rand(8, 16, 2, generator=gen) * 1.6 - 0.8
    mask = (torch.rand(8, 16, generator=gen) > 0.1).float()
    for reg in ('js', 'var'):
        ref = tp.head_loss_and_grad(z.float(), target, mask, reg, 1.0, 1.0, dtype=torch.float64, preact=preact)
        got = run_head(dp, z, target, mask, preact, reg)
        check(got, ref['loss'].item(), ref['coords'].numpy(), ref['dz'].numpy(), 'bf16 %s %s' % (preact, reg),
              dz_tol=4e-3)


def test_model_head_uses_fused_preact_kernels(dp, tp):
    """DSNTHead(preact=...) keeps the reference's forward_part2 / forward_loss surface and now takes the fused path."""
    from dsnt_pose2d_b200 import _lib
    gen = torch.Generator().manual_seed(13)
    z = torch.randn(4, 16, 28, 28, generator=gen)
    target = torch.rand(4, 16, 2, generator=gen) * 1.6 - 0.8
    mask = (torch.rand(4, 16, generator=gen) > 0.1).float()
    for preact in ('thresholded_softmax', 'abs', 'relu', 'sigmoid'):
        head = dp.DSNTHead(preact=preact, reg='js', reg_coeff=1.0, hm_sigma=1.0)
        zz = z.clone().to(DEV).requires_grad_(True)
        before = _lib.launch_count
        out = head.forward_part2(zz)
        loss = head.forward_loss(out, target.to(DEV), mask.to(DEV))
        loss.backward()
        assert _lib.launch_count - before == 5      # part2: coords fwd + finish; loss: fused fwd + finish; bwd
        ref = tp.head_loss_and_grad(z, target, mask, 'js', 1.0, 1.0, dtype=torch.float64, preact=preact)
        assert abs(loss.item() - ref['loss'].item()) < TOL * abs(ref['loss'].item())
        assert rel_l2(zz.grad.cpu().double().numpy(), ref['dz'].numpy()) < TOL
        assert np.abs(head.compute_coords(out).double().numpy() - ref['coords'].numpy()).max() < TOL
        # lazily materialised heatmaps = the reference's _hm_preact
        p_ref = tp.hm_preact(z.double(), preact)
        assert (head.heatmaps.detach().cpu().double() - p_ref).abs().max().item() < 1e-6


def test_unknown_preact_raises_like_the_reference(dp):
    z = torch.zeros(1, 1, 4, 4, device=DEV)
    with pytest.raises(Exception, match='unrecognised heatmap preactivation function'):
        dp.dsnt_head(z, None, None, preact='tanh')
