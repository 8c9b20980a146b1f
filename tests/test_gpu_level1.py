"""Level-1 API (`dsnt_pose2d_b200.nn`, the drop-in for the reference's `dsnt.nn`) on the GPU:
the reference's own known-answer tests (tests/test_nn.py, restated) and the golden vectors of the
unmodified reference for every function, forward and backward."""

import math

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_max

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5

SIMPLE_INPUT = torch.tensor([[[
    [0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.1, 0.0],
    [0.0, 0.0, 0.1, 0.6, 0.1],
    [0.0, 0.0, 0.0, 0.1, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0]]]])
SIMPLE_GRAD = torch.tensor([[[
    [0.4800, 0.4400, 0.4000, 0.3600, 0.3200],
    [0.2800, 0.2400, 0.2000, 0.1600, 0.1200],
    [0.0800, 0.0400, 0.0000, -0.0400, -0.0800],
    [-0.1200, -0.1600, -0.2000, -0.2400, -0.2800],
    [-0.3200, -0.3600, -0.4000, -0.4400, -0.4800]]]])


@pytest.fixture(scope='module')
def nn():
    import dsnt_pose2d_b200
    return dsnt_pose2d_b200.nn


def test_dsnt_forward_backward_known_answer(nn):
    """tests/test_nn.py:31-50,68-82 (test_forward, test_backward, test_cuda)."""
    p = SIMPLE_INPUT.to(DEV).requires_grad_(True)
    out = nn.dsnt(p)
    assert torch.allclose(out.cpu(), torch.tensor([[[0.4, 0.0]]]), atol=TOL)
    torch.nn.MSELoss()(out, torch.tensor([[[0.5, 0.5]]], device=DEV)).backward()
    assert torch.allclose(p.grad.cpu(), SIMPLE_GRAD, atol=TOL)


def test_dsnt_batchless(nn):
    """tests/test_nn.py:52-66."""
    p = SIMPLE_INPUT[0].to(DEV).requires_grad_(True)
    out = nn.dsnt(p)
    assert out.shape == (1, 2) and torch.allclose(out.cpu(), torch.tensor([[0.4, 0.0]]), atol=TOL)
    torch.nn.MSELoss()(out, torch.tensor([[0.5, 0.5]], device=DEV)).backward()
    assert torch.allclose(p.grad.cpu(), SIMPLE_GRAD[0], atol=TOL)
    q = SIMPLE_INPUT[0, 0].to(DEV)                                  # no leading dims at all
    assert nn.dsnt(q).shape == (2,)


def test_euclidean_loss_known_answer(nn):
    """tests/test_nn.py:86-130."""
    a = torch.tensor([[[3.0, 4.0]] * 2] * 2, device=DEV, requires_grad=True)
    loss = nn.euclidean_loss(a, torch.zeros(2, 2, 2, device=DEV))
    loss.backward()
    assert abs(loss.item() - 5.0) < TOL
    assert torch.allclose(a.grad.cpu(), torch.tensor([[[0.15, 0.20]] * 2] * 2), atol=TOL)
    out = torch.tensor([[[0, 0], [1, 1], [0, 0]], [[1, 1], [0, 0], [0, 0]]], dtype=torch.float32, device=DEV)
    mask = torch.tensor([[1, 0, 1], [0, 1, 1]], dtype=torch.float32, device=DEV)
    assert nn.euclidean_loss(out, torch.zeros_like(out), mask).item() == 0.0


def test_thresholded_softmax_known_answer_and_gradcheck(nn):
    """tests/test_nn.py:133-154 (gradcheck there is fp64; here finite differences in fp32 vs the analytic kernel)."""
    exp = torch.tensor([0.26894142, 0, 0.73105858])
    assert torch.allclose(nn.thresholded_softmax(torch.tensor([2.0, 1.0, 3.0], device=DEV), 1.5).cpu(), exp, atol=TOL)
    got = nn.thresholded_softmax(torch.tensor([[2.0, 1, 3], [4, 0, 0]], device=DEV), 1.5).cpu()
    assert torch.allclose(got, torch.stack([exp, torch.tensor([1.0, 0, 0])]), atol=TOL)
    torch.manual_seed(0)
    x = torch.randn(3, 20, device=DEV, requires_grad=True)
    w = torch.randn(3, 20, device=DEV)
    (nn.thresholded_softmax(x, 0) * w).sum().backward()
    x64 = x.detach().cpu().double().requires_grad_(True)
    from oracle import torch_port as tp
    (tp.thresholded_softmax(x64, 0) * w.cpu().double()).sum().backward()
    assert rel_l2(x.grad.cpu().double().numpy(), x64.grad.numpy()) < TOL


def test_make_gauss_known_answer(nn):
    """tests/test_nn.py:157-167."""
    exp = torch.tensor([
        [0.0030, 0.0133, 0.0219, 0.0133, 0.0030],
        [0.0133, 0.0596, 0.0983, 0.0596, 0.0133],
        [0.0219, 0.0983, 0.1621, 0.0983, 0.0219],
        [0.0133, 0.0596, 0.0983, 0.0596, 0.0133],
        [0.0030, 0.0133, 0.0219, 0.0133, 0.0030]])
    got = nn.make_gauss(torch.tensor([0.0, 0.0], device=DEV), 5, 5, sigma=0.4).cpu()
    assert got.shape == (5, 5) and (got - exp).abs().max() < 1e-4


@pytest.mark.parametrize('name,shift_mean', [('kl', True), ('mse', True), ('js', True), ('var', False)])
def test_reg_loss_minimum_properties(nn, name, shift_mean):
    """tests/test_nn.py:170-197,200-239."""
    fn = {'kl': nn.kl_reg_loss, 'mse': nn.mse_reg_loss, 'js': nn.js_reg_loss, 'var': nn.variance_reg_loss}[name]
    t_mean, t_std = torch.tensor([0.0, 0.0], device=DEV), 0.4

    def calc(mean, std):
        return fn(nn.make_gauss(mean, 5, 5, sigma=std), t_mean, t_std, mask=None).item()

    lo = calc(t_mean, t_std)
    assert abs(lo) < 1e-3
    assert calc(t_mean, t_std + 0.2) > lo + 1e-3
    assert calc(t_mean, t_std - 0.2) > lo + 1e-3
    if shift_mean:
        assert calc(t_mean + 0.1, t_std) > lo + 1e-3
        assert calc(t_mean - 0.1, t_std) > lo + 1e-3


def test_kl_mask_known_answer(nn):
    """tests/test_nn.py:204-224: expected 1.2228811717796824."""
    t = torch.tensor([
        [[0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.1], [0.0, 0.0, 0.1, 0.8]],
        [[0.8, 0.1, 0.0, 0.0], [0.1, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]], device=DEV)
    coords = torch.tensor([[1.0, 1.0], [0.0, 0.0]], device=DEV)
    mask = torch.tensor([1.0, 0.0], device=DEV)
    got = nn.kl_reg_loss(t, coords, 1, mask).item()
    assert abs(got - 1.2228811717796824) < TOL * 1.2228811717796824 * 2


# ------------------------------------------------------------------------------------------- golden vectors
def test_level1_dsnt_and_regs_match_golden(nn, golden_l1):
    g = golden_l1
    for name in g.cases:
        p32 = torch.from_numpy(g[name + '/p']).to(DEV)
        mu = torch.from_numpy(g[name + '/mu']).to(DEV)
        mask = torch.from_numpy(g[name + '/mask']).to(DEV)
        sigma = float(g[name + '/sigma'])
        p = p32.clone().requires_grad_(True)
        coords = nn.dsnt(p)
        coords.backward(torch.from_numpy(g[name + '/dsnt/g_coords']).float().to(DEV))
        assert np.abs(coords.detach().cpu().double().numpy() - g[name + '/dsnt/coords']).max() < TOL, name
        assert rel_l2(p.grad.cpu().double().numpy(), g[name + '/dsnt/dp']) < TOL, name
        for reg, fn in (('var', nn.variance_reg_loss), ('kl', nn.kl_reg_loss), ('js', nn.js_reg_loss),
                        ('mse', nn.mse_reg_loss)):
            for mtag in ('mask', 'nomask'):
                p = p32.clone().requires_grad_(True)
                val = fn(p, mu, sigma, mask if mtag == 'mask' else None)
                val.backward()
                gl = float(g['%s/%s/%s/loss' % (name, reg, mtag)])
                gdp = g['%s/%s/%s/dp' % (name, reg, mtag)]
                e_l = abs(val.item() - gl) / max(abs(gl), 1e-30)
                e_g = rel_l2(p.grad.cpu().double().numpy(), gdp)
                print('%-14s %-4s %-7s loss %.2e dp %.2e' % (name, reg, mtag, e_l, e_g))
                assert e_l < TOL, (name, reg, mtag, val.item(), gl)
                assert e_g < TOL and rel_max(p.grad.cpu().double().numpy(), gdp) < 4 * TOL, (name, reg, mtag)


def test_level1_euclid_softmax_gauss_match_golden(nn, golden_l1):
    g = golden_l1
    for tag in ('e2', 'e3', 'e_single'):
        actual = torch.from_numpy(g['euclid/%s/actual' % tag]).to(DEV)
        target = torch.from_numpy(g['euclid/%s/target' % tag]).to(DEV)
        mask = g.get('euclid/%s/mask' % tag)
        for mtag in ('mask', 'nomask'):
            mm = torch.from_numpy(mask).to(DEV) if (mtag == 'mask' and mask is not None) else None
            a = actual.clone().requires_grad_(True)
            val = nn.euclidean_loss(a, target, mm)
            val.backward()
            assert abs(val.item() - float(g['euclid/%s/%s/loss' % (tag, mtag)])) < TOL * max(1, abs(val.item()))
            assert rel_l2(a.grad.cpu().double().numpy(), g['euclid/%s/%s/grad' % (tag, mtag)]) < TOL
    z = torch.from_numpy(g['softmax2d/z']).to(DEV).requires_grad_(True)
    for fn in (nn.softmax_2d, nn.flat_softmax):
        z.grad = None
        out = fn(z)
        out.backward(torch.from_numpy(g['softmax2d/g']).float().to(DEV))
        assert rel_max(out.detach().cpu().double().numpy(), g['softmax2d/out']) < TOL
        assert rel_l2(z.grad.cpu().double().numpy(), g['softmax2d/dz']) < TOL
    x = torch.from_numpy(g['tsoftmax/x']).to(DEV)
    for tag in ('thr0', 'thrm05', 'thrinf'):
        thr = float(g['tsoftmax/%s/thr' % tag])
        xt = x.clone().requires_grad_(True)
        out = nn.thresholded_softmax(xt, thr)
        out.backward(torch.from_numpy(g['tsoftmax/%s/g' % tag]).float().to(DEV))
        assert rel_max(out.detach().cpu().double().numpy(), g['tsoftmax/%s/out' % tag]) < TOL
        assert rel_l2(xt.grad.cpu().double().numpy(), g['tsoftmax/%s/dx' % tag]) < TOL
    mu = torch.from_numpy(g['gauss/mu']).to(DEV).requires_grad_(True)
    out = nn.make_gauss(mu, 9, 6, 0.25)
    assert out.shape == (2, 3, 6, 9)
    out.backward(torch.from_numpy(g['gauss/g']).float().to(DEV))
    assert rel_max(out.detach().cpu().double().numpy(), g['gauss/out']) < TOL
    assert rel_l2(mu.grad.cpu().double().numpy(), g['gauss/dmu']) < 2 * TOL


def test_helpers_generate_xy_expectation_masked_average(nn):
    """Import-compat helpers (src/dsnt/nn.py:25-63,81-94)."""
    inp = torch.rand(2, 3, 4, 6, device=DEV)
    xs, ys = nn.generate_xy(inp)
    assert xs.shape == inp.shape and ys.shape == inp.shape
    assert torch.allclose(xs[0, 0, 0].cpu(), torch.linspace(-5 / 6, 5 / 6, 6), atol=1e-6)
    assert torch.allclose(ys[0, 0, :, 0].cpu(), torch.linspace(-3 / 4, 3 / 4, 4), atol=1e-6)
    p = torch.softmax(inp.view(6, -1), -1).view(2, 3, 4, 6)
    assert torch.allclose(nn.expectation_2d(xs, p), nn.dsnt(p)[..., 0], atol=1e-6)
    losses = torch.rand(2, 3, device=DEV, requires_grad=True)
    mask = torch.tensor([[1.0, 0, 1], [0, 0, 1]], device=DEV)
    val = nn.masked_average(losses, mask)
    val.backward()
    assert abs(val.item() - (losses.detach() * mask).sum().item() / 3.0) < 1e-6
    assert torch.allclose(losses.grad, mask / 3.0, atol=1e-7)
    assert abs(nn.masked_average(losses.detach()).item() - losses.mean().item()) < 1e-6
    assert nn.masked_average(losses.detach(), torch.zeros(2, 3, device=DEV)).item() == 0.0


def test_softmax_2d_large_rows_and_bf16(nn):
    z = torch.randn(3, 2, 64, 64, device=DEV)
    ref = torch.softmax(z.double().view(6, -1), -1).view_as(z)
    assert rel_max(nn.flat_softmax(z).cpu().double().numpy(), ref.cpu().numpy()) < TOL
    zb = z.to(torch.bfloat16)
    out = nn.flat_softmax(zb)
    assert out.dtype == torch.bfloat16
    refb = torch.softmax(zb.double().view(6, -1), -1).view_as(z)
    assert rel_max(out.float().cpu().double().numpy(), refb.cpu().numpy()) < 1e-2
    assert math.isclose(out.float().sum().item(), 6.0, rel_tol=2e-2)
