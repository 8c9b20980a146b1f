"""Pin the CPU oracle (oracle/) to the reference: its own known-answer tests
(/root/reference/tests/test_nn.py, restated here because that harness is bit-rotted under
torch 2.x -- SURVEY.md section 4) and the golden vectors produced by running the unmodified
reference (tests/golden/make_golden.py).  CPU only."""

import math

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import torch_port as tp
from conftest import head_case_params, rel_l2, rel_max

REGS = ['none', 'var', 'kl', 'js', 'mse']
D = torch.float64

# --------------------------------------------------------------- reference known answers
SIMPLE_INPUT = torch.tensor([[[
    [0.0, 0.0, 0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0, 0.1, 0.0],
    [0.0, 0.0, 0.1, 0.6, 0.1],
    [0.0, 0.0, 0.0, 0.1, 0.0],
    [0.0, 0.0, 0.0, 0.0, 0.0]]]], dtype=D)                 # tests/test_nn.py:10-16
SIMPLE_GRAD = torch.tensor([[[
    [0.4800, 0.4400, 0.4000, 0.3600, 0.3200],
    [0.2800, 0.2400, 0.2000, 0.1600, 0.1200],
    [0.0800, 0.0400, 0.0000, -0.0400, -0.0800],
    [-0.1200, -0.1600, -0.2000, -0.2400, -0.2800],
    [-0.3200, -0.3600, -0.4000, -0.4400, -0.4800]]]], dtype=D)  # tests/test_nn.py:23-29
GAUSS_5x5 = torch.tensor([
    [0.0030, 0.0133, 0.0219, 0.0133, 0.0030],
    [0.0133, 0.0596, 0.0983, 0.0596, 0.0133],
    [0.0219, 0.0983, 0.1621, 0.0983, 0.0219],
    [0.0133, 0.0596, 0.0983, 0.0596, 0.0133],
    [0.0030, 0.0133, 0.0219, 0.0133, 0.0030]], dtype=D)    # tests/test_nn.py:159-165
KL_MASK_HM = torch.tensor([
    [[0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.1], [0.0, 0.0, 0.1, 0.8]],
    [[0.8, 0.1, 0.0, 0.0], [0.1, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]]], dtype=D)


def test_dsnt_forward_backward_known_answer():
    """tests/test_nn.py:31-50."""
    p = SIMPLE_INPUT.clone().requires_grad_(True)
    out = tp.dsnt(p)
    assert torch.allclose(out, torch.tensor([[[0.4, 0.0]]], dtype=D), atol=1e-5)
    torch.nn.MSELoss()(out, torch.tensor([[[0.5, 0.5]]], dtype=D)).backward()
    assert torch.allclose(p.grad, SIMPLE_GRAD, atol=1e-5)
    # closed-form: d MSE/d coords = (coords - target) (mean over 2 elements, factor 2/2)
    res = cf.head(SIMPLE_INPUT[0].numpy(), None, reg='none', input_is_logits=False,
                  g_coords=np.array([[0.4 - 0.5, 0.0 - 0.5]]))
    assert np.allclose(res['coords'], [[0.4, 0.0]], atol=1e-12)
    assert np.allclose(res['dz'], SIMPLE_GRAD[0].numpy(), atol=1e-12)


def test_dsnt_batchless():
    """tests/test_nn.py:52-66 -- input without a batch dimension."""
    out = tp.dsnt(SIMPLE_INPUT[0])
    assert out.shape == (1, 2)
    assert torch.allclose(out, torch.tensor([[0.4, 0.0]], dtype=D), atol=1e-5)


def test_euclidean_loss_known_answer():
    """tests/test_nn.py:86-130."""
    a = torch.tensor([[[3.0, 4.0]] * 2] * 2, dtype=D, requires_grad=True)
    loss = tp.euclidean_loss(a, torch.zeros(2, 2, 2, dtype=D))
    loss.backward()
    assert abs(loss.item() - 5.0) < 1e-5
    assert torch.allclose(a.grad, torch.tensor([[[0.15, 0.20]] * 2] * 2, dtype=D), atol=1e-5)
    out = torch.tensor([[[0, 0], [1, 1], [0, 0]], [[1, 1], [0, 0], [0, 0]]], dtype=D)
    mask = torch.tensor([[1, 0, 1], [0, 1, 1]], dtype=D)
    assert tp.euclidean_loss(out, torch.zeros_like(out), mask).item() == 0.0


def test_thresholded_softmax_known_answer():
    """tests/test_nn.py:134-154."""
    exp = torch.tensor([0.26894142, 0, 0.73105858], dtype=D)
    assert torch.allclose(tp.thresholded_softmax(torch.tensor([2.0, 1.0, 3.0], dtype=D), 1.5), exp, atol=1e-5)
    got = tp.thresholded_softmax(torch.tensor([[2.0, 1, 3], [4, 0, 0]], dtype=D), 1.5)
    assert torch.allclose(got, torch.stack([exp, torch.tensor([1.0, 0, 0], dtype=D)]), atol=1e-5)
    assert np.allclose(cf.thresholded_softmax(np.array([[2.0, 1, 3], [4, 0, 0]]), 1.5), got.numpy(), atol=1e-12)
    torch.manual_seed(0)
    x = torch.randn(3, 20, dtype=D, requires_grad=True)
    assert torch.autograd.gradcheck(lambda t: tp.thresholded_softmax(t, 0), (x,))


def test_make_gauss_known_answer():
    """tests/test_nn.py:157-167 (tolerance 1e-4 there)."""
    g = tp.make_gauss(torch.tensor([0.0, 0.0], dtype=D), 5, 5, sigma=0.4)
    assert (g - GAUSS_5x5).abs().max() < 1e-4
    g2 = cf.gauss(np.zeros((1, 2)), 5, 5, 0.4)[0]
    assert np.abs(g2 - g.numpy()).max() < 1e-14


@pytest.mark.parametrize('name,shift_mean', [('kl', True), ('mse', True), ('js', True), ('var', False)])
def test_reg_loss_minimum_properties(name, shift_mean):
    """tests/test_nn.py:170-197,200-239."""
    fn = {'kl': tp.kl_reg_loss, 'mse': tp.mse_reg_loss, 'js': tp.js_reg_loss, 'var': tp.variance_reg_loss}[name]
    t_mean, t_std = torch.tensor([0.0, 0.0], dtype=D), 0.4

    def calc(mean, std):
        return fn(tp.make_gauss(mean, 5, 5, sigma=std), t_mean, t_std, mask=None).item()

    lo = calc(t_mean, t_std)
    assert abs(lo) < 1e-3
    assert calc(t_mean, t_std + 0.2) > lo + 1e-3
    assert calc(t_mean, t_std - 0.2) > lo + 1e-3
    if shift_mean:
        assert calc(t_mean + 0.1, t_std) > lo + 1e-3
        assert calc(t_mean - 0.1, t_std) > lo + 1e-3


def test_kl_mask_known_answer():
    """tests/test_nn.py:204-224."""
    coords = torch.tensor([[1.0, 1.0], [0.0, 0.0]], dtype=D)
    mask = torch.tensor([1.0, 0.0], dtype=D)
    got = tp.kl_reg_loss(KL_MASK_HM, coords, 1, mask).item()
    assert abs(got - 1.2228811717796824) < 1e-5
    res = cf.head(KL_MASK_HM.numpy(), coords.numpy(), mask.numpy(), reg='kl', sigma=1.0,
                  input_is_logits=False)
    assert abs(res['reg'] - 1.2228811717796824) < 1e-5


# --------------------------------------------------------------- golden vectors: fused head on logits
def _head_inputs(g, name):
    b, c, h, w, hm_sigma, coeff, with_mask = head_case_params(g, name)
    z = g[name + '/z']
    target = g[name + '/target']
    mask = g[name + '/mask'] if with_mask else None
    return (b, c, h, w, hm_sigma, coeff), z, target, mask


def test_golden_head_cases_present(golden_head):
    assert len(golden_head.cases) >= 9


@pytest.mark.parametrize('reg', REGS)
def test_torch_port_matches_golden_head(golden_head, reg):
    for name in golden_head.cases:
        (b, c, h, w, hm_sigma, coeff), z, target, mask = _head_inputs(golden_head, name)
        res = tp.head_loss_and_grad(torch.from_numpy(z), torch.from_numpy(target),
                                    None if mask is None else torch.from_numpy(mask),
                                    reg, hm_sigma, coeff, dtype=D)
        assert abs(res['loss'].item() - golden_head['%s/%s/loss' % (name, reg)]) < 1e-12, name
        assert rel_max(res['coords'].numpy(), golden_head[name + '/coords']) < 1e-12, name
        tol = 2e-7 if golden_head['%s/%s/dz' % (name, reg)].dtype == np.float32 else 1e-11
        assert rel_l2(res['dz'].numpy(), golden_head['%s/%s/dz' % (name, reg)]) < tol, name


@pytest.mark.parametrize('reg', REGS)
def test_closed_form_matches_golden_head(golden_head, reg):
    for name in golden_head.cases:
        (b, c, h, w, hm_sigma, coeff), z, target, mask = _head_inputs(golden_head, name)
        n = b * c
        res = cf.head(z.reshape(n, h, w).astype(np.float64), target.reshape(n, 2).astype(np.float64),
                      None if mask is None else mask.reshape(n).astype(np.float64),
                      reg=reg, sigma=2.0 * hm_sigma / w, reg_coeff=coeff)
        assert abs(res['loss'] - golden_head['%s/%s/loss' % (name, reg)]) < 1e-11, name
        assert abs(res['euclid'] - golden_head['%s/%s/euclid' % (name, reg)]) < 1e-11, name
        assert abs(res['reg'] - golden_head['%s/%s/reg' % (name, reg)]) < 1e-11, name
        assert rel_max(res['coords'], golden_head[name + '/coords'].reshape(n, 2)) < 1e-12, name
        gold = golden_head['%s/%s/dz' % (name, reg)].reshape(n, h, w)
        tol = 2e-7 if gold.dtype == np.float32 else 1e-9
        assert rel_l2(res['dz'], gold) < tol, (name, rel_l2(res['dz'], gold))


# --------------------------------------------------------------- golden vectors: level-1 API on P
def test_level1_dsnt_and_regs_match_golden(golden_l1):
    for name in golden_l1.cases:
        p = golden_l1[name + '/p'].astype(np.float64)
        mu = golden_l1[name + '/mu'].astype(np.float64)
        mask = golden_l1[name + '/mask'].astype(np.float64)
        sigma = float(golden_l1[name + '/sigma'])
        h, w = p.shape[-2:]
        n = p.size // (h * w)
        # torch port
        pt = torch.from_numpy(p).clone().requires_grad_(True)
        coords = tp.dsnt(pt)
        coords.backward(torch.from_numpy(golden_l1[name + '/dsnt/g_coords']))
        assert rel_max(coords.detach().numpy(), golden_l1[name + '/dsnt/coords']) < 1e-12
        assert rel_l2(pt.grad.numpy(), golden_l1[name + '/dsnt/dp']) < 1e-12
        # closed form
        res = cf.head(p.reshape(n, h, w), None, reg='none', input_is_logits=False,
                      g_coords=golden_l1[name + '/dsnt/g_coords'].reshape(n, 2))
        assert rel_l2(res['dz'], golden_l1[name + '/dsnt/dp'].reshape(n, h, w)) < 1e-12
        for reg, fn in (('var', tp.variance_reg_loss), ('kl', tp.kl_reg_loss), ('js', tp.js_reg_loss),
                        ('mse', tp.mse_reg_loss)):
            for mtag in ('mask', 'nomask'):
                mm = mask if mtag == 'mask' else None
                pt = torch.from_numpy(p).clone().requires_grad_(True)
                val = fn(pt, torch.from_numpy(mu), sigma, None if mm is None else torch.from_numpy(mm))
                val.backward()
                gl = float(golden_l1['%s/%s/%s/loss' % (name, reg, mtag)])
                gdp = golden_l1['%s/%s/%s/dp' % (name, reg, mtag)]
                assert abs(val.item() - gl) < 1e-12 * max(1, abs(gl)), (name, reg, mtag)
                assert rel_l2(pt.grad.numpy(), gdp) < 1e-11, (name, reg, mtag)
                # closed form: loss = reg only (no target for Euclid) -> use reg_coeff=1 and subtract euclid
                res = cf.head(p.reshape(n, h, w), mu.reshape(n, 2),
                              None if mm is None else mm.reshape(n), reg=reg, sigma=sigma,
                              input_is_logits=False, g_loss=1.0)
                assert abs(res['reg'] - gl) < 1e-11 * max(1, abs(gl)), (name, reg, mtag)
                # remove the Euclidean part of the closed-form gradient: a x + b y with g_loss weights
                res0 = cf.head(p.reshape(n, h, w), mu.reshape(n, 2), None if mm is None else mm.reshape(n),
                               reg='none', sigma=sigma, input_is_logits=False)
                assert rel_l2(res['dz'] - res0['dz'], gdp.reshape(n, h, w)) < 1e-9, (name, reg, mtag)


def test_level1_euclid_softmax_gauss_match_golden(golden_l1):
    g = golden_l1
    for tag in ('e2', 'e3', 'e_single'):
        actual = torch.from_numpy(g['euclid/%s/actual' % tag].astype(np.float64))
        target = torch.from_numpy(g['euclid/%s/target' % tag].astype(np.float64))
        mask = g.get('euclid/%s/mask' % tag)
        for mtag in ('mask', 'nomask'):
            mm = torch.from_numpy(mask.astype(np.float64)) if (mtag == 'mask' and mask is not None) else None
            a = actual.clone().requires_grad_(True)
            val = tp.euclidean_loss(a, target, mm)
            val.backward()
            assert abs(val.item() - float(g['euclid/%s/%s/loss' % (tag, mtag)])) < 1e-12
            assert rel_l2(a.grad.numpy(), g['euclid/%s/%s/grad' % (tag, mtag)]) < 1e-12
    z = torch.from_numpy(g['softmax2d/z'].astype(np.float64)).requires_grad_(True)
    out = tp.softmax_2d(z)
    out.backward(torch.from_numpy(g['softmax2d/g']))
    assert rel_max(out.detach().numpy(), g['softmax2d/out']) < 1e-12
    assert rel_l2(z.grad.numpy(), g['softmax2d/dz']) < 1e-11
    x = g['tsoftmax/x'].astype(np.float64)
    for tag in ('thr0', 'thrm05', 'thrinf'):
        thr = float(g['tsoftmax/%s/thr' % tag])
        xt = torch.from_numpy(x).clone().requires_grad_(True)
        out = tp.thresholded_softmax(xt, thr)
        out.backward(torch.from_numpy(g['tsoftmax/%s/g' % tag]))
        assert rel_max(out.detach().numpy(), g['tsoftmax/%s/out' % tag]) < 1e-12
        assert rel_l2(xt.grad.numpy(), g['tsoftmax/%s/dx' % tag]) < 1e-11
        o2 = cf.thresholded_softmax(x, thr)
        assert rel_max(o2, g['tsoftmax/%s/out' % tag]) < 1e-12
        assert rel_l2(cf.thresholded_softmax_grad(o2, g['tsoftmax/%s/g' % tag]), g['tsoftmax/%s/dx' % tag]) < 1e-11
    mu = torch.from_numpy(g['gauss/mu'].astype(np.float64)).requires_grad_(True)
    out = tp.make_gauss(mu, 9, 6, 0.25)
    out.backward(torch.from_numpy(g['gauss/g']))
    assert rel_max(out.detach().numpy(), g['gauss/out']) < 1e-12
    assert rel_l2(mu.grad.numpy(), g['gauss/dmu']) < 1e-11


def test_stacked_loss_matches_golden(golden_stacked):
    """Hourglass: sum of per-stack losses (src/dsnt/model.py:238-246)."""
    g = golden_stacked
    zs = [torch.from_numpy(g['z%d' % i].astype(np.float64)).requires_grad_(True) for i in range(3)]
    total, _ = tp.head_loss_stacked(zs, torch.from_numpy(g['target'].astype(np.float64)),
                                    torch.from_numpy(g['mask'].astype(np.float64)), 'js', 1.0, 1.0)
    total.backward()
    assert abs(total.item() - float(g['loss'])) < 1e-12
    for i, z in enumerate(zs):
        assert rel_l2(z.grad.numpy(), g['dz%d' % i]) < 1e-11


def test_all_zero_mask_and_none_mask_semantics():
    """all-zero mask -> loss 0; mask=None -> divide by numel."""
    torch.manual_seed(0)
    z = torch.randn(2, 3, 6, 6, dtype=D)
    t = torch.rand(2, 3, 2, dtype=D)
    res = tp.head_loss_and_grad(z, t, torch.zeros(2, 3, dtype=D), 'js', 1.0, 1.0)
    assert res['loss'].item() == 0.0 and res['dz'].abs().max().item() == 0.0
    a = tp.head_loss_and_grad(z, t, None, 'kl', 1.0, 1.0)
    b = tp.head_loss_and_grad(z, t, torch.ones(2, 3, dtype=D), 'kl', 1.0, 1.0)
    assert math.isclose(a['loss'].item(), b['loss'].item(), rel_tol=1e-14)
