"""Level-2 boundary: the `HumanPoseModel` head surface (forward_part2 / forward_loss / compute_coords /
.heatmaps) on the fused kernels -- mirrors /root/reference/tests/test_model.py (shapes, one SGD step moves
every parameter) with a small stand-in backbone, since the backbones themselves are out of scope."""

import numpy as np
import pytest
import torch
from torch import nn as tnn

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


class TinyPoseModel(tnn.Module):
    """Carries the attributes the reference models carry (src/dsnt/model.py:90-99) around a 2-layer FCN."""

    def __init__(self, stacks=0, preact='softmax', reg='js', reg_coeff=1.0, hm_sigma=1.0, n_chans=16):
        super().__init__()
        self.n_chans, self.output_strat, self.preact = n_chans, 'dsnt', preact
        self.reg, self.reg_coeff, self.hm_sigma = reg, reg_coeff, hm_sigma
        self.stacks = stacks
        self.fcn = tnn.Sequential(tnn.Conv2d(3, 8, 3, stride=2, padding=1), tnn.ReLU(), tnn.Conv2d(8, 8, 3, padding=1))
        self.hm_convs = tnn.ModuleList([tnn.Conv2d(8, n_chans, 1, bias=False) for _ in range(max(stacks, 1))])

    def forward_part1(self, x):
        f = self.fcn(x)
        outs = [conv(f) for conv in self.hm_convs]
        return outs if self.stacks else outs[0]


def make(dp, **kw):
    torch.manual_seed(0)
    return dp.attach_fused_head(TinyPoseModel(**kw).to(DEV))


@pytest.fixture(scope='module')
def dp():
    import dsnt_pose2d_b200
    return dsnt_pose2d_b200


def test_shapes_and_heatmaps_like_reference_test_model(dp):
    model = make(dp)
    out = model(torch.randn(2, 3, 56, 56, device=DEV))
    assert out.shape == (2, 16, 2)
    hm = model.heatmaps
    assert hm.shape == (2, 16, 28, 28)
    assert torch.allclose(hm.flatten(-2).sum(-1), torch.ones(2, 16, device=DEV), atol=1e-5)
    coords = model.compute_coords(out)
    assert coords.device.type == 'cpu' and coords.dtype == torch.float32 and coords.shape == (2, 16, 2)


@pytest.mark.parametrize('reg', ['none', 'var', 'kl', 'js', 'mse'])
def test_model_loss_and_backbone_gradients_match_oracle(dp, reg):
    from oracle import torch_port as tp
    model = make(dp, reg=reg, reg_coeff=0.7, hm_sigma=1.3)
    x = torch.randn(3, 3, 32, 32, device=DEV)
    target = torch.rand(3, 16, 2, device=DEV) * 1.6 - 0.8
    mask = (torch.rand(3, 16, device=DEV) > 0.2).float()
    grabbed = {}

    def grab(_module, _inp, output):                 # the logits entering the head, and their gradient
        output.retain_grad()
        grabbed['z'] = output

    hook = model.hm_convs[0].register_forward_hook(grab)
    out = model(x)
    hook.remove()
    loss = model.forward_loss(out, target, mask)
    loss.backward()
    z = grabbed['z']
    ref = tp.head_loss_and_grad(z.detach(), target, mask, reg, 1.3, 0.7)
    assert abs(loss.item() - ref['loss'].item()) / abs(ref['loss'].item()) < TOL
    assert rel_l2(z.grad.cpu().double().numpy(), ref['dz'].numpy()) < TOL      # dL/dZ handed to the backbone
    for p_name, p in model.named_parameters():                                # ... and it reaches every parameter
        assert p.grad is not None and torch.isfinite(p.grad).all(), p_name
    assert model.hm_convs[0].weight.grad.abs().max().item() > 0


def test_training_step_changes_every_parameter(dp):
    """tests/test_model.py:39-63: reg='js', mask_var=None, one SGD step updates all parameter groups."""
    model = make(dp, reg='js')
    old = [p.detach().clone() for p in model.parameters()]
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    x = torch.rand(1, 3, 28, 28, device=DEV)
    target = torch.rand(1, 16, 2, device=DEV) * 2 - 1
    out = model(x)
    loss = model.forward_loss(out, target, mask_var=None)
    loss.backward()
    opt.step()
    for p, o in zip(model.parameters(), old):
        assert not torch.equal(p.detach(), o)


def test_hourglass_style_stacks(dp):
    """List in / list out, losses summed over stacks, .heatmaps = first stack, compute_coords = last stack
    (src/dsnt/model.py:229-246,262-267,286-292)."""
    from oracle import torch_port as tp
    model = make(dp, stacks=3, reg='js')
    x = torch.randn(2, 3, 32, 32, device=DEV)
    target = torch.rand(2, 16, 2, device=DEV) * 1.6 - 0.8
    mask = torch.ones(2, 16, device=DEV)
    outs = model(x)
    assert isinstance(outs, list) and len(outs) == 3
    loss = model.forward_loss(outs, target, mask)
    loss.backward()
    zs = [z.detach() for z in model.forward_part1(x)]
    total, coords = tp.head_loss_stacked([z.cpu().double() for z in zs], target.cpu().double(), mask.cpu().double(),
                                         'js', 1.0, 1.0)
    assert abs(loss.item() - total.item()) / total.item() < TOL
    assert torch.allclose(model.compute_coords(outs).double(), coords[-1], atol=TOL)
    assert torch.allclose(model.heatmaps.cpu().double(), tp.hm_preact_softmax(zs[0].cpu().double()), atol=1e-6)
    assert len(model.heatmaps_array) == 3


def test_foreign_coords_fall_back_to_reference_composition(dp):
    """forward_loss on coords that did not come from forward_part2 (e.g. flip-averaged, inference.py:38-48)."""
    from oracle import torch_port as tp
    model = make(dp, reg='kl')
    x = torch.randn(2, 3, 32, 32, device=DEV)
    target = torch.rand(2, 16, 2, device=DEV) - 0.5
    out = model(x)
    other = out.detach().clone().requires_grad_(True)
    loss = model.forward_loss(other, target, None)
    loss.backward()
    z = model.forward_part1(x).detach()
    ref, _, _, _ = tp.head_loss(z.cpu().double(), target.cpu().double(), None, 'kl', 1.0, 1.0)
    assert abs(loss.item() - ref.item()) / ref.item() < TOL
    assert other.grad is not None


@pytest.mark.parametrize('preact', ['thresholded_softmax', 'abs', 'relu', 'sigmoid'])
def test_other_preactivations(dp, preact):
    """src/dsnt/model.py:31-41 -- not fused yet, composed from the level-1 operators."""
    model = make(dp, preact=preact, reg='js')
    x = torch.randn(2, 3, 32, 32, device=DEV)
    out = model(x)
    z = model.forward_part1(x).detach().cpu().double()
    flat = z.view(-1, 16 * 16)
    if preact == 'thresholded_softmax':
        from oracle import torch_port as tp
        p = tp.thresholded_softmax(flat, -0.5)
    else:
        a = {'abs': torch.abs, 'relu': torch.relu, 'sigmoid': torch.sigmoid}[preact](flat)
        p = a / (a.sum(-1, keepdim=True) + 1e-12)
    from oracle import torch_port as tp
    ref = tp.dsnt(p.view(2, 16, 16, 16))
    assert torch.allclose(out.detach().cpu().double(), ref, atol=2e-5)
    loss = model.forward_loss(out, torch.zeros(2, 16, 2, device=DEV), None)
    loss.backward()
    assert all(p.grad is not None for p in model.parameters())


def test_install_as_dsnt_nn_binding(dp):
    import sys
    saved = {k: sys.modules.get(k) for k in ('dsnt', 'dsnt.nn')}
    try:
        mod = dp.install_as_dsnt_nn()
        import dsnt.nn as bound
        assert bound is mod and hasattr(bound, 'js_reg_loss') and hasattr(bound, 'flat_softmax')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
