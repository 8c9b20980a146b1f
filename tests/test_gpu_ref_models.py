"""The reference's OWN model classes (src/dsnt/model.py: ResNetHumanPoseModel, HourglassHumanPoseModel) driven on the GPU
with this library's head -- BASELINE configs 2 and 3 -- against the same model objects with the reference's own head.

The unmodified reference is pip-installed under baseline/_ref by tools/install_ref.sh (git-ignored, travels to the GPU box);
without it these tests skip.  `dsnt.model` imports `dsnt.data`, which needs anibali/torchdata's `torchdata.mpii`
(requirements.txt:18; absent from the image): a five-name stub is registered for it, as SURVEY.md 8c describes.
Mirrors /root/reference/tests/test_model.py:11-63 (shapes with truncate / dilate, one SGD step moves every parameter).

Tolerances: coordinates 1e-5 max-abs, loss 1e-5 relative, dL/dZ at the head input 1e-5 L2-relative, against the reference's
head evaluated in fp64 on the same logits; parameter gradients: no worse than the reference's own fp32 head (see _compare)."""

import copy
import os
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
DEV = 'cuda:0'

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'dsnt', 'model.py')),
                                 reason='baseline/_ref is absent (tools/install_ref.sh needs /root/reference)')]

STUB = '''
import sys, types
mp = types.ModuleType('torchdata.mpii')
class MpiiData:            # src/dsnt/data.py:15 only needs the names at import time
    def __init__(self, *a, **k): raise RuntimeError('stub')
mp.MpiiData = MpiiData
mp.MPII_Joint_Horizontal_Flips = [5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10]
mp.MPII_Image_Mean = [0.44, 0.40, 0.37]
mp.MPII_Image_Stddev = [0.25, 0.24, 0.24]
mp.transform_keypoints = lambda kp, m: kp
try:
    import torchdata as td
except ImportError:
    td = types.ModuleType('torchdata'); td.__path__ = []; sys.modules['torchdata'] = td
td.mpii = mp
sys.modules['torchdata.mpii'] = mp
'''


@pytest.fixture(scope='module')
def ref_model():
    """The reference's dsnt.model, bound to the reference's own dsnt.nn."""
    exec(STUB, {})
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings('ignore')
    for name in [m for m in sys.modules if m == 'dsnt' or m.startswith('dsnt.')]:
        del sys.modules[name]
    import dsnt.model as rm
    assert os.path.realpath(rm.__file__).startswith(os.path.realpath(REF))
    assert os.path.realpath(rm.dsnt.nn.__file__).startswith(os.path.realpath(REF))      # the reference's own operators
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    return rm


@pytest.fixture(scope='module')
def dp():
    import dsnt_pose2d_b200
    return dsnt_pose2d_b200


def _rel(a, b, floor=0.0):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(max(floor, 1e-30))).item()


def _train_step(model, x, target, mask, head_dtype=None):
    """model(x) -> forward_loss -> backward, as src/dsnt/bin/train.py:355-381; returns what the comparison needs.
    head_dtype=torch.float64 evaluates the HEAD (forward_part2 + forward_loss) in fp64 on the fp32 logits of the fp32
    backbone: the arbiter for the head alone, with the backbone's own rounding out of the picture."""
    model.zero_grad(set_to_none=True)
    z = model.forward_part1(x)
    zs = z if isinstance(z, (list, tuple)) else [z]
    for t in zs:
        t.retain_grad()
    if head_dtype is not None:
        zin = [t.to(head_dtype) for t in zs]
        zin = zin if isinstance(z, (list, tuple)) else zin[0]
        target, mask = target.to(head_dtype), mask.to(head_dtype)
    else:
        zin = z
    out = model.forward_part2(zin)
    loss = model.forward_loss(out, target, mask)
    loss.backward()
    torch.cuda.synchronize()
    coords = model.compute_coords(out)
    return {'loss': loss.item(), 'coords': coords.cpu(), 'dz': [t.grad.detach().cpu() for t in zs],
            'grads': {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None},
            'heatmaps': model.heatmaps.detach().cpu()}


def _errors(a, b):
    e_loss = abs(a['loss'] - b['loss']) / abs(b['loss'])
    e_coords = (a['coords'].double() - b['coords'].double()).abs().max().item()
    e_dz = max(_rel(x, y) for x, y in zip(a['dz'], b['dz']))
    assert set(a['grads']) == set(b['grads']) and len(b['grads']) > 0
    # a conv bias in front of a batch norm has an analytically ZERO gradient: its relative error means nothing
    # (its computed value is pure rounding noise of the backbone): per-parameter norms are floored at 1 % of the largest
    # parameter-gradient norm, and the error over ALL parameters taken as one vector is reported as well
    floor = 1e-2 * max(g.double().norm().item() for g in b['grads'].values())
    worst = max(((_rel(a['grads'][n], b['grads'][n], floor), n) for n in b['grads']))
    num = sum((a['grads'][n].double() - b['grads'][n].double()).pow(2).sum().item() for n in b['grads']) ** 0.5
    den = sum(b['grads'][n].double().pow(2).sum().item() for n in b['grads']) ** 0.5
    worst = (max(worst[0], num / den), worst[1] + '; all parameters as one vector %.2e' % (num / den))
    e_hm = (a['heatmaps'].double() - b['heatmaps'].double()).abs().max().item()
    return e_loss, e_coords, e_dz, worst, e_hm


def _compare(ours, ref32, ref64, what):
    """ours (fused head) against the same model object with the reference's own head, both on the same fp32 backbone, and
    both against that model with the reference's head evaluated in fp64 (the arbiter for the head).  The head's outputs --
    loss, coords, dL/dZ at the head input -- carry north_star's 1e-5 bar against the arbiter.  A parameter gradient is
    dL/dZ pushed through dozens of cuDNN layers, which amplifies any rounding of the head by the backbone's conditioning:
    the bar there is the reference head's OWN fp32 rounding pushed through the same backbone (printed beside ours) --
    ours-vs-arbiter <= 1.5 x reference-vs-arbiter (or 2e-5 where that is smaller): the fused head adds no noise."""
    o32 = _errors(ours, ref32)
    o64 = _errors(ours, ref64)
    r64 = _errors(ref32, ref64)
    print('%s\n  ours vs reference head fp32          : loss %.2e coords %.2e dZ %.2e worst parameter gradient %.2e (%s); heatmaps %.2e'
          % ((what,) + o32[:3] + (o32[3][0], o32[3][1], o32[4])))
    print('  ours vs reference head in fp64       : loss %.2e coords %.2e dZ %.2e worst parameter gradient %.2e (%s)'
          % (o64[:3] + (o64[3][0], o64[3][1])))
    print('  reference head fp32 vs the same fp64 : loss %.2e coords %.2e dZ %.2e worst parameter gradient %.2e (%s) of %d parameters'
          % (r64[:3] + (r64[3][0], r64[3][1], len(ref64['grads']))))
    assert o64[0] < 1e-5 and o64[1] < 1e-5 and o64[2] < 1e-5 and o64[4] < 1e-6          # the head itself, against the arbiter
    assert o32[0] < 1e-5 and o32[1] < 1e-5 and o32[2] < 1e-5
    bar = max(2e-5, 1.5 * r64[3][0])
    assert o64[3][0] < bar, (o64[3], r64[3])


def _three_way(ref, dp, x, target, mask):
    """the reference's head in fp64 on the fp32 backbone (arbiter), the reference in fp32, ours: one backbone, three heads."""
    import gc
    ours = dp.attach_fused_head(copy.deepcopy(ref))
    r64 = _train_step(ref, x, target, mask, head_dtype=torch.float64)
    r32 = _train_step(ref, x, target, mask)
    ref.zero_grad(set_to_none=True)
    for name in ('heatmaps_array', 'heatmaps'):          # drop the reference's materialised heatmaps (hourglass: a property)
        try:
            setattr(ref, name, None)
        except AttributeError:
            pass
    gc.collect()
    torch.cuda.empty_cache()
    o32 = _train_step(ours, x, target, mask)
    return ours, o32, r32, r64


def _head_cost(dp, model, x, target, mask, steps=5):
    """Head time and launches inside the full training step (CUDA events around every entry point of the library)."""
    from dsnt_pose2d_b200 import _lib
    for _ in range(2):
        _train_step(model, x, target, mask)
    names = [n for n in _lib.SIGNATURES if n.startswith('dsnt_') and 'supported' not in n and 'bytes' not in n
             and n not in ('dsnt_b200_version', 'dsnt_b200_last_error')]
    _lib.event_log = {n: [] for n in names}
    before = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        _train_step(model, x, target, mask)
    e1.record()
    torch.cuda.synchronize()
    logs, _lib.event_log = _lib.event_log, None
    head_us = sum(sum(a.elapsed_time(b) for a, b in v) for v in logs.values()) / steps * 1e3
    used = {k: len(v) // steps for k, v in logs.items() if v}
    return head_us, (_lib.launch_count - before) / steps, e0.elapsed_time(e1) / steps, used


def test_resnet34_cfg2_matches_reference_head(ref_model, dp):
    """BASELINE config 2: ResNet-34, dilate=2 (28x28 heatmaps), batch 64 at 224x224, JS regulariser."""
    import torchvision
    torch.manual_seed(0)
    ref = ref_model.ResNetHumanPoseModel(torchvision.models.resnet34(), n_chans=16, dilate=2, output_strat='dsnt',
                                         reg='js', reg_coeff=1.0, hm_sigma=1.0).to(DEV)
    x = torch.rand(64, 3, 224, 224, device=DEV)
    target = torch.rand(64, 16, 2, device=DEV) * 1.6 - 0.8
    mask = (torch.rand(64, 16, device=DEV) > 0.1).float()
    ours, o32, r32, r64 = _three_way(ref, dp, x, target, mask)
    assert type(ours).__name__ == 'ResNetHumanPoseModelB200' and isinstance(ours, ref_model.ResNetHumanPoseModel)
    assert o32['heatmaps'].shape == (64, 16, 28, 28)
    _compare(o32, r32, r64, 'resnet34 dilate=2 batch 64 (cfg 2)')
    head_us, launches, step_ms, used = _head_cost(dp, ours, x, target, mask)
    print('cfg 2 full training step %.1f ms; head: %.0f us in %d launches per step %r' % (step_ms, head_us, launches, used))


def test_hourglass8_cfg3_matches_reference_head(ref_model, dp):
    """BASELINE config 3: 8-stack hourglass, 64x64 heatmaps per stack, JS sigma = 1, batch 32 at 256x256."""
    torch.manual_seed(0)
    ref = ref_model.build_mpii_pose_model('hg8', output_strat='dsnt', reg='js', reg_coeff=1.0, hm_sigma=1.0).to(DEV)
    x = torch.rand(32, 3, 256, 256, device=DEV)
    target = torch.rand(32, 16, 2, device=DEV) * 1.6 - 0.8
    mask = (torch.rand(32, 16, device=DEV) > 0.1).float()
    ours, o32, r32, r64 = _three_way(ref, dp, x, target, mask)
    assert len(r32['dz']) == 8 and r32['heatmaps'].shape == (32, 16, 64, 64)
    _compare(o32, r32, r64, 'hourglass hg8 batch 32 (cfg 3)')
    head_us, launches, step_ms, used = _head_cost(dp, ours, x, target, mask, steps=3)
    print('cfg 3 full training step %.1f ms; head: %.0f us in %d launches per step %r' % (step_ms, head_us, launches, used))
    # forward_part2: ONE coordinate launch for the 8 stacks; forward_loss: ONE fused launch; backward: the scale check;
    # + the lazily materialised `.heatmaps` this test reads (one softmax launch): 4 launches against the 24+ of per-stack calls
    assert launches <= 4


@pytest.mark.parametrize('kw,hm', [({'truncate': 1}, 14), ({'dilate': 2}, 28), ({}, 7)])
def test_shapes_like_reference_test_model(ref_model, dp, kw, hm):
    """tests/test_model.py:11-39 with the fused head attached."""
    import torchvision
    model = dp.attach_fused_head(ref_model.ResNetHumanPoseModel(torchvision.models.resnet18(), n_chans=16, **kw).to(DEV))
    sz = model.image_specs.size
    assert sz == 224
    out = model(torch.randn(1, 3, sz, sz, device=DEV))
    assert out.shape == (1, 16, 2)
    assert model.heatmaps.shape == (1, 16, hm, hm)
    c = model.compute_coords(out)
    assert c.device.type == 'cpu' and c.dtype == torch.float32


def test_training_step_moves_every_parameter(ref_model, dp):
    """tests/test_model.py:41-63: resnet18, reg='js', mask None, one SGD step."""
    import torchvision
    torch.manual_seed(1)
    model = dp.attach_fused_head(ref_model.ResNetHumanPoseModel(torchvision.models.resnet18(), n_chans=16,
                                                                output_strat='dsnt', reg='js').to(DEV))
    old = [p.detach().clone() for p in model.parameters()]
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    x = torch.rand(1, 3, 224, 224, device=DEV)
    target = torch.rand(1, 16, 2, device=DEV) * 2 - 1
    loss = model.forward_loss(model(x), target, None)
    loss.backward()
    opt.step()
    for p, o in zip(model.parameters(), old):
        assert not torch.equal(p.detach(), o)


INSTALL_SCRIPT = STUB + '''
import sys, warnings
warnings.filterwarnings('ignore')
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import torch, torchvision
import dsnt_pose2d_b200 as dp
ours_nn = dp.install_as_dsnt_nn()                 # BEFORE dsnt.model is imported (SURVEY.md 8b binding note)
import dsnt.model as rm
assert rm.dsnt.nn is ours_nn and rm.euclidean_loss is ours_nn.euclidean_loss
assert rm.__file__.startswith(%(ref)r)
from oracle import torch_port as tp
torch.manual_seed(0)
for build in ('resnet', 'hg2'):
    if build == 'resnet':
        model = rm.ResNetHumanPoseModel(torchvision.models.resnet18(), n_chans=16, dilate=2, reg='js').cuda()
        x = torch.rand(2, 3, 224, 224, device='cuda')
    else:
        model = rm.build_mpii_pose_model('hg2', output_strat='dsnt', reg='js').cuda()
        x = torch.rand(2, 3, 256, 256, device='cuda')
    target = torch.rand(2, 16, 2, device='cuda') * 1.6 - 0.8
    mask = (torch.rand(2, 16, device='cuda') > 0.2).float()
    z = model.forward_part1(x)
    zs = z if isinstance(z, list) else [z]
    for t in zs: t.retain_grad()
    out = model.forward_part2(z)                   # the reference's own code, our operators underneath
    loss = model.forward_loss(out, target, mask)
    loss.backward()
    want = 0.0
    for i, t in enumerate(zs):
        r = tp.head_loss_and_grad(t.detach(), target, mask, 'js', 1.0, 1.0, dtype=torch.float64)
        want += r['loss'].item()
        if i == len(zs) - 1:       # an earlier stack's logits also feed the next stack: only the last one's gradient is the head's alone
            e = ((t.grad.cpu().double() - r['dz']).norm() / r['dz'].norm()).item()
            assert e < 1e-5, (build, 'dz', e)
    assert abs(loss.item() - want) / want < 1e-5, (build, loss.item(), want)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    print(build, 'ok', loss.item(), want)
print('INSTALL-OK')
'''


def test_install_as_dsnt_nn_runs_the_reference_model_code_on_our_operators():
    """`install_as_dsnt_nn()` before `import dsnt.model`: the reference's unmodified forward_part2 / forward_loss call this
    library's level-1 operators (dsnt, euclidean_loss, js_reg_loss); loss and dL/dZ against the fp64 oracle."""
    script = INSTALL_SCRIPT % {'root': ROOT, 'ref': REF}
    r = subprocess.run([sys.executable, '-c', script], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and 'INSTALL-OK' in r.stdout
