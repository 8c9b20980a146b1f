"""Parity of the sm_100a kernels (called through the C ABI via the package) against
  (a) the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  (b) the fp64 CPU oracle on seeded synthetic inputs at the BASELINE shapes.

Tolerance (north_star): 1e-5 relative for fp32 coordinates, losses and gradients, measured against the
fp64 arbiter -- coords: max-abs (values live in [-1,1]); loss terms: relative; gradients: L2-relative per
tensor AND max-abs/max-abs.  bf16 input: the oracle is evaluated on the bf16-rounded logits; coords/loss
keep the 1e-5 bar, dZ is emitted in bf16 so its bar is 4e-3 L2-relative (1 bf16 ulp = 2^-8 per element).
"""

import numpy as np
import pytest
import torch

from conftest import head_case_params, rel_l2, rel_max

pytestmark = pytest.mark.gpu

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


def run_head(dp, z, target, mask, reg, hm_sigma=1.0, coeff=1.0, variant=0):
    zz = z.detach().clone().to(DEV).requires_grad_(True)
    tt = None if target is None else target.to(DEV)
    mm = None if mask is None else mask.to(DEV)
    out = dp.dsnt_head(zz, tt, mm, reg=reg, hm_sigma=hm_sigma, reg_coeff=coeff, variant=variant)
    out.loss.backward()
    torch.cuda.synchronize()
    return {'loss': out.loss.item(), 'euclid': out.euclid.item(), 'reg': out.reg.item(),
            'coords': out.coords.detach().cpu().double().numpy(), 'dz': zz.grad.detach().cpu().double().numpy()}


def check(got, ref_loss, ref_coords, ref_dz, what, tol=TOL, dz_tol=None, ref32=None):
    dz_tol = tol if dz_tol is None else dz_tol
    e_loss = abs(got['loss'] - ref_loss) / max(abs(ref_loss), 1e-30)
    e_coords = float(np.abs(got['coords'] - ref_coords).max())
    e_l2 = rel_l2(got['dz'], ref_dz)
    e_max = rel_max(got['dz'], ref_dz)
    extra = ''
    if ref32 is not None:
        extra = ' | reference-fp32 vs fp64: loss %.1e dz %.1e' % ref32
    print('%-42s loss %.2e coords %.2e dz L2 %.2e max %.2e%s' % (what, e_loss, e_coords, e_l2, e_max, extra))
    assert e_loss < tol, (what, 'loss', got['loss'], ref_loss)
    assert e_coords < tol, (what, 'coords', e_coords)
    assert e_l2 < dz_tol and e_max < dz_tol * 4, (what, 'dz', e_l2, e_max)


# ------------------------------------------------------------------------------------------- golden
@pytest.mark.parametrize('reg', REGS)
def test_fused_head_matches_reference_golden(dp, golden_head, reg):
    for name in golden_head.cases:
        b, c, h, w, hm_sigma, coeff, with_mask = head_case_params(golden_head, name)
        z = torch.from_numpy(golden_head[name + '/z'])
        target = torch.from_numpy(golden_head[name + '/target'])
        mask = torch.from_numpy(golden_head[name + '/mask']) if with_mask else None
        got = run_head(dp, z, target, mask, reg, hm_sigma, coeff)
        check(got, float(golden_head['%s/%s/loss' % (name, reg)]), golden_head[name + '/coords'],
              golden_head['%s/%s/dz' % (name, reg)].astype(np.float64), 'golden %s %s' % (name, reg))
        assert abs(got['euclid'] - float(golden_head['%s/%s/euclid' % (name, reg)])) < TOL
        assert abs(got['reg'] - float(golden_head['%s/%s/reg' % (name, reg)])) < TOL * max(1.0, abs(got['reg']))


def test_stacked_hourglass_loss_matches_golden(dp, golden_stacked):
    g = golden_stacked
    zs = [torch.from_numpy(g['z%d' % i]).to(DEV).requires_grad_(True) for i in range(3)]
    coords, total = dp.dsnt_head_stacked(zs, torch.from_numpy(g['target']).to(DEV),
                                         torch.from_numpy(g['mask']).to(DEV), reg='js', hm_sigma=1.0, reg_coeff=1.0)
    total.backward()
    assert abs(total.item() - float(g['loss'])) / float(g['loss']) < TOL
    for i, z in enumerate(zs):
        assert rel_l2(z.grad.cpu().double().numpy(), g['dz%d' % i]) < TOL


@pytest.mark.parametrize('reg', REGS)
@pytest.mark.parametrize('shape,dtype', [((8, 2, 16, 64, 64), torch.float32), ((3, 2, 5, 28, 28), torch.float32),
                                         ((2, 1, 3, 7, 7), torch.float32), ((4, 1, 4, 256, 256), torch.float32),
                                         ((8, 2, 16, 64, 64), torch.bfloat16)])
def test_stacked_launch_equals_per_stack_calls(dp, reg, shape, dtype):
    """One launch over all hourglass stacks (cfg 3 has 8) must give exactly what the per-stack calls give:
    same kernels, same per-heatmap arithmetic -- coords and dZ bitwise, the summed loss to rounding."""
    stacks, b, c, h, w = shape
    gen = torch.Generator().manual_seed(31)
    zs = [(torch.randn(b, c, h, w, generator=gen) * 2).to(dtype).to(DEV) for _ in range(stacks)]
    target = (torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8).to(DEV)
    mask = (torch.rand(b, c, generator=gen) > 0.2).float().to(DEV)
    wts = [torch.randn(b, c, 2, generator=gen).to(DEV) for _ in range(stacks)]
    za = [z.clone().requires_grad_(True) for z in zs]
    coords_a, total_a = dp.dsnt_head_stacked(za, target, mask, reg=reg, hm_sigma=1.0, reg_coeff=0.7)
    (total_a + sum((ca * wt).sum() for ca, wt in zip(coords_a, wts))).backward()
    zb = [z.clone().requires_grad_(True) for z in zs]
    outs = [dp.dsnt_head(z, target, mask, reg=reg, hm_sigma=1.0, reg_coeff=0.7) for z in zb]
    total_b = sum(o.loss for o in outs)
    (total_b + sum((o.coords * wt).sum() for o, wt in zip(outs, wts))).backward()
    assert abs(total_a.item() - total_b.item()) <= 2e-6 * abs(total_b.item())
    for i in range(stacks):
        assert torch.equal(coords_a[i], outs[i].coords), i
        assert torch.equal(za[i].grad, zb[i].grad), i


# ------------------------------------------------------------------------------------------- synthetic, BASELINE shapes
def synth(b, c, h, w, scale, seed=0, trained=False, tp=None):
    gen = torch.Generator().manual_seed(seed)
    target = torch.rand(b, c, 2, generator=gen) * 1.6 - 0.8
    if trained:
        g = tp.make_gauss(target + 0.05 * torch.randn(b, c, 2, generator=gen), w, h, 2.0 / w)
        z = (g + 1e-6).log() + 0.1 * torch.randn(b, c, h, w, generator=gen)
    else:
        z = torch.randn(b, c, h, w, generator=gen) * scale
    mask = (torch.rand(b, c, generator=gen) > 0.1).float()
    return z.float(), target.float(), mask


def oracle(tp, z, target, mask, reg, hm_sigma=1.0, coeff=1.0, dtype=torch.float64):
    r = tp.head_loss_and_grad(z, target, mask, reg, hm_sigma, coeff, dtype=dtype)
    return r['loss'].item(), r['coords'].double().numpy(), r['dz'].double().numpy()


@pytest.mark.parametrize('reg', REGS)
@pytest.mark.parametrize('kind', ['randn', 'randn_x5', 'trained'])
def test_cfg1_shape_vs_fp64_oracle(dp, tp, reg, kind):
    """BASELINE cfg 1: 32 x 16 joints x 64x64 fp32, Euclid + reg, mask."""
    z, target, mask = synth(32, 16, 64, 64, 5.0 if kind == 'randn_x5' else 1.0, trained=(kind == 'trained'), tp=tp)
    l64, c64, d64 = oracle(tp, z, target, mask, reg)
    l32, c32, d32 = oracle(tp, z, target, mask, reg, dtype=torch.float32)
    got = run_head(dp, z, target, mask, reg)
    check(got, l64, c64, d64, 'cfg1 %s %s' % (kind, reg), ref32=(abs(l32 - l64) / abs(l64), rel_l2(d32, d64)))


@pytest.mark.parametrize('reg', ['js', 'var'])
def test_cfg2_head_shape_28x28(dp, tp, reg):
    """BASELINE cfg 2 head: 64 x 16 x 28x28 (ResNet-34, dilate=2)."""
    z, target, mask = synth(64, 16, 28, 28, 1.0, seed=1)
    l64, c64, d64 = oracle(tp, z, target, mask, reg)
    check(run_head(dp, z, target, mask, reg), l64, c64, d64, 'cfg2 28x28 %s' % reg)


@pytest.mark.parametrize('reg', REGS)
def test_cfg5_shape_256x256_streaming_kernel(dp, tp, reg):
    """BASELINE cfg 5 shape (256x256, variance regulariser) on a small batch; all regularisers."""
    z, target, mask = synth(2, 4, 256, 256, 1.0, seed=2, trained=(reg in ('kl', 'js')), tp=tp)
    l64, c64, d64 = oracle(tp, z, target, mask, reg)
    check(run_head(dp, z, target, mask, reg), l64, c64, d64, 'cfg5 256x256 %s' % reg)


@pytest.mark.parametrize('shape', [(3, 5, 56, 56), (2, 3, 128, 128), (2, 2, 14, 14), (2, 3, 7, 7), (1, 2, 9, 33),
                                   (2, 2, 12, 20), (1, 3, 100, 60)])
def test_assorted_sizes_js_and_var(dp, tp, shape):
    """Sizes the reference uses (7,14,28,56,64) plus odd / non-square ones (scalar path, chunked backward)."""
    for reg in ('js', 'var', 'kl'):
        z, target, mask = synth(*shape, 2.0, seed=3)
        l64, c64, d64 = oracle(tp, z, target, mask, reg, hm_sigma=1.5)
        check(run_head(dp, z, target, mask, reg, hm_sigma=1.5), l64, c64, d64, 'size %s %s' % (shape, reg))


@pytest.mark.parametrize('reg', REGS)
@pytest.mark.parametrize('variant', [1, 2, 3])
def test_alternative_kernels_agree_with_oracle(dp, tp, reg, variant):
    """variant=1 forces the two-pass L2 kernel, variant=2 the register-resident kernels (forward and backward),
    variant=3 the generic streaming kernels (head_stream.cuh) on a 64x64 map; the default (variant 0) is the
    tuned fast path (head_fast.cuh).  All must match the oracle."""
    z, target, mask = synth(4, 4, 64, 64, 3.0, seed=4)
    l64, c64, d64 = oracle(tp, z, target, mask, reg)
    check(run_head(dp, z, target, mask, reg, variant=variant), l64, c64, d64, 'variant %d 64x64 %s' % (variant, reg))


@pytest.mark.parametrize('reg', ['kl', 'js'])
@pytest.mark.parametrize('hm_sigma', [0.5, 1.0, 4.0, 40.0])
@pytest.mark.parametrize('where', ['inside', 'edge', 'outside', 'far'])
def test_gaussian_window_is_exact_for_any_sigma_and_target(dp, tp, reg, hm_sigma, where):
    """The divergence is evaluated only on the window where the target Gaussian is non-negligible; the window
    is derived from sigma and the target, so tiny / huge sigma and targets at or beyond the border must give
    the same answer as the dense reference arithmetic."""
    z, target, mask = synth(3, 4, 64, 64, 2.0, seed=11)
    if where == 'edge':
        target = target.sign() * 0.99
    elif where == 'outside':
        target = target.sign() * 1.3
    elif where == 'far':
        target = target.sign() * 25.0
    l64, c64, d64 = oracle(tp, z, target, mask, reg, hm_sigma=hm_sigma)
    check(run_head(dp, z, target, mask, reg, hm_sigma=hm_sigma), l64, c64, d64,
          'window %s sigma=%g %s' % (reg, hm_sigma, where))


@pytest.mark.parametrize('reg', ['kl', 'js', 'mse'])
@pytest.mark.parametrize('shape,hm_sigma', [((2, 3, 64, 64), 1.0), ((2, 3, 64, 64), 2.5), ((1, 2, 128, 128), 1.0),
                                            ((1, 2, 256, 256), 1.0), ((1, 2, 256, 256), 3.0), ((3, 2, 32, 32), 0.7),
                                            ((2, 2, 48, 64), 1.0)])
def test_fast_path_stash_and_fallback(dp, tp, reg, shape, hm_sigma):
    """The tuned kernels stash the Gaussian window in shared memory; a window that does not fit (large sigma)
    takes their global-memory fallback.  Both must match the dense reference arithmetic, fp32 and bf16."""
    z, target, mask = synth(*shape, 2.0, seed=21, trained=(hm_sigma == 1.0 and shape[-1] == 64), tp=tp)
    l64, c64, d64 = oracle(tp, z, target, mask, reg, hm_sigma=hm_sigma)
    check(run_head(dp, z, target, mask, reg, hm_sigma=hm_sigma), l64, c64, d64,
          'fast %s %s sigma=%g' % (shape, reg, hm_sigma))
    zb = z.to(torch.bfloat16)
    l64, c64, d64 = oracle(tp, zb.float(), target, mask, reg, hm_sigma=hm_sigma)
    zz = zb.to(DEV).requires_grad_(True)
    out = dp.dsnt_head(zz, target.to(DEV), mask.to(DEV), reg=reg, hm_sigma=hm_sigma)
    out.loss.backward()
    got = {'loss': out.loss.item(), 'coords': out.coords.detach().cpu().double().numpy(),
           'dz': zz.grad.float().cpu().double().numpy()}
    check(got, l64, c64, d64, 'fast bf16 %s %s sigma=%g' % (shape, reg, hm_sigma), dz_tol=4e-3)


def test_unaligned_base_pointer_takes_scalar_path(dp, tp):
    buf = torch.randn(2 * 3 * 16 * 16 + 1)
    z = buf[1:].view(2, 3, 16, 16)                      # 4-byte aligned only once on the device
    target = torch.rand(2, 3, 2) - 0.5
    zz = torch.empty(2 * 3 * 16 * 16 + 1, device=DEV)
    zz[1:] = z.reshape(-1).to(DEV)
    zv = zz[1:].view(2, 3, 16, 16).requires_grad_(True)
    assert zv.data_ptr() % 16 != 0
    out = dp.dsnt_head(zv, target.to(DEV), None, reg='js', hm_sigma=1.0)
    l64, c64, _ = oracle(tp, z, target, None, 'js')
    assert abs(out.loss.item() - l64) / l64 < TOL
    assert np.abs(out.coords.detach().cpu().double().numpy() - c64).max() < TOL


# ------------------------------------------------------------------------------------------- bf16
@pytest.mark.parametrize('reg', ['none', 'var', 'js', 'kl', 'mse'])
@pytest.mark.parametrize('shape', [(8, 16, 64, 64), (4, 16, 28, 28), (2, 2, 7, 7)])
def test_bf16_logits(dp, tp, reg, shape):
    z, target, mask = synth(*shape, 1.0, seed=5)
    zb = z.to(torch.bfloat16)
    l64, c64, d64 = oracle(tp, zb.float(), target, mask, reg)      # oracle on the bf16-rounded logits
    zz = zb.to(DEV).requires_grad_(True)
    out = dp.dsnt_head(zz, target.to(DEV), mask.to(DEV), reg=reg, hm_sigma=1.0)
    out.loss.backward()
    assert zz.grad.dtype == torch.bfloat16 and out.coords.dtype == torch.float32
    got = {'loss': out.loss.item(), 'coords': out.coords.detach().cpu().double().numpy(),
           'dz': zz.grad.float().cpu().double().numpy()}
    check(got, l64, c64, d64, 'bf16 %s %s' % (shape, reg), dz_tol=4e-3)


# ------------------------------------------------------------------------------------------- edge cases
def test_all_zero_mask_and_no_mask(dp, tp):
    z, target, _ = synth(2, 3, 16, 16, 1.0, seed=6)
    got = run_head(dp, z, target, torch.zeros(2, 3), 'js')
    assert got['loss'] == 0.0 and np.abs(got['dz']).max() == 0.0
    l64, c64, d64 = oracle(tp, z, target, None, 'kl')
    check(run_head(dp, z, target, None, 'kl'), l64, c64, d64, 'mask=None kl')


def test_zero_distance_gradient_guard_and_strict_nan(dp):
    """coords == target gives NaN gradients in the reference; we default to 0."""
    from dsnt_pose2d_b200 import head
    z = torch.zeros(1, 1, 4, 4, device=DEV, requires_grad=True)     # uniform heatmap -> coords exactly (0,0)
    target = torch.zeros(1, 1, 2, device=DEV)
    out = dp.dsnt_head(z, target, None, reg='none')
    out.loss.backward()
    assert out.loss.item() == 0.0 and torch.isfinite(z.grad).all() and z.grad.abs().max().item() == 0.0
    head.STRICT_NAN = True
    try:
        z2 = torch.zeros(1, 1, 4, 4, device=DEV, requires_grad=True)
        dp.dsnt_head(z2, target, None, reg='none').loss.backward()
        assert torch.isnan(z2.grad).all()
    finally:
        head.STRICT_NAN = False


def test_empty_batch_and_error_paths(dp):
    z = torch.zeros(0, 16, 8, 8, device=DEV, requires_grad=True)
    out = dp.dsnt_head(z, torch.zeros(0, 16, 2, device=DEV), None, reg='js')
    assert out.coords.shape == (0, 16, 2) and out.loss.item() == 0.0
    with pytest.raises(NotImplementedError):
        dp.dsnt_head(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 2), None)               # CPU tensor
    with pytest.raises(NotImplementedError):
        dp.dsnt_head(torch.zeros(1, 1, 4, 4, device=DEV, dtype=torch.float16), torch.zeros(1, 1, 2, device=DEV))
    with pytest.raises(NotImplementedError):
        dp.nn.dsnt(torch.zeros(1, 4, 4, device=DEV, dtype=torch.float64))
    with pytest.raises(ValueError):
        dp.dsnt_head(torch.zeros(1, 1, 4, 4, device=DEV), torch.zeros(1, 1, 2, device=DEV), reg='bogus')


def test_coords_gradient_path_and_mixed_upstream(dp, tp):
    """Gradients arriving through `coords` (not the loss) and through both at once."""
    z, target, mask = synth(2, 3, 16, 16, 1.0, seed=7)
    w = torch.randn(2, 3, 2)
    zz = z.to(DEV).requires_grad_(True)
    out = dp.dsnt_head(zz, target.to(DEV), mask.to(DEV), reg='js')
    (2.0 * out.loss + (out.coords * w.to(DEV)).sum()).backward()
    z64 = z.double().requires_grad_(True)
    loss, coords, _, _ = tp.head_loss(z64, target.double(), mask.double(), 'js', 1.0, 1.0)
    (2.0 * loss + (coords * w.double()).sum()).backward()
    assert rel_l2(zz.grad.cpu().double().numpy(), z64.grad.numpy()) < TOL


def test_deterministic_bitwise(dp):
    z, target, mask = synth(8, 16, 64, 64, 1.0, seed=8)
    a = run_head(dp, z, target, mask, 'js')
    b = run_head(dp, z, target, mask, 'js')
    assert a['loss'] == b['loss'] and np.array_equal(a['dz'], b['dz']) and np.array_equal(a['coords'], b['coords'])


def test_cuda_graph_capture_of_fwd_bwd(dp):
    """No host synchronisation inside the operators: forward+backward can be captured and replayed."""
    z, target, mask = synth(4, 16, 64, 64, 1.0, seed=9)
    zz, tt, mm = z.to(DEV).requires_grad_(True), target.to(DEV), mask.to(DEV)
    ref = run_head(dp, z, target, mask, 'js')
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                                         # warm-up on the side stream
            zz.grad = None
            dp.dsnt_head(zz, tt, mm, reg='js').loss.backward()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    zz.grad = None
    with torch.cuda.graph(g):
        out = dp.dsnt_head(zz, tt, mm, reg='js')
        out.loss.backward()
    zz.grad.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert out.loss.item() == ref['loss']
    assert np.array_equal(zz.grad.cpu().double().numpy(), ref['dz'])


# ------------------------------------------------------------------------------------------- full-size properties
def test_cfg4_full_size_properties(dp, tp):
    """BASELINE cfg 4: 4096 x 16 x 64x64 fp32 (1 GiB).  The oracle cannot run this in seconds, so check
    (1) a random sample of heatmaps against the fp64 oracle through per-heatmap quantities,
    (2) softmax shift invariance, (3) sum_ij dZ = 0 per heatmap, (4) linearity of dZ in the upstream gradient."""
    torch.manual_seed(0)
    b, c, h, w = 4096, 16, 64, 64
    z = torch.randn(b, c, h, w, device=DEV)
    target = torch.rand(b, c, 2, device=DEV) * 1.6 - 0.8
    mask = (torch.rand(b, c, device=DEV) > 0.1).float()
    z.requires_grad_(True)
    out = dp.dsnt_head(z, target, mask, reg='js', hm_sigma=1.0)
    out.loss.backward()
    dz = z.grad
    # (3) every heatmap's gradient sums to zero (softmax Jacobian annihilates constants)
    sums = dz.view(b * c, -1).sum(-1)
    scale = dz.view(b * c, -1).abs().sum(-1).clamp_min(1e-30)
    assert (sums.abs() / scale).max().item() < 1e-4
    # (1) sample: slice 8 samples, run the oracle on the slice with the GLOBAL denominator
    idx = torch.tensor([0, 1, 777, 2048, 3000, 4094, 4095, 1234], device=DEV)
    zs, ts, ms = z.detach()[idx].cpu(), target[idx].cpu(), mask[idx].cpu()
    z64 = zs.double().requires_grad_(True)
    coords, p = tp.head_forward(z64)
    denom = mask.sum().item()
    dist = (coords - ts.double()).pow(2).sum(-1).sqrt()
    js = tp._js_2d(p, tp.make_gauss(ts.double(), w, h, 2.0 / w))
    ((dist * ms.double()).sum() / denom + (js * ms.double()).sum() / denom).backward()
    assert (out.coords.detach()[idx].cpu().double() - coords.detach()).abs().max().item() < TOL
    assert rel_l2(dz[idx].cpu().double().numpy(), z64.grad.numpy()) < TOL
    # global loss against a float64 accumulation of per-heatmap terms is covered by cfg1; here check shift invariance
    z2 = (z.detach() + 3.25).requires_grad_(True)
    out2 = dp.dsnt_head(z2, target, mask, reg='js', hm_sigma=1.0)
    out2.loss.backward()
    assert abs(out2.loss.item() - out.loss.item()) / out.loss.item() < 2e-6
    assert ((z2.grad - dz).norm() / dz.norm()).item() < 2e-5
    # (4) linearity in the upstream gradient
    z3 = z.detach().clone().requires_grad_(True)
    (dp.dsnt_head(z3, target, mask, reg='js', hm_sigma=1.0).loss * 3.0).backward()
    assert ((z3.grad - 3.0 * dz).norm() / (3.0 * dz.norm())).item() < 1e-6
