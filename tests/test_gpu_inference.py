"""GPU parity of the fused flip-TTA + forward-only head (src/dsnt/inference.py:36-48).
coords: 1e-5 max-abs against the reference golden vectors and the fp64 oracle; averaged heatmaps: exact in fp32
((a+b)/2 is one rounding, the same one the reference performs)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PREACTS = ['softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid']
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


@pytest.mark.parametrize('preact', PREACTS)
def test_flip_tta_matches_reference_golden(dp, golden_flip, preact):
    g = golden_flip
    for name in g.cases:
        pair = torch.from_numpy(g[name + '/hm_pair']).to(DEV)
        flips = [int(i) for i in g[name + '/flips']]
        coords, hm = dp.flip_tta_coords(pair, flips, preact=preact, return_heatmaps=True)
        err = np.abs(coords.cpu().double().numpy() - g['%s/%s/coords' % (name, preact)][0:1].reshape(coords.shape)).max()
        print('%-10s %-20s coords %.2e' % (name, preact, err))
        assert err < TOL
        assert np.array_equal(hm.cpu().numpy(), g[name + '/hm'].astype(np.float32))
        # without materialising the average: same coordinates, bitwise
        assert torch.equal(dp.flip_tta_coords(pair, flips, preact=preact), coords)


@pytest.mark.parametrize('preact', ['softmax', 'thresholded_softmax', 'sigmoid'])
@pytest.mark.parametrize('shape', [(1, 16, 64, 64), (8, 16, 64, 64), (4, 16, 28, 28), (2, 16, 7, 7), (2, 16, 130, 132),
                                   (1, 16, 256, 256), (3, 16, 33, 31)])
def test_flip_tta_matches_fp64_oracle(dp, tp, shape, preact):
    b, c, h, w = shape
    gen = torch.Generator().manual_seed(21)
    pair = torch.randn(2 * b, c, h, w, generator=gen) * 2
    ref, hm_ref = tp.flip_tta_coords(pair.double(), tp.MPII_HFLIP_INDICES, preact)
    coords, hm = dp.flip_tta_coords(pair.to(DEV), preact=preact, return_heatmaps=True)
    assert coords.shape == (b, c, 2)
    err = (coords.cpu().double() - ref).abs().max().item()
    print('%-16s %-20s coords %.2e' % ('x'.join(map(str, shape)), preact, err))
    assert err < TOL
    assert torch.equal(hm.cpu(), hm_ref.float())


def test_flip_tta_equals_unfused_composition_on_gpu(dp, tp):
    """The fused launch against the reference's own op sequence run with OUR head on the averaged heatmaps."""
    gen = torch.Generator().manual_seed(22)
    pair = (torch.randn(8, 16, 64, 64, generator=gen) * 3).to(DEV)
    hm = tp.flip_tta_heatmaps(pair)                      # torch ops on the GPU: reverse, index_select, mean
    two_step = dp.dsnt_head(hm, None, None, reg='none').coords
    fused = dp.flip_tta_coords(pair)
    assert (fused - two_step).abs().max().item() < 2e-6


def test_flip_tta_bf16(dp, tp):
    gen = torch.Generator().manual_seed(23)
    pair = (torch.randn(4, 16, 64, 64, generator=gen) * 2).to(torch.bfloat16)
    # the kernel averages in fp32 (the reference would round (a+b) to bf16 first): oracle on the widened values
    ref, hm_ref = tp.flip_tta_coords(pair.double(), tp.MPII_HFLIP_INDICES, 'softmax')
    coords, hm = dp.flip_tta_coords(pair.to(DEV), return_heatmaps=True)
    assert (coords.cpu().double() - ref).abs().max().item() < TOL
    assert hm.dtype == torch.bfloat16
    assert (hm.cpu().double() - hm_ref).abs().max().item() <= 2.0 ** -8 * hm_ref.abs().max().item()


def test_flip_tta_identity_permutation_and_errors(dp, tp):
    gen = torch.Generator().manual_seed(24)
    pair = torch.randn(2, 5, 12, 12, generator=gen)
    ref, _ = tp.flip_tta_coords(pair.double(), list(range(5)), 'softmax')
    got = dp.flip_tta_coords(pair.to(DEV), None)
    assert (got.cpu().double() - ref).abs().max().item() < TOL
    with pytest.raises(ValueError):
        dp.flip_tta_coords(pair.to(DEV), [0, 1, 2])             # wrong length
    with pytest.raises(ValueError):
        dp.flip_tta_coords(pair[:1].to(DEV), None)              # odd batch
    with pytest.raises(NotImplementedError):
        dp.flip_tta_coords(pair, None)                           # CPU tensor: no fallback


def test_predict_flipped_runs_a_model_end_to_end(dp, tp):
    """`generate_predictions(use_flipped=True)` for one batch: backbone on [x, flip(x)] + fused TTA head."""
    class Toy(torch.nn.Module):
        preact = 'softmax'

        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(3, 16, 3, padding=1)

        def forward_part1(self, x):
            return [self.conv(x) * 0.5, self.conv(x)]        # hourglass-style list: the LAST entry is used

    torch.manual_seed(25)
    model = Toy().to(DEV).eval()
    x = torch.randn(3, 3, 32, 32, device=DEV)
    got = dp.predict_flipped(model, x)
    assert got.device.type == 'cpu' and got.dtype == torch.float32 and got.shape == (3, 16, 2)
    with torch.no_grad():
        hm = model.conv(torch.cat([x, x.flip(-1)], 0)).double().cpu()
    ref, _ = tp.flip_tta_coords(hm, tp.MPII_HFLIP_INDICES, 'softmax')
    assert (got.double() - ref).abs().max().item() < TOL
