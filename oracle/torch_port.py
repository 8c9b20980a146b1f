"""CPU torch restatement of the reference DSNT head (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function names the reference lines it restates (paths relative to the
reference checkout).  The op *sequence* is kept the same as the reference's
(materialised coordinate grids, materialised Gaussian, separate log terms ...)
so that timing this port on the host cores is an honest stand-in for timing the
reference itself on the GPU box, where `/root/reference` does not exist.

Works for any floating dtype; the parity tests use float64 as arbiter.
"""

import math

import torch
import torch.nn.functional as F

EPS_DIV = 1e-24      # src/dsnt/nn.py:208,214 (inside each log) and :202 (Gaussian normaliser)


# --------------------------------------------------------------------------- grids
def _axis(n, like):
    """Pixel-centre coordinates of an n-pixel axis: linspace(-(n-1)/n, (n-1)/n, n).

    src/dsnt/nn.py:30-37 (and :180-187 for make_gauss).  The reference builds the
    linspace in the *default* dtype and then `type_as`; its tests set the default
    to double (tests/common.py:18).  We build it in the tensor's own dtype for
    fp32/fp64 (identical to the reference whenever default dtype == input dtype,
    which is every case its tests and scripts exercise) and in fp32 for narrower
    inputs such as bf16 (what `type_as` from an fp32 default gives).
    """
    lim = (n - 1) / n
    wide = like.dtype if like.dtype in (torch.float32, torch.float64) else torch.float32
    return torch.linspace(-lim, lim, n, dtype=wide).to(dtype=like.dtype, device=like.device)


def generate_xy(inp):
    """src/dsnt/nn.py:25-46 -- X and Y grids broadcast to the shape of `inp`."""
    h, w = inp.shape[-2], inp.shape[-1]
    lead = [1] * (inp.dim() - 2)
    gx = _axis(w, inp).view(*lead, 1, w).expand_as(inp)
    gy = _axis(h, inp).view(*lead, h, 1).expand_as(inp)
    return gx, gy


def expectation_2d(values, probabilities):
    """src/dsnt/nn.py:49-63 -- sum of values*prob over the last two dims."""
    weighted = values * probabilities
    return weighted.flatten(-2).sum(-1)


def dsnt(heatmaps):
    """src/dsnt/nn.py:66-78 -- coords[..., 0] = E[x], coords[..., 1] = E[y]."""
    gx, gy = generate_xy(heatmaps)
    return torch.stack((expectation_2d(gx, heatmaps), expectation_2d(gy, heatmaps)), dim=-1)


# --------------------------------------------------------------------------- losses
def masked_average(losses, mask=None):
    """src/dsnt/nn.py:81-94 -- sum(loss*mask)/max(sum(mask),1); no mask: mean over numel."""
    if mask is None:
        return losses.sum() / max(losses.numel(), 1)
    return (losses * mask).sum() / mask.sum().clamp(min=1)


def euclidean_loss(actual, target, mask=None):
    """src/dsnt/nn.py:97-116 -- per-point L2 distance, then masked_average."""
    delta = actual - target
    return masked_average(delta.pow(2).sum(-1).sqrt(), mask)


# --------------------------------------------------------------------------- softmaxes
class _ThresholdedSoftmax(torch.autograd.Function):
    """src/dsnt/nn.py:119-139."""

    @staticmethod
    def forward(ctx, inp, threshold, eps):
        keep = (inp >= threshold).to(inp.dtype)                       # :122
        shifted = inp - inp.max(-1, keepdim=True)[0]                  # :124 (max over ALL entries)
        e = shifted.exp() * keep                                      # :125
        out = e / (e.sum(-1, keepdim=True) + eps)                     # :126
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        inner = (grad_output * out).sum(-1, keepdim=True)             # :136
        return out * (grad_output - inner), None, None                # :137


def thresholded_softmax(inp, threshold=-math.inf, eps=1e-12):
    """src/dsnt/nn.py:142-157."""
    return _ThresholdedSoftmax.apply(inp, threshold, eps)


def softmax_2d(inp):
    """src/dsnt/nn.py:160-165 -- softmax over the last two dims taken together."""
    shape = inp.shape
    return F.softmax(inp.reshape(-1, shape[-1] * shape[-2]), dim=1).view(shape)


flat_softmax = softmax_2d       # the name used by north_star / dsntnn


def hm_preact_softmax(z):
    """src/dsnt/model.py:24-30,44-45 -- what the models actually call."""
    c, h, w = z.shape[-3], z.shape[-2], z.shape[-1]
    return F.softmax(z.reshape(-1, h * w), dim=-1).view(-1, c, h, w)


def hm_preact(z, preact='softmax'):
    """src/dsnt/model.py:24-45 -- every `--preact` choice; returns normalised heatmaps [-1, C, H, W]."""
    c, h, w = z.shape[-3], z.shape[-2], z.shape[-1]
    x = z.reshape(-1, h * w)
    if preact == 'softmax':
        x = F.softmax(x, dim=-1)                                      # :29-30
    elif preact == 'thresholded_softmax':
        x = thresholded_softmax(x, -0.5)                              # :31-32
    elif preact == 'abs':
        x = x.abs()                                                   # :33-35
        x = x / (x.sum(-1, keepdim=True) + 1e-12)
    elif preact == 'relu':
        x = F.relu(x)                                                 # :36-38
        x = x / (x.sum(-1, keepdim=True) + 1e-12)
    elif preact == 'sigmoid':
        x = torch.sigmoid(x)                                          # :39-41
        x = x / (x.sum(-1, keepdim=True) + 1e-12)
    else:
        raise Exception('unrecognised heatmap preactivation function: {}'.format(preact))   # :42-43
    return x.view(-1, c, h, w)


# --------------------------------------------------------------------------- Gaussian target
def make_gauss(coords, width, height, sigma):
    """src/dsnt/nn.py:168-205 -- normalised 2-D Gaussian, (width, height) argument order."""
    lead = [1] * (coords.dim() - 1)
    gx = _axis(width, coords).view(*lead, 1, width).expand(*lead, height, width)
    gy = _axis(height, coords).view(*lead, height, 1).expand(*lead, height, width)
    k = -0.5 * (1 / sigma) ** 2                                        # :196
    dx2 = (gx - coords[..., 0:1].unsqueeze(-1)) ** 2                   # :197
    dy2 = (gy - coords[..., 1:2].unsqueeze(-1)) ** 2                   # :198
    g = ((dx2 + dy2) * k).exp()                                        # :199
    total = g.sum(-1, keepdim=True).sum(-2, keepdim=True) + EPS_DIV    # :202
    return g / total


# --------------------------------------------------------------------------- divergences
def _kl_2d(p, q, eps=EPS_DIV):
    """src/dsnt/nn.py:208-211."""
    return (p * ((p + eps).log() - (q + eps).log())).sum(-1).sum(-1)


def _js_2d(p, q, eps=EPS_DIV):
    """src/dsnt/nn.py:214-216."""
    m = 0.5 * (p + q)
    return 0.5 * _kl_2d(p, m, eps) + 0.5 * _kl_2d(q, m, eps)


def kl_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """src/dsnt/nn.py:219-234."""
    g = make_gauss(mu_t, heatmaps.size(-1), heatmaps.size(-2), sigma_t)
    return masked_average(_kl_2d(heatmaps, g), mask)


def js_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """src/dsnt/nn.py:237-252."""
    g = make_gauss(mu_t, heatmaps.size(-1), heatmaps.size(-2), sigma_t)
    return masked_average(_js_2d(heatmaps, g), mask)


def mse_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """src/dsnt/nn.py:255-271."""
    g = make_gauss(mu_t, heatmaps.size(-1), heatmaps.size(-2), sigma_t)
    return masked_average(((heatmaps - g) ** 2).sum(-1).sum(-1), mask)


def variance_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """src/dsnt/nn.py:274-298 (mu_t unused there as well)."""
    gx, gy = generate_xy(heatmaps)
    mx = expectation_2d(gx, heatmaps)[..., None, None]
    my = expectation_2d(gy, heatmaps)[..., None, None]
    var = torch.stack((expectation_2d((gx - mx) ** 2, heatmaps),
                       expectation_2d((gy - my) ** 2, heatmaps)), dim=-1)
    return masked_average(((var - sigma_t ** 2) ** 2).sum(-1), mask)


_REG_FUNCS = {'var': variance_reg_loss, 'kl': kl_reg_loss, 'js': js_reg_loss, 'mse': mse_reg_loss}


def calculate_reg_loss(target, mask, reg, heatmaps, hm_sigma):
    """src/dsnt/model.py:47-63 -- sigma px -> normalised units uses the WIDTH only."""
    sigma = 2.0 * hm_sigma / heatmaps.size(-1)
    fn = _REG_FUNCS.get(reg)
    return fn(heatmaps, target, sigma, mask) if fn is not None else 0


# --------------------------------------------------------------------------- the whole head
def head_forward(z, preact='softmax'):
    """src/dsnt/model.py:176-183 (forward_part2, 'dsnt' strategy).

    Returns (coords [B,C,2], heatmaps P [B,C,H,W]).
    """
    p = hm_preact_softmax(z) if preact == 'softmax' else hm_preact(z, preact)
    return dsnt(p), p


def head_loss(z, target, mask=None, reg='none', hm_sigma=1.0, reg_coeff=1.0, preact='softmax'):
    """forward_part2 followed by forward_loss for one heatmap tensor.

    src/dsnt/model.py:138-145: loss = euclidean_loss + reg_coeff * reg_loss.
    Returns (loss, coords, euclid, reg_value).
    """
    coords, p = head_forward(z, preact)
    euc = euclidean_loss(coords, target, mask)
    rv = calculate_reg_loss(target, mask, reg, p, hm_sigma)
    return euc + reg_coeff * rv, coords, euc, rv


def head_loss_stacked(zs, target, mask=None, reg='none', hm_sigma=1.0, reg_coeff=1.0):
    """Hourglass variant, src/dsnt/model.py:233-246,286-292: sum of per-stack losses."""
    total = 0
    coords_per_stack = []
    for z in zs:
        loss, coords, _, _ = head_loss(z, target, mask, reg, hm_sigma, reg_coeff)
        total = total + loss
        coords_per_stack.append(coords)
    return total, coords_per_stack


def head_loss_and_grad(z, target, mask=None, reg='none', hm_sigma=1.0, reg_coeff=1.0,
                       dtype=torch.float64, preact='softmax'):
    """Convenience for parity tests: evaluate in `dtype` on CPU and return
    dict(loss, coords, euclid, reg, dz) as CPU tensors of that dtype."""
    zz = z.detach().to('cpu', dtype).clone().requires_grad_(True)
    tt = target.detach().to('cpu', dtype)
    mm = None if mask is None else mask.detach().to('cpu', dtype)
    loss, coords, euc, rv = head_loss(zz, tt, mm, reg, hm_sigma, reg_coeff, preact)
    loss.backward()
    rv_t = rv if torch.is_tensor(rv) else torch.tensor(float(rv), dtype=dtype)
    return {'loss': loss.detach(), 'coords': coords.detach(), 'euclid': euc.detach(),
            'reg': rv_t.detach(), 'dz': zz.grad.detach()}


# --------------------------------------------------------------------------- inference: flip test-time augmentation
# joint permutation under a horizontal flip for the MPII joint order (torchdata.mpii.MPII_Joint_Horizontal_Flips,
# bound to MPIIDataset.HFLIP_INDICES at src/dsnt/data.py:97)
MPII_HFLIP_INDICES = (5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 15, 14, 13, 12, 11, 10)


def reverse_tensor(tensor, dim):
    """src/dsnt/util.py:207-210."""
    idx = torch.arange(tensor.size(dim) - 1, -1, -1, device=tensor.device)
    return tensor.index_select(dim, idx)


def flip_tta_heatmaps(hm_pair, hflip_indices=MPII_HFLIP_INDICES):
    """src/dsnt/inference.py:43-46 -- `hm_pair` holds the raw heatmaps of [images, mirrored images] (the cat of
    :36); the reference does this for one image (`split(1)`), the same lines are applied per half here."""
    hm1, hm2 = hm_pair.chunk(2, 0)                                    # :43 (split(1) when the batch is one image)
    hm2 = reverse_tensor(hm2, -1)                                     # :44
    hm2 = hm2.index_select(-3, torch.as_tensor(hflip_indices, dtype=torch.long, device=hm2.device))   # :45
    return (hm1 + hm2) / 2                                            # :46


def flip_tta_coords(hm_pair, hflip_indices=MPII_HFLIP_INDICES, preact='softmax'):
    """src/dsnt/inference.py:43-48: averaged heatmaps -> forward_part2 ('dsnt') -> coords [B,C,2]."""
    hm = flip_tta_heatmaps(hm_pair, hflip_indices)
    return head_forward(hm, preact)[0], hm
