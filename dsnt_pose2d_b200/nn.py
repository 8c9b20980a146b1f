"""Drop-in for the reference's `dsnt.nn` (src/dsnt/nn.py): same names, arguments and semantics, every
tensor operation executed by the sm_100a kernels of libdsnt_b200.so.

Level-1 API: the functions here take *normalised heatmaps P*, exactly like the reference.
The fully fused logits -> loss path lives in `dsnt_pose2d_b200.head.dsnt_head`.

Differences from the reference, all deliberate:
  * CUDA float32 / bfloat16 tensors only; CPU, float16 and float64 raise NotImplementedError (no fallback).
  * scalar results (coords, losses) are float32 even for bfloat16 heatmaps (the reference returns bf16).
  * d sqrt(0) = 0 instead of NaN in `euclidean_loss` unless `dsnt_pose2d_b200.head.STRICT_NAN` is set.
  * gradients w.r.t. `mu_t` / `target` exist (as in the reference, where make_gauss is differentiable) but take one
    extra small launch (dsnt_reg_dmu); the reference's training loop never asks for them.
"""

import math

import torch

from . import _lib
from . import head as _head

__all__ = ['generate_xy', 'expectation_2d', 'dsnt', 'masked_average', 'euclidean_loss', 'thresholded_softmax',
           'softmax_2d', 'flat_softmax', 'make_gauss', 'kl_reg_loss', 'js_reg_loss', 'mse_reg_loss',
           'variance_reg_loss', 'ThresholdedSoftmax']


# --------------------------------------------------------------------------------------------------
# import-compat helpers (src/dsnt/nn.py:25-63,81-94).  Not on the hot path: the kernels compute the grid
# from the pixel index and never materialise it.  These stay tiny torch expressions on the input's device.
def generate_xy(inp):
    """src/dsnt/nn.py:25-46 -- stride-0 views, built on the input's device (no host round trip)."""
    h, w = inp.shape[-2], inp.shape[-1]
    lead = [1] * (inp.dim() - 2)
    xs = ((2.0 * torch.arange(w, device=inp.device, dtype=torch.float32) + 1.0) / w - 1.0).to(inp.dtype)
    ys = ((2.0 * torch.arange(h, device=inp.device, dtype=torch.float32) + 1.0) / h - 1.0).to(inp.dtype)
    return xs.view(*lead, 1, w).expand_as(inp), ys.view(*lead, h, 1).expand_as(inp)


def expectation_2d(values, probabilities):
    """src/dsnt/nn.py:49-63."""
    prod = values * probabilities
    return prod.flatten(-2).sum(-1)


def _finish(terms, mask, n, dev, stream):
    out8 = torch.empty(8, dtype=torch.float32, device=dev)
    ws = _lib.finish_workspace(dev)
    _lib.call('dsnt_finish_loss', terms.data_ptr(), _lib.ptr(mask), n, 1.0, out8.data_ptr(), ws.data_ptr(), stream)
    return out8


class _MaskedAverage(torch.autograd.Function):
    @staticmethod
    def forward(ctx, losses, mask):
        n = losses.numel()
        dev = losses.device
        with torch.cuda.device(dev):
            terms = torch.zeros(n, 2, dtype=torch.float32, device=dev)
            terms[:, 0] = losses.reshape(-1).to(torch.float32)
            out8 = _finish(terms, mask, n, dev, _lib.stream_of(losses))
        ctx.save_for_backward(mask, out8)
        ctx.shape = losses.shape
        ctx.in_dtype = losses.dtype
        return out8[4]

    @staticmethod
    def backward(ctx, g):
        mask, out8 = ctx.saved_tensors
        scale = g.to(torch.float32) / out8[3]
        grad = scale.expand(ctx.shape) if mask is None else mask.view(ctx.shape) * scale
        return grad.to(ctx.in_dtype), None


def masked_average(losses, mask=None):
    """src/dsnt/nn.py:81-94 -- sum(losses*mask)/max(sum(mask),1); deterministic reduction kernel."""
    _lib.require_cuda(losses, 'losses')
    mask = _head._as_f32(mask, losses.numel(), 1, 'mask')
    return _MaskedAverage.apply(losses, mask)


# --------------------------------------------------------------------------------------------------
class _DSNT(torch.autograd.Function):
    """coords = (sum P x, sum P y);  dP = g_x x_j + g_y y_i  (src/dsnt/nn.py:66-78)."""

    @staticmethod
    def forward(ctx, heatmaps):
        pc, n, h, w = _head._flat_heatmaps(heatmaps)
        dev = pc.device
        with torch.cuda.device(dev):
            coords = torch.empty(n, 2, dtype=torch.float32, device=dev)
            stats = torch.empty(n, _lib.STATS_K, dtype=torch.float32, device=dev)
            _lib.call('dsnt_head_fwd', pc.data_ptr(), _lib.dtype_id(pc), 0, n, h, w, None, 0, 1.0,
                      coords.data_ptr(), stats.data_ptr(), None, 0, _lib.stream_of(pc))
        ctx.save_for_backward(stats)
        ctx.meta = (n, h, w, heatmaps.shape, pc.dtype)
        return coords.view(*heatmaps.shape[:-2], 2)

    @staticmethod
    def backward(ctx, g_coords):
        (stats,) = ctx.saved_tensors
        n, h, w, shape, dtype = ctx.meta
        dev = stats.device
        with torch.cuda.device(dev):
            gc = g_coords.to(torch.float32).contiguous()
            dp = torch.empty(shape, dtype=dtype, device=dev)
            _lib.call('dsnt_head_bwd', None, _lib.dtype_id(dp), 0, n, h, w, None, None, stats.data_ptr(),
                      gc.data_ptr(), None, None, None, 0.0, 0, 1.0, 0, dp.data_ptr(), 0, _lib.stream_of(dp))
        return dp


def dsnt(heatmaps):
    """Differentiable spatial to numerical transform (src/dsnt/nn.py:66-78).

    heatmaps: [..., H, W] (any leading dims, including none) -> coords [..., 2], x first."""
    _lib.require_cuda(heatmaps, 'heatmaps')
    return _DSNT.apply(heatmaps)


# --------------------------------------------------------------------------------------------------
class _EuclideanLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, actual, target, mask, flags):
        d = actual.shape[-1]
        n = actual.numel() // d
        dev = actual.device
        a = actual.to(torch.float32).contiguous()
        t = target.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = _lib.stream_of(a)
            terms = torch.empty(n, 2, dtype=torch.float32, device=dev)
            _lib.call('dsnt_euclid_fwd', a.data_ptr(), t.data_ptr(), n, d, terms.data_ptr(), stream)
            out8 = _finish(terms, mask, n, dev, stream)
        ctx.save_for_backward(a, t, mask, terms, out8)
        ctx.meta = (n, d, flags, actual.shape, actual.dtype)
        return out8[4]

    @staticmethod
    def backward(ctx, g):
        a, t, mask, terms, out8 = ctx.saved_tensors
        n, d, flags, shape, dtype = ctx.meta
        dev = a.device
        with torch.cuda.device(dev):
            g = g.to(torch.float32).contiguous()
            ga = torch.empty(n, d, dtype=torch.float32, device=dev)
            _lib.call('dsnt_euclid_bwd', a.data_ptr(), t.data_ptr(), terms.data_ptr(), _lib.ptr(mask),
                      g.data_ptr(), out8[3:4].data_ptr(), n, d, flags, ga.data_ptr(), _lib.stream_of(a))
        g_actual = ga.view(shape).to(dtype) if ctx.needs_input_grad[0] else None
        g_target = (-ga).view(shape).to(dtype) if ctx.needs_input_grad[1] else None      # d/dtarget = -d/dactual
        return g_actual, g_target, None, None


def euclidean_loss(actual, target, mask=None):
    """Average Euclidean distance for multi-point samples (src/dsnt/nn.py:97-116).

    actual, target: [..., n, d]; mask: [..., n] or None."""
    _lib.require_cuda(actual, 'actual')
    _lib.require_cuda(target, 'target')
    if actual.shape != target.shape:
        target = target.expand_as(actual)
    d = actual.shape[-1]
    mask = _head._as_f32(mask, actual.numel() // d, 1, 'mask')
    flags = _lib.FLAG_STRICT_NAN if _head.STRICT_NAN else 0
    return _EuclideanLoss.apply(actual, target, mask, flags)


# --------------------------------------------------------------------------------------------------
class ThresholdedSoftmax(torch.autograd.Function):
    """src/dsnt/nn.py:119-139 -- same static forward/backward shape as the reference's Function."""

    @staticmethod
    def forward(ctx, inp, threshold=-math.inf, eps=1e-12):
        _lib.require_cuda(inp, 'inp')
        x = inp.contiguous()
        length = x.shape[-1] if x.dim() > 0 else 1
        rows = x.numel() // max(length, 1)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('dsnt_tsoftmax_fwd', x.data_ptr(), _lib.dtype_id(x), rows, length, float(threshold),
                      float(eps), out.data_ptr(), _lib.stream_of(x))
        ctx.save_for_backward(out)
        ctx.meta = (rows, length)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        (out,) = ctx.saved_tensors
        rows, length = ctx.meta
        g = grad_output.to(out.dtype).contiguous()
        dx = torch.empty_like(out)
        with torch.cuda.device(out.device):
            _lib.call('dsnt_tsoftmax_bwd', out.data_ptr(), g.data_ptr(), _lib.dtype_id(out), rows, length,
                      dx.data_ptr(), _lib.stream_of(out))
        return dx, None, None


def thresholded_softmax(inp, threshold=-math.inf, eps=1e-12):
    """Softmax over the last dim with inputs below `threshold` zeroed (src/dsnt/nn.py:142-157)."""
    return ThresholdedSoftmax.apply(inp, threshold, eps)


def softmax_2d(inp):
    """Softmax with the last two dimensions combined (src/dsnt/nn.py:160-165)."""
    _lib.require_cuda(inp, 'inp')
    shape = inp.shape
    flat = inp.reshape(-1, shape[-1] * shape[-2])
    return ThresholdedSoftmax.apply(flat, -math.inf, 0.0).view(shape)


flat_softmax = softmax_2d      # north_star / dsntnn name for the same operation


# --------------------------------------------------------------------------------------------------
class _MakeGauss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, width, height, sigma):
        mu = coords.to(torch.float32).contiguous()
        n = mu.numel() // 2
        dev = mu.device
        out = torch.empty(*coords.shape[:-1], height, width, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call('dsnt_make_gauss_fwd', mu.data_ptr(), n, width, height, float(sigma), out.data_ptr(),
                      _lib.stream_of(mu))
        ctx.save_for_backward(mu)
        ctx.meta = (n, width, height, float(sigma), coords.shape, coords.dtype)
        return out.to(coords.dtype)

    @staticmethod
    def backward(ctx, g):
        (mu,) = ctx.saved_tensors
        n, width, height, sigma, shape, dtype = ctx.meta
        g = g.to(torch.float32).contiguous()
        dmu = torch.empty(n, 2, dtype=torch.float32, device=mu.device)
        with torch.cuda.device(mu.device):
            _lib.call('dsnt_make_gauss_bwd', mu.data_ptr(), g.data_ptr(), n, width, height, sigma, dmu.data_ptr(),
                      _lib.stream_of(mu))
        return dmu.view(shape).to(dtype), None, None, None


def make_gauss(coords, width, height, sigma):
    """Normalised 2-D Gaussians, differentiable w.r.t. coords (src/dsnt/nn.py:168-205; width before height)."""
    _lib.require_cuda(coords, 'coords')
    if coords.shape[-1] != 2:
        raise ValueError('coords must have a trailing dimension of 2')
    return _MakeGauss.apply(coords, int(width), int(height), sigma)


# --------------------------------------------------------------------------------------------------
class _RegLoss(torch.autograd.Function):
    """Per-heatmap divergence + masked average in two launches; backward is one streaming launch for the heatmaps and,
    when the target centres require grad (they do not in the reference's training loop, but make_gauss is differentiable:
    src/dsnt/nn.py:170,232), one small launch for d(loss)/d(mu_t) (dsnt_reg_dmu)."""

    @staticmethod
    def forward(ctx, heatmaps, mu_t, mask, reg_id, sigma):
        pc, n, h, w = _head._flat_heatmaps(heatmaps)
        dev = pc.device
        mu = None if mu_t is None else mu_t.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            stream = _lib.stream_of(pc)
            coords = torch.empty(n, 2, dtype=torch.float32, device=dev)
            stats = torch.empty(n, _lib.STATS_K, dtype=torch.float32, device=dev)
            terms = torch.empty(n, 2, dtype=torch.float32, device=dev)
            _lib.call('dsnt_head_fwd', pc.data_ptr(), _lib.dtype_id(pc), 0, n, h, w, _lib.ptr(mu), reg_id, sigma,
                      coords.data_ptr(), stats.data_ptr(), terms.data_ptr(), 0, stream)
            out8 = _finish(terms, mask, n, dev, stream)
        ctx.save_for_backward(pc, mu, mask, stats, out8)
        ctx.meta = (n, h, w, reg_id, sigma, heatmaps.shape, None if mu_t is None else (mu_t.shape, mu_t.dtype))
        return out8[5]

    @staticmethod
    def backward(ctx, g):
        pc, mu, mask, stats, out8 = ctx.saved_tensors
        n, h, w, reg_id, sigma, shape, mu_meta = ctx.meta
        dev = pc.device
        dp = dmu = None
        with torch.cuda.device(dev):
            g = g.to(torch.float32).contiguous()
            stream = _lib.stream_of(pc)
            if ctx.needs_input_grad[0]:
                dp = torch.empty_like(pc)
                _lib.call('dsnt_head_bwd', pc.data_ptr(), _lib.dtype_id(pc), 0, n, h, w, _lib.ptr(mu), _lib.ptr(mask),
                          stats.data_ptr(), None, None, g.data_ptr(), out8[3:4].data_ptr(), 1.0, reg_id, sigma,
                          _lib.FLAG_NO_EUCLID, dp.data_ptr(), 0, stream)
                dp = dp.view(shape)
            if mu is not None and ctx.needs_input_grad[1]:
                dmu = torch.zeros(n, 2, dtype=torch.float32, device=dev)
                if reg_id != _lib.REG_IDS['var']:
                    _lib.call('dsnt_reg_dmu', pc.data_ptr(), _lib.dtype_id(pc), 0, n, h, w, None, mu.data_ptr(),
                              _lib.ptr(mask), g.data_ptr(), out8[3:4].data_ptr(), 1.0, reg_id, sigma, dmu.data_ptr(), stream)
                dmu = dmu.view(mu_meta[0]).to(mu_meta[1])
        return dp, dmu, None, None, None


def _reg_loss(name, heatmaps, mu_t, sigma_t, mask):
    _lib.require_cuda(heatmaps, 'heatmaps')
    h, w = heatmaps.shape[-2], heatmaps.shape[-1]
    n = heatmaps.numel() // max(h * w, 1)
    if name != 'var':
        if mu_t is None:
            raise ValueError('%s_reg_loss needs mu_t' % name)
        _lib.require_cuda(mu_t, 'mu_t')
        if mu_t.numel() != 2 * n:
            raise ValueError('mu_t has %d elements, expected %d' % (mu_t.numel(), 2 * n))
    else:
        mu_t = None                      # unused by the reference as well (src/dsnt/nn.py:274-298)
    mask = _head._as_f32(mask, n, 1, 'mask')
    return _RegLoss.apply(heatmaps, mu_t, mask, _lib.REG_IDS[name], float(sigma_t))


def kl_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """Average KL(P || Gaussian(mu_t, sigma_t)) (src/dsnt/nn.py:219-234)."""
    return _reg_loss('kl', heatmaps, mu_t, sigma_t, mask)


def js_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """Average Jensen-Shannon divergence to the target Gaussians (src/dsnt/nn.py:237-252)."""
    return _reg_loss('js', heatmaps, mu_t, sigma_t, mask)


def mse_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """Mean squared error to the target Gaussians (src/dsnt/nn.py:255-271)."""
    return _reg_loss('mse', heatmaps, mu_t, sigma_t, mask)


def variance_reg_loss(heatmaps, mu_t, sigma_t, mask=None):
    """Squared error between heatmap variances and sigma_t^2; mu_t unused (src/dsnt/nn.py:274-298)."""
    return _reg_loss('var', heatmaps, None, sigma_t, mask)
