"""numpy fp64 closed-form oracle for the DSNT head (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Forward values and *analytic* gradients, written without autograd, following
SURVEY.md Appendix A.  This is the arithmetic contract of the CUDA kernels:

    forward   m = max z, S = sum exp(z-m), P = exp(z-m)/S           (src/dsnt/model.py:28-30)
              mu = (sum P x_j, sum P y_i)                            (src/dsnt/nn.py:66-78)
              d  = |mu - t|                                          (src/dsnt/nn.py:112-114)
              D  = regulariser(P, gauss(t, sigma))                   (src/dsnt/nn.py:208-298)
              L  = sum_n w_n (d_n + coeff * D_n),  w = mask/max(sum mask,1)   (nn.py:81-94, model.py:145)
    backward  g_ij = a x_j + b y_i + rho r_ij ;  dZ = P (g - sum P g)

All inputs are float64 numpy arrays; heatmaps are [N, H, W], targets [N, 2], mask [N] or None.
"""

import numpy as np

EPS = 1e-24
REGS = ('none', 'var', 'kl', 'js', 'mse')


def axis_coords(n):
    """x_j = (2j+1)/n - 1, the closed form of linspace(-(n-1)/n, (n-1)/n, n) (src/dsnt/nn.py:30-37)."""
    return (2.0 * np.arange(n, dtype=np.float64) + 1.0) / n - 1.0


def softmax_flat(z):
    """Flat spatial softmax per heatmap (src/dsnt/model.py:28-30)."""
    n, h, w = z.shape
    f = z.reshape(n, h * w)
    e = np.exp(f - f.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).reshape(n, h, w)


def gauss(target, width, height, sigma):
    """Normalised separable Gaussian (src/dsnt/nn.py:168-205), [N, H, W]."""
    k = -0.5 / (sigma * sigma)
    gx = np.exp(k * (axis_coords(width)[None, :] - target[:, 0:1]) ** 2)      # [N, W]
    gy = np.exp(k * (axis_coords(height)[None, :] - target[:, 1:2]) ** 2)     # [N, H]
    g = gy[:, :, None] * gx[:, None, :]
    return g / (g.sum(axis=(1, 2), keepdims=True) + EPS)


def weights(n, mask):
    """Per-heatmap weight of masked_average (src/dsnt/nn.py:81-94)."""
    if mask is None:
        return np.full(n, 1.0 / max(n, 1))
    return mask / max(mask.sum(), 1.0)


def reg_value_and_grad(p, target, sigma, reg):
    """Per-heatmap regulariser D [N] and r = dD/dP [N,H,W] for an arbitrary (not necessarily
    normalised) heatmap P -- SURVEY.md Appendix A.1/A.2 with every epsilon kept."""
    n, h, w = p.shape
    xs, ys = axis_coords(w), axis_coords(h)
    if reg == 'none':
        return np.zeros(n), np.zeros_like(p)
    if reg == 'var':
        s0 = p.sum(axis=(1, 2))
        mx = (p * xs[None, None, :]).sum(axis=(1, 2))
        my = (p * ys[None, :, None]).sum(axis=(1, 2))
        dx = xs[None, None, :] - mx[:, None, None]
        dy = ys[None, :, None] - my[:, None, None]
        vx = (p * dx ** 2).sum(axis=(1, 2))
        vy = (p * dy ** 2).sum(axis=(1, 2))
        s2 = sigma * sigma
        d = (vx - s2) ** 2 + (vy - s2) ** 2
        # d vx / d P_kl = (x_l - mx)^2 - 2 x_l sum P (x - mx),  sum P (x - mx) = mx (1 - s0)
        cx = (mx * (1.0 - s0))[:, None, None]
        cy = (my * (1.0 - s0))[:, None, None]
        r = (2.0 * (vx - s2))[:, None, None] * (dx ** 2 - 2.0 * xs[None, None, :] * cx) \
            + (2.0 * (vy - s2))[:, None, None] * (dy ** 2 - 2.0 * ys[None, :, None] * cy)
        return d, r
    g = gauss(target, w, h, sigma)
    if reg == 'kl':
        lp, lg = np.log(p + EPS), np.log(g + EPS)
        return (p * (lp - lg)).sum(axis=(1, 2)), lp - lg + p / (p + EPS)
    if reg == 'js':
        m = 0.5 * (p + g)
        lp, lg, lm = np.log(p + EPS), np.log(g + EPS), np.log(m + EPS)
        d = 0.5 * (p * (lp - lm)).sum(axis=(1, 2)) + 0.5 * (g * (lg - lm)).sum(axis=(1, 2))
        # dD/dP = 1/2 [ lp - lm + P/(P+e) ] - 1/2 * (P + G)/2 /(M+e)  ... both KL terms depend on M
        r = 0.5 * (lp - lm + p / (p + EPS) - m / (m + EPS))
        return d, r
    if reg == 'mse':
        return ((p - g) ** 2).sum(axis=(1, 2)), 2.0 * (p - g)
    raise ValueError(reg)


def head(z, target, mask=None, reg='none', sigma=None, reg_coeff=1.0, input_is_logits=True,
         g_loss=1.0, g_coords=None):
    """Fused head forward + backward.

    z: [N,H,W] logits (or heatmaps P when input_is_logits=False); target [N,2]; mask [N]|None;
    sigma: NORMALISED std-dev (= 2*hm_sigma/W, src/dsnt/model.py:49).
    g_coords: optional extra upstream gradient on coords [N,2] (for standalone `dsnt`).
    Returns dict(coords, dist, reg_terms, euclid, reg, loss, dz).
    """
    n, h, w = z.shape
    xs, ys = axis_coords(w), axis_coords(h)
    p = softmax_flat(z) if input_is_logits else z
    mx = (p * xs[None, None, :]).sum(axis=(1, 2))
    my = (p * ys[None, :, None]).sum(axis=(1, 2))
    coords = np.stack([mx, my], axis=-1)
    wt = weights(n, mask)
    if target is None:
        dist = np.zeros(n)
    else:
        dist = np.sqrt(((coords - target) ** 2).sum(axis=-1))
    dterm, r = reg_value_and_grad(p, target, sigma if sigma is not None else 1.0, reg)
    euclid = (wt * dist).sum()
    regv = (wt * dterm).sum()
    loss = euclid + reg_coeff * regv

    # backward (Appendix A.2 / A.3)
    with np.errstate(divide='ignore', invalid='ignore'):
        if target is None:
            a = np.zeros(n)
            b = np.zeros(n)
        else:
            a = g_loss * wt * (mx - target[:, 0]) / dist      # NaN when dist == 0, like the reference
            b = g_loss * wt * (my - target[:, 1]) / dist
    if g_coords is not None:
        a = a + g_coords[:, 0]
        b = b + g_coords[:, 1]
    rho = g_loss * reg_coeff * wt
    g = a[:, None, None] * xs[None, None, :] + b[:, None, None] * ys[None, :, None] + rho[:, None, None] * r
    if input_is_logits:
        c = (p * g).sum(axis=(1, 2), keepdims=True)
        dz = p * (g - c)
    else:
        dz = g
    return {'coords': coords, 'dist': dist, 'reg_terms': dterm, 'euclid': euclid, 'reg': regv,
            'loss': loss, 'dz': dz}


def thresholded_softmax(x, threshold=-np.inf, eps=1e-12):
    """src/dsnt/nn.py:119-130 over the last axis; the max is over ALL entries, kept or not."""
    keep = (x >= threshold).astype(np.float64)
    e = np.exp(x - x.max(axis=-1, keepdims=True)) * keep
    return e / (e.sum(axis=-1, keepdims=True) + eps)


def thresholded_softmax_grad(out, grad_out):
    """src/dsnt/nn.py:131-139."""
    return out * (grad_out - (grad_out * out).sum(axis=-1, keepdims=True))
