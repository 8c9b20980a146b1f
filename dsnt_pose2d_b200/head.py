"""The fused DSNT head: logits Z -> (coords, loss) with ONE forward kernel, one tiny finishing
reduction and ONE streaming backward kernel (12 bytes/pixel for fp32 instead of the reference's ~400).

This op is not in the reference; it is what `forward_part2` + `forward_loss` of
`HumanPoseModel` compute together (src/dsnt/model.py:24-63,138-145,176-183):

    P      = softmax over H*W of Z                      (model.py:28-30)
    coords = dsnt(P)                                    (nn.py:66-78)
    loss   = euclidean_loss(coords, target, mask)       (nn.py:97-116)
             + reg_coeff * {var,kl,js,mse}_reg_loss(P, target, 2*hm_sigma/W, mask)   (model.py:47-63,145)

P is never materialised; the autograd node saves Z and 8 floats per heatmap.

Batch sharding: pass `group=` and every rank calls the op on its slice of the batch.
The only coupling between heatmaps is the scalar normalisation of `masked_average`, so the op all-reduces
three floats (sum mask*dist, sum mask*D, sum mask) and the result -- value and gradients -- equals the
single-process op on the concatenated batch.
"""

import collections
import ctypes
import os

import torch

from . import _lib
from .parallel import PeerExchange

class HeadOutput:
    """(coords, loss, euclid, reg) of the fused head; behaves like the 4-tuple it used to be.  `euclid` and `reg` are views
    into the loss block, made only when somebody asks (two tensor slices per call are host time the small configs feel)."""

    __slots__ = ('coords', 'loss', '_out8')
    _fields = ('coords', 'loss', 'euclid', 'reg')

    def __init__(self, coords, loss, out8, off=0):
        # out8: a float32 tensor holding the loss block [sum m*d, sum m*D, count, denom, euclid, reg, loss, 0] at offset `off`
        self.coords, self.loss, self._out8 = coords, loss, (out8, off)

    @property
    def euclid(self):
        return self._out8[0][self._out8[1] + 4]

    @property
    def reg(self):
        return self._out8[0][self._out8[1] + 5]

    def __iter__(self):
        return iter((self.coords, self.loss, self.euclid, self.reg))

    def __len__(self):
        return 4

    def __getitem__(self, i):
        return (self.coords, self.loss, self.euclid, self.reg)[i]

    def __repr__(self):
        return 'HeadOutput(coords=%r, loss=%r, euclid=%r, reg=%r)' % tuple(self)

# Reproduce the reference's NaN gradient at coords == target (sqrt'(0)) instead of
# the default zero gradient.  Test-only switch; real training never wants the NaN.
STRICT_NAN = False


def _flat_heatmaps(z):
    if z.dim() < 2:
        raise ValueError('heatmaps need at least 2 dimensions, got shape %s' % (tuple(z.shape),))
    h, w = z.shape[-2], z.shape[-1]
    n = z.numel() // (h * w) if h * w > 0 else 0
    return z.contiguous(), n, h, w


def _as_f32(t, n, last, what):
    if t is None:
        return None
    _lib.require_cuda(t, what)
    if t.dtype is not torch.float32 or t.requires_grad or not t.is_contiguous():     # the usual case converts nothing
        t = t.detach().to(torch.float32).contiguous()
    expect = n * last
    if t.numel() != expect:
        raise ValueError('%s has %d elements, expected %d' % (what, t.numel(), expect))
    return t


def _is_sharded(group):
    if group is None:
        return False
    import torch.distributed as dist
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def all_reduce_sums(out8, group):
    """Sum the three partial sums out8[0:3] over the ranks of `group` (no-op for a single rank)."""
    import torch.distributed as dist
    if group is None or not dist.is_available() or not dist.is_initialized():
        return
    if dist.get_world_size(group) > 1:
        dist.all_reduce(out8[0:3], op=dist.ReduceOp.SUM, group=group)


def finish_loss(terms, mask, n_per_stack, n_stacks, reg_coeff, out8, ws, group, dev, stream):
    """masked_average + loss composition over the per-heatmap terms; with a sharded batch the partial sums of the ranks
    are exchanged inside the kernel over peer memory (one node) or all-reduced by NCCL.  terms=None: mask count only."""
    finish_loss_ptr(None if terms is None else terms.data_ptr(), mask, n_per_stack, n_stacks, reg_coeff,
                    out8.data_ptr(), out8, ws, group, dev, stream)


def finish_loss_ptr(terms_ptr, mask, n_per_stack, n_stacks, reg_coeff, out8_ptr, out8_owner, ws, group, dev, stream):
    """`finish_loss` on raw device addresses (the one-pass step keeps all its small outputs in one allocation);
    `out8_owner` is a float32 tensor that contains the 8 floats at `out8_ptr` (needed for the NCCL all-reduce only)."""
    peer = PeerExchange.get(group, dev)
    if peer is not None:
        if terms_ptr is None:
            _lib.call('dsnt_mask_count_peer', _lib.ptr(mask), n_per_stack, out8_ptr, ws.data_ptr(), *peer.args(), stream)
        else:
            _lib.call('dsnt_finish_loss_peer', terms_ptr, _lib.ptr(mask), n_per_stack, n_stacks, reg_coeff,
                      out8_ptr, ws.data_ptr(), *peer.args(), stream)
        return
    if terms_ptr is None:
        _lib.call('dsnt_mask_count', _lib.ptr(mask), n_per_stack, out8_ptr, ws.data_ptr(), stream)
    elif n_stacks == 1:
        _lib.call('dsnt_finish_loss', terms_ptr, _lib.ptr(mask), n_per_stack, reg_coeff, out8_ptr, ws.data_ptr(), stream)
    else:
        _lib.call('dsnt_finish_loss_stacked', terms_ptr, _lib.ptr(mask), n_per_stack, n_stacks, reg_coeff,
                  out8_ptr, ws.data_ptr(), stream)
    if _is_sharded(group):
        off = (out8_ptr - out8_owner.data_ptr()) // 4
        all_reduce_sums(out8_owner[off:off + 8], group)
        _lib.call('dsnt_combine_loss', out8_ptr, reg_coeff, stream)


_no_loss = {}


def _no_loss_block(dev):
    """The loss block of a call without a target and without a regulariser (forward_part2: coordinates only): every
    term is zero, the denominator 1 -- a per-device constant instead of a finishing launch over n zeros."""
    blk = _no_loss.get(dev)
    if blk is None:
        blk = torch.tensor([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        _no_loss[dev] = blk
    return blk


class _FusedHead(torch.autograd.Function):
    """forward: dsnt_head_fwd + dsnt_finish_loss;  backward: dsnt_head_bwd (include/dsnt_b200.h)."""

    @staticmethod
    def forward(ctx, z, target, mask, reg_id, sigma, reg_coeff, flags, group, variant, input_is_logits, aux):
        zc, n, h, w = _flat_heatmaps(z)
        dev = zc.device
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            coords = torch.empty(n, 2, dtype=torch.float32, device=dev)
            stats = torch.empty(n, _lib.STATS_K, dtype=torch.float32, device=dev)
            terms = torch.empty(n, 2, dtype=torch.float32, device=dev)
            _lib.call('dsnt_head_fwd', zc.data_ptr(), _lib.dtype_id(zc), int(input_is_logits), n, h, w,
                      _lib.ptr(target), reg_id, sigma, coords.data_ptr(), stats.data_ptr(), terms.data_ptr(),
                      variant, stream)
            if target is None and reg_id == 0 and not _is_sharded(group):
                out8 = _no_loss_block(dev)
            else:
                out8 = torch.empty(8, dtype=torch.float32, device=dev)
                ws = _lib.finish_workspace(dev, stream)
                finish_loss(terms, mask, n, 1, reg_coeff, out8, ws, group, dev, stream)
        ctx.save_for_backward(zc, target, mask, stats, out8)
        ctx.meta = (n, h, w, reg_id, sigma, reg_coeff, flags, variant, input_is_logits, z.shape)
        ctx.set_materialize_grads(False)
        aux['out8'] = out8          # side channel: [sum m*d, sum m*D, count, denom, euclid, reg, loss, 0]
        return coords.view(*z.shape[:-2], 2), out8[6]

    @staticmethod
    def backward(ctx, g_coords, g_loss):
        zc, target, mask, stats, out8 = ctx.saved_tensors
        n, h, w, reg_id, sigma, reg_coeff, flags, variant, input_is_logits, shape = ctx.meta
        dev = zc.device
        if g_coords is None and g_loss is None:
            return (None,) * 11
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            if g_coords is not None:
                g_coords = g_coords.to(torch.float32).contiguous()
            if g_loss is not None:
                g_loss = g_loss.to(torch.float32).contiguous()
            dz = torch.empty_like(zc)
            _lib.call('dsnt_head_bwd', zc.data_ptr(), _lib.dtype_id(zc), int(input_is_logits), n, h, w,
                      _lib.ptr(target), _lib.ptr(mask), stats.data_ptr(), _lib.ptr(g_coords), None,
                      _lib.ptr(g_loss), out8[3:4].data_ptr() if g_loss is not None else None,
                      reg_coeff, reg_id, sigma, flags, dz.data_ptr(), variant, stream)
        return (dz.view(shape),) + (None,) * 10


class _FusedHeadPreact(torch.autograd.Function):
    """The fused head for the reference's other pre-activations (src/dsnt/model.py:31-41):
    forward: dsnt_head_preact_fwd + dsnt_finish_loss;  backward: dsnt_head_preact_bwd (include/dsnt_b200.h)."""

    @staticmethod
    def forward(ctx, z, target, mask, reg_id, sigma, reg_coeff, flags, group, preact_id, threshold, eps, variant, aux):
        zc, n, h, w = _flat_heatmaps(z)
        dev = zc.device
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            coords = torch.empty(n, 2, dtype=torch.float32, device=dev)
            stats = torch.empty(n, _lib.STATS_K, dtype=torch.float32, device=dev)
            terms = torch.empty(n, 2, dtype=torch.float32, device=dev)
            out8 = torch.empty(8, dtype=torch.float32, device=dev)
            _lib.call('dsnt_head_preact_fwd', zc.data_ptr(), _lib.dtype_id(zc), preact_id, threshold, eps, n, h, w,
                      _lib.ptr(target), reg_id, sigma, coords.data_ptr(), stats.data_ptr(), terms.data_ptr(), variant,
                      stream)
            ws = _lib.finish_workspace(dev, stream)
            finish_loss(terms, mask, n, 1, reg_coeff, out8, ws, group, dev, stream)
        ctx.save_for_backward(zc, target, mask, stats, out8)
        ctx.meta = (n, h, w, reg_id, sigma, reg_coeff, flags, preact_id, threshold, variant, z.shape)
        ctx.set_materialize_grads(False)
        aux['out8'] = out8
        return coords.view(*z.shape[:-2], 2), out8[6]

    @staticmethod
    def backward(ctx, g_coords, g_loss):
        zc, target, mask, stats, out8 = ctx.saved_tensors
        n, h, w, reg_id, sigma, reg_coeff, flags, preact_id, threshold, variant, shape = ctx.meta
        dev = zc.device
        if g_coords is None and g_loss is None:
            return (None,) * 13
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            if g_coords is not None:
                g_coords = g_coords.to(torch.float32).contiguous()
            if g_loss is not None:
                g_loss = g_loss.to(torch.float32).contiguous()
            dz = torch.empty_like(zc)
            _lib.call('dsnt_head_preact_bwd', zc.data_ptr(), _lib.dtype_id(zc), preact_id, threshold, n, h, w,
                      _lib.ptr(target), _lib.ptr(mask), stats.data_ptr(), _lib.ptr(g_coords), None,
                      _lib.ptr(g_loss), out8[3:4].data_ptr() if g_loss is not None else None,
                      reg_coeff, reg_id, sigma, flags, dz.data_ptr(), variant, stream)
        return (dz.view(shape),) + (None,) * 12


# what HumanPoseModel._hm_preact passes for each `--preact` choice (src/dsnt/model.py:29-41)
PREACT_DEFAULTS = {
    'softmax': (float('-inf'), 0.0),
    'thresholded_softmax': (-0.5, 1e-12),
    'abs': (0.0, 1e-12),
    'relu': (0.0, 1e-12),
    'sigmoid': (0.0, 1e-12),
}


def dsnt_head(z, target, mask=None, reg='none', sigma=None, reg_coeff=1.0, hm_sigma=None, group=None,
              variant=0, input_is_logits=True, preact='softmax', threshold=None, eps=None, one_pass=False):
    """Fused head.

    Args:
        z: logits [..., H, W] (CUDA, float32 or bfloat16); any leading dims, usually [B, C, H, W].
        target: [..., 2] target coordinates (x, y) in normalised [-1, 1] units.
        mask: [...] joint visibility weights or None (src/dsnt/nn.py:81-94 semantics).
        reg: 'none' | 'var' | 'kl' | 'js' | 'mse'  (src/dsnt/model.py:52-61).
        sigma: target std-dev in normalised units; or give `hm_sigma` in pixels and the reference's
            conversion sigma = 2*hm_sigma/W is applied (src/dsnt/model.py:49).
        reg_coeff: weight of the regulariser (src/dsnt/model.py:145).
        group: torch.distributed process group when the batch is sharded across ranks.
        preact: how raw heatmaps become a distribution (src/dsnt/model.py:24-45): 'softmax' (the tuned kernels),
            'thresholded_softmax', 'abs', 'relu', 'sigmoid' (the epsilon-exact kernels of csrc/head_preact.cuh).
            `threshold` / `eps` default to what the reference passes (-0.5 / 1e-12).
        one_pass: evaluate forward AND the gradient for d(loss) = 1 in one pass over the logits (`dsnt_head_step`,
            8 instead of 12 bytes per fp32 pixel); `loss.backward()` then only hands the stored gradient out.  Needs
            the softmax pre-activation and heatmaps of which at least four fit in shared memory (`step_supported`);
            otherwise the two-kernel path is taken silently.  Costs one extra heatmap-sized buffer if backward is
            never called, so leave it off for inference.
    Returns:
        HeadOutput(coords [..., 2] float32, loss 0-dim float32, euclid 0-dim, reg 0-dim);
        `loss` and `coords` are differentiable w.r.t. `z`.
    """
    _lib.require_cuda(z, 'z')
    if reg not in _lib.REG_IDS:
        raise ValueError('unrecognised regulariser: %r' % (reg,))
    h, w = z.shape[-2], z.shape[-1]
    n = z.numel() // max(h * w, 1)
    if sigma is None:
        sigma = 2.0 * (1.0 if hm_sigma is None else hm_sigma) / w
    if target is not None and target.requires_grad:
        # Not the hot path (the reference's training loop never differentiates its targets): the reference's own composition
        # (src/dsnt/model.py:138-145) on this library's level-1 operators, every one of which is differentiable w.r.t. the target
        return _head_with_target_grad(z, target, mask, reg, float(sigma), float(reg_coeff), preact, threshold, eps,
                                      input_is_logits, group)
    target = _as_f32(target, n, 2, 'target')
    mask = _as_f32(mask, n, 1, 'mask')
    flags = _lib.FLAG_STRICT_NAN if STRICT_NAN else 0
    aux = {}
    if preact not in _lib.PREACT_IDS:
        raise Exception('unrecognised heatmap preactivation function: {}'.format(preact))   # model.py:42-43
    if (one_pass and preact == 'softmax' and threshold is None and eps is None and input_is_logits
            and z.requires_grad and torch.is_grad_enabled()
            and takes_one_pass(z, reg, float(sigma), _is_sharded(group))):
        coords, loss = _FusedHeadStep.apply(z, target, mask, (_lib.REG_IDS[reg], float(sigma), float(reg_coeff), flags, group, aux))
    elif preact == 'softmax' and threshold is None and eps is None:
        coords, loss = _FusedHead.apply(z, target, mask, _lib.REG_IDS[reg], float(sigma), float(reg_coeff),
                                        flags, group, int(variant), bool(input_is_logits), aux)
    else:
        if not input_is_logits:
            raise ValueError('preact=%r needs raw heatmaps (input_is_logits=True)' % (preact,))
        d_thr, d_eps = PREACT_DEFAULTS[preact]
        coords, loss = _FusedHeadPreact.apply(z, target, mask, _lib.REG_IDS[reg], float(sigma), float(reg_coeff),
                                              flags, group, _lib.PREACT_IDS[preact],
                                              float(d_thr if threshold is None else threshold),
                                              float(d_eps if eps is None else eps), int(variant), aux)
    return HeadOutput(coords, loss, aux['out8'], aux.get('off', 0))


def _head_with_target_grad(z, target, mask, reg, sigma, reg_coeff, preact, threshold, eps, input_is_logits, group):
    """`dsnt_head` for a target that requires grad: heatmaps are materialised and the loss composed as the reference does."""
    from . import nn as dnn
    from .model import hm_preact
    if _is_sharded(group):
        raise NotImplementedError('dsnt_head: gradients w.r.t. the target are not implemented for a sharded batch')
    if not input_is_logits:
        p = z
    elif preact == 'softmax' and threshold is None and eps is None:
        p = dnn.softmax_2d(z)
    elif threshold is None and eps is None and z.dim() >= 3:
        p = hm_preact(z, preact).view(z.shape)
    else:
        raise NotImplementedError('dsnt_head: gradients w.r.t. the target need the default threshold / eps of the pre-activation')
    coords = dnn.dsnt(p)
    euclid = dnn.euclidean_loss(coords, target, mask)
    fn = {'var': dnn.variance_reg_loss, 'kl': dnn.kl_reg_loss, 'js': dnn.js_reg_loss, 'mse': dnn.mse_reg_loss}.get(reg)
    regv = fn(p, target, sigma, mask) if fn is not None else torch.zeros((), dtype=torch.float32, device=z.device)
    loss = euclid + reg_coeff * regv
    zero = torch.zeros((), dtype=torch.float32, device=z.device)
    return HeadOutput(coords, loss, torch.stack([zero, zero, zero, zero, euclid.float(), regv.float(), loss.float(), zero]))


USE_PAIR_STEP = True          # 256x256 fp32 (not KL): the cluster-of-two-CTAs one-pass step (csrc/step_pair.cu),
                              # 666 us against 970 us of forward + backward at BASELINE config 5 (variance regulariser)
USE_PAIR_STEP_BF16 = False    # 256x256 bf16: the same kernel serves it at 4 B/px but is bound by the SM (MUFU + exchange
                              # latency per heatmap), 594 us (variance) against 510 us of the two-kernel path at 6 B/px
# Sharded batch: take the single-launch step with both exchanges inside the kernel (dsnt_head_step_fused_peer).  The count
# is published by the last CTA to arrive and picked up by every warp only before its first backward, so the exchange --
# and the wait for a rank that started later -- hides behind the first loads and the first forward.  DSNT_FUSED_PEER_STEP=0
# keeps dsnt_mask_count_peer + dsnt_head_step + dsnt_finish_loss_peer (three launches, same results).
FUSED_PEER_STEP = os.environ.get('DSNT_FUSED_PEER_STEP', '1') != '0'
STEP_MIN_BYTES = 32 << 20     # logits smaller than this take the one-pass step only in its single-launch form


def _step_pays(z, h, w, reg_id, sigma, group=None, sharded=None):
    """The one-pass step saves a read of the logits.  Where the single-launch form serves the case it also saves launches;
    where it does not (other shapes, KL) it takes one launch MORE than the two-kernel path, which only pays once the logits
    no longer sit in L2 (small batches are bound by launches, tools/stepbench.py).
    A sharded batch never decides from the size of the LOCAL shard: the one-pass step exchanges twice per step (mask count,
    loss sums), the two-kernel path once, and shards may be uneven or empty -- every rank must take the same path."""
    if _is_sharded(group) if sharded is None else sharded:
        return True
    if _lib.LIB.dsnt_head_step_fused_supported(_lib.dtype_id(z), h, w, reg_id, sigma):
        return True
    return z.numel() * z.element_size() >= STEP_MIN_BYTES


def takes_one_pass(z, reg, sigma, sharded):
    """The kernel-selection rule of `dsnt_head(one_pass=True)`: a pure function of dtype, heatmap shape, regulariser and --
    single process only -- the size of the batch."""
    h, w = int(z.shape[-2]), int(z.shape[-1])
    key = (z.dtype, h, w, reg, sigma, sharded, z.numel() * z.element_size() >= STEP_MIN_BYTES, STEP_MIN_BYTES,
           USE_PAIR_STEP, USE_PAIR_STEP_BF16)
    hit = _one_pass_cache.get(key)
    if hit is None:
        hit = step_supported(z, reg) and _step_pays(z, h, w, _lib.REG_IDS[reg], sigma, group=None, sharded=sharded)
        _one_pass_cache[key] = hit
    return hit


_one_pass_cache = {}


def step_supported(z, reg=None):
    """True when `dsnt_head_step` (one pass over the logits) can take heatmaps of this dtype and SHAPE (the batch size plays
    no part, so that the ranks of a sharded batch agree): at least four of them fit in shared memory, or -- for larger
    ones, given the regulariser -- the cluster-of-two-CTAs kernel serves the case (256x256 fp32; not KL; bf16 only with
    `USE_PAIR_STEP_BF16`, where the two-kernel path is the faster one)."""
    if z.dtype not in (torch.float32, torch.bfloat16) or z.dim() < 2 or z.shape[-1] == 0 or z.shape[-2] == 0:
        return False
    if reg is not None and USE_PAIR_STEP and (z.dtype == torch.float32 or USE_PAIR_STEP_BF16) and \
            _lib.LIB.dsnt_head_step_pair_supported(_lib.dtype_id(z), int(z.shape[-2]), int(z.shape[-1]), _lib.REG_IDS[reg]):
        return True
    return bool(_lib.LIB.dsnt_head_step_supported(_lib.dtype_id(z), int(z.shape[-2]), int(z.shape[-1])))


class _StepArena:
    """One allocation for the per-heatmap outputs and the loss blocks of a one-pass step (six torch.empty calls cost more
    host time than the kernel takes at the small BASELINE configs): floats [stats n*8 | coords n*2 | terms n*2 | out8 | cnt8]."""

    __slots__ = ('buf', 'n', 'base')

    def __init__(self, n, dev):
        self.n = n
        self.buf = torch.empty(12 * n + 16, dtype=torch.float32, device=dev)
        self.base = self.buf.data_ptr()

    @property
    def stats_ptr(self):
        return self.base

    @property
    def coords_ptr(self):
        return self.base + 32 * self.n

    @property
    def terms_ptr(self):
        return self.base + 40 * self.n

    @property
    def out8_ptr(self):
        return self.base + 48 * self.n

    @property
    def cnt8_ptr(self):
        return self.base + 48 * self.n + 32

    def coords(self):
        return self.buf[8 * self.n:10 * self.n]

    def out8(self):
        return self.buf[12 * self.n:12 * self.n + 8]


class _FusedHeadStep(torch.autograd.Function):
    """One pass over the logits for the whole training step (include/dsnt_b200.h: dsnt_head_step_fused, or dsnt_mask_count +
    dsnt_head_step + dsnt_finish_loss).  The forward already writes dL/dz for d(loss) = 1; the FIRST backward hands that
    buffer out, scaled in place by the actual d(loss) (`dsnt_scale_unless_one`: no traffic when it is 1, as in
    `loss.backward()`), and forgets it: the buffer now belongs to the caller (it usually becomes `z.grad`).  Any further
    backward through the node (retain_graph=True, two losses sharing the head) and any gradient w.r.t. the coordinates go
    through the regular backward kernel on the saved statistics, into a fresh tensor."""

    @staticmethod
    def forward(ctx, z, target, mask, cfg):
        reg_id, sigma, reg_coeff, flags, group, aux = cfg      # one argument: Function.apply walks every one of them
        zc, n, h, w = _flat_heatmaps(z)
        dev = zc.device
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            ar = _StepArena(n, dev)
            coords = torch.empty(z.shape[:-2] + (2,), dtype=torch.float32, device=dev)     # in its final shape: no views
            coords_ptr = coords.data_ptr()
            dz = torch.empty_like(zc)
            ws = _lib.finish_workspace(dev, stream)
            dt = _lib.dtype_id(zc)
            sharded = _is_sharded(group)
            fused = n > 0 and bool(_lib.LIB.dsnt_head_step_fused_supported(dt, h, w, reg_id, sigma))
            peer = PeerExchange.get(group, dev) if sharded and FUSED_PEER_STEP else None
            if sharded and peer is not None and fused:
                # one launch per rank: mask count and loss sums cross the ranks inside the kernel (peer memory)
                _lib.call('dsnt_head_step_fused_peer', zc.data_ptr(), dt, n, h, w, _lib.ptr(target),
                          _lib.ptr(mask), None, reg_coeff, reg_id, sigma, flags, coords_ptr, ar.stats_ptr,
                          dz.data_ptr(), ar.out8_ptr, ws.data_ptr(), *peer.args(), stream)
            elif not sharded and fused:
                # one launch: the kernel adds up the mask itself and its last CTA composes the loss
                _lib.call('dsnt_head_step_fused', zc.data_ptr(), dt, n, h, w, _lib.ptr(target), _lib.ptr(mask),
                          None, reg_coeff, reg_id, sigma, flags, coords_ptr, ar.stats_ptr, dz.data_ptr(),
                          ar.out8_ptr, ws.data_ptr(), stream)
            else:
                # the denominator of masked_average depends on the mask alone: known before the forward.  A rank whose
                # shard is empty (or that the single-launch kernel does not serve) meets the others in the same two exchanges.
                finish_loss_ptr(None, mask, n, 1, reg_coeff, ar.cnt8_ptr, ar.buf, ws, group, dev, stream)
                _lib.call('dsnt_head_step', zc.data_ptr(), dt, n, h, w, _lib.ptr(target), _lib.ptr(mask),
                          ar.cnt8_ptr + 12, None, reg_coeff, reg_id, sigma, flags, coords_ptr, ar.stats_ptr,
                          ar.terms_ptr, dz.data_ptr(), stream)
                finish_loss_ptr(ar.terms_ptr, mask, n, 1, reg_coeff, ar.out8_ptr, ar.buf, ws, group, dev, stream)
        ctx.save_for_backward(zc, target, mask, ar.buf)
        ctx.dz_box = [dz]          # NOT a saved tensor: handed out once, then gone (see the class docstring)
        ctx.meta = (n, h, w, reg_id, sigma, reg_coeff, flags, z.shape)
        ctx.set_materialize_grads(False)
        aux['out8'] = ar.buf       # the loss block sits at float offset 12 n of the arena
        aux['off'] = 12 * n
        aux['dz'] = dz
        return coords, ar.buf[12 * n + 6]

    @staticmethod
    def backward(ctx, g_coords, g_loss):
        zc, target, mask, arena = ctx.saved_tensors
        n, h, w, reg_id, sigma, reg_coeff, flags, shape = ctx.meta
        dev = zc.device
        if g_coords is None and g_loss is None:
            return (None,) * 4
        with _lib.on_device(dev):
            stream = _lib.stream_of(zc)
            if g_loss is not None and (g_loss.dtype is not torch.float32 or not g_loss.is_contiguous()):
                g_loss = g_loss.to(torch.float32).contiguous()
            dz = ctx.dz_box[0]
            if g_coords is None and dz is not None:
                # the usual case (train.py:381): the gradient is already there
                ctx.dz_box[0] = None
                _lib.call('dsnt_scale_unless_one', dz.data_ptr(), _lib.dtype_id(dz), dz.numel(), g_loss.data_ptr(), stream)
                return (dz if dz.shape == shape else dz.view(shape)), None, None, None
            if g_coords is not None:
                g_coords = g_coords.to(torch.float32).contiguous()
            full = torch.empty_like(zc)
            base = arena.data_ptr()
            _lib.call('dsnt_head_bwd', zc.data_ptr(), _lib.dtype_id(zc), 1, n, h, w,
                      _lib.ptr(target), _lib.ptr(mask), base, _lib.ptr(g_coords), None,
                      _lib.ptr(g_loss), base + 48 * n + 12 if g_loss is not None else None,
                      reg_coeff, reg_id, sigma, flags, full.data_ptr(), 0, stream)
        return full.view(shape), None, None, None


class _FusedHeadStacked(torch.autograd.Function):
    """All hourglass stacks in ONE forward launch, one finishing reduction and ONE backward launch
    (dsnt_head_fwd_stacked / dsnt_finish_loss_stacked / dsnt_head_bwd_stacked, include/dsnt_b200.h)."""

    @staticmethod
    def forward(ctx, target, mask, reg_id, sigma, reg_coeff, flags, group, variant, aux, *zs):
        flat = [_flat_heatmaps(z) for z in zs]
        zcs = [f[0] for f in flat]
        n, h, w = flat[0][1], flat[0][2], flat[0][3]
        s_count = len(zcs)
        dev = zcs[0].device
        with _lib.on_device(dev):
            stream = _lib.stream_of(zcs[0])
            coords = torch.empty(s_count, n, 2, dtype=torch.float32, device=dev)
            stats = torch.empty(s_count * n, _lib.STATS_K, dtype=torch.float32, device=dev)
            terms = torch.empty(s_count * n, 2, dtype=torch.float32, device=dev)
            _lib.call('dsnt_head_fwd_stacked', _lib.ptr_array(zcs), s_count, _lib.dtype_id(zcs[0]), 1, n, h, w,
                      _lib.ptr(target), reg_id, sigma, coords.data_ptr(), stats.data_ptr(), terms.data_ptr(),
                      variant, stream)
            if target is None and reg_id == 0 and not _is_sharded(group):
                out8 = _no_loss_block(dev)
            else:
                out8 = torch.empty(8, dtype=torch.float32, device=dev)
                ws = _lib.finish_workspace(dev, stream)
                finish_loss(terms, mask, n, s_count, reg_coeff, out8, ws, group, dev, stream)
        ctx.save_for_backward(target, mask, stats, out8, *zcs)
        ctx.meta = (n, h, w, reg_id, sigma, reg_coeff, flags, variant, [z.shape for z in zs])
        ctx.set_materialize_grads(False)
        aux['out8'] = out8
        lead = zs[0].shape[:-2]
        return (out8[6],) + coords.view(s_count, *lead, 2).unbind(0)

    @staticmethod
    def backward(ctx, g_loss, *g_coords):
        target, mask, stats, out8 = ctx.saved_tensors[:4]
        zcs = ctx.saved_tensors[4:]
        n, h, w, reg_id, sigma, reg_coeff, flags, variant, shapes = ctx.meta
        s_count = len(zcs)
        dev = zcs[0].device
        if g_loss is None and all(g is None for g in g_coords):
            return (None,) * (9 + s_count)
        with _lib.on_device(dev):
            stream = _lib.stream_of(zcs[0])
            gc = None
            if any(g is not None for g in g_coords):
                gc = torch.zeros(s_count, n, 2, dtype=torch.float32, device=dev)
                for i, g in enumerate(g_coords):
                    if g is not None:
                        gc[i] = g.reshape(n, 2).to(torch.float32)
            if g_loss is not None:
                g_loss = g_loss.to(torch.float32).contiguous()
            dzs = [torch.empty_like(z) for z in zcs]
            _lib.call('dsnt_head_bwd_stacked', _lib.ptr_array(zcs), _lib.ptr_array(dzs), s_count,
                      _lib.dtype_id(zcs[0]), 1, n, h, w, _lib.ptr(target), _lib.ptr(mask), stats.data_ptr(),
                      _lib.ptr(gc), None, _lib.ptr(g_loss), out8[3:4].data_ptr() if g_loss is not None else None,
                      reg_coeff, reg_id, sigma, flags, variant, stream)
        return (None,) * 9 + tuple(dz.view(shape) for dz, shape in zip(dzs, shapes))


class _FusedHeadStackedStep(torch.autograd.Function):
    """All hourglass stacks in ONE launch that is forward and backward at once (dsnt_head_step_fused_stacked,
    include/dsnt_b200.h): coordinates, the summed loss and dL/dz of every stack, each heatmap read once.  The first backward
    hands the stored gradients out (one contiguous buffer for all stacks, scaled in place when d(loss) != 1) and forgets
    them; a further backward through the node, or gradients w.r.t. the coordinates, go through dsnt_head_bwd_stacked on the
    saved statistics (see _FusedHeadStep)."""

    @staticmethod
    def forward(ctx, target, mask, reg_id, sigma, reg_coeff, flags, aux, *zs):
        flat = [_flat_heatmaps(z) for z in zs]
        zcs = [f[0] for f in flat]
        n, h, w = flat[0][1], flat[0][2], flat[0][3]
        s_count = len(zcs)
        dev = zcs[0].device
        with _lib.on_device(dev):
            stream = _lib.stream_of(zcs[0])
            ar = _StepArena(s_count * n, dev)
            dz = torch.empty((s_count,) + tuple(zcs[0].shape), dtype=zcs[0].dtype, device=dev)
            ws = _lib.finish_workspace(dev, stream)
            hm_bytes = dz[0].numel() * dz.element_size()
            dz_ptrs = (ctypes.c_void_p * s_count)(*[dz.data_ptr() + i * hm_bytes for i in range(s_count)])
            _lib.call('dsnt_head_step_fused_stacked', _lib.ptr_array(zcs), dz_ptrs,
                      s_count, _lib.dtype_id(zcs[0]), n, h, w, _lib.ptr(target), _lib.ptr(mask), None, reg_coeff, reg_id,
                      sigma, flags, ar.coords_ptr, ar.stats_ptr, ar.out8_ptr, ws.data_ptr(), stream)
        ctx.save_for_backward(target, mask, ar.buf, *zcs)
        ctx.dz_box = [dz]
        ctx.meta = (n, h, w, reg_id, sigma, reg_coeff, flags, [z.shape for z in zs])
        ctx.set_materialize_grads(False)
        aux['out8'] = ar.out8()
        lead = zs[0].shape[:-2]
        return (aux['out8'][6],) + ar.coords().view(s_count, *lead, 2).unbind(0)

    @staticmethod
    def backward(ctx, g_loss, *g_coords):
        target, mask, arena = ctx.saved_tensors[:3]
        zcs = ctx.saved_tensors[3:]
        n, h, w, reg_id, sigma, reg_coeff, flags, shapes = ctx.meta
        s_count = len(zcs)
        dev = zcs[0].device
        if g_loss is None and all(g is None for g in g_coords):
            return (None,) * (7 + s_count)
        with _lib.on_device(dev):
            stream = _lib.stream_of(zcs[0])
            if g_loss is not None:
                g_loss = g_loss.to(torch.float32).contiguous()
            dz = ctx.dz_box[0]
            no_gc = all(g is None for g in g_coords)
            if no_gc and dz is not None:
                ctx.dz_box[0] = None
                _lib.call('dsnt_scale_unless_one', dz.data_ptr(), _lib.dtype_id(dz), dz.numel(), g_loss.data_ptr(), stream)
                parts = dz.unbind(0)
                if any(prt.shape != shape for prt, shape in zip(parts, shapes)):
                    parts = tuple(prt.view(shape) for prt, shape in zip(parts, shapes))
                return (None,) * 7 + parts
            gc = None
            if not no_gc:
                gc = torch.zeros(s_count, n, 2, dtype=torch.float32, device=dev)
                for i, g in enumerate(g_coords):
                    if g is not None:
                        gc[i] = g.reshape(n, 2).to(torch.float32)
            dzs = [torch.empty_like(z) for z in zcs]
            base = arena.data_ptr()
            _lib.call('dsnt_head_bwd_stacked', _lib.ptr_array(zcs), _lib.ptr_array(dzs), s_count,
                      _lib.dtype_id(zcs[0]), 1, n, h, w, _lib.ptr(target), _lib.ptr(mask), base,
                      _lib.ptr(gc), None, _lib.ptr(g_loss), base + 48 * s_count * n + 12 if g_loss is not None else None,
                      reg_coeff, reg_id, sigma, flags, 0, stream)
        return (None,) * 7 + tuple(d.view(shape) for d, shape in zip(dzs, shapes))


def dsnt_head_stacked(zs, target, mask=None, reg='none', sigma=None, reg_coeff=1.0, hm_sigma=None, group=None,
                      variant=0, one_pass=False):
    """Hourglass form (src/dsnt/model.py:233-246,286-292): every stack's head evaluated by one launch, losses summed.

    zs: list of logits tensors [..., H, W] of identical shape/dtype (one per stack, as hourglass.py:166-177 returns).
    one_pass: a training step (only d(loss) flows back): one launch writes every stack's dL/dz while its heatmaps are on
    chip (64x64 heatmaps, single process; other cases take the forward / backward launches as before).
    Returns (list of coords per stack, total loss = sum_s euclid_s + reg_coeff * reg_s)."""
    zs = list(zs)
    if not zs:
        raise ValueError('dsnt_head_stacked needs at least one stack')
    if len(zs) > _lib.MAX_STACKS:
        raise ValueError('at most %d stacks per call, got %d' % (_lib.MAX_STACKS, len(zs)))
    for z in zs:
        _lib.require_cuda(z, 'z')
        if z.shape != zs[0].shape or z.dtype != zs[0].dtype or z.device != zs[0].device:
            raise ValueError('all stacks must share shape, dtype and device')
    if reg not in _lib.REG_IDS:
        raise ValueError('unrecognised regulariser: %r' % (reg,))
    h, w = zs[0].shape[-2], zs[0].shape[-1]
    n = zs[0].numel() // max(h * w, 1)
    if sigma is None:
        sigma = 2.0 * (1.0 if hm_sigma is None else hm_sigma) / w
    if target is not None and target.requires_grad:
        outs = [_head_with_target_grad(z, target, mask, reg, float(sigma), float(reg_coeff), 'softmax', None, None, True, group)
                for z in zs]
        total = outs[0].loss
        for o in outs[1:]:
            total = total + o.loss
        return [o.coords for o in outs], total
    target = _as_f32(target, n, 2, 'target')
    mask = _as_f32(mask, n, 1, 'mask')
    flags = _lib.FLAG_STRICT_NAN if STRICT_NAN else 0
    aux = {}
    sharded = _is_sharded(group)
    if (one_pass and not sharded and target is not None and n > 0
            and torch.is_grad_enabled() and any(z.requires_grad for z in zs)      # validation under no_grad: forward only
            and _lib.LIB.dsnt_head_step_fused_supported(_lib.dtype_id(zs[0]), h, w, _lib.REG_IDS[reg], float(sigma))):
        out = _FusedHeadStackedStep.apply(target, mask, _lib.REG_IDS[reg], float(sigma), float(reg_coeff), flags, aux, *zs)
    else:
        out = _FusedHeadStacked.apply(target, mask, _lib.REG_IDS[reg], float(sigma), float(reg_coeff), flags, group,
                                      int(variant), aux, *zs)
    return list(out[1:]), out[0]
