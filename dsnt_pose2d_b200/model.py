"""Level-2 boundary: the head surface of the reference's `HumanPoseModel`
(src/dsnt/model.py:21-76,138-201,229-314) on top of the fused kernels.

The reference splits the head over two calls with the target arriving only in the second one
(`train.py:355-358`):

    out  = model.forward_part2(Z)                 # softmax + dsnt, remembers self.heatmaps = P
    loss = model.forward_loss(out, target, mask)  # euclid(out) + reg_coeff * reg(self.heatmaps)

`DSNTHead` keeps exactly that surface.  `forward_part2` runs the coordinate kernel (one read of Z);
`forward_loss`, when handed the very tensor `forward_part2` returned, evaluates the fused loss on the
remembered logits, so backward is ONE streaming kernel writing dL/dZ.  P is never stored: `.heatmaps`
materialises it on demand (consumers: visualisation in train.py:394,438, gui.py:31, tests).

`attach_fused_head(model)` grafts these methods onto an existing reference model object (ResNet or
hourglass flavour) so `train.py` / `infer.py` loops run unchanged with the backbone left to cuDNN.
"""

import types

import torch
from torch import nn

from . import nn as dnn
from . import _lib
from .head import dsnt_head, dsnt_head_stacked

_FUSED_PREACTS = ('softmax', 'thresholded_softmax', 'abs', 'relu', 'sigmoid')   # src/dsnt/model.py:29-41
_STACKED_PREACTS = ('softmax',)


def hm_preact(x, preact):
    """src/dsnt/model.py:24-45 for every `--preact` choice, returning normalised heatmaps [-1,C,H,W]."""
    n_chans, height, width = x.size(-3), x.size(-2), x.size(-1)
    flat = x.reshape(-1, height * width)
    if preact == 'softmax':
        flat = dnn.thresholded_softmax(flat, float('-inf'), 0.0)
    elif preact == 'thresholded_softmax':
        flat = dnn.thresholded_softmax(flat, -0.5)
    elif preact in ('abs', 'relu', 'sigmoid'):
        flat = {'abs': torch.abs, 'relu': torch.relu, 'sigmoid': torch.sigmoid}[preact](flat)
        flat = flat / (flat.sum(-1, keepdim=True) + 1e-12)
    else:
        raise Exception('unrecognised heatmap preactivation function: {}'.format(preact))
    return flat.view(-1, n_chans, height, width)


class DSNTHead(nn.Module):
    """Parameter-free DSNT head with the reference model's head interface.

    Args mirror `ResNetHumanPoseModel` / `HourglassHumanPoseModel` (src/dsnt/model.py:90-99,205-215).
    """

    def __init__(self, n_chans=16, preact='softmax', reg='none', reg_coeff=1.0, hm_sigma=1.0, group=None,
                 one_pass=True):
        super().__init__()
        self.n_chans = n_chans
        self.output_strat = 'dsnt'
        self.preact = preact
        self.reg = reg
        self.reg_coeff = reg_coeff
        self.hm_sigma = hm_sigma
        self.group = group
        self.one_pass = one_pass   # forward_loss also writes dL/dZ (dsnt_head_step): backward re-reads nothing
        self._logits = []          # raw heatmap tensors of the last forward_part2 (one per stack)
        self._coords = []          # what forward_part2 returned for them
        self._heatmaps = {}        # lazily materialised P per stack

    # ---- forward_part2 (src/dsnt/model.py:176-194, 278-307)
    def _part2_one(self, z):
        if self.preact in _FUSED_PREACTS:
            coords = dsnt_head(z, None, None, reg='none', preact=self.preact).coords
        else:
            coords = dnn.dsnt(hm_preact(z, self.preact))
        self._logits.append(z)
        self._coords.append(coords)
        return coords

    def _stackable(self, zs):
        return (self.preact in _STACKED_PREACTS and 1 < len(zs) <= _lib.MAX_STACKS
                and all(z.shape == zs[0].shape and z.dtype == zs[0].dtype and z.device == zs[0].device for z in zs))

    def forward_part2(self, x):
        self._logits, self._coords, self._heatmaps = [], [], {}
        if isinstance(x, (list, tuple)):
            zs = list(x)
            if self._stackable(zs):
                # hourglass: every stack in ONE launch (src/dsnt/model.py:286-292 loops in Python)
                coords, _ = dsnt_head_stacked(zs, None, None, reg='none')
                self._logits, self._coords = zs, coords
                return coords
            return [self._part2_one(z) for z in zs]
        return self._part2_one(x)

    forward = forward_part2

    # ---- lazily materialised heatmaps (src/dsnt/model.py:181,229-231,287-290)
    def _heatmap(self, i):
        if i not in self._heatmaps:
            self._heatmaps[i] = hm_preact(self._logits[i], self.preact)
        return self._heatmaps[i]

    @property
    def heatmaps(self):
        return self._heatmap(0)                      # hourglass: FIRST stack (model.py:229-231)

    @property
    def heatmaps_array(self):
        return [self._heatmap(i) for i in range(len(self._logits))]

    # ---- forward_loss (src/dsnt/model.py:138-145, 233-246)
    def _loss_one(self, i, out, target, mask):
        if i < len(self._coords) and out is self._coords[i] and self.preact in _FUSED_PREACTS:
            return dsnt_head(self._logits[i], target, mask, reg=self.reg, hm_sigma=self.hm_sigma,
                             reg_coeff=self.reg_coeff, group=self.group, preact=self.preact,
                             one_pass=self.one_pass).loss
        # coords that did not come from forward_part2 (or a non-fused preact): the reference's composition
        loss = dnn.euclidean_loss(out, target, mask)
        sigma = 2.0 * self.hm_sigma / self._logits[i].size(-1)
        fn = {'var': dnn.variance_reg_loss, 'kl': dnn.kl_reg_loss, 'js': dnn.js_reg_loss,
              'mse': dnn.mse_reg_loss}.get(self.reg)
        if fn is not None:
            loss = loss + self.reg_coeff * fn(self._heatmap(i), target, sigma, mask)
        return loss

    def forward_loss(self, out_var, target_var, mask_var):
        if isinstance(out_var, (list, tuple)):
            if (len(out_var) == len(self._coords) and self._stackable(self._logits)
                    and all(o is c for o, c in zip(out_var, self._coords))):
                # one fused forward over all stacks, one finishing reduction; backward is one launch too
                return dsnt_head_stacked(self._logits, target_var, mask_var, reg=self.reg, hm_sigma=self.hm_sigma,
                                         reg_coeff=self.reg_coeff, group=self.group, one_pass=self.one_pass)[1]
            total = 0
            for i, out in enumerate(out_var):          # sum over stacks (model.py:238-246)
                total = total + self._loss_one(i, out, target_var, mask_var)
            return total
        return self._loss_one(0, out_var, target_var, mask_var)

    # ---- compute_coords (src/dsnt/model.py:161-163, 262-267)
    def compute_coords(self, out_var):
        if isinstance(out_var, (list, tuple)):
            out_var = out_var[-1]                       # hourglass: LAST stack
        return out_var.detach().to('cpu', torch.float32)


def attach_fused_head(model, group=None):
    """Route `forward_part2` / `forward_loss` / `compute_coords` / `.heatmaps` of a reference-style pose model
    through `DSNTHead`.  `model` only needs the attributes the reference models carry
    (`output_strat`, `preact`, `reg`, `reg_coeff`, `hm_sigma`, `n_chans`, `forward_part1`).

    Only the 'dsnt' output strategy is the hot path; other strategies are left untouched.
    """
    if getattr(model, 'output_strat', 'dsnt') != 'dsnt':
        return model
    head = DSNTHead(n_chans=getattr(model, 'n_chans', 16), preact=model.preact, reg=model.reg,
                    reg_coeff=model.reg_coeff, hm_sigma=model.hm_sigma, group=group)
    object.__setattr__(model, '_dsnt_b200_head', head)          # not a registered submodule: no parameters

    def forward_part2(self, x):
        return self._dsnt_b200_head.forward_part2(x)

    def forward_loss(self, out_var, target_var, mask_var):
        return self._dsnt_b200_head.forward_loss(out_var, target_var, mask_var)

    def compute_coords(self, out_var):
        return self._dsnt_b200_head.compute_coords(out_var)

    def forward(self, *inputs):
        return self.forward_part2(self.forward_part1(inputs[0]))

    cls = type(model)
    patched = type(cls.__name__ + 'B200', (cls,), {
        'forward_part2': forward_part2, 'forward_loss': forward_loss, 'compute_coords': compute_coords,
        'forward': forward,
        'heatmaps': property(lambda self: self._dsnt_b200_head.heatmaps),
        'heatmaps_array': property(lambda self: self._dsnt_b200_head.heatmaps_array),
    })
    model.__dict__.pop('heatmaps', None)
    model.__dict__.pop('heatmaps_array', None)
    model.__class__ = patched
    return model


def install_as_dsnt_nn():
    """Register this package's `nn` module as `dsnt.nn` so the reference's own `dsnt/model.py` (which does
    `import dsnt.nn` and `from dsnt.nn import euclidean_loss, thresholded_softmax`, model.py:15-16) binds to the
    CUDA operators.  Must run BEFORE `dsnt.model` is imported (SURVEY.md 8b binding note)."""
    import importlib
    import sys
    pkg = sys.modules.get('dsnt')
    if pkg is None:
        try:
            pkg = importlib.import_module('dsnt')      # the reference's package (its __init__ is empty): keeps dsnt.model importable
        except ImportError:
            pkg = types.ModuleType('dsnt')
            pkg.__path__ = []
            sys.modules['dsnt'] = pkg
    sys.modules['dsnt.nn'] = dnn
    pkg.nn = dnn
    return dnn
