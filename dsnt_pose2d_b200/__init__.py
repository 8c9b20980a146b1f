"""dsnt_pose2d_b200 -- the DSNT head of anibali/dsnt-pose2d as hand-written sm_100a CUDA kernels.

    from dsnt_pose2d_b200 import nn          # drop-in for the reference's `dsnt.nn`
    from dsnt_pose2d_b200 import dsnt_head   # fused logits -> (coords, loss), 12 B/pixel fwd+bwd
    from dsnt_pose2d_b200.model import DSNTHead, attach_fused_head

Importing the package loads `libdsnt_b200.so`; if it has not been built the import fails (no fallback).
"""

from . import _lib
from . import nn
from .head import HeadOutput, dsnt_head, dsnt_head_stacked
from .model import DSNTHead, attach_fused_head, install_as_dsnt_nn
from .inference import MPII_HFLIP_INDICES, flip_tta_coords, predict_flipped

__all__ = ['nn', 'dsnt_head', 'dsnt_head_stacked', 'HeadOutput', 'DSNTHead', 'attach_fused_head',
           'install_as_dsnt_nn', 'library_version', 'flip_tta_coords', 'predict_flipped', 'MPII_HFLIP_INDICES']


def library_version():
    return _lib.version()
