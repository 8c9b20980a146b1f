#!/usr/bin/env python
"""Golden vectors at BASELINE config 5's heatmap size (256x256) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_cfg5.py

Same recipe as make_golden.py (the reference's head evaluated in float64 on CPU, tests/common.py:18), on the inputs of
cfg5_inputs.py; writes head_256.npz: loss / euclid / reg / coords, dL/dZ at the sampled pixels and its per-heatmap L2 norm,
for every regulariser."""

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from cfg5_inputs import CASES, make_case  # noqa: E402
from make_golden import import_reference, ref_head  # noqa: E402


def main():
    torch.set_default_dtype(torch.float64)
    warnings.simplefilter('ignore')
    ref_nn, ref_model = import_reference()
    out = {}
    for name, (b, c, kind, hm_sigma, coeff, seed) in CASES.items():
        z32, target, mask, idx = make_case(name)
        for reg in ('none', 'var', 'kl', 'js', 'mse'):
            z = torch.from_numpy(z32).double().requires_grad_(True)
            loss, coords, euc, rv, _ = ref_head(ref_nn, ref_model, z, torch.from_numpy(target).double(),
                                                torch.from_numpy(mask).double(), reg, hm_sigma, coeff)
            loss.backward()
            dz = z.grad.numpy()
            out['%s/%s/loss' % (name, reg)] = np.float64(loss.item())
            out['%s/%s/euclid' % (name, reg)] = np.float64(euc.item())
            out['%s/%s/reg' % (name, reg)] = np.float64(float(rv))
            out['%s/%s/dz_samples' % (name, reg)] = dz.reshape(-1)[idx].copy()
            out['%s/%s/dz_norms' % (name, reg)] = np.sqrt((dz.reshape(b * c, -1) ** 2).sum(-1))
            if reg == 'none':
                out[name + '/coords'] = coords.detach().numpy()
            print('%-14s %-4s loss %.9f reg %.6e |dz| %.3e' % (name, reg, loss.item(), float(rv), np.linalg.norm(dz)))
    out['__cases__'] = np.array(list(CASES))
    np.savez_compressed(os.path.join(HERE, 'head_256.npz'), **out)


if __name__ == '__main__':
    main()
