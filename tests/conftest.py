import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def ensure_built():
    """The .so is a build artefact (git-ignored): compile it (nvcc cross-compiles without a GPU) if absent."""
    lib = os.path.join(ROOT, 'dsnt_pose2d_b200', 'libdsnt_b200.so')
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(['make', '-C', os.path.join(ROOT, 'dsnt_pose2d_b200', 'csrc'), '-j', '8'], check=True)
    return lib


@pytest.fixture(scope='session')
def libpath():
    return ensure_built()


class Golden:
    """Lazy view over an .npz fixture: g['case/key']."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)

    def __getitem__(self, key):
        return self._z[key]

    def __contains__(self, key):
        return key in self._z.files

    def get(self, key, default=None):
        return self._z[key] if key in self._z.files else default

    @property
    def cases(self):
        return [str(c) for c in self._z['__cases__']]


@pytest.fixture(scope='session')
def golden_head():
    return Golden('head_logits.npz')


@pytest.fixture(scope='session')
def golden_l1():
    return Golden('level1_api.npz')


@pytest.fixture(scope='session')
def golden_stacked():
    return Golden('stacked_js.npz')


def head_case_params(g, name):
    idx = g.cases.index(name)
    b, c, h, w, hm_sigma, coeff, with_mask = g['__params__'][idx]
    return int(b), int(c), int(h), int(w), float(hm_sigma), float(coeff), bool(with_mask)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max() if b.size else 0.0
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0)) if b.size else 0.0


@pytest.fixture(scope='session')
def golden_preact():
    return Golden('preact_heads.npz')


@pytest.fixture(scope='session')
def golden_flip():
    return Golden('flip_tta.npz')


@pytest.fixture(scope='session')
def golden_gauss():
    return Golden('gauss_util.npz')
