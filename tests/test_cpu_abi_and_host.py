"""CPU-side checks: the C-ABI library loads and exports every symbol include/dsnt_b200.h declares (no compute
calls without a GPU), the host logic (argument validation, sigma conversion, sharding) behaves, and the
product never falls back to a CPU path."""

import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'dsnt_b200.h')).read()
    return re.findall(r'DSNT_API\s+[\w\s\*]+?\b(dsnt_\w+)\s*\(', text)


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ('dsnt_head_fwd', 'dsnt_head_bwd', 'dsnt_finish_loss', 'dsnt_combine_loss', 'dsnt_euclid_fwd',
                 'dsnt_euclid_bwd', 'dsnt_tsoftmax_fwd', 'dsnt_tsoftmax_bwd', 'dsnt_make_gauss_fwd',
                 'dsnt_make_gauss_bwd', 'dsnt_b200_version', 'dsnt_b200_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_symbols():
        assert hasattr(lib, name), 'libdsnt_b200.so does not export %s' % name
    lib.dsnt_b200_version.restype = ctypes.c_int
    assert lib.dsnt_b200_version() >= 100


def test_binding_table_matches_header(libpath):
    from dsnt_pose2d_b200 import _lib
    assert sorted(_lib.SIGNATURES) == sorted(declared_symbols())
    text = open(os.path.join(ROOT, 'include', 'dsnt_b200.h')).read()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        proto = re.search(r'DSNT_API[^;(]*\b%s\s*\(([^;]*)\)\s*;' % name, text).group(1)
        n_args = 0 if proto.strip() == 'void' else proto.count(',') + 1
        assert n_args == len(argtypes), (name, n_args, len(argtypes))


def test_argument_validation_without_gpu(libpath):
    """Rejected calls return before any CUDA work, so they are safe on a GPU-less host."""
    from dsnt_pose2d_b200 import _lib
    lib = _lib.LIB
    assert lib.dsnt_head_fwd(None, 0, 1, 4, 8, 8, None, 0, 1.0, None, None, None, 0, None) == -1
    assert 'null heatmap' in _lib.last_error()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    assert lib.dsnt_head_fwd(addr, 7, 1, 1, 4, 4, None, 0, 1.0, addr, None, None, 0, None) == -2   # dtype
    assert lib.dsnt_head_fwd(addr, 0, 1, 1, 4, 4, None, 3, 1.0, addr, None, None, 0, None) == -1   # JS w/o target
    assert 'needs a target' in _lib.last_error()
    assert lib.dsnt_head_fwd(addr, 0, 1, 1, 4, 4, addr, 3, 0.0, addr, None, None, 0, None) == -1   # sigma
    assert lib.dsnt_head_fwd(addr, 0, 1, 1, 4, 4, None, 9, 1.0, addr, None, None, 0, None) == -1   # reg id
    assert lib.dsnt_head_fwd(addr, 0, 1, 0, 4, 4, None, 0, 1.0, addr, None, None, 0, None) == 0    # n = 0: no-op
    assert lib.dsnt_head_bwd(addr, 0, 1, 1, 4, 4, None, None, None, None, None, None, None, 1.0, 0, 1.0, 0,
                             addr, 0, None) == -1                                                   # no stats
    assert lib.dsnt_head_bwd(addr, 0, 1, 1, 4, 4, None, None, addr, None, None, addr, None, 1.0, 0, 1.0, 0,
                             addr, 0, None) == -1                                                   # g_loss w/o denom
    assert lib.dsnt_finish_workspace_bytes() >= 4 * 128 * 4 + 4


def test_no_cpu_fallback_anywhere():
    import dsnt_pose2d_b200 as dp
    z = torch.randn(1, 2, 8, 8)
    t = torch.zeros(1, 2, 2)
    for call in (lambda: dp.dsnt_head(z, t), lambda: dp.nn.dsnt(z), lambda: dp.nn.flat_softmax(z),
                 lambda: dp.nn.softmax_2d(z), lambda: dp.nn.thresholded_softmax(z, 0.0),
                 lambda: dp.nn.js_reg_loss(z, t, 0.1), lambda: dp.nn.kl_reg_loss(z, t, 0.1),
                 lambda: dp.nn.mse_reg_loss(z, t, 0.1), lambda: dp.nn.variance_reg_loss(z, t, 0.1),
                 lambda: dp.nn.euclidean_loss(t, t), lambda: dp.nn.make_gauss(t, 8, 8, 0.1),
                 lambda: dp.nn.masked_average(t)):
        with pytest.raises(NotImplementedError):
            call()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'dsnt_pose2d_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src, fn


def test_api_names_match_reference_nn():
    """Every public name of src/dsnt/nn.py (plus north_star's flat_softmax) exists with the same signature."""
    import inspect
    from dsnt_pose2d_b200 import nn
    expect = {
        'generate_xy': ['inp'], 'expectation_2d': ['values', 'probabilities'], 'dsnt': ['heatmaps'],
        'masked_average': ['losses', 'mask'], 'euclidean_loss': ['actual', 'target', 'mask'],
        'thresholded_softmax': ['inp', 'threshold', 'eps'], 'softmax_2d': ['inp'], 'flat_softmax': ['inp'],
        'make_gauss': ['coords', 'width', 'height', 'sigma'],
        'kl_reg_loss': ['heatmaps', 'mu_t', 'sigma_t', 'mask'], 'js_reg_loss': ['heatmaps', 'mu_t', 'sigma_t', 'mask'],
        'mse_reg_loss': ['heatmaps', 'mu_t', 'sigma_t', 'mask'],
        'variance_reg_loss': ['heatmaps', 'mu_t', 'sigma_t', 'mask'],
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(nn, name)).parameters) == params, name
    assert hasattr(nn, 'ThresholdedSoftmax')


def test_shard_range_partitions_the_batch():
    from dsnt_pose2d_b200.parallel import shard_range
    for batch in (0, 1, 7, 32, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_step_support_queries_are_host_logic(libpath):
    """Which kernel serves which case is decided on the host (no GPU needed): the shared-memory ring, the 64x64
    single-launch kernel with its Gaussian-window register slots, the cluster-of-two-CTAs kernel for 256x256 fp32."""
    from dsnt_pose2d_b200 import _lib
    lib, reg = _lib.LIB, _lib.REG_IDS
    f32, bf16 = _lib.DTYPE_F32, _lib.DTYPE_BF16
    # at least four heatmaps in 224 KiB of shared memory, 16-byte vectors
    assert lib.dsnt_head_step_supported(f32, 64, 64) == 1 and lib.dsnt_head_step_supported(bf16, 128, 128) == 1
    assert lib.dsnt_head_step_supported(f32, 128, 128) == 0 and lib.dsnt_head_step_supported(f32, 256, 256) == 0
    assert lib.dsnt_head_step_supported(f32, 7, 7) == 0 and lib.dsnt_head_step_supported(5, 64, 64) == 0
    # single-launch form: 64x64; JS / MSE: the widest possible Gaussian window must fit the register slots (KL walks its window)
    sigma_1px = 2.0 / 64
    for r in ('none', 'var', 'js', 'mse'):
        assert lib.dsnt_head_step_fused_supported(f32, 64, 64, reg[r], sigma_1px) == 1
        assert lib.dsnt_head_step_fused_supported(bf16, 64, 64, reg[r], sigma_1px) == 1
        assert lib.dsnt_head_step_fused_supported(f32, 32, 32, reg[r], 2.0 / 32) == 0
    assert lib.dsnt_head_step_fused_supported(f32, 64, 64, reg['kl'], sigma_1px) == 1
    assert lib.dsnt_head_step_fused_supported(bf16, 64, 64, reg['kl'], 20 * sigma_1px) == 1     # any sigma
    assert lib.dsnt_head_step_fused_supported(f32, 32, 32, reg['kl'], 2.0 / 32) == 0
    assert lib.dsnt_head_step_fused_supported(f32, 64, 64, reg['js'], 20 * sigma_1px) == 0      # window = whole image
    assert lib.dsnt_head_step_fused_supported(f32, 64, 64, reg['var'], 20 * sigma_1px) == 1     # no window to fit
    # 256x256 fp32: a pair of CTAs, without a Gaussian window
    assert lib.dsnt_head_step_pair_supported(f32, 256, 256, reg['var']) == 1
    assert lib.dsnt_head_step_pair_supported(f32, 256, 256, reg['none']) == 1
    assert lib.dsnt_head_step_pair_supported(f32, 256, 256, reg['js']) == 1                     # window terms from the registers
    assert lib.dsnt_head_step_pair_supported(f32, 256, 256, reg['kl']) == 0
    assert lib.dsnt_head_step_pair_supported(bf16, 256, 256, reg['var']) == 1                   # round 2: bf16 too
    assert lib.dsnt_head_step_pair_supported(bf16, 256, 256, reg['kl']) == 0
    assert lib.dsnt_head_step_pair_supported(f32, 128, 128, reg['var']) == 0
    assert lib.dsnt_head_step_supported_reg(f32, 256, 256, reg['var']) == 1
    assert lib.dsnt_head_step_supported_reg(f32, 256, 256, reg['js']) == 1
    assert lib.dsnt_head_step_supported_reg(f32, 256, 256, reg['kl']) == 0
    assert lib.dsnt_head_step_supported_reg(bf16, 256, 256, reg['js']) == 1                      # served (the dispatcher still prefers
                                                                                                 # the two-kernel path: head.USE_PAIR_STEP_BF16)
    assert lib.dsnt_head_step_supported_reg(bf16, 256, 256, reg['kl']) == 0                      # two-kernel path
    # exchange buffer of the peer reductions: two parities x 16 ranks x float4
    assert lib.dsnt_peer_exchange_bytes() == 2 * 16 * 4 * 8     # two parities x 16 ranks x 4 words of 8 bytes
    assert lib.dsnt_finish_workspace_bytes() >= (256 * 4 + 4 + 256) * 4


def test_peer_entry_points_validate_their_arguments(libpath):
    from dsnt_pose2d_b200 import _lib
    lib = _lib.LIB
    buf = (ctypes.c_float * 2048)()
    addr = ctypes.addressof(buf)
    peers = (ctypes.c_void_p * 2)(addr, addr + 1024)
    pp = ctypes.cast(peers, ctypes.c_void_p)
    assert lib.dsnt_mask_count_peer(None, 4, addr, addr, pp, 0, 17, addr, addr, None) == -1      # too many ranks
    assert 'at most 16' in _lib.last_error()
    assert lib.dsnt_mask_count_peer(None, 4, addr, addr, pp, 2, 2, addr, addr, None) == -1       # rank out of range
    assert lib.dsnt_mask_count_peer(None, 4, addr, addr, None, 0, 2, addr, addr, None) == -1     # no peer table
    bad = (ctypes.c_void_p * 2)(addr, None)
    assert lib.dsnt_finish_loss_peer(addr, None, 4, 1, 1.0, addr, addr, ctypes.cast(bad, ctypes.c_void_p), 0, 2, addr, addr,
                                     None) == -1
    assert 'rank 1' in _lib.last_error()
    assert lib.dsnt_head_step_fused(addr, 0, 4, 32, 32, addr, None, None, 1.0, _lib.REG_IDS['kl'], 2.0 / 32, 0, addr, addr,
                                    addr, addr, addr, None) == -2                               # 32x32: three-launch form
    assert lib.dsnt_head_step_fused(addr, 0, 4, 64, 64, addr, None, None, 1.0, 0, 1.0, 0, addr, addr, addr, None, addr,
                                    None) == -1                                                 # no loss block


def test_one_pass_dispatch_rule_without_gpu():
    """head._step_pays: the one-pass step is taken where it saves launches or where the logits no longer sit in L2."""
    from dsnt_pose2d_b200 import head, _lib

    class FakeZ:                        # only what the rule looks at
        def __init__(self, n, h, w, dtype):
            self.shape, self.dtype, self._n = (n, h, w), dtype, n * h * w
            self.device = torch.device('cpu')

        def numel(self):
            return self._n

        def element_size(self):
            return 4 if self.dtype == torch.float32 else 2

    js, kl = _lib.REG_IDS['js'], _lib.REG_IDS['kl']
    small, big = FakeZ(512, 64, 64, torch.float32), FakeZ(65536, 64, 64, torch.float32)
    assert head._step_pays(small, 64, 64, js, 2.0 / 64, None)            # single-launch form: always
    assert head._step_pays(small, 64, 64, kl, 2.0 / 64, None)            # ... KL included (it walks its window)
    wide = 20 * 2.0 / 64                                                  # JS window = whole image: generic kernel, three launches
    assert not head._step_pays(small, 64, 64, js, wide, None)            # 8 MiB: two kernels are one launch fewer
    assert head._step_pays(big, 64, 64, js, wide, None)                  # 1 GiB: the saved read pays
    odd = FakeZ(1024, 28, 28, torch.float32)
    assert not head._step_pays(odd, 28, 28, js, 2.0 / 28, None)


def test_sharded_path_decision_does_not_depend_on_the_local_shard():
    """the one-pass step exchanges twice per step (mask count, loss sums), the two-kernel path once; with uneven
    or empty shards a decision taken from the LOCAL shard size made the ranks issue different numbers of exchanges.
    Sharded, the decision is a function of dtype / shape / regulariser only."""
    from dsnt_pose2d_b200 import head
    from dsnt_pose2d_b200.parallel import shard_range
    sigma = 2.0 / 64
    # the advisor's example: 255 samples of 16x64x64 fp32 on 2 ranks = 128 | 127 samples, astride the 32 MiB threshold
    for batch, world in ((255, 2), (7, 8), (1, 2), (4099, 8), (0, 2)):
        for reg in ('js', 'kl', 'var', 'none', 'mse'):
            for dtype in (torch.float32, torch.bfloat16):
                picks = set()
                for r in range(world):
                    lo, hi = shard_range(batch, r, world)
                    z = torch.empty(hi - lo, 16, 64, 64, dtype=dtype)
                    picks.add(head.takes_one_pass(z, reg, sigma, sharded=True))
                assert len(picks) == 1, (batch, world, reg, dtype, picks)
    # single process: the size of the batch still matters where the single-launch kernel does not serve the case
    # (a JS window as wide as the image: 8 MiB of logits take two kernels, one launch fewer than three)
    assert not head.takes_one_pass(torch.empty(32, 16, 64, 64), 'js', 20 * sigma, sharded=False)
    assert head.takes_one_pass(torch.empty(32, 16, 64, 64), 'js', 20 * sigma, sharded=True)
    assert head.takes_one_pass(torch.empty(32, 16, 64, 64), 'kl', sigma, sharded=False)
    # the shape decides everywhere: 7x7 has no 16-byte vectors
    assert not head.takes_one_pass(torch.empty(4, 16, 7, 7), 'js', 2.0 / 7, sharded=True)


def test_step_arena_layout():
    """One allocation for the small outputs of a one-pass step: stats 16-byte aligned, coords / terms 8-byte aligned."""
    from dsnt_pose2d_b200.head import _StepArena
    for n in (0, 1, 3, 512, 65536):
        ar = _StepArena(n, torch.device('cpu'))
        assert ar.buf.numel() == 12 * n + 16
        assert ar.stats_ptr % 16 == 0 and ar.coords_ptr % 8 == 0 and ar.terms_ptr % 8 == 0 and ar.out8_ptr % 16 == 0
        assert ar.coords_ptr == ar.buf[8 * n:].data_ptr() and ar.terms_ptr == ar.buf[10 * n:].data_ptr()
        assert ar.out8_ptr == ar.out8().data_ptr() and ar.cnt8_ptr == ar.out8_ptr + 32
        assert ar.coords().numel() == 2 * n and ar.out8().numel() == 8


def test_bench_config_is_shared_by_both_arms():
    """The reference arm is timed 'on your arm's config': bench.py builds the config dict of both arms with one function."""
    import importlib
    bench = importlib.import_module('bench')
    for name in bench.WORKLOADS:
        for world, scaling in ((1, 'weak'), (8, 'weak'), (8, 'strong')):
            cfg = bench.workload_config(name, world, scaling)
            assert cfg['workload'] == name and 'model' not in cfg
            b = bench.WORKLOADS[name][0]
            assert cfg['global_batch'] == (b * world if scaling == 'weak' else b)
    assert bench.rotation(65536, 64, 64, 4) == 1 and bench.rotation(512, 64, 64, 4) > 8
    assert bench.DEFAULT_WORKLOAD == 'cfg4_64x64_f32_js'
