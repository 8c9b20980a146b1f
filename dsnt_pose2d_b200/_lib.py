"""ctypes binding of libdsnt_b200.so (C ABI declared in include/dsnt_b200.h).

There is deliberately NO fallback: if the shared library is missing the import fails loudly, and
every operator refuses non-CUDA tensors (north_star: "no CPU fallback, no multi-backend dispatch").
"""

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSNT_B200_LIB: another build of the same library (developer measurements: tools/probe builds one with the phase trace)
LIB_PATH = os.environ.get('DSNT_B200_LIB') or os.path.join(_HERE, 'libdsnt_b200.so')

DTYPE_F32, DTYPE_BF16 = 0, 1
REG_IDS = {'none': 0, 'var': 1, 'kl': 2, 'js': 3, 'mse': 4}
PREACT_IDS = {'softmax': 0, 'thresholded_softmax': 1, 'abs': 2, 'relu': 3, 'sigmoid': 4}
STATS_K = 8
FLAG_STRICT_NAN = 1
FLAG_NO_EUCLID = 2

_c_long, _c_int, _c_float, _c_ptr = ctypes.c_long, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/dsnt_b200.h one to one
SIGNATURES = {
    'dsnt_b200_version': (_c_int, []),
    'dsnt_b200_last_error': (ctypes.c_char_p, []),
    'dsnt_head_fwd': (_c_int, [_c_ptr, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_int, _c_float,
                               _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]),
    'dsnt_head_bwd': (_c_int, [_c_ptr, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr,
                               _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int, _c_float, _c_int,
                               _c_ptr, _c_int, _c_ptr]),
    'dsnt_head_fwd_stacked': (_c_int, [_c_ptr, _c_int, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_int, _c_float,
                                       _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]),
    'dsnt_head_bwd_stacked': (_c_int, [_c_ptr, _c_ptr, _c_int, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr,
                                       _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int, _c_float, _c_int,
                                       _c_int, _c_ptr]),
    'dsnt_head_preact_fwd': (_c_int, [_c_ptr, _c_int, _c_int, _c_float, _c_float, _c_long, _c_int, _c_int, _c_ptr, _c_int,
                                      _c_float, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]),
    'dsnt_head_preact_bwd': (_c_int, [_c_ptr, _c_int, _c_int, _c_float, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr,
                                      _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int, _c_float, _c_int, _c_ptr,
                                      _c_int, _c_ptr]),
    'dsnt_flip_tta_fwd': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, _c_int, _c_ptr, _c_int, _c_float, _c_float,
                                   _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_draw_gaussians': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, ctypes.c_double, ctypes.c_double, _c_int,
                                     _c_ptr, _c_ptr]),
    'dsnt_decode_heatmaps': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, _c_int, _c_ptr, _c_ptr]),
    'dsnt_head_step_supported': (_c_int, [_c_int, _c_int, _c_int]),
    'dsnt_head_step_supported_reg': (_c_int, [_c_int, _c_int, _c_int, _c_int]),
    'dsnt_head_step_pair_supported': (_c_int, [_c_int, _c_int, _c_int, _c_int]),
    'dsnt_head_step': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int,
                                _c_float, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_mask_count': (_c_int, [_c_ptr, _c_long, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_head_step_fused_stacked': (_c_int, [_c_ptr, _c_ptr, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr,
                                              _c_float, _c_int, _c_float, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_head_step_fused_supported': (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_float]),
    'dsnt_head_step_fused': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int,
                                      _c_float, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_scale_unless_one': (_c_int, [_c_ptr, _c_int, _c_long, _c_ptr, _c_ptr]),
    'dsnt_finish_loss_stacked': (_c_int, [_c_ptr, _c_ptr, _c_long, _c_int, _c_float, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_finish_workspace_bytes': (_c_int, []),
    'dsnt_finish_trace_offset_bytes': (_c_int, []),
    'dsnt_finish_loss': (_c_int, [_c_ptr, _c_ptr, _c_long, _c_float, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_combine_loss': (_c_int, [_c_ptr, _c_float, _c_ptr]),
    'dsnt_peer_exchange_bytes': (_c_int, []),
    'dsnt_finish_loss_peer': (_c_int, [_c_ptr, _c_ptr, _c_long, _c_int, _c_float, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int,
                                       _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_head_step_fused_peer': (_c_int, [_c_ptr, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_float, _c_int,
                                           _c_float, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int,
                                           _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_mask_count_peer': (_c_int, [_c_ptr, _c_long, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr]),
    'dsnt_euclid_fwd': (_c_int, [_c_ptr, _c_ptr, _c_long, _c_int, _c_ptr, _c_ptr]),
    'dsnt_euclid_bwd': (_c_int, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_long, _c_int, _c_int, _c_ptr,
                                 _c_ptr]),
    'dsnt_tsoftmax_fwd': (_c_int, [_c_ptr, _c_int, _c_long, _c_long, _c_float, _c_float, _c_ptr, _c_ptr]),
    'dsnt_tsoftmax_bwd': (_c_int, [_c_ptr, _c_ptr, _c_int, _c_long, _c_long, _c_ptr, _c_ptr]),
    'dsnt_make_gauss_fwd': (_c_int, [_c_ptr, _c_long, _c_int, _c_int, _c_float, _c_ptr, _c_ptr]),
    'dsnt_make_gauss_bwd': (_c_int, [_c_ptr, _c_ptr, _c_long, _c_int, _c_int, _c_float, _c_ptr, _c_ptr]),
    'dsnt_reg_dmu': (_c_int, [_c_ptr, _c_int, _c_int, _c_long, _c_int, _c_int, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_float,
                              _c_int, _c_float, _c_ptr, _c_ptr]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'dsnt_pose2d_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C dsnt_pose2d_b200/csrc`. There is no CPU / PyTorch fallback for this path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


LIB = _load()
launch_count = 0          # kernels enqueued through this binding (bench.py reports it as gpu_launches)
event_log = None          # bench.py sets this to {entry_point_name: []} to get CUDA-event pairs per launch


def version():
    return LIB.dsnt_b200_version()


def last_error():
    return LIB.dsnt_b200_last_error().decode('utf-8', 'replace')


_FN = {name: getattr(LIB, name) for name in SIGNATURES}        # bound once: no attribute lookup per launch


def call(name, *args, launches=1):
    """Invoke an entry point; a non-zero return raises RuntimeError with the library's message."""
    global launch_count
    if event_log is not None:
        log = event_log.get(name)
        if log is not None:            # time this launch on the stream it is enqueued on (torch's current stream)
            start = torch.cuda.Event(enable_timing=True)
            stop = torch.cuda.Event(enable_timing=True)
            start.record()
            rc = _FN[name](*args)
            stop.record()
            log.append((start, stop))
            if rc != 0:
                raise RuntimeError('%s failed (%d): %s' % (name, rc, last_error()))
            launch_count += launches
            return
    rc = _FN[name](*args)
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, last_error()))
    launch_count += launches


def ptr(t):
    return None if t is None else t.data_ptr()


MAX_STACKS = 16


def ptr_array(tensors):
    """Host array of device pointers for the *_stacked entry points."""
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_of(t):
    """The raw cudaStream_t of torch's current stream on the tensor's device."""
    if _raw_stream is not None:
        idx = t.device.index
        return _raw_stream(torch.cuda.current_device() if idx is None else idx)
    return torch.cuda.current_stream(t.device).cuda_stream


def dtype_id(t):
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    raise NotImplementedError(
        'dsnt_pose2d_b200 supports float32 and bfloat16 heatmaps only (got %s); float16 already yields NaN in '
        'the reference because 1e-24 underflows, float64 has no kernel' % (t.dtype,))


def require_cuda(t, what):
    if not torch.is_tensor(t):
        raise TypeError('%s must be a tensor' % what)
    if not t.is_cuda:
        raise NotImplementedError('%s must be a CUDA tensor: dsnt_pose2d_b200 has no CPU fallback' % what)


_workspaces = {}


def finish_workspace(device, stream=None):
    """Per-(device, stream) zero-initialised scratch for dsnt_finish_loss (the kernel re-zeroes its ticket)."""
    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(LIB.dsnt_finish_workspace_bytes() // 4, dtype=torch.float32, device=device)
        _workspaces[key] = ws
    return ws


class on_device:
    """`with on_device(dev):` = torch.cuda.device(dev) that costs nothing when dev is already current (the usual case)."""

    __slots__ = ('guard',)

    def __init__(self, dev):
        idx = dev.index
        self.guard = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False
