"""ctypes binding of the C ABI in include/istgcn_b200.h (libistgcn_b200.so).

There is NO fallback: if the library is missing (run ``python __graft_entry__.py`` /
``__graft_entry__.build()`` first) or the device is not an sm_100 GPU, calls raise."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libistgcn_b200.so')

_lib = None
_checked_device = False

# every symbol include/istgcn_b200.h declares (tests/test_cabi.py checks the export list)
SYMBOLS = (
    'istgcn_last_error', 'istgcn_version', 'istgcn_check_device',
    'istgcn_data_bn_stats', 'istgcn_data_bn_apply', 'istgcn_data_bn_bwd',
    'istgcn_bn_finalize', 'istgcn_bn_eval_coeffs', 'istgcn_bn_bwd_coeffs',
    'istgcn_gcn_fwd', 'istgcn_gcn_bwd_x', 'istgcn_gcn_bwd_w', 'istgcn_gcn_tc', 'istgcn_gcn_tc_dvals', 'istgcn_gcn_tc_dw', 'istgcn_gcn_pair_grads', 'istgcn_tcn2_down_bn', 'istgcn_tcn2_bwd_up_bn', 'istgcn_bn_back_colsum_bn',
    'istgcn_block_tail_fwd_bn', 'istgcn_joint_colsum', 'istgcn_gcn_small_bwd_post',
    'istgcn_gcn_small_fwd', 'istgcn_gcn_small_bwd',
    'istgcn_tcn_fwd', 'istgcn_tcn_bwd',
    'istgcn_tcn2_down', 'istgcn_tcn2_conv', 'istgcn_tcn2_up', 'istgcn_tcn2_bwd_up', 'istgcn_tcn2_bwd_conv',
    'istgcn_tcn2_bwd_down',
    'istgcn_bn_relu_apply', 'istgcn_bn_back_apply', 'istgcn_bn_back_colsum', 'istgcn_relu_bn_bwd', 'istgcn_tconv_tc',
    'istgcn_tconv_dw_tc',
    'istgcn_block_tail_fwd', 'istgcn_block_tail_bwd', 'istgcn_dropout_mask',
    'istgcn_pool_fwd', 'istgcn_pool_bwd',
    'istgcn_sgd_step', 'istgcn_feeder_augment', 'istgcn_block_prep_fwd', 'istgcn_block_prep_bwd',
)


def load():
    """Load the shared library (once).  Raises with build instructions when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                'istgcn_b200: %s not found - the CUDA library has not been built. Run '
                '`python -c "import __graft_entry__ as g; g.build()"` at the repo root. '
                'There is no CPU or PyTorch fallback.' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.istgcn_last_error.restype = ctypes.c_char_p
        for name in SYMBOLS:
            getattr(lib, name)                  # AttributeError if the export list drifts
        _lib = lib
    return _lib


class _Ptr(object):
    """Marks an argument as a device pointer (tensor or None)."""
    __slots__ = ()


def _conv(a):
    import torch
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError('istgcn_b200: CPU tensor passed to a CUDA kernel (no CPU fallback)')
        if not a.is_contiguous():
            raise RuntimeError('istgcn_b200: non-contiguous tensor passed to the C ABI')
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    return a                                     # already a ctypes value


# kernels launched per C-ABI call (tcn_fwd = down + up, tcn_bwd = up + temporal + down,
# pool_fwd = kernel behind a memset) -- used for the ``gpu_launches`` count of bench.py
KERNELS_PER_CALL = {'tcn_fwd': 2, 'tcn_bwd': 3, 'gcn_tc_dw': 2, 'tconv_dw_tc': 2, 'tcn2_bwd_conv': 2, 'gcn_pair_grads': 3}
launch_count = 0          # kernels launched through this binding since import
timing = None             # set to a dict by bench.py: name -> list of (start_event, end_event)


def call(name, *args):
    """Invoke ``istgcn_<name>`` on torch's current CUDA stream; raise on a non-zero status."""
    import torch
    global _checked_device, launch_count
    lib = load()
    if not _checked_device:
        if not torch.cuda.is_available():
            raise RuntimeError('istgcn_b200 needs a CUDA device (sm_100a); none is visible and '
                               'there is no CPU fallback')
        rc = lib.istgcn_check_device()
        if rc != 0:
            raise RuntimeError('istgcn_b200: ' + lib.istgcn_last_error().decode())
        _checked_device = True
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    launch_count += KERNELS_PER_CALL.get(name, 1)
    if name == 'gcn_tc_dw' and args[8] is None:        # no bias-term column sums: one kernel
        launch_count -= 1
    if name == 'tcn2_bwd_conv' and (args[3] is None or args[4] is None):   # one of its two kernels
        launch_count -= 1
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, 'istgcn_' + name)(*[_conv(a) for a in args], stream)
        e1.record()
        timing.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, 'istgcn_' + name)(*[_conv(a) for a in args], stream)
    if rc != 0:
        raise RuntimeError('istgcn_%s failed (%d): %s' % (name, rc,
                                                          lib.istgcn_last_error().decode()))


def i64(v):
    return ctypes.c_longlong(int(v))


def u64(v):
    return ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)


def f64(v):
    return ctypes.c_double(float(v))
