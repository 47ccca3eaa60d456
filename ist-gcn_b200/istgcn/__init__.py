"""istgcn -- Python side of the B200-native IST-GCN hot path.

``_lib``   ctypes binding of libistgcn_b200.so (the C ABI of include/istgcn_b200.h)
``sparse`` static non-zero lists of the partitioned adjacency
``ops``    torch.autograd Functions that drive the CUDA kernels
``dp``     one-process-per-GPU data-parallel trainer (NCCL gradient buckets)

The drop-in modules under ``net/`` (same import paths, constructors and state_dict layout as
the reference) are thin nn.Module shells over ``ops``.  There is no CPU fallback."""
import os

MATH_TF32 = 0
MATH_3XTF32 = 1
_MODES = {'tf32': MATH_TF32, '3xtf32': MATH_3XTF32, 'fp32': MATH_3XTF32}
_math = _MODES[os.environ.get('ISTGCN_MATH', 'tf32').lower()]


def set_math(mode):
    """'tf32' (one TF32 tensor-core pass, fast, <=2e-2 budget) or '3xtf32' (error-compensated,
    fp32-grade, <=1e-4 budget).  Returns the previous mode name."""
    global _math
    old = get_math()
    _math = _MODES[mode.lower()]
    return old


def get_math():
    return 'tf32' if _math == MATH_TF32 else '3xtf32'


def math_flag():
    return _math


# tensor-core (tcgen05/TMA/TMEM) engine of the graph convolution; ISTGCN_TC=0 selects the
# mma.sync engine for every layer (both are this library's own sm_100a kernels).  The '3xtf32'
# mode always uses the mma.sync engine (error-compensated split).
_use_tc = os.environ.get('ISTGCN_TC', '1') != '0'


def set_tensor_core_engine(flag):
    global _use_tc
    old, _use_tc = _use_tc, bool(flag)
    return old


def use_tc():
    return _use_tc and _math == MATH_TF32
