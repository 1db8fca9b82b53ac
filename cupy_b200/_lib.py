"""ctypes binding of libcupy_b200.so (the C ABI declared in include/cupy_b200.h).

The library is the product: every compute entry point of the package goes
through it, and importing fails loudly when it has not been built (there is no
CPU / NumPy fallback anywhere in the package).

Reference counterpart: the Cython `cdef extern` blocks that bind the reference's
native code -- cupy/cuda/cub.pyx:30-63 (cupy_cub.h), cupy_backends/cuda/api/
driver.pyx (cuLaunchKernel, cuModuleLoadData), cupy_backends/cuda/libs/nvrtc.pyx.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int,
                    c_int32, c_int64, c_size_t, c_uint, c_uint32, c_void_p)

MAX_NDIM = 10
MAX_ARGS = 12

# dtype ids (== cupy/_core/include/cupy/type_dispatcher.cuh:15-28)
TYPE_INT8, TYPE_UINT8, TYPE_INT16, TYPE_UINT16, TYPE_INT32, TYPE_UINT32, \
    TYPE_INT64, TYPE_UINT64, TYPE_FLOAT16, TYPE_FLOAT32, TYPE_FLOAT64, \
    TYPE_COMPLEX64, TYPE_COMPLEX128, TYPE_BOOL = range(14)

# op codes (== cupy/cuda/cupy_cub.h:4-11, then extensions)
OP_SUM, OP_MIN, OP_MAX, OP_ARGMIN, OP_ARGMAX, OP_CUMSUM, OP_CUMPROD, OP_PROD, \
    OP_MEAN, OP_VAR, OP_MOMENTS = range(11)

OK, E_INVALID, E_UNSUPPORTED, E_WORKSPACE, E_NOLIB, E_COMPILE = 0, -1, -2, -3, -4, -5

KIND_ARRAY, KIND_SCALAR, KIND_RAW = 0, 1, 2
PLAN_KEEP_ORDER = 1
EW_FLAT, EW_ROWWISE, EW_TILED, EW_TILED_TMA, EW_TILED_REG = 0, 1, 2, 3, 4
RED_FULL, RED_ROWS, RED_COLS = 0, 1, 2

UFUNC_IDS = {name: i for i, name in enumerate((
    'copy', 'add', 'subtract', 'multiply', 'true_divide', 'negative',
    'absolute', 'square', 'sqrt', 'exp', 'log', 'maximum', 'minimum', 'fma'))}


class Operand(Structure):
    _fields_ = [('data', c_void_p),
                ('scalar', c_int64 * 2),
                ('kind', c_int32),
                ('dtype', c_int32),
                ('ndim', c_int32),
                ('is_output', c_int32),
                ('shape', c_int64 * MAX_NDIM),
                ('strides', c_int64 * MAX_NDIM)]


class EwPlan(Structure):
    _fields_ = [('variant', c_int32),
                ('ndim', c_int32),
                ('vec', c_int32),
                ('idx32', c_int32),
                ('tile_axis', c_int32),
                ('nargs', c_int32),
                ('staged_mask', c_uint32),
                ('reserved', c_uint32),
                ('size', c_int64),
                ('shape', c_int64 * MAX_NDIM),
                ('strides', (c_int64 * MAX_NDIM) * MAX_ARGS)]


class ReduceDesc(Structure):
    _fields_ = [('op', c_int32),
                ('layout', c_int32),
                ('in_dtype', c_int32),
                ('out_dtype', c_int32),
                ('batch', c_int64),
                ('n_reduce', c_int64),
                ('n_out', c_int64),
                ('param', c_double)]


MAX_PEERS = 16
EXCHANGE_BYTES = 2048


class PeerExchange(Structure):
    _fields_ = [('rank', c_int32), ('nranks', c_int32), ('tag', c_uint32), ('reserved', c_uint32),
                ('n_total', c_int64), ('slots', c_void_p * MAX_PEERS)]


class B200Error(RuntimeError):
    """A failing C-ABI call (status != 0)."""

    def __init__(self, status, message):
        super().__init__('[cupy_b200 status %d] %s' % (status, message))
        self.status = status


class UnsupportedError(B200Error):
    """B200_E_UNSUPPORTED: no prebuilt kernel; the host takes the JIT route."""


class CompileException(B200Error):
    """NVRTC rejected a generated / user kernel (cf. cupy.cuda.compiler.CompileException)."""

    def __init__(self, status, message, log='', source=''):
        super().__init__(status, message + '\n' + log)
        self.log = log
        self.source = source


_LIB_NAME = 'libcupy_b200.so'
_EXPORTS = {
    # name: (restype, argtypes)
    'b200_abi_version': (c_int, []),
    'b200_last_error_string': (c_char_p, []),
    'b200_device_info': (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    'b200_dtype_itemsize': (c_int, [c_int]),
    'b200_workspace_init': (c_int, [c_void_p, c_size_t, c_void_p]),
    'b200_ew_plan': (c_int, [c_int, POINTER(Operand), POINTER(EwPlan)]),
    'b200_ew_plan_ex': (c_int, [c_int, POINTER(Operand), c_uint32, POINTER(EwPlan)]),
    'b200_ufunc_supported': (c_int, [c_int, c_int, POINTER(c_int32), c_int32]),
    'b200_ufunc_launch': (c_int, [c_int, POINTER(EwPlan), c_int, POINTER(Operand), c_void_p]),
    'b200_reduce_supported': (c_int, [POINTER(ReduceDesc)]),
    'b200_reduce_workspace_bytes': (c_int, [POINTER(ReduceDesc), POINTER(c_size_t)]),
    'b200_reduce_run': (c_int, [POINTER(ReduceDesc), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'b200_reduce_run_sharded': (c_int, [POINTER(ReduceDesc), c_void_p, c_void_p, c_void_p, c_size_t, POINTER(PeerExchange),
                                        c_void_p]),
    'b200_moments_merge': (c_int, [c_void_p, c_int, c_double, c_void_p, c_void_p]),
    'b200_scan_supported': (c_int, [c_int, c_int, c_int]),
    'b200_scan_workspace_bytes': (c_int, [c_int64, c_int, POINTER(c_size_t)]),
    'b200_scan_run': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    'b200_scan_axis_workspace_bytes': (c_int, [c_int64, c_int64, c_int64, POINTER(c_size_t)]),
    'b200_scan_axis_run': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_size_t,
                                   c_void_p]),
    'b200_jit_compile': (c_int, [c_char_p, c_char_p, c_int, POINTER(c_char_p), POINTER(c_void_p), POINTER(c_size_t)]),
    'b200_jit_last_log': (c_char_p, []),
    'b200_jit_free_image': (None, [c_void_p]),
    'b200_module_load': (c_int, [c_void_p, POINTER(c_void_p)]),
    'b200_module_unload': (c_int, [c_void_p]),
    'b200_module_get_function': (c_int, [c_void_p, c_char_p, POINTER(c_void_p)]),
    'b200_jit_ew_launch': (c_int, [c_void_p, POINTER(EwPlan), c_int, POINTER(Operand), c_int, c_void_p]),
    'b200_jit_ew_launch_ex': (c_int, [c_void_p, POINTER(EwPlan), c_int, POINTER(Operand), c_int, c_int, POINTER(c_int64),
                                     c_void_p]),
    'b200_jit_launch': (c_int, [c_void_p, c_uint, c_uint, c_uint, c_uint, c_uint, c_void_p, c_size_t, c_void_p]),
}


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def _load():
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            '%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(or `make -C cupy_b200/csrc`). cupy_b200 has no CPU fallback.' % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.b200_abi_version() != 1:
        raise ImportError('libcupy_b200.so ABI version mismatch')
    return lib


lib = _load()


def last_error():
    return (lib.b200_last_error_string() or b'').decode('utf-8', 'replace')


def check(status):
    if status == 0:
        return
    msg = last_error()
    if status == E_UNSUPPORTED:
        raise UnsupportedError(status, msg)
    raise B200Error(status, msg)


_device_info_cache = {}


def device_info():
    """(sm_count, cc_major, cc_minor, l2_bytes) of the current device."""
    import torch
    dev = torch.cuda.current_device()
    info = _device_info_cache.get(dev)
    if info is None:
        sm, ma, mi, l2 = c_int(), c_int(), c_int(), c_size_t()
        check(lib.b200_device_info(byref(sm), byref(ma), byref(mi), byref(l2)))
        info = _device_info_cache[dev] = (sm.value, ma.value, mi.value, l2.value)
    return info
