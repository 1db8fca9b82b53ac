"""TEST INFRASTRUCTURE -- never imported by the product (cupy_b200/).

GPU-side harness for the REFERENCE's own kernels, used by bench.py's `reference_gpu` leg and by the
GPU parity tests: it launches
  * the cubins rendered from the reference's JIT templates by oracle/render_ref_jit.py
    (oracle/_ref/jit/*.cubin + manifest.json: elementwise, generic reduction, CUB-block reduction) with the
    reference's launch geometry
    (`linear_launch` cupy/cuda/function.pyx:153-171, `_get_block_specs` cupy/_core/_reduction.pyx:239-253,
    `_launch` :481-508), through libcuda (cuModuleLoadData / cuLaunchKernel, what
    cupy_backends/cuda/api/driver.pyx:273-286 calls), and
  * the reference's pre-compiled CUB path (oracle/_ref/libcupy_cub_ref.so = cupy/cuda/cupy_cub.cu)
    the way cupy/cuda/cub.pyx:137-306 drives it (workspace query, then run).
Everything takes raw device pointers (ints); the caller owns the memory.  Needs no reference tree.
"""
from __future__ import annotations

import ctypes
import json
import os
from ctypes import POINTER, byref, c_int, c_int64, c_size_t, c_uint, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
JIT_DIR = os.path.join(HERE, '_ref', 'jit')
CUB_SO = os.path.join(HERE, '_ref', 'libcupy_cub_ref.so')

# op codes cupy/cuda/cupy_cub.h:4-11, dtype ids cupy/_core/include/cupy/type_dispatcher.cuh:15-28
CUB_SUM, CUB_MIN, CUB_MAX, CUB_ARGMIN, CUB_ARGMAX, CUB_CUMSUM, CUB_CUMPROD, CUB_PROD = range(8)
DTYPE_ID = {'int8': 0, 'uint8': 1, 'int16': 2, 'uint16': 3, 'int32': 4, 'uint32': 5, 'int64': 6, 'uint64': 7,
            'float16': 8, 'float32': 9, 'float64': 10, 'bool': 13}


def available():
    return os.path.exists(os.path.join(JIT_DIR, 'manifest.json')) and os.path.exists(CUB_SO)


def _carray_struct(ndim):
    class CArray(ctypes.Structure):      # cupy/_core/include/cupy/carray.cuh:237-240, filled as _carray.pyx:90-120
        _fields_ = [('data', c_void_p), ('size', c_int64), ('shape', c_int64 * ndim), ('strides', c_int64 * ndim)]
    return CArray


def _cindexer_struct(ndim):
    class CIndexer(ctypes.Structure):    # carray.cuh:523-525, filled as _carray.pyx:131-150
        _fields_ = [('size', c_int64), ('shape', c_int64 * ndim), ('index', c_int64 * ndim)]
    return CIndexer


def carray(ptr, shape, strides):
    n = 1
    for s in shape:
        n *= s
    st = _carray_struct(len(shape))()
    st.data, st.size = ptr, n
    for i, (s, b) in enumerate(zip(shape, strides)):
        st.shape[i], st.strides[i] = s, b
    return st


def cindexer(shape):
    n = 1
    for s in shape:
        n *= s
    st = _cindexer_struct(len(shape))()
    st.size = n
    for i, s in enumerate(shape):
        st.shape[i] = s
    return st


class RefJit:
    """The rendered reference JIT kernels, loaded into the current CUDA context."""

    def __init__(self):
        with open(os.path.join(JIT_DIR, 'manifest.json')) as f:
            self.manifest = json.load(f)
        self.cu = ctypes.CDLL('libcuda.so.1')
        self.cu.cuModuleLoadData.argtypes = [POINTER(c_void_p), c_void_p]
        self.cu.cuModuleGetFunction.argtypes = [POINTER(c_void_p), c_void_p, ctypes.c_char_p]
        self.cu.cuLaunchKernel.argtypes = [c_void_p] + [c_uint] * 6 + [c_uint, c_void_p, POINTER(c_void_p), c_void_p]
        self._fn = {}

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError('%s failed: CUresult %d' % (what, rc))

    def function(self, name):
        fn = self._fn.get(name)
        if fn is None:
            with open(os.path.join(JIT_DIR, name + '.cubin'), 'rb') as f:
                image = f.read()
            mod, fn = c_void_p(), c_void_p()
            self._check(self.cu.cuModuleLoadData(byref(mod), image), 'cuModuleLoadData(%s)' % name)
            self._check(self.cu.cuModuleGetFunction(byref(fn), mod, name.encode()), 'cuModuleGetFunction')
            self._fn[name] = fn
        return fn

    def prepare(self, name, args, grid, block, stream=0):
        """Returns a zero-argument callable that launches `name` (args = ctypes values in parameter order)."""
        fn = self.function(name)
        spec = self.manifest[name]['params']
        assert len(spec) == len(args), (name, len(spec), len(args))
        keep = list(args)
        ptrs = (c_void_p * len(keep))(*[ctypes.cast(byref(a), c_void_p) for a in keep])
        cu, stream_p = self.cu, c_void_p(stream)

        def launch():
            rc = cu.cuLaunchKernel(fn, grid, 1, 1, block, 1, 1, 0, stream_p, ptrs, None)
            if rc:
                raise RuntimeError('cuLaunchKernel(%s) failed: CUresult %d' % (name, rc))
        launch.keep = keep
        return launch

    # ---- the reference's launch geometries -------------------------------------------------------
    def elementwise(self, name, args, size, stream=0, block=128):
        """Function.linear_launch (cupy/cuda/function.pyx:153-171): one element per thread per grid-stride step."""
        grid = min(0x7fffffff, (size + block - 1) // block)
        return self.prepare(name, args, grid, min(block, size), stream)

    def reduction(self, name, in_arrays, scalars_mid, out_array, in_shape, out_shape, contiguous_size, stream=0):
        """_AbstractReductionKernel._launch (cupy/_core/_reduction.pyx:481-508) with _get_block_specs geometry."""
        from oracle.render_ref_jit import block_specs
        in_size, out_size = 1, 1
        for s in in_shape:
            in_size *= s
        for s in out_shape:
            out_size *= s
        block_size, block_stride, out_block_num = block_specs(in_size, out_size, contiguous_size)
        args = list(in_arrays) + list(scalars_mid) + [out_array, cindexer(in_shape), cindexer(out_shape),
                                                      ctypes.c_int32(block_stride)]
        launch = self.prepare(name, args, out_block_num, block_size, stream)
        launch.geometry = {'block_size': block_size, 'block_stride': block_stride, 'grid': out_block_num}
        return launch


    def cub_block(self, name, in_ptr, out_ptr, n_segments, segment_size, stream=0):
        """_launch_cub, one-pass branch (cupy/_core/_cub_reduction.pyx:545-563): one block per segment,
        linear_launch(out_block_num * block_size, ..., block_size)."""
        block = self.manifest[name]['block_size']
        args = [c_void_p(in_ptr), c_void_p(out_ptr), ctypes.c_int32(segment_size), ctypes.c_int32(0)]
        return self.prepare(name, args, n_segments, block, stream)


class RefCub:
    """cupy/cuda/cub.pyx's way of driving cupy_cub.cu: query the workspace size, allocate, run."""

    def __init__(self):
        lib = ctypes.CDLL(CUB_SO)
        for nm, extra in (('reduce', [c_int]), ('segmented_reduce', [c_int, c_int]), ('scan', [c_int])):
            ws = getattr(lib, 'ref_cub_%s_workspace' % nm)
            ws.restype = c_size_t
            ws.argtypes = [c_void_p, c_void_p] + extra + [c_void_p, c_int, c_int]
            run = getattr(lib, 'ref_cub_' + nm)
            run.restype = None
            run.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p] + extra + [c_void_p, c_int, c_int]
        self.lib = lib

    def reduce(self, x, y, n, op, dtype, alloc, stream=0):
        ws_bytes = self.lib.ref_cub_reduce_workspace(x, y, n, stream, op, DTYPE_ID[dtype])
        ws = alloc(max(ws_bytes, 1))
        return lambda: self.lib.ref_cub_reduce(ws, ws_bytes, x, y, n, stream, op, DTYPE_ID[dtype])

    def segmented_reduce(self, x, y, n_seg, seg_size, op, dtype, alloc, stream=0):
        ws_bytes = self.lib.ref_cub_segmented_reduce_workspace(x, y, n_seg, seg_size, stream, op, DTYPE_ID[dtype])
        ws = alloc(max(ws_bytes, 1))
        return lambda: self.lib.ref_cub_segmented_reduce(ws, ws_bytes, x, y, n_seg, seg_size, stream, op,
                                                         DTYPE_ID[dtype])

    def scan(self, x, y, n, op, dtype, alloc, stream=0):
        ws_bytes = self.lib.ref_cub_scan_workspace(x, y, n, stream, op, DTYPE_ID[dtype])
        ws = alloc(max(ws_bytes, 1))
        return lambda: self.lib.ref_cub_scan(ws, ws_bytes, x, y, n, stream, op, DTYPE_ID[dtype])
