"""CPU tier: the C-ABI library loads, exports every symbol include/cupy_b200.h
declares, and its host-only entry points (planner, workspace queries, support
tables, NVRTC) behave.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from cupy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'cupy_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(b200_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(_lib.library_path())
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'include/cupy_b200.h declares %s but the library does not export it' % n


def test_python_binding_covers_header():
    assert set(_declared_symbols()) == set(_lib._EXPORTS)


def test_abi_version_and_itemsize():
    assert _lib.lib.b200_abi_version() == 1
    sizes = [1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8, 8, 16, 1]   # type_dispatcher.cuh:15-28 order
    for i, s in enumerate(sizes):
        assert _lib.lib.b200_dtype_itemsize(i) == s
    assert _lib.lib.b200_dtype_itemsize(99) == 0


def _operand(ptr, dtype_id, shape, strides, kind=_lib.KIND_ARRAY, out=False):
    o = _lib.Operand()
    o.data = ptr
    o.kind = kind
    o.dtype = dtype_id
    o.ndim = len(shape)
    o.is_output = int(out)
    for d, (s, t) in enumerate(zip(shape, strides)):
        o.shape[d] = s
        o.strides[d] = t
    return o


def _plan(ops):
    arr = (_lib.Operand * len(ops))(*ops)
    plan = _lib.EwPlan()
    st = _lib.lib.b200_ew_plan(len(ops), arr, ctypes.byref(plan))
    return st, plan


F32 = _lib.TYPE_FLOAT32


def test_plan_contiguous_collapses_to_flat_vec4():
    shape, st = (8, 16, 32), (2048, 128, 4)
    s, p = _plan([_operand(0x1000, F32, shape, st), _operand(0x2000, F32, shape, st),
                  _operand(0x3000, F32, shape, st, out=True)])
    assert s == 0
    assert (p.variant, p.ndim, p.vec, p.idx32, p.size) == (_lib.EW_FLAT, 1, 4, 1, 8 * 16 * 32)


def test_plan_misaligned_pointer_drops_vector_width():
    s, p = _plan([_operand(0x1004, F32, (1024,), (4,)), _operand(0x3000, F32, (1024,), (4,), out=True)])
    assert s == 0 and p.variant == _lib.EW_FLAT and p.vec == 1
    s, p = _plan([_operand(0x1008, F32, (1024,), (4,)), _operand(0x3000, F32, (1024,), (4,), out=True)])
    assert p.vec == 2


def test_plan_row_broadcast_is_flat_with_a_periodic_operand():
    # x[1024,256] + v[256]: v has stride 0 along dim 0 and its length divides 256 * vec -> the dense
    # operands are walked 1-D and v is indexed by (i % 256), constant per thread (FlatTiler)
    s, p = _plan([_operand(0x10000, F32, (1024, 256), (1024, 4)), _operand(0x90000, F32, (1024, 256), (0, 4)),
                  _operand(0xa0000, F32, (1024, 256), (1024, 4), out=True)])
    assert s == 0
    assert (p.variant, p.ndim, p.vec, p.size) == (_lib.EW_FLAT, 1, 4, 1024 * 256)
    assert p.staged_mask == 0b010 and p.tile_axis == 256
    # a row length that does not divide 1024 elements stays ROWWISE (vectors along the row)
    s, p = _plan([_operand(0x10000, F32, (1024, 768), (3072, 4)), _operand(0x90000, F32, (1024, 768), (0, 4)),
                  _operand(0xa00000, F32, (1024, 768), (3072, 4), out=True)])
    assert s == 0
    assert (p.variant, p.ndim, p.vec) == (_lib.EW_ROWWISE, 2, 4)
    assert list(p.shape[:2]) == [1024, 768]
    # a padded (non-dense) left operand too
    s, p = _plan([_operand(0x10000, F32, (1024, 256), (2048, 4)), _operand(0x90000, F32, (1024, 256), (0, 4)),
                  _operand(0xa00000, F32, (1024, 256), (1024, 4), out=True)])
    assert s == 0 and p.variant == _lib.EW_ROWWISE


def test_plan_column_broadcast():
    s, p = _plan([_operand(0x10000, F32, (1024, 256), (1024, 4)), _operand(0x90000, F32, (1024, 256), (4, 0)),
                  _operand(0xa0000, F32, (1024, 256), (1024, 4), out=True)])
    assert p.variant == _lib.EW_ROWWISE and p.vec == 4


def test_plan_transposed_input_is_tiled():
    # BASELINE config 4a: x strides (4, 4096, 4194304) shape (1024,1024,256), v broadcast, out contiguous
    shape = (1024, 1024, 256)
    s, p = _plan([_operand(0x100000, F32, shape, (4, 4096, 4194304)),
                  _operand(0x200000, F32, shape, (0, 0, 4)),
                  _operand(0x300000, F32, shape, (1048576, 1024, 4), out=True)])
    assert s == 0
    # equal item sizes, 16-byte aligned bases and strides -> the register-block tiler
    assert p.variant == _lib.EW_TILED_REG and p.tile_axis == 0 and p.staged_mask == 0b001 and p.ndim == 3
    assert p.vec == 4 and (p.reserved >> 24) == 4
    # a misaligned base (or mixed item sizes) keeps the plain shared-memory tile
    s, p = _plan([_operand(0x100004, F32, shape, (4, 4096, 4194304)),
                  _operand(0x200000, F32, shape, (0, 0, 4)),
                  _operand(0x300000, F32, shape, (1048576, 1024, 4), out=True)])
    assert s == 0 and p.variant == _lib.EW_TILED and p.staged_mask == 0b001
    s, p = _plan([_operand(0x100000, F32, shape, (4, 4096, 4194304)),
                  _operand(0x300000, _lib.TYPE_FLOAT64, shape, (2097152, 2048, 8), out=True)])
    assert s == 0 and p.variant == _lib.EW_TILED
    # innermost extent not a multiple of the vector width
    s, p = _plan([_operand(0x100000, F32, (64, 30), (4, 256)),
                  _operand(0x300000, F32, (64, 30), (120, 4), out=True)])
    assert s == 0 and p.variant == _lib.EW_TILED


def test_plan_2d_transpose_collapse_and_64bit():
    # 2-D transposed view: no collapse possible, tiled on axis 0
    s, p = _plan([_operand(0x1000, F32, (4096, 4096), (4, 16384)),
                  _operand(0x8000000, F32, (4096, 4096), (16384, 4), out=True)])
    assert p.variant == _lib.EW_TILED_REG and p.idx32 == 1
    # > 2^31 elements -> 64-bit indexing
    s, p = _plan([_operand(0x1000, _lib.TYPE_INT8, (1 << 32,), (1,)), _operand(0x1000, _lib.TYPE_INT8, (1 << 32,), (1,), out=True)])
    assert p.variant == _lib.EW_FLAT and p.idx32 == 0 and p.size == 1 << 32


def test_plan_drops_unit_dims_and_merges_partial():
    # (4,1,8,16) with a strided outer dim: dims (8,16) merge, dim 0 does not
    s, p = _plan([_operand(0x1000, F32, (4, 1, 8, 16), (1024, 512, 64, 4)),
                  _operand(0x9000, F32, (4, 1, 8, 16), (512, 512, 64, 4), out=True)])
    assert s == 0 and p.ndim == 2 and list(p.shape[:2]) == [4, 128]
    assert p.variant == _lib.EW_ROWWISE


def test_plan_errors():
    st, _ = _plan([_operand(0, F32, (), (), kind=_lib.KIND_SCALAR)])
    assert st == _lib.E_INVALID and b'undecided' in _lib.lib.b200_last_error_string()
    st, _ = _plan([_operand(0x1000, 77, (4,), (4,))])
    assert st == _lib.E_INVALID
    st, _ = _plan([_operand(0x1000, F32, (4,), (4,)), _operand(0x2000, F32, (5,), (4,), out=True)])
    assert st == _lib.E_INVALID


def test_ufunc_support_table():
    ids = (ctypes.c_int32 * 2)(F32, F32)
    assert _lib.lib.b200_ufunc_supported(_lib.UFUNC_IDS['add'], 2, ids, F32) == 1
    assert _lib.lib.b200_ufunc_supported(_lib.UFUNC_IDS['exp'], 1, ids, F32) == 1
    assert _lib.lib.b200_ufunc_supported(_lib.UFUNC_IDS['exp'], 1, (ctypes.c_int32 * 1)(_lib.TYPE_INT32), _lib.TYPE_INT32) == 0
    assert _lib.lib.b200_ufunc_supported(_lib.UFUNC_IDS['add'], 2, (ctypes.c_int32 * 2)(_lib.TYPE_INT8, _lib.TYPE_INT8), _lib.TYPE_INT8) == 0
    assert _lib.lib.b200_ufunc_supported(_lib.UFUNC_IDS['copy'], 1, (ctypes.c_int32 * 1)(_lib.TYPE_FLOAT16), F32) == 1


@pytest.mark.parametrize('op', [_lib.OP_SUM, _lib.OP_MAX, _lib.OP_ARGMAX, _lib.OP_MEAN, _lib.OP_VAR])
@pytest.mark.parametrize('layout', [_lib.RED_FULL, _lib.RED_ROWS, _lib.RED_COLS])
def test_reduce_workspace_query(op, layout):
    out = {_lib.OP_ARGMAX: _lib.TYPE_INT64}.get(op, F32)
    n_out = 1 if layout == _lib.RED_FULL else 32768
    d = _lib.ReduceDesc(op, layout, F32, out, 1, 32768, n_out, 0.0)
    assert _lib.lib.b200_reduce_supported(ctypes.byref(d)) == 1
    need = ctypes.c_size_t()
    assert _lib.lib.b200_reduce_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == 0
    if layout == _lib.RED_ROWS:
        assert need.value == 0
    if layout == _lib.RED_FULL:
        assert need.value >= 16384


def test_reduce_rejects_bad_descriptors():
    d = _lib.ReduceDesc(_lib.OP_SUM, _lib.RED_ROWS, F32, _lib.TYPE_FLOAT64, 1, 100, 100, 0.0)
    assert _lib.lib.b200_reduce_supported(ctypes.byref(d)) == 0       # result dtype not the loop's
    d = _lib.ReduceDesc(_lib.OP_SUM, _lib.RED_ROWS, F32, F32, 1, 0, 100, 0.0)
    need = ctypes.c_size_t()
    assert _lib.lib.b200_reduce_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == _lib.E_INVALID
    d = _lib.ReduceDesc(_lib.OP_CUMSUM, _lib.RED_FULL, F32, F32, 1, 10, 1, 0.0)
    assert _lib.lib.b200_reduce_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == _lib.E_INVALID
    d = _lib.ReduceDesc(_lib.OP_SUM, _lib.RED_ROWS, _lib.TYPE_UINT16, _lib.TYPE_UINT64, 1, 10, 10, 0.0)
    assert _lib.lib.b200_reduce_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == _lib.E_UNSUPPORTED


def test_scan_support_and_workspace():
    assert _lib.lib.b200_scan_supported(_lib.OP_CUMSUM, _lib.TYPE_INT64, _lib.TYPE_INT64) == 1
    assert _lib.lib.b200_scan_supported(_lib.OP_CUMSUM, _lib.TYPE_INT32, _lib.TYPE_INT64) == 1
    assert _lib.lib.b200_scan_supported(_lib.OP_CUMPROD, F32, F32) == 1
    # the pair a compaction ranks its flags with (cupy_b200/_core/_compaction.py)
    assert _lib.lib.b200_scan_supported(_lib.OP_CUMSUM, _lib.TYPE_BOOL, _lib.TYPE_INT32) == 1
    assert _lib.lib.b200_scan_supported(_lib.OP_SUM, F32, F32) == 0
    need = ctypes.c_size_t()
    assert _lib.lib.b200_scan_workspace_bytes(1 << 28, _lib.TYPE_INT64, ctypes.byref(need)) == 0
    tiles = (1 << 28) // 4096
    assert need.value >= tiles * (4 + 16)
    assert _lib.lib.b200_scan_workspace_bytes(-1, _lib.TYPE_INT64, ctypes.byref(need)) == _lib.E_INVALID


def test_nvrtc_compiles_for_sm100a_and_reports_errors():
    from cupy_b200._core import _jit
    src = '#include <b200/reduce.cuh>\nextern "C" __global__ void k(float* p) { p[0] = 1.f; }\n'
    cubin = _jit.compile_to_cubin(src, (), 'k.cu')
    assert cubin[:4] == b'\x7fELF'
    with pytest.raises(_lib.CompileException) as e:
        _jit.compile_to_cubin('extern "C" __global__ void k() { undefined_symbol(); }', (), 'bad.cu')
    assert 'undefined_symbol' in str(e.value)


def test_plan_ex_free_loop_order_collapses_f_ordered_operands():
    # F-ordered 2-D float32 -> float16 copy (x.T.astype): both operands unit-stride along dim 0
    shape = (4096, 2048)
    ops = [_operand(0x100000, F32, shape, (4, 16384)), _operand(0x4000000, _lib.TYPE_FLOAT16, shape, (2, 8192), out=True)]
    arr = (_lib.Operand * 2)(*ops)
    keep, free = _lib.EwPlan(), _lib.EwPlan()
    assert _lib.lib.b200_ew_plan_ex(2, arr, _lib.PLAN_KEEP_ORDER, ctypes.byref(keep)) == 0
    assert _lib.lib.b200_ew_plan_ex(2, arr, 0, ctypes.byref(free)) == 0
    assert keep.variant == _lib.EW_ROWWISE and keep.ndim == 2           # the linear index pins the order
    assert free.variant == _lib.EW_FLAT and free.ndim == 1 and free.vec == 4 and free.size == 4096 * 2048
    # b200_ew_plan is the order-keeping form
    s, p = _plan(ops)
    assert s == 0 and p.variant == keep.variant and p.ndim == keep.ndim
    # C-ordered inputs into an F-ordered output: sorted by the OUTPUT's strides -> the inputs are the transposed ones
    ops = [_operand(0x100000, F32, shape, (8192, 4)), _operand(0x4000000, F32, shape, (4, 16384), out=True)]
    arr = (_lib.Operand * 2)(*ops)
    assert _lib.lib.b200_ew_plan_ex(2, arr, 0, ctypes.byref(free)) == 0
    assert free.variant == _lib.EW_TILED_REG and free.staged_mask == 0b01
    assert tuple(free.shape[:2]) == (2048, 4096)
    # permuted 3-D: (2, 0, 1) transpose of a C array, same permutation on both sides -> flat again
    shape3, st3 = (64, 16, 32), (4, 8192, 256)
    ops = [_operand(0x1000, F32, shape3, st3), _operand(0x800000, F32, shape3, st3, out=True)]
    arr = (_lib.Operand * 2)(*ops)
    assert _lib.lib.b200_ew_plan_ex(2, arr, 0, ctypes.byref(free)) == 0
    assert free.variant == _lib.EW_FLAT and free.ndim == 1
