"""GPU tier: the C ABI called directly (ctypes, torch buffers only as device memory) -- the
calls a reference maintainer's binding would make.  Mirrors how the
reference drives cupy_cub.h: query workspace, allocate, run (cupy/cuda/cub.pyx:137-306)."""
import ctypes

import numpy as np
import pytest

from cupy_b200 import _lib
from oracle import oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

IDS = {torch.float32: _lib.TYPE_FLOAT32, torch.float16: _lib.TYPE_FLOAT16, torch.float64: _lib.TYPE_FLOAT64,
       torch.int32: _lib.TYPE_INT32, torch.int64: _lib.TYPE_INT64, torch.int8: _lib.TYPE_INT8}


def stream():
    return torch.cuda.current_stream().cuda_stream


def reduce_abi(x, op, layout, batch, n_reduce, n_out, out_dtype, param=0.0):
    d = _lib.ReduceDesc(op, layout, IDS[x.dtype], IDS[out_dtype], batch, n_reduce, n_out, param)
    need = ctypes.c_size_t()
    assert _lib.lib.b200_reduce_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == 0
    ws = torch.zeros(max(need.value, 16), dtype=torch.uint8, device='cuda')
    y = torch.empty(batch * n_out, dtype=out_dtype, device='cuda')
    for _ in range(2):                 # twice: the kernels must leave the ticket header zeroed
        y.fill_(-7)
        st = _lib.lib.b200_reduce_run(ctypes.byref(d), x.data_ptr(), y.data_ptr(), ws.data_ptr(), ws.numel(), stream())
        assert st == 0, _lib.last_error()
    assert int(ws[:16384].to(torch.int64).sum().item()) == 0 if need.value else True
    return y.cpu().numpy()


@pytest.mark.parametrize('tdt', [torch.float32, torch.float16, torch.int32])
def test_reduce_layouts_through_the_abi(tdt):
    g = torch.Generator(device='cuda').manual_seed(3)
    if tdt.is_floating_point:
        x = (torch.rand(6, 700, 260, device='cuda', generator=g) * 2 - 1).to(tdt)
    else:
        x = torch.randint(-50, 50, (6, 700, 260), device='cuda', generator=g, dtype=tdt)
    h = x.cpu().numpy()
    sum_dt = {torch.float32: torch.float32, torch.float16: torch.float16, torch.int32: torch.int64}[tdt]
    rt = 3e-3 if tdt == torch.float16 else 1e-5
    got = reduce_abi(x, _lib.OP_SUM, _lib.RED_FULL, 1, x.numel(), 1, sum_dt)
    np.testing.assert_allclose(got[0].astype(np.float64), oracle.sum(h).astype(np.float64), rtol=rt, atol=1e-2)
    got = reduce_abi(x, _lib.OP_SUM, _lib.RED_ROWS, 1, 260, 6 * 700, sum_dt).reshape(6, 700)
    np.testing.assert_allclose(got.astype(np.float64), oracle.sum(h, axis=2).astype(np.float64), rtol=rt, atol=1e-2)
    got = reduce_abi(x, _lib.OP_SUM, _lib.RED_COLS, 6, 700, 260, sum_dt).reshape(6, 260)
    np.testing.assert_allclose(got.astype(np.float64), oracle.sum(h, axis=1).astype(np.float64), rtol=rt, atol=5e-2)
    got = reduce_abi(x, _lib.OP_ARGMAX, _lib.RED_COLS, 6, 700, 260, torch.int64).reshape(6, 260)
    np.testing.assert_array_equal(got, oracle.argmax(h, axis=1))
    got = reduce_abi(x, _lib.OP_MAX, _lib.RED_ROWS, 1, 700 * 260, 6, tdt)
    np.testing.assert_array_equal(got, oracle.amax(h.reshape(6, -1), axis=1))
    if tdt.is_floating_point:
        got = reduce_abi(x, _lib.OP_VAR, _lib.RED_COLS, 1, 6 * 700, 260, tdt, param=1.0)
        np.testing.assert_allclose(got.astype(np.float64), oracle.var(h.reshape(-1, 260), axis=0, ddof=1).astype(np.float64),
                                   rtol=3e-3 if tdt == torch.float16 else 1e-5)


def test_reduce_errors_are_status_codes():
    x = torch.zeros(100, device='cuda')
    y = torch.zeros(1, device='cuda')
    d = _lib.ReduceDesc(_lib.OP_SUM, _lib.RED_FULL, _lib.TYPE_FLOAT32, _lib.TYPE_FLOAT32, 1, 1 << 24, 1, 0.0)
    st = _lib.lib.b200_reduce_run(ctypes.byref(d), x.data_ptr(), y.data_ptr(), None, 0, stream())
    assert st == _lib.E_WORKSPACE and b'workspace' in _lib.lib.b200_last_error_string()
    st = _lib.lib.b200_reduce_run(ctypes.byref(d), None, y.data_ptr(), None, 0, stream())
    assert st == _lib.E_INVALID


@pytest.mark.parametrize('n', [1, 4096, 4097, 1 << 20, (1 << 22) + 123])
def test_scan_through_the_abi(n):
    g = torch.Generator(device='cuda').manual_seed(n)
    x = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', generator=g, dtype=torch.int64)
    y = torch.empty_like(x)
    need = ctypes.c_size_t()
    assert _lib.lib.b200_scan_workspace_bytes(n, _lib.TYPE_INT64, ctypes.byref(need)) == 0
    ws = torch.full((need.value,), 255, dtype=torch.uint8, device='cuda')     # dirty: scan zeroes what it needs
    for _ in range(2):
        st = _lib.lib.b200_scan_run(_lib.OP_CUMSUM, _lib.TYPE_INT64, _lib.TYPE_INT64, x.data_ptr(), y.data_ptr(), n,
                                    ws.data_ptr(), ws.numel(), stream())
        assert st == 0, _lib.last_error()
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.cumsum(x.cpu().numpy()))
    xi = x.to(torch.int32)
    st = _lib.lib.b200_scan_run(_lib.OP_CUMSUM, _lib.TYPE_INT32, _lib.TYPE_INT64, xi.data_ptr(), y.data_ptr(), n,
                                ws.data_ptr(), ws.numel(), stream())
    assert st == 0
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.cumsum(xi.cpu().numpy()))
    st = _lib.lib.b200_scan_run(_lib.OP_CUMSUM, _lib.TYPE_INT64, _lib.TYPE_INT64, x.data_ptr() + 8, y.data_ptr(), n - 1,
                                ws.data_ptr(), ws.numel(), stream()) if n > 1 else _lib.E_UNSUPPORTED
    assert st == _lib.E_UNSUPPORTED                                           # misaligned: host must stage


def test_ufunc_through_the_abi():
    n = 1 << 20
    x = torch.rand(n, device='cuda')
    y = torch.rand(n, device='cuda')
    z = torch.empty(n, device='cuda')
    ops = (_lib.Operand * 3)()
    for o, t, out in zip(ops, (x, y, z), (0, 0, 1)):
        o.data, o.kind, o.dtype, o.ndim, o.is_output = t.data_ptr(), _lib.KIND_ARRAY, _lib.TYPE_FLOAT32, 1, out
        o.shape[0], o.strides[0] = n, 4
    plan = _lib.EwPlan()
    assert _lib.lib.b200_ew_plan(3, ops, ctypes.byref(plan)) == 0
    assert plan.variant == _lib.EW_FLAT and plan.vec == 4
    assert _lib.lib.b200_ufunc_launch(_lib.UFUNC_IDS['add'], ctypes.byref(plan), 3, ops, stream()) == 0
    np.testing.assert_array_equal(z.cpu().numpy(), x.cpu().numpy() + y.cpu().numpy())
    # scalar operand by value
    ops[1].kind = _lib.KIND_SCALAR
    ops[1].scalar[0] = int(np.float32(2.5).view(np.uint32))
    assert _lib.lib.b200_ew_plan(3, ops, ctypes.byref(plan)) == 0
    assert _lib.lib.b200_ufunc_launch(_lib.UFUNC_IDS['multiply'], ctypes.byref(plan), 3, ops, stream()) == 0
    np.testing.assert_array_equal(z.cpu().numpy(), x.cpu().numpy() * np.float32(2.5))
    assert _lib.lib.b200_ufunc_launch(999, ctypes.byref(plan), 3, ops, stream()) == _lib.E_INVALID


def test_jit_compile_load_launch_through_the_abi():
    src = b'''#include <b200/elementwise.cuh>
struct P { float* p; float v; int n; };
extern "C" __global__ void fill(const P a) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < a.n) a.p[i] = a.v; }
'''
    from cupy_b200._core import _jit
    opts = [o.encode() for o in _jit.default_options()]
    c_opts = (ctypes.c_char_p * len(opts))(*opts)
    image, size = ctypes.c_void_p(), ctypes.c_size_t()
    assert _lib.lib.b200_jit_compile(src, b'fill.cu', len(opts), c_opts, ctypes.byref(image), ctypes.byref(size)) == 0
    mod, fn = ctypes.c_void_p(), ctypes.c_void_p()
    assert _lib.lib.b200_module_load(image, ctypes.byref(mod)) == 0
    assert _lib.lib.b200_module_get_function(mod, b'fill', ctypes.byref(fn)) == 0
    assert _lib.lib.b200_module_get_function(mod, b'nope', ctypes.byref(ctypes.c_void_p())) > 0       # CUresult
    t = torch.zeros(1000, device='cuda')

    class P(ctypes.Structure):
        _fields_ = [('p', ctypes.c_void_p), ('v', ctypes.c_float), ('n', ctypes.c_int)]
    p = P(t.data_ptr(), 3.0, 1000)
    assert _lib.lib.b200_jit_launch(fn, 4, 1, 1, 256, 0, ctypes.byref(p), ctypes.sizeof(p), stream()) == 0
    assert float(t.sum().item()) == 3000.0
    assert _lib.lib.b200_module_unload(mod) == 0
    _lib.lib.b200_jit_free_image(image)
