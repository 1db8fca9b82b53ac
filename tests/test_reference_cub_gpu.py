"""GPU tier: pin against the REFERENCE ITSELF -- its native CUB path
(cupy/cuda/cupy_cub.cu, compiled from the reference's own sources into
oracle/_ref/libcupy_cub_ref.so by oracle/Makefile) run on the same device buffers.
This is what the reference calls for sum/max/argmax(axis=None / axis=-1) and
cumsum (cupy/cuda/cub.pyx:137-306)."""
import ctypes
import os

import numpy as np
import pytest

from cupy_b200 import _lib

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', 'libcupy_cub_ref.so')


@pytest.fixture(scope='module')
def ref():
    if not os.path.exists(REF):
        pytest.skip('oracle/_ref/libcupy_cub_ref.so not built (needs the reference tree at build time)')
    lib = ctypes.CDLL(REF)
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.ref_cub_reduce_workspace.restype = sz
    lib.ref_cub_reduce_workspace.argtypes = [vp, vp, i, vp, i, i]
    lib.ref_cub_reduce.argtypes = [vp, sz, vp, vp, i, vp, i, i]
    lib.ref_cub_segmented_reduce_workspace.restype = sz
    lib.ref_cub_segmented_reduce_workspace.argtypes = [vp, vp, i, i, vp, i, i]
    lib.ref_cub_segmented_reduce.argtypes = [vp, sz, vp, vp, i, i, vp, i, i]
    lib.ref_cub_scan_workspace.restype = sz
    lib.ref_cub_scan_workspace.argtypes = [vp, vp, i, vp, i, i]
    lib.ref_cub_scan.argtypes = [vp, sz, vp, vp, i, vp, i, i]
    return lib


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


def stream():
    return torch.cuda.current_stream().cuda_stream


def test_full_and_segmented_sum_max_vs_reference_cub(ref, cp):
    g = torch.Generator(device='cuda').manual_seed(0)
    t = torch.rand(2048, 4096, device='cuda', generator=g) * 2 - 1
    x = cp.from_torch(t)
    n = t.numel()
    for op, ours in ((_lib.OP_SUM, 'sum'), (_lib.OP_MAX, 'max'), (_lib.OP_MIN, 'min')):
        y = torch.empty(1, device='cuda')
        wsb = ref.ref_cub_reduce_workspace(t.data_ptr(), y.data_ptr(), n, stream(), op, _lib.TYPE_FLOAT32)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
        ref.ref_cub_reduce(ws.data_ptr(), wsb, t.data_ptr(), y.data_ptr(), n, stream(), op, _lib.TYPE_FLOAT32)
        got = float(getattr(x, ours)().get())
        if op == _lib.OP_SUM:
            assert abs(got - float(y.item())) <= 1e-5 * float(t.abs().sum().item())
        else:
            assert got == float(y.item())
        ys = torch.empty(2048, device='cuda')
        wsb = ref.ref_cub_segmented_reduce_workspace(t.data_ptr(), ys.data_ptr(), 2048, 4096, stream(), op, _lib.TYPE_FLOAT32)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
        ref.ref_cub_segmented_reduce(ws.data_ptr(), wsb, t.data_ptr(), ys.data_ptr(), 2048, 4096, stream(), op, _lib.TYPE_FLOAT32)
        gs = getattr(x, ours)(axis=1).to_torch()
        if op == _lib.OP_SUM:
            assert float((gs - ys).abs().max().item()) <= 1e-5 * float(t.abs().sum(dim=1).max().item())
        else:
            assert bool((gs == ys).all().item())


def test_argmax_vs_reference_cub(ref, cp):
    g = torch.Generator(device='cuda').manual_seed(1)
    t = torch.rand(1 << 22, device='cuda', generator=g)
    t[[17, 1 << 21, (1 << 22) - 1]] = 2.0                       # three-way tie: the lowest index wins
    n = t.numel()
    kv = torch.empty(2, dtype=torch.int32, device='cuda')      # cub::KeyValuePair<int, float>
    wsb = ref.ref_cub_reduce_workspace(t.data_ptr(), kv.data_ptr(), n, stream(), _lib.OP_ARGMAX, _lib.TYPE_FLOAT32)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
    ref.ref_cub_reduce(ws.data_ptr(), wsb, t.data_ptr(), kv.data_ptr(), n, stream(), _lib.OP_ARGMAX, _lib.TYPE_FLOAT32)
    assert int(cp.from_torch(t).argmax().get()) == int(kv[0].item()) == 17


def test_cumsum_int64_bit_exact_vs_reference_cub(ref, cp):
    g = torch.Generator(device='cuda').manual_seed(2)
    n = (1 << 24) + 77
    t = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', generator=g, dtype=torch.int64)
    y = t.clone()                                               # the reference scans in place after its astype copy
    wsb = ref.ref_cub_scan_workspace(y.data_ptr(), y.data_ptr(), n, stream(), _lib.OP_CUMSUM, _lib.TYPE_INT64)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
    ref.ref_cub_scan(ws.data_ptr(), wsb, y.data_ptr(), y.data_ptr(), n, stream(), _lib.OP_CUMSUM, _lib.TYPE_INT64)
    ours = cp.cumsum(cp.from_torch(t)).to_torch()
    assert bool((ours == y).all().item())
    tf = torch.rand(1 << 20, device='cuda', generator=g)
    yf = tf.clone()
    wsb = ref.ref_cub_scan_workspace(yf.data_ptr(), yf.data_ptr(), 1 << 20, stream(), _lib.OP_CUMSUM, _lib.TYPE_FLOAT32)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
    ref.ref_cub_scan(ws.data_ptr(), wsb, yf.data_ptr(), yf.data_ptr(), 1 << 20, stream(), _lib.OP_CUMSUM, _lib.TYPE_FLOAT32)
    of = cp.cumsum(cp.from_torch(tf)).to_torch()
    assert float(((of - yf).abs() / yf).max().item()) <= 1e-5


def test_accelerator_switch_routes_whole_calls_through_the_reference_in_tests(ref, cp):
    """SURVEY section 5 / VERDICT r1 #7: one setter switches a call between the new engine and the reference
    oracle.  The backend is registered HERE (test infrastructure); the package itself has none."""
    from cupy_b200._core import _accelerator
    g = torch.Generator(device='cuda').manual_seed(5)
    t = torch.rand(1 << 22, device='cuda', generator=g) * 2 - 1
    x = cp.from_torch(t)
    calls = []

    def backend(kind, name, a, axis=None, dtype=None, out=None, keepdims=False):
        if kind != 'reduction' or name != 'cupy_sum' or axis is not None or a.dtype != np.float32:
            return None
        y = cp.empty((), np.float32)
        wsb = ref.ref_cub_reduce_workspace(a.ptr, y.ptr, a.size, stream(), _lib.OP_SUM, _lib.TYPE_FLOAT32)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
        ref.ref_cub_reduce(ws.data_ptr(), wsb, a.ptr, y.ptr, a.size, stream(), _lib.OP_SUM, _lib.TYPE_FLOAT32)
        torch.cuda.synchronize()
        calls.append(name)
        return y

    ours = float(x.sum().get())
    _accelerator.register_reference_backend(backend)
    cp.set_reduction_accelerators(['reference', 'b200'])
    try:
        theirs = float(x.sum().get())
        assert calls == ['cupy_sum']
        assert float(x.max().get()) == float(t.max())          # declined by the backend: the engine runs
    finally:
        _accelerator.register_reference_backend(None)
        cp.set_reduction_accelerators(['b200'])
    assert abs(ours - theirs) <= 1e-5 * float(t.abs().sum())
