"""GPU tier: the transposing tilers of cupy_b200/csrc/include/b200/elementwise.cuh against
NumPy on the same inputs -- RegTileTiler (B200_EW_TILED_REG, the default: register-block
transpose), TmaTileTiler (B200_EW_TILED_TMA: TMA-pipelined swizzled tiles, A/B knob) and the
plain shared-memory TileTiler (B200_EW_TILED: mixed item sizes / misaligned views).  Every
test runs under all three planner modes (B200_EW_TILED_MODE).

The calls are the reference's transposed-operand ufunc / ElementwiseKernel launches
(cupy/_core/_kernel.pyx:1241-1401, 862-1002).  Copies and add/mul are IEEE-exact -> bit
equality; exp within 2 ulp.  Shapes cover partial tiles along both tile axes, batch dims,
ring wrap-around (more tiles per block than ring stages), 2/4/8-byte items, several staged
operands, read-modify-write outputs and direct (non-staged) operands of every stride kind."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


@pytest.fixture(scope='module')
def lib():
    from cupy_b200 import _lib
    return _lib


RS = np.random.RandomState(11)


@pytest.fixture(autouse=True, params=['reg', 'tma', 'smem'])
def mode(request):
    import os
    os.environ['B200_EW_TILED_MODE'] = request.param
    yield request.param
    del os.environ['B200_EW_TILED_MODE']


def rnd(shape, dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        return (RS.rand(*shape) * 2 - 1).astype(dt)
    return RS.randint(-1000, 1000, size=shape).astype(dt)


def _variant_of(cp, fn):
    from cupy_b200._core import _dryrun
    with _dryrun.dry_run() as dry:
        fn()
    return dry[-1]['variant']


SHAPES_2D = [(32, 128), (64, 256), (33, 132), (31, 4), (1000, 1000), (4096, 512), (17, 2048), (2048, 20), (5000, 36)]


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64', 'int32', 'int64', 'int16'])
@pytest.mark.parametrize('shape', SHAPES_2D)
def test_transposed_copy_2d(cp, lib, shape, dt):
    a = rnd(shape, dt)                       # (O, I) in memory; the view is (I, O)^T
    d = cp.asarray(a)
    if min(shape) >= 16:
        assert _variant_of(cp, lambda: cp.empty(shape, dt).T.copy()) in (lib.EW_TILED_REG, lib.EW_TILED_TMA, lib.EW_TILED)
    np.testing.assert_array_equal(d.T.copy().get(), a.T)


@pytest.mark.parametrize('shape,perm', [
    ((8, 40, 64), (2, 1, 0)), ((3, 5, 7, 64, 48), (3, 1, 2, 0, 4)), ((2, 100, 36), (0, 2, 1)),
    ((6, 257, 128), (2, 0, 1)), ((4, 4, 1000, 24), (0, 3, 2, 1)), ((300, 8, 260), (2, 1, 0)),
])
@pytest.mark.parametrize('dt', ['float32', 'float16', 'int64'])
def test_transposed_nd(cp, shape, perm, dt):
    a = rnd(shape, dt)
    d = cp.asarray(a)
    np.testing.assert_array_equal(d.transpose(perm).copy().get(), a.transpose(perm))
    b = rnd(tuple(shape[p] for p in perm), dt)
    np.testing.assert_array_equal((d.transpose(perm) + cp.asarray(b)).get(), a.transpose(perm) + b)


def test_variant_for_config4a_shape(cp, lib, mode):
    want = {'reg': lib.EW_TILED_REG, 'tma': lib.EW_TILED_TMA, 'smem': lib.EW_TILED}[mode]
    xt = cp.empty((16, 64, 128), 'f').transpose(2, 1, 0)
    assert _variant_of(cp, lambda: cp.exp(xt)) == want
    k = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'fused_tiled_probe')
    assert _variant_of(cp, lambda: k(xt, cp.empty((16,), 'f'))) == want


@pytest.mark.parametrize('shape', [(16, 64, 128), (36, 50, 70), (256, 8, 1024)])
def test_config4a_exp_plus_row(cp, shape):
    a = rnd(shape, 'float32')
    v = rnd((shape[0],), 'float32')
    xt = cp.asarray(a).transpose(2, 1, 0)
    dv = cp.asarray(v)
    fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'fused_tma')
    got = fused(xt, dv).get()
    e = oracle.exp_exact(a.transpose(2, 1, 0))
    want = e.astype(np.float64) + v
    # <= 2 ulp on exp, one more rounding on the add
    tol = 2 * np.spacing(e) + np.spacing(np.abs(want).astype(np.float32))
    assert (np.abs(got - want) <= tol).all()
    two = (cp.exp(xt) + dv).get()
    np.testing.assert_array_equal(two, (cp.exp(xt).get() + v))
    assert oracle.ulp_diff(cp.exp(xt).get(), e).max() <= 2


def test_ring_wraparound_many_tiles_per_block(cp):
    # 2^25 elements -> 8192 tiles of 32x128 over <= 296 blocks: every block wraps its ring several times
    a = rnd((4096, 8192), 'float32')
    d = cp.asarray(a)
    np.testing.assert_array_equal(d.T.copy().get(), a.T)
    np.testing.assert_array_equal(cp.multiply(d.T, np.float32(3)).get(), a.T * np.float32(3))


def test_two_staged_operands_and_direct_operands(cp):
    a, b = rnd((520, 200), 'float32'), rnd((520, 200), 'float32')
    c = rnd((200, 520), 'float32')
    da, db, dc = cp.asarray(a), cp.asarray(b), cp.asarray(c)
    np.testing.assert_array_equal((da.T + db.T).get(), a.T + b.T)             # ROWWISE-or-FLAT after collapse
    np.testing.assert_array_equal((da.T + dc).get(), a.T + c)                 # staged + direct unit-stride
    k3 = cp.ElementwiseKernel('T x, T y, T z', 'T w', 'w = x * y + z', 'fma3')
    # integer-valued operands: exact whether or not the compiler contracts to an FMA
    ia, ib, ic = (RS.randint(-500, 500, size=s).astype(np.float32) for s in ((520, 200), (520, 200), (200, 520)))
    got = k3(cp.asarray(ia).T, cp.asarray(ic), cp.asarray(ib).T).get()        # two staged, one direct
    np.testing.assert_array_equal(got, ia.T * ic + ib.T)
    col = rnd((200, 1), 'float32')
    row = rnd((1, 520), 'float32')
    np.testing.assert_array_equal((da.T + cp.asarray(col)).get(), a.T + col)  # stride 0 along O
    np.testing.assert_array_equal((da.T + cp.asarray(row)).get(), a.T + row)  # stride 0 along I
    wide = rnd((200, 1040), 'float32')
    np.testing.assert_array_equal((da.T + cp.asarray(wide)[:, ::2]).get(), a.T + wide[:, ::2])   # strided along O


def test_read_modify_write_output(cp):
    a = rnd((260, 96), 'float32')
    acc = rnd((96, 260), 'float32')
    k = cp.ElementwiseKernel('T x', 'T y', 'y += x', 'accumulate_t')
    d = cp.asarray(acc.copy())
    k(cp.asarray(a).T, d)
    np.testing.assert_array_equal(d.get(), acc + a.T)


def test_kernel_index_i_matches_c_order(cp):
    a = rnd((64, 130, 36), 'float32')
    k = cp.ElementwiseKernel('T x', 'int64 y', 'y = i', 'lin_index')
    got = k(cp.asarray(a).transpose(2, 1, 0)).get()
    np.testing.assert_array_equal(got, np.arange(a.size, dtype=np.int64).reshape(36, 130, 64))
    kf = cp.ElementwiseKernel('T x', 'T y', 'y = i', 'lin_index_f')       # same item size -> TMA tiler
    gotf = kf(cp.asarray(a).transpose(2, 1, 0)).get()
    np.testing.assert_array_equal(gotf, np.arange(a.size, dtype=np.float32).reshape(36, 130, 64))


def test_where_mask_and_out_argument(cp):
    a = rnd((128, 68), 'float32')
    m = RS.rand(68, 128) > 0.5
    out = rnd((68, 128), 'float32')
    d_out = cp.asarray(out.copy())
    cp.add(cp.asarray(a).T, np.float32(1), out=d_out, _where=cp.asarray(m))   # the reference spells it _where (_kernel.pyx:1268)
    want = out.copy()
    np.add(a.T, np.float32(1), out=want, where=m)
    np.testing.assert_array_equal(d_out.get(), want)
