"""GPU tier: cupy_b200.fuse (drop-in for cupy.fuse, cupy/_core/fusion.pyx) against the same
Python function evaluated by NumPy, the way the reference tests it
(tests/cupy_tests/core_tests/fusion_tests/*: `fusion_utils.check_fusion` compares the fused
CuPy function with the un-fused NumPy call).  IEEE-exact chains compare bit-exactly."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


RS = np.random.RandomState(3)


def rnd(shape, dt='float32'):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        return (RS.rand(*shape) * 2 - 1).astype(dt)
    return RS.randint(-50, 50, size=shape).astype(dt)


def test_config1_x_times_2_plus_1_is_one_kernel(cp):
    from cupy_b200._core import _dryrun

    @cp.fuse()
    def f(x):
        return x * 2 + 1

    a = rnd((4096, 4096))
    np.testing.assert_array_equal(f(cp.asarray(a)).get(), a * 2 + 1)
    with _dryrun.dry_run() as dry:
        f(cp.empty((4096, 4096), 'f'))
    assert len(dry) == 1 and dry[0]['kind'] == 'jit_elementwise'
    # NEP 50 weak scalars inside the fused function: int8 stays int8, float16 stays float16
    i8 = rnd((1000,), 'int8')
    assert f(cp.asarray(i8)).dtype == np.int8
    np.testing.assert_array_equal(f(cp.asarray(i8)).get(), i8 * 2 + 1)
    h = rnd((1000,), 'float16')
    np.testing.assert_array_equal(f(cp.asarray(h)).get(), h * 2 + 1)


def test_config4a_exp_plus_row_fused_on_transposed_input(cp):
    from cupy_b200 import _lib
    from cupy_b200._core import _dryrun

    @cp.fuse(kernel_name='exp_plus_row')
    def f(x, v):
        return cp.exp(x) + v

    a = rnd((36, 50, 72))
    v = rnd((36,))
    xt = cp.asarray(a).transpose(2, 1, 0)
    got = f(xt, cp.asarray(v)).get()
    e = oracle.exp_exact(a.transpose(2, 1, 0))
    want = e.astype(np.float64) + v
    tol = 2 * np.spacing(e) + np.spacing(np.abs(want).astype(np.float32))
    assert (np.abs(got - want) <= tol).all()
    # identical to the two separate ufunc launches, bit for bit
    np.testing.assert_array_equal(got, (cp.exp(xt) + cp.asarray(v)).get())
    with _dryrun.dry_run() as dry:
        f(cp.empty((16, 64, 128), 'f').transpose(2, 1, 0), cp.empty((16,), 'f'))
    assert len(dry) == 1 and dry[0]['variant'] == _lib.EW_TILED_REG
    assert 'load<_FULL>(2' not in dry[0]['source']          # the output is never read


@pytest.mark.parametrize('dt', ['float32', 'float64', 'int32', 'int64'])
def test_chain_of_operators_and_ufuncs(cp, dt):
    @cp.fuse()
    def f(x, y, z):
        t = cp.maximum(x, y) - z
        return cp.absolute(t) * 3 + cp.minimum(x, z), t * t

    a, b, c = rnd((300, 77), dt), rnd((300, 77), dt), rnd((77,), dt)
    r0, r1 = f(cp.asarray(a), cp.asarray(b), cp.asarray(c))
    t = np.maximum(a, b) - c
    if np.dtype(dt).kind == 'f':      # abs(t)*3 + min is contracted into one FMA, as in the reference's NVRTC build
        tol = 1e-6 if dt == 'float32' else 1e-14
        np.testing.assert_allclose(r0.get(), np.absolute(t) * 3 + np.minimum(a, c), rtol=tol, atol=tol)
    else:
        np.testing.assert_array_equal(r0.get(), np.absolute(t) * 3 + np.minimum(a, c))
    np.testing.assert_array_equal(r1.get(), t * t)
    assert r0.dtype == np.dtype(dt) and r1.dtype == np.dtype(dt)


def test_scalars_casts_comparisons_and_inplace(cp):
    @cp.fuse()
    def f(a, x, y):
        y += a * x                    # in place: writes back into the argument
        return (y > 0).astype('float32') * x.astype('float64')

    a = np.float32(1.5)
    x, y = rnd((1000,)), rnd((1000,))
    dy = cp.asarray(y)
    r = f(a, cp.asarray(x), dy)
    y2 = y + a * x
    np.testing.assert_allclose(dy.get(), y2, rtol=1e-6, atol=1e-6)     # a*x+y may be contracted to one FMA
    np.testing.assert_array_equal(r.get(), (dy.get() > 0).astype('float32') * x.astype('float64'))
    assert r.dtype == np.float64

    @cp.fuse()
    def g(x, s):
        return x / s + 2.5

    xi = rnd((50, 40), 'int32')
    np.testing.assert_array_equal(g(cp.asarray(xi), 4).get(), xi / 4 + 2.5)       # python int scalar argument
    np.testing.assert_allclose(g(cp.asarray(x), np.float32(3)).get(), x / np.float32(3) + 2.5, rtol=1e-6)


@pytest.mark.parametrize('axis', [None, 0, 1])
def test_trailing_reduction(cp, axis):
    @cp.fuse(kernel_name='sq_diff_sum')
    def f(x, y):
        return cp.sum((x - y) * (x - y), axis=axis)

    @cp.fuse()
    def g(x):
        return abs(x).max(axis=axis)

    a, b = rnd((257, 1030)), rnd((257, 1030))
    got = f(cp.asarray(a), cp.asarray(b)).get()
    want = ((a.astype(np.float64) - b) ** 2).sum(axis=axis)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    np.testing.assert_array_equal(g(cp.asarray(a)).get(), np.abs(a).max(axis=axis))
    i = rnd((64, 100), 'int8')
    r = cp.fuse(lambda x: cp.sum(x * x, axis=axis))(cp.asarray(i))
    assert r.dtype == np.int64                                   # int8 products wrap, then sum promotes
    np.testing.assert_array_equal(r.get(), (i * i).sum(axis=axis))


def test_errors_and_fallbacks(cp):
    @cp.fuse()
    def f(x):
        return x + 1

    assert f(3) == 4                                   # no device array: plain Python call
    with pytest.raises(NotImplementedError):
        cp.fuse(lambda x: cp.sum(x) + 1)(cp.asarray(rnd((10,))))
    with pytest.raises(TypeError):
        cp.fuse(lambda x: x if x > 0 else -x)(cp.asarray(rnd((10,))))
    outside = cp.asarray(rnd((10,)))
    with pytest.raises(TypeError):
        cp.fuse(lambda x: x + outside)(cp.asarray(rnd((10,))))


def test_out_argument_then_read(cp):
    """`cp.add(a, b, out=a); return a * 2` computes (a + b) * 2 and updates a."""
    @cp.fuse(kernel_name='fused_out_read')
    def f(a, b):
        cp.add(a, b, out=a)
        return a * 2

    rs = np.random.RandomState(3)
    a, b = rs.rand(1000).astype('f'), rs.rand(1000).astype('f')
    da, db = cp.asarray(a), cp.asarray(b)
    r = f(da, db)
    np.testing.assert_array_equal(r.get(), (a + b) * 2)
    np.testing.assert_array_equal(da.get(), a + b)
