"""GPU tier: the ufunc methods beside __call__ -- `at`, `reduceat`, `outer`, `reduce`, `accumulate`
(cupy/_core/_kernel.pyx:1432-1493) -- and the bitwise ufuncs (cupy/_core/_routines_binary.pyx), against NumPy's
methods of the same name, the way tests/cupy_tests/core_tests/test_ufunc_methods.py and
tests/cupy_tests/core_tests/test_ndarray_scatter.py do.  Integer results are bit-exact; float `add.at` sums in
atomic arrival order, so it is compared within a few ulp of the accumulated magnitude."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


RS = np.random.RandomState(11)
SCATTER_DTYPES = ['int32', 'int64', 'uint32', 'uint64', 'float32', 'float64']


def _vals(shape, dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        return (RS.rand(*shape) * 2 - 1).astype(dt)
    if dt.kind == 'u':
        return RS.randint(0, 100, size=shape).astype(dt)
    return RS.randint(-100, 100, size=shape).astype(dt)


def _close(got, want, dt, scale=1.0):
    if np.dtype(dt).kind == 'f':
        np.testing.assert_allclose(got, want, rtol=0, atol=64 * scale * np.finfo(dt).eps)
    else:
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('dt', SCATTER_DTYPES + ['float16'])
def test_add_at_repeated_indices_1d(cp, dt):
    a = _vals((1000,), dt)
    idx = RS.randint(-1000, 1000, size=50000)            # negative indices wrap; ~50 hits per slot
    b = _vals((50000,), dt)
    want = a.copy()
    np.add.at(want, idx, b)
    d = cp.asarray(a)
    cp.add.at(d, cp.asarray(idx), cp.asarray(b))
    if dt == 'float16':
        np.testing.assert_allclose(d.get().astype('f8'), want.astype('f8'), atol=0.5)   # fp16 running sums round
    else:
        _close(d.get(), want, dt, scale=50)


@pytest.mark.parametrize('dt', ['int64', 'float32'])
def test_add_at_rows_scalar_and_broadcast_values(cp, dt):
    a = _vals((64, 33), dt)
    idx = np.array([3, 3, -1, 0, 3, 63])
    want = a.copy()
    np.add.at(want, idx, 2)
    d = cp.asarray(a)
    cp.add.at(d, idx, 2)                                  # host index + python scalar, as numpy takes them
    _close(d.get(), want, dt)
    row = _vals((33,), dt)
    np.add.at(want, idx, row)
    cp.add.at(d, cp.asarray(idx), cp.asarray(row))
    _close(d.get(), want, dt, scale=8)
    block = _vals((6, 33), dt)
    np.add.at(want, idx, block)
    cp.add.at(d, cp.asarray(idx), cp.asarray(block))
    _close(d.get(), want, dt, scale=16)


def test_add_at_tuple_of_index_arrays_and_2d_index(cp):
    a = _vals((20, 30, 4), 'int64')
    i0 = RS.randint(-20, 20, size=(7, 1))
    i1 = RS.randint(-30, 30, size=(1, 9))
    v = _vals((7, 9, 4), 'int64')
    want = a.copy()
    np.add.at(want, (i0, i1), v)
    d = cp.asarray(a)
    cp.add.at(d, (cp.asarray(i0), cp.asarray(i1)), cp.asarray(v))
    np.testing.assert_array_equal(d.get(), want)
    i2 = RS.randint(0, 20, size=(5, 6))                   # one 2-d index array along axis 0
    v2 = _vals((5, 6, 30, 4), 'int64')
    np.add.at(want, i2, v2)
    cp.add.at(d, cp.asarray(i2), cp.asarray(v2))
    np.testing.assert_array_equal(d.get(), want)


def test_add_at_boolean_mask(cp):
    a = _vals((50, 8), 'int32')
    m = RS.rand(50) < 0.4
    want = a.copy()
    np.add.at(want, m, 5)
    d = cp.asarray(a)
    cp.add.at(d, cp.asarray(m), 5)
    np.testing.assert_array_equal(d.get(), want)
    m2 = RS.rand(50, 8) < 0.3
    v = _vals((int(m2.sum()),), 'int32')
    np.add.at(want, m2, v)
    cp.add.at(d, cp.asarray(m2), cp.asarray(v))
    np.testing.assert_array_equal(d.get(), want)
    none = np.zeros(50, bool)
    cp.add.at(d, cp.asarray(none), 1)                     # nothing selected: unchanged
    np.testing.assert_array_equal(d.get(), want)


@pytest.mark.parametrize('name', ['maximum', 'minimum'])
@pytest.mark.parametrize('dt', SCATTER_DTYPES)
def test_maximum_minimum_at(cp, name, dt):
    a = _vals((300,), dt)
    idx = RS.randint(0, 300, size=20000)
    b = _vals((20000,), dt)
    if np.dtype(dt).kind == 'f':
        b[::977] = np.nan                                  # NaN propagates, as numpy.maximum / minimum
    want = a.copy()
    getattr(np, name).at(want, idx, b)
    d = cp.asarray(a)
    getattr(cp, name).at(d, cp.asarray(idx), cp.asarray(b))
    np.testing.assert_array_equal(d.get(), want)


@pytest.mark.parametrize('name', ['bitwise_and', 'bitwise_or', 'bitwise_xor'])
@pytest.mark.parametrize('dt', ['int32', 'int64', 'uint32', 'uint64'])
def test_bitwise_at(cp, name, dt):
    a = RS.randint(0, 1 << 30, size=200).astype(dt)
    idx = RS.randint(0, 200, size=5000)
    b = RS.randint(0, 1 << 30, size=5000).astype(dt)
    want = a.copy()
    getattr(np, name).at(want, idx, b)
    d = cp.asarray(a)
    getattr(cp, name).at(d, cp.asarray(idx), cp.asarray(b))
    np.testing.assert_array_equal(d.get(), want)


def test_subtract_at_int32_only_and_unsupported(cp):
    a = _vals((100,), 'int32')
    idx = RS.randint(0, 100, size=3000)
    b = _vals((3000,), 'int32')
    want = a.copy()
    np.subtract.at(want, idx, b)
    d = cp.asarray(a)
    cp.subtract.at(d, cp.asarray(idx), cp.asarray(b))
    np.testing.assert_array_equal(d.get(), want)
    with pytest.raises(TypeError):
        cp.subtract.at(cp.zeros((4,), 'float32'), cp.asarray(np.array([0])), 1.0)
    with pytest.raises(TypeError):
        cp.add.at(cp.zeros((4,), 'int8'), cp.asarray(np.array([0])), 1)
    with pytest.raises(NotImplementedError):
        cp.multiply.at(cp.zeros((4,), 'float32'), cp.asarray(np.array([0])), 1.0)
    with pytest.raises(ValueError):
        cp.add.at(cp.zeros((4,), 'float32'), cp.asarray(np.array([0])))


def test_numpy_ufunc_at_dispatches_to_the_device(cp):
    a = _vals((40,), 'int64')
    idx = np.array([1, 1, 1, 39, 0])
    want = a.copy()
    np.add.at(want, idx, 7)
    d = cp.asarray(a)
    np.add.at(d, idx, 7)                                   # __array_ufunc__ (method 'at')
    np.testing.assert_array_equal(d.get(), want)


@pytest.mark.parametrize('dt', ['int32', 'int64', 'float32', 'float64', 'bool', 'float16'])
@pytest.mark.parametrize('shape,axis', [((1000,), 0), ((37, 50), 0), ((37, 50), 1), ((5, 40, 6), 1), ((5, 40, 6), -1)])
def test_add_reduceat(cp, dt, shape, axis):
    a = (RS.rand(*shape) < 0.5) if dt == 'bool' else _vals(shape, dt)
    n = shape[axis]
    idx = np.array([0, n // 3, n // 3, n // 2, 2, n - 1, 1])      # empty, decreasing and last-element segments
    want = np.add.reduceat(a, idx, axis=axis)
    got = cp.add.reduceat(cp.asarray(a), cp.asarray(idx), axis=axis)
    assert got.dtype == want.dtype and got.shape == want.shape
    if np.dtype(dt).kind == 'f':
        # the reference's formulation: a difference of two prefix sums, each up to n terms long
        np.testing.assert_allclose(got.get().astype('f8'), want.astype('f8'), rtol=0,
                                   atol=4 * n * float(np.finfo(dt).eps))
    else:
        np.testing.assert_array_equal(got.get(), want)


def test_add_reduceat_out_dtype_and_errors(cp):
    a = _vals((30, 8), 'int32')
    idx = [0, 4, 10]
    out = cp.empty((3, 8), 'int64')
    r = cp.add.reduceat(cp.asarray(a), idx, axis=0, dtype='int64', out=out)
    assert r is out
    np.testing.assert_array_equal(out.get(), np.add.reduceat(a, idx, axis=0, dtype='int64'))
    with pytest.raises(IndexError):
        cp.add.reduceat(cp.asarray(a), [0, 30], axis=0)
    with pytest.raises(NotImplementedError):
        cp.multiply.reduceat(cp.asarray(a), idx)


def test_outer_reduce_accumulate(cp):
    a, b = _vals((7, 3), 'float32'), _vals((5,), 'float32')
    np.testing.assert_array_equal(cp.multiply.outer(cp.asarray(a), cp.asarray(b)).get(), np.multiply.outer(a, b))
    x = _vals((40, 9), 'int64')
    d = cp.asarray(x)
    np.testing.assert_array_equal(cp.add.reduce(d, axis=1).get(), np.add.reduce(x, axis=1))
    np.testing.assert_array_equal(cp.multiply.reduce(d[:, :3], axis=0).get(), np.multiply.reduce(x[:, :3], axis=0))
    np.testing.assert_array_equal(cp.maximum.reduce(d, axis=0).get(), np.maximum.reduce(x, axis=0))
    np.testing.assert_array_equal(cp.add.accumulate(d, axis=1).get(), np.add.accumulate(x, axis=1))
    np.testing.assert_array_equal(cp.multiply.accumulate(d[:, :4], axis=1).get(), np.multiply.accumulate(x[:, :4], axis=1))


@pytest.mark.parametrize('dt', ['bool', 'int8', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'uint64'])
def test_bitwise_ufuncs_and_operators(cp, dt):
    if dt == 'bool':
        a, b = RS.rand(500) < 0.5, RS.rand(500) < 0.5
    else:
        a = RS.randint(0, 120, size=500).astype(dt)
        b = RS.randint(0, 120, size=500).astype(dt)
    da, db = cp.asarray(a), cp.asarray(b)
    for name, op in (('bitwise_and', '&'), ('bitwise_or', '|'), ('bitwise_xor', '^')):
        want = getattr(np, name)(a, b)
        got = getattr(cp, name)(da, db)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got.get(), want)
        np.testing.assert_array_equal(eval('da %s db' % op).get(), want)
    np.testing.assert_array_equal((~da).get(), ~a)
    np.testing.assert_array_equal(cp.invert(da).get(), np.invert(a))
    if dt != 'bool':
        s = (b % 5).astype(dt)
        np.testing.assert_array_equal((da << cp.asarray(s)).get(), a << s)
        np.testing.assert_array_equal((da >> cp.asarray(s)).get(), a >> s)
        np.testing.assert_array_equal((da >> 2).get(), a >> 2)
        acc = cp.asarray(a)
        acc |= db
        np.testing.assert_array_equal(acc.get(), a | b)
    else:
        got = cp.left_shift(da, db)                          # no bool loop: the first integer loop that fits (int8)
        assert got.dtype == np.left_shift(a, b).dtype
        np.testing.assert_array_equal(got.get(), np.left_shift(a, b))
    with pytest.raises(TypeError):
        cp.bitwise_and(cp.ones((3,), 'float32'), cp.ones((3,), 'float32'))
