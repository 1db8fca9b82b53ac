"""GPU tier: cumsum / cumprod along an axis (b200_scan_axis_run, csrc/scan_axis.cu) against
NumPy -- the reference's own oracle for scans (tests/cupy_tests/math_tests/test_sumprod.py
`TestCumsum/TestCumprod`: shaped_arange inputs, every axis, all dtypes; ints bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


RS = np.random.RandomState(5)


def rnd(shape, dt):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        return (RS.rand(*shape) * 2 - 1).astype(dt)
    if dt.kind == 'b':
        return RS.rand(*shape) > 0.5
    if dt.kind == 'u':
        return RS.randint(0, 9, size=shape).astype(dt)
    return RS.randint(-9, 9, size=shape).astype(dt)


SHAPES = [(2, 3, 4), (20, 30, 40), (1, 1000), (1000, 1), (7, 4100), (4100, 7), (3, 513, 5), (64, 64, 64),
          (5, 70000), (300, 300), (2, 2, 2, 33)]


@pytest.mark.parametrize('dt', ['int8', 'uint8', 'int16', 'int32', 'uint32', 'int64', 'uint64', 'bool'])
@pytest.mark.parametrize('shape', SHAPES)
def test_cumsum_axis_integers_bit_exact(cp, shape, dt):
    a = rnd(shape, dt)
    d = cp.asarray(a)
    for ax in range(a.ndim):
        got = d.cumsum(axis=ax).get()
        want = a.cumsum(axis=ax)
        assert got.dtype == want.dtype and got.shape == want.shape
        np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cp.cumsum(d, axis=-1).get(), a.cumsum(axis=-1))


@pytest.mark.parametrize('dt,rtol', [('float32', 2e-5), ('float64', 1e-12), ('float16', 2e-2)])
@pytest.mark.parametrize('shape', SHAPES)
def test_cumsum_axis_floats(cp, shape, dt, rtol):
    a = (np.abs(rnd(shape, dt)) + 0.01).astype(dt)        # positive: a relative bound is meaningful
    if dt == 'float16':
        a = (a / 64).astype(dt)
    d = cp.asarray(a)
    for ax in range(a.ndim):
        got = d.cumsum(axis=ax).get()
        want = a.astype(np.float64).cumsum(axis=ax)
        assert got.dtype == np.dtype(dt)
        np.testing.assert_allclose(got, want, rtol=rtol)


def test_shaped_arange_known_answers(cp):
    # testing.shaped_arange((2,3,4)) = 1..24 (cupy/testing/_helper.py)
    a = np.arange(1, 25, dtype=np.float32).reshape(2, 3, 4)
    for ax in range(3):
        np.testing.assert_array_equal(cp.cumsum(cp.asarray(a), axis=ax).get(), np.cumsum(a, axis=ax))
    ones = np.ones((3, 10000), np.int32)
    np.testing.assert_array_equal(cp.cumsum(cp.asarray(ones), axis=1).get(), np.tile(np.arange(1, 10001), (3, 1)))
    np.testing.assert_array_equal(cp.cumsum(cp.asarray(ones), axis=0).get(), np.cumsum(ones, axis=0))


@pytest.mark.parametrize('dt', ['int32', 'int64', 'float64'])
def test_cumprod_axis(cp, dt):
    a = rnd((6, 9, 11), dt)
    if np.dtype(dt).kind == 'f':
        a = 1 + a / 8
    else:
        a = np.where(a == 0, 1, np.sign(a) * (1 + (np.abs(a) % 2))).astype(dt)
    d = cp.asarray(a)
    for ax in range(3):
        got, want = cp.cumprod(d, axis=ax).get(), np.cumprod(a, axis=ax)
        assert got.dtype == want.dtype
        if np.dtype(dt).kind == 'f':
            np.testing.assert_allclose(got, want, rtol=1e-12)
        else:
            np.testing.assert_array_equal(got, want)


def test_dtype_out_and_views(cp):
    a = rnd((33, 65, 17), 'int32')
    d = cp.asarray(a)
    np.testing.assert_array_equal(cp.cumsum(d, axis=1, dtype=np.int32).get(), np.cumsum(a, axis=1, dtype=np.int32))
    np.testing.assert_array_equal(cp.cumsum(d, axis=0, dtype=np.float64).get(), np.cumsum(a, axis=0, dtype=np.float64))
    out = cp.empty(a.shape, np.int64)
    r = cp.cumsum(d, axis=2, out=out)
    assert r is out
    np.testing.assert_array_equal(out.get(), np.cumsum(a, axis=2))
    # non-contiguous input and output
    v, nv = d.transpose(2, 0, 1)[::2], a.transpose(2, 0, 1)[::2]
    np.testing.assert_array_equal(cp.cumsum(v, axis=1).get(), np.cumsum(nv, axis=1))
    big = cp.empty((33, 130, 17), np.int64)
    o2 = big[:, ::2]
    cp.cumsum(d, axis=1, out=o2)
    np.testing.assert_array_equal(o2.get(), np.cumsum(a, axis=1))
    # misaligned lines (odd n, 1-byte items) and an offset base
    b = rnd((9, 1001), 'int8')
    np.testing.assert_array_equal(cp.cumsum(cp.asarray(b), axis=1).get(), np.cumsum(b, axis=1))
    f = rnd((64, 1026), 'float32')
    np.testing.assert_allclose(cp.cumsum(cp.asarray(f)[:, 1:], axis=1).get(), np.cumsum(f[:, 1:].astype(np.float64), axis=1),
                               rtol=1e-4, atol=1e-4)
    with pytest.raises(ValueError):
        cp.cumsum(d, axis=1, out=cp.empty((3, 3), np.int64))
    with pytest.raises(Exception):
        cp.cumsum(d, axis=3)


def test_large_axis_scans(cp):
    a = rnd((4096, 8192), 'int32')
    d = cp.asarray(a)
    np.testing.assert_array_equal(d.cumsum(axis=1).get(), a.cumsum(axis=1))
    np.testing.assert_array_equal(d.cumsum(axis=0).get(), a.cumsum(axis=0))


# shapes that take the column-march kernel (csrc/include/b200/scan_march.cuh): a few thousand to ~150 thousand column
# vectors, scanned axis not last; ragged row tiles (n % R != 0), ragged strips (inner % W != 0), several outer slices,
# every strip width (512 / 256 / 128 bytes per row)
MARCH_SHAPES = [(700, 16384), (1030, 4096), (300, 9000), (513, 2052), (3, 600, 4100), (2, 2049, 1540), (260, 40000)]


@pytest.mark.parametrize('dt', ['int64', 'int32', 'uint64', 'float32', 'float64', 'bool', 'int16', 'uint8', 'float16'])
@pytest.mark.parametrize('shape', MARCH_SHAPES)
def test_column_march_scans(cp, shape, dt):
    a = rnd(shape, dt)
    if dt == 'float16':
        a = (a / 64).astype(dt)
    d = cp.asarray(a)
    ax = len(shape) - 2
    got = d.cumsum(axis=ax).get()
    if np.dtype(dt).kind == 'f':
        want = a.astype(np.float64).cumsum(axis=ax)
        np.testing.assert_allclose(got, want, rtol=0, atol={'float32': 1e-3, 'float64': 1e-10, 'float16': 2e-2}[dt])
        assert got.dtype == np.dtype(dt)
    else:
        want = a.cumsum(axis=ax)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want)
    # cumprod: rows / columns past the edge of a tile must count as ones, not as the zeros TMA fills in
    p = np.where(RS.rand(*shape) < 0.002, 2, 1).astype(dt)
    gotp = cp.asarray(p).cumprod(axis=ax).get()
    wantp = p.cumprod(axis=ax)
    if np.dtype(dt).kind == 'f':
        np.testing.assert_allclose(gotp, wantp.astype(np.float64), rtol=1e-6 if dt != 'float16' else 1e-3)
    else:
        np.testing.assert_array_equal(gotp, wantp)
    # in place (out= the input itself), where the result dtype is the input's
    if got.dtype == a.dtype:
        d2 = cp.asarray(a)
        cp.cumsum(d2, axis=ax, out=d2)
        np.testing.assert_array_equal(d2.get(), got)
