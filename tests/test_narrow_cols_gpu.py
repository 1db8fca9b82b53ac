"""GPU tier: axis-0 reductions of tall, narrow matrices (rows of at most 64 elements: point clouds, feature
tables) -- the flat-stream kernel `reduce_narrow_body` (csrc/include/b200/reduce.cuh) -- against NumPy, over
column counts that do and do not divide the vector width, ragged row counts, every prebuilt functor, planted
ties and NaNs (first occurrence, NaN wins: tests/cupy_tests/sorting_tests/test_search.py:26-66)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


RS = np.random.RandomState(77)
SHAPES = [(100000, 3), (70001, 4), (50000, 7), (33333, 16), (20011, 100), (9000, 128), (40000, 1), (300007, 2),
          (12345, 33), (8193, 64), (1 << 20, 5), (17, 128), (1100, 30)]


def _data(shape, dt):
    dt = np.dtype(dt)
    if dt.kind == 'b':
        return RS.rand(*shape) < 0.999
    if dt.kind == 'f':
        return (RS.rand(*shape) * 2 - 1).astype(dt)
    if dt.kind == 'u':
        return RS.randint(0, 100, size=shape).astype(dt)
    return RS.randint(-100, 100, size=shape).astype(dt)


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64', 'int32', 'int64', 'int8', 'uint8', 'bool'])
@pytest.mark.parametrize('shape', SHAPES)
def test_axis0_reductions_of_narrow_matrices(cp, shape, dt):
    a = _data(shape, dt)
    d = cp.asarray(a)
    n = shape[0]
    kind = np.dtype(dt).kind
    want = a.sum(axis=0)
    got = d.sum(axis=0).get()
    assert got.dtype == want.dtype and got.shape == want.shape
    if kind == 'f':
        tol = {2: 2e-3, 4: 1e-6, 8: 1e-14}[np.dtype(dt).itemsize]
        ref = a.astype(np.float64).sum(axis=0)
        np.testing.assert_allclose(got.astype('f8'), ref, rtol=tol, atol=tol * np.abs(a.astype('f8')).sum(axis=0).max())
        np.testing.assert_allclose(d.mean(axis=0).get().astype('f8'), ref / n, rtol=tol, atol=tol)
        np.testing.assert_allclose(d.var(axis=0).get().astype('f8'), a.astype('f8').var(axis=0), rtol=max(tol * 10, 1e-5))
        np.testing.assert_allclose(d.var(axis=0, ddof=1).get().astype('f8'), a.astype('f8').var(axis=0, ddof=1), rtol=max(tol * 10, 1e-5))
    else:
        np.testing.assert_array_equal(got, want)
        np.testing.assert_allclose(d.mean(axis=0).get(), a.mean(axis=0), rtol=1e-12)
        np.testing.assert_allclose(d.var(axis=0).get(), a.var(axis=0), rtol=1e-9)
    for name in ('max', 'min', 'argmax', 'argmin'):
        np.testing.assert_array_equal(getattr(d, name)(axis=0).get(), getattr(a, name)(axis=0), err_msg=name)
    np.testing.assert_array_equal(d.any(axis=0).get(), a.any(axis=0))
    np.testing.assert_array_equal(d.all(axis=0).get(), a.all(axis=0))
    # keepdims and a 3-d array whose leading axes are reduced together
    np.testing.assert_array_equal(d.max(axis=0, keepdims=True).get(), a.max(axis=0, keepdims=True))
    if n % 10 == 0:
        a3, d3 = a.reshape(10, n // 10, shape[1]), d.reshape(10, n // 10, shape[1])
        np.testing.assert_array_equal(d3.argmax(axis=1).get(), a3.argmax(axis=1))       # batch > 1: strip kernel
        np.testing.assert_array_equal(d3.max(axis=(0, 1)).get(), a3.max(axis=(0, 1)))   # collapses to (n, cols)


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64'])
@pytest.mark.parametrize('cols', [3, 4, 16, 100])
def test_narrow_arg_reductions_ties_and_nans(cp, cols, dt):
    n = 200003
    a = (RS.randint(-5, 6, size=(n, cols))).astype(dt)          # many ties: first occurrence must win
    a[RS.randint(0, n, size=50), RS.randint(0, cols, size=50)] = 7     # planted maxima, repeated rows possible
    d = cp.asarray(a)
    np.testing.assert_array_equal(d.argmax(axis=0).get(), a.argmax(axis=0))
    np.testing.assert_array_equal(d.argmin(axis=0).get(), a.argmin(axis=0))
    b = a.copy()
    b[n - 5, 0] = np.nan                                          # NaN beats everything, first NaN wins
    b[n // 2, 0] = np.nan
    b[3, cols - 1] = np.nan
    b[n - 1, cols // 2] = np.nan                                  # in the ragged tail
    d = cp.asarray(b)
    np.testing.assert_array_equal(d.argmax(axis=0).get(), b.argmax(axis=0))
    np.testing.assert_array_equal(d.argmin(axis=0).get(), b.argmin(axis=0))
    np.testing.assert_array_equal(d.max(axis=0).get(), b.max(axis=0))
    np.testing.assert_array_equal(np.isnan(d.sum(axis=0).get()), np.isnan(b.sum(axis=0)))
    c = np.full((n, cols), -np.inf, dtype=dt)                     # every element equals the running start value
    np.testing.assert_array_equal(cp.asarray(c).argmax(axis=0).get(), c.argmax(axis=0))
    c[n - 2, 1 % cols] = 0
    np.testing.assert_array_equal(cp.asarray(c).argmax(axis=0).get(), c.argmax(axis=0))


def test_narrow_misaligned_base_and_views(cp):
    """A base pointer that is not vector-aligned (a row-sliced view) falls back to the strip kernel: same results."""
    a = (RS.rand(50001, 3) * 2 - 1).astype(np.float32)
    d = cp.asarray(a)
    np.testing.assert_allclose(d[1:].sum(axis=0).get(), a[1:].astype('f8').sum(axis=0), rtol=1e-5)
    np.testing.assert_array_equal(d[1:].argmax(axis=0).get(), a[1:].argmax(axis=0))
    np.testing.assert_array_equal(d[4:].argmax(axis=0).get(), a[4:].argmax(axis=0))         # 48-byte offset: aligned again
    np.testing.assert_array_equal(d[:, :2].max(axis=0).get(), a[:, :2].max(axis=0))         # strided rows: generic route


SCAN_SHAPES = [(100000, 3), (70001, 4), (50000, 7), (33333, 16), (20011, 33), (8193, 64), (300007, 2), (1 << 20, 5),
               (65536, 1), (1100, 60), (2048, 32), (30011, 100), (9000, 128), (5000, 250), (4099, 256), (3000, 255)]


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64', 'int32', 'int64', 'int8', 'uint8', 'uint32', 'bool'])
@pytest.mark.parametrize('shape', SCAN_SHAPES)
def test_axis0_scans_of_narrow_matrices(cp, shape, dt):
    """cumsum / cumprod along axis 0 of rows of at most 128 elements (scan_narrow.cuh; wider ones take the strip routes): flat-stream tiles, one
    segment per block, segment totals first.  Integers bit-exact; floats against a float64 scan."""
    a = _data(shape, dt)
    if np.dtype(dt).kind == 'f':
        a = (a / 8).astype(dt)
    d = cp.asarray(a)
    want = np.cumsum(a, axis=0)
    got = cp.cumsum(d, axis=0)
    assert got.dtype == want.dtype and got.shape == want.shape
    if np.dtype(dt).kind == 'f':
        ref = np.cumsum(a.astype(np.float64), axis=0)
        eps = {2: 1e-3, 4: 1.2e-7, 8: 2.3e-16}[np.dtype(dt).itemsize]
        bound = eps * (np.abs(ref) + 1) + 4 * {2: 1.2e-7, 4: 1.2e-7, 8: 2.3e-16}[np.dtype(dt).itemsize] * np.cumsum(np.abs(a.astype('f8')), axis=0)
        assert np.all(np.abs(got.get().astype('f8') - ref) <= bound)
    else:
        np.testing.assert_array_equal(got.get(), want)
        # cumprod of {-1, 0, 1, 2}-ish values wraps in the result dtype exactly as NumPy's
        b = (np.abs(a.astype(np.int64)) % 3 - 1 + (a.astype(np.int64) % 7 == 0)).astype(dt)
        np.testing.assert_array_equal(cp.cumprod(cp.asarray(b), axis=0).get(), np.cumprod(b, axis=0))
    # out= of another dtype, dtype=, in place, negative axis of the 2-d array
    np.testing.assert_array_equal(cp.cumsum(d, axis=-2).get(), got.get())
    if np.dtype(dt).kind in 'iu':
        np.testing.assert_array_equal(cp.cumsum(d, axis=0, dtype=dt).get(), np.cumsum(a, axis=0, dtype=dt))
    if dt in ('float32', 'int64'):
        e = cp.asarray(a)
        r = cp.cumsum(e, axis=0, out=e)
        assert r is e
        np.testing.assert_array_equal(e.get(), got.get())


def test_narrow_scan_is_deterministic_and_matches_the_strip_route(cp, monkeypatch):
    a = (RS.rand(1 << 19, 6) * 2 - 1).astype(np.float32)
    d = cp.asarray(a)
    first = cp.cumsum(d, axis=0).get()
    for _ in range(3):
        np.testing.assert_array_equal(cp.cumsum(d, axis=0).get(), first)         # bit-identical run to run
    ref = np.cumsum(a.astype(np.float64), axis=0)
    assert np.abs(first - ref).max() <= 1e-6 * np.abs(a).sum(axis=0).max()
    # a row-offset view is not 16-byte aligned for every column count: the general route takes it
    np.testing.assert_allclose(cp.cumsum(d[1:], axis=0).get(), np.cumsum(a[1:].astype('f8'), axis=0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64', 'int32', 'int64', 'int8', 'uint32', 'bool'])
@pytest.mark.parametrize('shape', [(100000, 3), (70001, 4), (50000, 7), (33333, 16), (20011, 33), (8193, 64), (300007, 2),
                                   (1 << 20, 5), (17, 4000, 6), (1100, 60)])
def test_scans_along_short_rows(cp, shape, dt):
    """cumsum / cumprod along the contiguous axis when rows have at most 64 elements (scan_short_rows_body: whole
    rows staged in shared memory, a thread per row).  Rows are independent: every row is checked."""
    a = _data(shape, dt)
    if np.dtype(dt).kind == 'f':
        a = (a / 2).astype(dt)
    d = cp.asarray(a)
    want = np.cumsum(a, axis=-1)
    got = cp.cumsum(d, axis=-1)
    assert got.dtype == want.dtype and got.shape == want.shape
    if np.dtype(dt).kind == 'f':
        ref = np.cumsum(a.astype(np.float64), axis=-1)
        eps = {2: 1e-3, 4: 1.2e-7, 8: 2.3e-16}[np.dtype(dt).itemsize]
        acc_eps = 1.2e-7 if np.dtype(dt).itemsize <= 4 else 2.3e-16
        bound = eps * (np.abs(ref) + 1) + 2 * acc_eps * np.cumsum(np.abs(a.astype('f8')), axis=-1)
        assert np.all(np.abs(got.get().astype('f8') - ref) <= bound)
    else:
        np.testing.assert_array_equal(got.get(), want)
        b = (np.abs(a.astype(np.int64)) % 3 - 1 + (a.astype(np.int64) % 7 == 0)).astype(dt)
        np.testing.assert_array_equal(cp.cumprod(cp.asarray(b), axis=-1).get(), np.cumprod(b, axis=-1))
    if dt in ('float32', 'int64'):
        e = cp.asarray(a)
        assert cp.cumsum(e, axis=-1, out=e) is e
        np.testing.assert_array_equal(e.get(), got.get())
        o = cp.empty(shape, np.float64)
        cp.cumsum(d, axis=-1, out=o)
        np.testing.assert_allclose(o.get(), np.cumsum(a.astype('f8'), axis=-1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('dt', ['float32', 'float16', 'float64', 'int32', 'int64', 'int8', 'uint8', 'bool'])
@pytest.mark.parametrize('shape', [(100000, 3), (70001, 4), (50000, 7), (33333, 16), (20011, 33), (8193, 64), (300007, 2),
                                   (1 << 20, 5), (17, 4000, 6), (1100, 60), (40000, 32)])
def test_reductions_along_short_rows(cp, shape, dt):
    """sum / max / argmax / mean / var ... along the contiguous axis when rows have at most 64 elements
    (reduce_short_rows_body: whole rows staged in shared memory, a thread per row)."""
    a = _data(shape, dt)
    d = cp.asarray(a)
    n = shape[-1]
    kind = np.dtype(dt).kind
    want = a.sum(axis=-1)
    got = d.sum(axis=-1).get()
    assert got.dtype == want.dtype and got.shape == want.shape
    if kind == 'f':
        tol = {2: 2e-3, 4: 1e-6, 8: 1e-14}[np.dtype(dt).itemsize]
        ref = a.astype(np.float64).sum(axis=-1)
        np.testing.assert_allclose(got.astype('f8'), ref, rtol=tol, atol=tol * n)
        np.testing.assert_allclose(d.mean(axis=-1).get().astype('f8'), ref / n, rtol=tol, atol=tol)
        np.testing.assert_allclose(d.var(axis=-1).get().astype('f8'), a.astype('f8').var(axis=-1), rtol=max(tol * 10, 1e-5), atol=tol)
        np.testing.assert_allclose(d.var(axis=-1, ddof=1).get().astype('f8'), a.astype('f8').var(axis=-1, ddof=1),
                                   rtol=max(tol * 10, 1e-5), atol=tol)
    else:
        np.testing.assert_array_equal(got, want)
        np.testing.assert_allclose(d.mean(axis=-1).get(), a.mean(axis=-1), rtol=1e-12)
        np.testing.assert_allclose(d.var(axis=-1).get(), a.var(axis=-1), rtol=1e-9, atol=1e-12)
    for name in ('max', 'min', 'argmax', 'argmin', 'any', 'all'):
        np.testing.assert_array_equal(getattr(d, name)(axis=-1).get(), getattr(a, name)(axis=-1), err_msg=name)
    np.testing.assert_array_equal(d.prod(axis=-1).get() if kind != 'f' else 0, a.prod(axis=-1) if kind != 'f' else 0)
    np.testing.assert_array_equal(d.max(axis=-1, keepdims=True).get(), a.max(axis=-1, keepdims=True))
    if kind == 'f':
        b = a.copy()
        b[::7, n // 2] = np.nan
        b[3::11, 0] = np.nan
        b[5::13, n - 1] = np.nan
        e = cp.asarray(b)
        np.testing.assert_array_equal(e.argmax(axis=-1).get(), b.argmax(axis=-1))
        np.testing.assert_array_equal(e.argmin(axis=-1).get(), b.argmin(axis=-1))
        np.testing.assert_array_equal(e.max(axis=-1).get(), b.max(axis=-1))
        c = np.round(a.astype('f8') * 2).astype(dt)            # few distinct values: ties, first occurrence wins
        np.testing.assert_array_equal(cp.asarray(c).argmax(axis=-1).get(), c.argmax(axis=-1))
