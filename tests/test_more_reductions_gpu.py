"""GPU tier: all / any / count_nonzero / nansum / nanprod / nanmin / nanmax / nanarg* / ptp against
NumPy (the reference's oracle: tests/cupy_tests/logic_tests/test_truth.py,
math_tests/test_sumprod.py `TestNansumNanprod*`, statistics_tests/test_order.py,
sorting_tests/test_search.py `TestNanArgMin/Max`)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


RS = np.random.RandomState(9)


def rnd(shape, dt, nan_frac=0.0):
    dt = np.dtype(dt)
    if dt.kind == 'f':
        a = (RS.rand(*shape) * 2 - 1).astype(dt)
        if nan_frac:
            a[RS.rand(*shape) < nan_frac] = np.nan
        return a
    if dt.kind == 'b':
        return RS.rand(*shape) > 0.3
    return RS.randint(-3, 4, size=shape).astype(dt)


AXES = [None, 0, 1, 2, (0, 2)]


@pytest.mark.parametrize('dt', ['bool', 'int8', 'int32', 'int64', 'float16', 'float32', 'float64'])
@pytest.mark.parametrize('axis', AXES)
def test_all_any_count_nonzero(cp, dt, axis):
    a = rnd((7, 65, 33), dt)
    a[:, 3] = 1                                # an all-true line and an all-false line
    a[:, 5] = 0
    d = cp.asarray(a)
    for name in ('all', 'any'):
        got, want = getattr(cp, name)(d, axis=axis).get(), getattr(np, name)(a, axis=axis)
        assert got.dtype == np.bool_
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(getattr(d, name)(axis=axis, keepdims=True).get(), getattr(a, name)(axis=axis, keepdims=True))
    got = cp.count_nonzero(d, axis=axis).get()
    np.testing.assert_array_equal(got, np.count_nonzero(a, axis=axis))
    assert got.dtype == np.int64


@pytest.mark.parametrize('dt,rtol', [('float32', 1e-5), ('float64', 1e-12), ('float16', 1e-2)])
@pytest.mark.parametrize('axis', AXES)
def test_nansum_nanprod(cp, dt, rtol, axis):
    a = rnd((6, 40, 50), dt, nan_frac=0.1)
    d = cp.asarray(a)
    want = np.nansum(a.astype(np.float64), axis=axis)
    np.testing.assert_allclose(cp.nansum(d, axis=axis).get(), want, rtol=rtol, atol=rtol * 30)
    p = (1 + rnd((4, 9, 10), dt, nan_frac=0.2) / 4).astype(dt)
    np.testing.assert_allclose(cp.nanprod(cp.asarray(p), axis=axis).get(), np.nanprod(p.astype(np.float64), axis=axis),
                               rtol=max(rtol * 10, 1e-11))
    i = rnd((5, 6, 7), 'int32')
    np.testing.assert_array_equal(cp.nansum(cp.asarray(i), axis=axis).get(), np.nansum(i, axis=axis))


@pytest.mark.parametrize('dt', ['float32', 'float64', 'float16', 'int32'])
@pytest.mark.parametrize('axis', [None, 0, 1, 2])
def test_nanmin_nanmax_nanarg_ptp(cp, dt, axis):
    a = rnd((5, 300, 40), dt, nan_frac=0.15)
    clean = rnd((5, 300, 40), dt)
    a[0], a[:, 0], a[:, :, 0] = clean[0], clean[:, 0], clean[:, :, 0]      # no all-NaN slice along any axis
    d = cp.asarray(a)
    np.testing.assert_array_equal(cp.nanmin(d, axis=axis).get(), np.nanmin(a, axis=axis))
    np.testing.assert_array_equal(cp.nanmax(d, axis=axis).get(), np.nanmax(a, axis=axis))
    np.testing.assert_array_equal(cp.nanargmin(d, axis=axis).get(), np.nanargmin(a, axis=axis))
    np.testing.assert_array_equal(cp.nanargmax(d, axis=axis).get(), np.nanargmax(a, axis=axis))
    b = rnd((5, 300, 40), dt)                     # ptp propagates NaN like max - min: test on clean data
    got, want = cp.ptp(cp.asarray(b), axis=axis).get(), np.ptp(b, axis=axis)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(cp.asarray(b).ptp(axis=axis, keepdims=True).get(), np.ptp(b, axis=axis, keepdims=True))


def test_all_nan_slices(cp):
    a = rnd((4, 50), 'float32')
    a[2] = np.nan
    d = cp.asarray(a)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.testing.assert_array_equal(cp.nanmax(d, axis=1).get(), np.nanmax(a, axis=1))    # NaN for the all-NaN row
        np.testing.assert_array_equal(cp.nanmin(d, axis=1).get(), np.nanmin(a, axis=1))
    with pytest.raises(ValueError):
        cp.nanargmax(d, axis=1)
    with pytest.raises(ValueError):
        cp.nanargmin(cp.asarray(np.full(5, np.nan, np.float32)))
    np.testing.assert_array_equal(cp.nanargmax(d, axis=0).get(), np.nanargmax(a, axis=0))
    np.testing.assert_allclose(cp.nansum(d, axis=1).get(), np.nansum(a, axis=1), rtol=1e-5, atol=1e-5)
