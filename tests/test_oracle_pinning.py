"""CPU tier: pin the oracle (oracle/oracle.py, oracle/oracle.c) against the known-answer
tests of the reference's own suite (tests/ref_cases.py lists each with file:line) and
against independent exact arithmetic.  An unpinned oracle proves nothing."""
import fractions
import math

import numpy as np
import pytest

from oracle import oracle
from tests import ref_cases


def apply_oracle(op, a, kwargs):
    if op == 'argmax':
        return oracle.argmax(a, **kwargs)
    if op == 'argmin':
        return oracle.argmin(a, **kwargs)
    if op == 'max':
        return oracle.amax(a, **kwargs)
    if op == 'min':
        return oracle.amin(a, **kwargs)
    if op == 'sum':
        return oracle.sum(a, **kwargs)
    if op == 'mean':
        return oracle.mean(a, **kwargs)
    if op == 'var':
        return oracle.var(a, **kwargs)
    if op == 'cumsum_same_dtype':
        return oracle.cumsum(a, dtype=a.dtype if a.dtype != np.bool_ else None)
    raise KeyError(op)


@pytest.mark.parametrize('case', ref_cases.KNOWN_ANSWERS, ids=[c[0] for c in ref_cases.KNOWN_ANSWERS])
def test_oracle_reproduces_reference_known_answers(case):
    cid, ref, build, op, kwargs, expect = case
    a = build()
    got = apply_oracle(op, a, kwargs)
    if expect is not None:
        if isinstance(expect, float) and math.isnan(expect):
            assert np.isnan(got).all(), (cid, ref)
        else:
            np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(expect, dtype=np.float64),
                                       rtol=2e-3 if a.dtype == np.float16 else 1e-6, err_msg='%s (%s)' % (cid, ref))
    # NumPy is the reference's oracle for every one of these tests: same values and result dtype
    np_op = {'argmax': np.argmax, 'argmin': np.argmin, 'max': np.max, 'min': np.min, 'sum': np.sum,
             'mean': np.mean, 'var': np.var}.get(op)
    if np_op is not None:
        want = np_op(a, **kwargs)
        assert np.asarray(got).dtype == np.asarray(want).dtype, (cid, np.asarray(got).dtype, np.asarray(want).dtype)
        np.testing.assert_allclose(np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64),
                                   rtol=5e-3 if a.dtype == np.float16 else 1e-6, equal_nan=True)


def _exact_fma_f32(a, x, y):
    """Correctly rounded float32 fma through exact rational arithmetic."""
    r = fractions.Fraction(float(a)) * fractions.Fraction(float(x)) + fractions.Fraction(float(y))
    if r == 0:
        return np.float32(0.0)
    # round-to-nearest-even to 24 significant bits
    lo = np.float32(float(r))                     # double rounding is possible here, so verify the neighbours
    cands = [lo, np.nextafter(lo, np.float32(np.inf)), np.nextafter(lo, np.float32(-np.inf))]
    best = min(cands, key=lambda c: (abs(fractions.Fraction(float(c)) - r), int(c.view(np.uint32)) & 1))
    return np.float32(best)


def test_axpy_oracle_is_single_rounding_fma():
    """oracle.axpy == fmaf (one rounding), which differs from NumPy's two roundings on a
    measurable fraction of inputs -- the reference GPU kernel is FFMA (SURVEY.md probe 2)."""
    rs = np.random.RandomState(0)
    x = (rs.rand(4000) * 2 - 1).astype(np.float32)
    y = (rs.rand(4000) * 2 - 1).astype(np.float32)
    a = np.float32(1.5)
    z = oracle.axpy(a, x, y)
    exact = np.array([_exact_fma_f32(a, xi, yi) for xi, yi in zip(x, y)], np.float32)
    np.testing.assert_array_equal(z, exact)
    two_roundings = a * x + y
    # the two forms differ by at most the rounding of the product (not 1 ulp of the RESULT:
    # under cancellation the result's ulp is much smaller than the product's)
    assert (np.abs(z.astype(np.float64) - two_roundings) <= np.spacing(np.abs(a * x)) + np.spacing(np.abs(z))).all()
    assert (z != two_roundings).any()            # the distinction is real on this data


def test_flush_to_zero_is_restated():
    tiny = np.float32(1e-40)                       # denormal
    assert tiny != 0
    z = oracle.axpy(1.0, np.array([tiny, -tiny, 1.0], np.float32), np.zeros(3, np.float32))
    assert z[0] == 0 and z[1] == 0 and z[2] == 1.0
    assert np.signbit(oracle.ftz32(np.array([-tiny], np.float32))[0])          # flush keeps the sign
    assert oracle.binary_f32('add', np.array([tiny], np.float32), np.array([tiny], np.float32))[0] == 0
    np.testing.assert_array_equal(oracle.ftz32(np.array([tiny, 2.0], np.float32)), np.array([0, 2.0], np.float32))


def test_ulp_diff():
    one = np.float32(1.0)
    nxt = np.nextafter(one, np.float32(2))
    assert oracle.ulp_diff(np.array([one]), np.array([nxt]))[0] == 1
    assert oracle.ulp_diff(np.array([np.float32(-0.0)]), np.array([np.float32(0.0)]))[0] == 0
    assert oracle.ulp_diff(np.array([np.float32(np.nan)]), np.array([np.float32(np.nan)]))[0] == 0
    assert oracle.ulp_diff(np.float16([1.0]), np.float16([1.001]))[0] == 1


def test_dtype_rules_follow_the_reference_tables():
    # cupy/_core/_routines_math.pyx:762-775 (sum), :704-714 (scan), _routines_statistics.pyx:132-146 (mean)
    for dt in ref_cases.ALL_DTYPES:
        a = np.ones((4,), dt)
        assert oracle.sum_dtype(dt) == a.sum().dtype
        assert oracle.scan_dtype(dt) == a.cumsum().dtype
        if dt != '?':
            assert oracle.mean_dtype(dt) == a.mean().dtype
            assert oracle.var(a).dtype == a.var().dtype


def test_float16_sum_accumulates_in_float32():
    """('e->e', (None, None, None, 'float')): 4096 * 0.1 is far beyond float16 serial accumulation."""
    a = np.full(70000, 0.1, np.float16)
    want = np.float16(np.float32(np.float16(0.1)) * 70000)         # inf in float16
    got = oracle.sum(a)
    assert got.dtype == np.float16 and got == want


def test_var_is_the_reference_two_pass_formula():
    rs = np.random.RandomState(1)
    a = (rs.rand(50, 60) * 10).astype(np.float32)
    for axis, ddof in ((None, 0), (0, 1), (1, 0), ((0, 1), 2)):
        np.testing.assert_allclose(oracle.var(a, axis=axis, ddof=ddof),
                                   np.var(a.astype(np.float64), axis=axis, ddof=ddof).astype(np.float32), rtol=1e-6)
    assert np.isnan(oracle.var(np.ones(3, np.float32), ddof=3))    # alpha = NaN when n - ddof <= 0


def test_cumsum_int_is_exact_and_wraps():
    a = np.array([2 ** 62, 2 ** 62, 2 ** 62], np.int64)
    got = oracle.cumsum(a)
    np.testing.assert_array_equal(got, np.array([2 ** 62, -2 ** 63, -2 ** 62], np.int64))
    import ctypes
    out = np.empty_like(a)
    oracle.clib().oracle_cumsum_i64(a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 3)
    np.testing.assert_array_equal(out, got)


# ---------------------------------------------------------------------------------------------
# division / rounding members of the elementwise family: the oracle restates the REFERENCE's routine text
# (which is not always NumPy's algorithm); pinned on the reference's own border cases and on NumPy wherever the
# two definitions must agree -- the domain the GPU tier (tests/test_elementwise_family_gpu.py) draws from
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('value,decimals', [(14, -1), (15, -1), (16, -1), (-14, -1), (-15, -1), (-16, -1),
                                            (25, -1), (35, -1), (149, -2), (150, -2), (250, -2), (99999, -2), (7, -3)])
def test_oracle_around_int_border_cases(value, decimals):
    """tests/cupy_tests/math_tests/test_rounding.py:131-166 `TestRoundBorder` (14 / 15 / 16 at -1, both signs):
    the digit-splitting routine agrees with NumPy's half-to-even."""
    for dt in ('int32', 'int64'):
        a = np.array([value], dt)
        np.testing.assert_array_equal(oracle.around_int(a, decimals), np.around(a, decimals))


def test_oracle_division_family_agrees_with_numpy_on_the_test_domain():
    rs = np.random.RandomState(3)
    for dt in ('int8', 'uint8', 'int16', 'int32', 'uint32', 'int64', 'uint64'):
        lo = 0 if np.dtype(dt).kind == 'u' else -60
        x = rs.randint(lo, 60, size=5000).astype(dt)
        y = rs.randint(lo, 60, size=5000).astype(dt)              # zeros included: both give 0
        with np.errstate(divide='ignore'):
            np.testing.assert_array_equal(oracle.floor_divide(x, y), np.floor_divide(x, y))
            np.testing.assert_array_equal(oracle.remainder(x, y), np.remainder(x, y))
    for dt in ('float16', 'float32', 'float64'):
        x = (rs.randint(-50, 50, size=5000) / 4).astype(dt)       # quarter-valued: quotients exact or far from an integer
        y = rs.randint(1, 9, size=5000).astype(dt) * rs.choice([-1, 1], size=5000).astype(dt)
        np.testing.assert_array_equal(oracle.floor_divide(x, y), np.floor_divide(x, y))
        np.testing.assert_array_equal(oracle.remainder(x, y), np.remainder(x, y))


def test_oracle_records_where_the_reference_leaves_numpy():
    """floor(x / y) against NumPy's fmod-based quotient: when x / y rounds UP to an integer the reference's answer
    is one larger (and its remainder is then about zero instead of about y).  The GPU tier stays off
    this set; the oracle states which side the engine is on."""
    x, y = np.float32(0.7), np.float32(0.1)                      # 0.7f / 0.1f rounds to 7.0f; exactly it is 6.9999998
    assert fractions.Fraction(float(x)) / fractions.Fraction(float(y)) < 7
    assert np.floor_divide(x, y) == 6.0
    assert oracle.floor_divide(x, y) == 7.0
    assert abs(oracle.remainder(x, y)) < 1e-6 and np.remainder(x, y) > 0.09      # 0.7 - 7 * 0.1 against 0.7 - 6 * 0.1
    # sign(-0.) is +0. (x - x), NaN stays NaN; clip with lo > hi is hi everywhere -- both as NumPy
    s = np.array([0.0, -0.0, np.nan, -2.5, 3.0], 'float32')
    np.testing.assert_array_equal(oracle.sign(s), np.sign(s))
    assert not np.signbit(oracle.sign(s)[1])
    for dt in ('int8', 'uint16', 'int64'):
        a = np.arange(0, 100, 7).astype(dt)
        np.testing.assert_array_equal(oracle.sign(a), np.sign(a))
        np.testing.assert_array_equal(oracle.clip(a, 20, 60), np.clip(a, 20, 60))
        np.testing.assert_array_equal(oracle.clip(a, 60, 20), np.clip(a, 60, 20))
