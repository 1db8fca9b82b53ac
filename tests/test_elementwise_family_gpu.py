"""GPU tier: the rest of the elementwise family (cupy_b200/_core/_routines_elementwise.py) and the
NaN-ignoring moments against NumPy -- the reference's own oracle for these functions
(tests/cupy_tests/math_tests/test_{trigonometric,hyperbolic,explog,rounding,arithmetic,floating,misc,
rational}.py, logic_tests/test_{ops,content,comparison}.py: `numpy_cupy_allclose` / `numpy_cupy_array_equal`).

Bars: bit-exact for integer, boolean, rounding and sign/exponent work; <= 2 ulp for transcendentals
(<= 4 for the inverse hyperbolics, CUDA's documented bound); float16 results within 1 float16 ulp of the
float32-computed, float16-rounded answer."""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu

RS = np.random.RandomState(11)
SHAPE = (67, 129)
FLOATS = ('float16', 'float32', 'float64')
INTS = ('int8', 'uint8', 'int16', 'int32', 'uint32', 'int64', 'uint64')


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


def uniform(lo, hi, dt, shape=SHAPE):
    return (RS.rand(*shape) * (hi - lo) + lo).astype(dt)


def ints(dt, lo=-60, hi=60, shape=SHAPE, nonzero=False):
    dt = np.dtype(dt)
    if dt.kind == 'u':
        lo = max(lo, 0)
    a = RS.randint(lo, hi, size=shape)
    if nonzero:
        a[a == 0] = 3
    return a.astype(dt)


def check_ulp(got, want, ulps, msg=''):
    assert got.dtype == want.dtype and got.shape == want.shape, (msg, got.dtype, want.dtype)
    worst = oracle.ulp_diff(got, want).max()
    assert worst <= ulps, '%s: %d ulp > %d' % (msg, worst, ulps)


# ---------------------------------------------------------------------------------------------
# float -> float functions
# ---------------------------------------------------------------------------------------------
UNARY = [  # name, domain, ulp bound (float32 / float64)
    ('arcsin', (-1, 1), 2), ('arccos', (-1, 1), 2), ('arctan', (-20, 20), 2),
    ('arcsinh', (-20, 20), 4), ('arccosh', (1, 40), 4), ('arctanh', (-0.99, 0.99), 4),
    ('cbrt', (-30, 30), 2), ('deg2rad', (-720, 720), 1), ('rad2deg', (-7, 7), 1),
    ('fabs', (-5, 5), 0), ('rint', (-40, 40), 0), ('floor', (-40, 40), 0), ('ceil', (-40, 40), 0),
    ('trunc', (-40, 40), 0), ('fix', (-40, 40), 0), ('reciprocal', (0.5, 9), 0), ('positive', (-5, 5), 0),
]


@pytest.mark.parametrize('dt', FLOATS)
@pytest.mark.parametrize('name,dom,ulps', UNARY, ids=[u[0] for u in UNARY])
def test_unary_float(cp, name, dom, ulps, dt):
    a = uniform(dom[0], dom[1], dt)
    if ulps == 0 and name not in ('reciprocal',):
        a.ravel()[:6] = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5], dtype=dt)       # ties
    got = getattr(cp, name)(cp.asarray(a)).get()
    if ulps == 0:
        np.testing.assert_array_equal(got, getattr(np, name)(a), err_msg=name)
        assert got.dtype == a.dtype
        return
    want = getattr(np, name)(a.astype(np.float64)).astype(dt)
    check_ulp(got, want, 1 if dt == 'float16' else ulps, name)


@pytest.mark.parametrize('dt', FLOATS)
def test_binary_float(cp, dt):
    a, b = uniform(-8, 8, dt), uniform(-8, 8, dt)
    da, db = cp.asarray(a), cp.asarray(b)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    # max(a, b) + log1p(exp(-|a - b|)): the rounding error scales with the larger operand, not with the result
    bound = 4 * np.finfo(dt).eps * np.maximum(np.maximum(np.abs(a64), np.abs(b64)), 1)
    for name in ('logaddexp', 'logaddexp2'):
        got = getattr(cp, name)(da, db).get()
        assert got.dtype == np.dtype(dt)
        assert (np.abs(got.astype(np.float64) - getattr(np, name)(a64, b64)) <= bound).all(), name
    np.testing.assert_array_equal(cp.copysign(da, db).get(), np.copysign(a, b))
    np.testing.assert_array_equal(cp.fmax(da, db).get(), np.fmax(a, b))
    np.testing.assert_array_equal(cp.fmin(da, db).get(), np.fmin(a, b))
    big, to = uniform(1, 100, dt), uniform(-100, 200, dt)
    np.testing.assert_array_equal(cp.nextafter(cp.asarray(big), cp.asarray(to)).get(), np.nextafter(big, to))
    h = a.copy()
    h.ravel()[:3] = 0
    np.testing.assert_array_equal(cp.heaviside(cp.asarray(h), db).get(), np.heaviside(h, b))
    if dt == 'float64':
        p = uniform(0.1, 9, dt)
        check_ulp(cp.float_power(cp.asarray(p), db).get(), np.float_power(p, b), 2, 'float_power')
    # equal operands (the branch that adds log 2), infinities of one sign
    s = np.array([0.0, 3.0, -np.inf, np.inf], dtype=dt)
    with np.errstate(invalid='ignore'):
        check_ulp(cp.logaddexp(cp.asarray(s), cp.asarray(s)).get(), np.logaddexp(s, s), 1, 'logaddexp equal')
        check_ulp(cp.logaddexp2(cp.asarray(s), cp.asarray(s)).get(), np.logaddexp2(s, s), 1, 'logaddexp2 equal')


@pytest.mark.parametrize('dt', FLOATS)
def test_special_values(cp, dt):
    s = np.array([0.0, -0.0, 1.5, -2.5, np.inf, -np.inf, np.nan, 7.0], dtype=dt)
    d = cp.asarray(s)
    for name in ('isnan', 'isinf', 'isfinite', 'signbit', 'isposinf', 'isneginf'):
        got = getattr(cp, name)(d).get()
        assert got.dtype == np.bool_
        np.testing.assert_array_equal(got, getattr(np, name)(s), err_msg=name)
    got = cp.sign(d).get()
    want = np.sign(s)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(np.signbit(got), np.signbit(want))                 # sign(-0.) is +0.
    o = np.full_like(s, 0.5)
    np.testing.assert_array_equal(cp.fmax(d, cp.asarray(o)).get(), np.fmax(s, o))     # a NaN operand loses
    np.testing.assert_array_equal(cp.fmin(d, cp.asarray(o)).get(), np.fmin(s, o))
    np.testing.assert_array_equal(cp.heaviside(d, cp.asarray(o)).get(), np.heaviside(s, o))
    np.testing.assert_array_equal(cp.nan_to_num(d).get(), np.nan_to_num(s))
    np.testing.assert_array_equal(cp.nan_to_num(d, nan=-1, posinf=5, neginf=-5).get(),
                                  np.nan_to_num(s, nan=-1, posinf=5, neginf=-5))
    c = d.copy()
    assert cp.nan_to_num(c, copy=False) is c
    np.testing.assert_array_equal(c.get(), np.nan_to_num(s))
    np.testing.assert_array_equal(cp.logical_not(d).get(), np.logical_not(s))         # NaN is truthy
    z = np.array([0.0, -0.0, 0.0, -0.0], dtype=dt)
    y = np.array([1.0, 1.0, -1.0, -1.0], dtype=dt)
    if dt == 'float16':                                                              # float16 subnormals survive -ftz
        np.testing.assert_array_equal(cp.nextafter(cp.asarray(z), cp.asarray(y)).get(), np.nextafter(z, y))


@pytest.mark.parametrize('dt', FLOATS)
def test_modf_frexp_ldexp(cp, dt):
    a = uniform(-300, 300, dt)
    d = cp.asarray(a)
    f, i = cp.modf(d)
    wf, wi = np.modf(a)
    np.testing.assert_array_equal(f.get(), wf)
    np.testing.assert_array_equal(i.get(), wi)
    m, e = cp.frexp(d)
    wm, we = np.frexp(a)
    assert m.dtype == a.dtype and e.dtype == np.int32
    np.testing.assert_array_equal(m.get(), wm)
    np.testing.assert_array_equal(e.get(), we)
    small = uniform(-2, 2, dt)
    for et in ('int32', 'int64'):
        k = ints(et, -6, 6)
        got = cp.ldexp(cp.asarray(small), cp.asarray(k)).get()
        assert got.dtype == small.dtype
        np.testing.assert_array_equal(got, np.ldexp(small, k))
    np.testing.assert_array_equal(cp.ldexp(cp.asarray(small), 3).get(), np.ldexp(small, 3))


# ---------------------------------------------------------------------------------------------
# division family
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', INTS)
def test_integer_division_family(cp, dt):
    a, b = ints(dt), ints(dt, nonzero=True)
    da, db = cp.asarray(a), cp.asarray(b)
    for name in ('floor_divide', 'remainder', 'fmod', 'gcd'):
        got = getattr(cp, name)(da, db).get()
        want = getattr(np, name)(a, b)
        assert got.dtype == want.dtype, name
        np.testing.assert_array_equal(got, want, err_msg=name)
    s, t = ints(dt, -11, 12), ints(dt, -11, 12)                    # products stay inside int8
    np.testing.assert_array_equal(cp.lcm(cp.asarray(s), cp.asarray(t)).get(), np.lcm(s, t))
    q, r = cp.divmod(da, db)
    wq, wr = np.divmod(a, b)
    np.testing.assert_array_equal(q.get(), wq)
    np.testing.assert_array_equal(r.get(), wr)
    np.testing.assert_array_equal((da // db).get(), a // b)
    np.testing.assert_array_equal((da % db).get(), a % b)
    np.testing.assert_array_equal((da // 7).get(), a // 7)
    np.testing.assert_array_equal((50 % db).get(), 50 % b)
    q2, r2 = divmod(da, db)
    np.testing.assert_array_equal(q2.get(), wq)
    np.testing.assert_array_equal(r2.get(), wr)
    c = da.copy()
    c //= db
    np.testing.assert_array_equal(c.get(), a // b)
    c = da.copy()
    c %= db
    np.testing.assert_array_equal(c.get(), a % b)
    # zero divisors: 0 everywhere, as NumPy (which warns) and the reference
    z = np.zeros_like(b)
    with np.errstate(divide='ignore'):
        for name in ('floor_divide', 'remainder', 'fmod'):
            np.testing.assert_array_equal(getattr(cp, name)(da, cp.asarray(z)).get(), getattr(np, name)(a, z), err_msg=name)
    nz = ints(dt, nonzero=True)
    np.testing.assert_array_equal(cp.reciprocal(cp.asarray(nz)).get(), np.reciprocal(nz))
    np.testing.assert_array_equal(cp.sign(da).get(), np.sign(a))
    np.testing.assert_array_equal(cp.positive(da).get(), np.positive(a))


@pytest.mark.parametrize('dt', FLOATS)
def test_float_division_family(cp, dt):
    # integer-valued and quarter-valued operands: quotients are exact or far from an integer, so
    # floor(x / y) -- the reference's definition -- and NumPy's fmod-based one agree bit for bit
    a = (ints('int32', -50, 50) / 4).astype(dt)
    b = ints('int32', -9, 9, nonzero=True).astype(dt)
    da, db = cp.asarray(a), cp.asarray(b)
    for name in ('floor_divide', 'remainder', 'fmod'):
        got = getattr(cp, name)(da, db).get()
        want = getattr(np, name)(a, b)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want, err_msg=name)
    q, r = cp.divmod(da, db)
    wq, wr = np.divmod(a, b)
    np.testing.assert_array_equal(q.get(), wq)
    np.testing.assert_array_equal(r.get(), wr)
    np.testing.assert_array_equal((da // db).get(), a // b)
    np.testing.assert_array_equal((da % 3).get(), a % 3)


def test_mixed_dtypes_and_layouts(cp):
    a, b = ints('int32'), ints('int64', nonzero=True)
    np.testing.assert_array_equal((cp.asarray(a) // cp.asarray(b)).get(), a // b)
    f = (ints('int32', -40, 40) / 2).astype('float32')
    np.testing.assert_array_equal((cp.asarray(a) % cp.asarray(np.abs(f) + 1)).get(), a % (np.abs(f) + 1))
    # transposed, strided and broadcast operands go through the same tilers as any ufunc
    np.testing.assert_array_equal(cp.floor(cp.asarray(f).T[::2]).get(), np.floor(f.T[::2]))
    np.testing.assert_array_equal(cp.fmax(cp.asarray(f), cp.asarray(f[:1])).get(), np.fmax(f, f[:1]))
    np.testing.assert_array_equal(cp.sign(cp.asarray(f)[:, ::3]).get(), np.sign(f[:, ::3]))
    m = RS.rand(*SHAPE) > 0.5
    out = cp.asarray(f.copy())
    cp.floor_divide(cp.asarray(f), 3, out=out, _where=cp.asarray(m))
    np.testing.assert_array_equal(out.get(), np.where(m, f // 3, f))
    # NumPy's dispatch protocol reaches the same kernels
    got = np.floor(cp.asarray(f))
    assert isinstance(got, cp.ndarray)
    np.testing.assert_array_equal(got.get(), np.floor(f))
    np.testing.assert_array_equal(np.remainder(cp.asarray(a), 7).get(), np.remainder(a, 7))
    # integer input to a float-only function promotes as in NumPy
    s8 = ints('int8', -1, 2)
    got = cp.arcsin(cp.asarray(s8)).get()
    want = np.arcsin(s8)
    assert got.dtype == want.dtype == np.float16
    check_ulp(got, want, 1, 'arcsin(int8)')
    got = cp.arctan(cp.asarray(a)).get()
    assert got.dtype == np.float64
    check_ulp(got, np.arctan(a), 2, 'arctan(int32)')
    for name in ('floor', 'ceil', 'trunc', 'fix'):
        got = getattr(cp, name)(cp.asarray(a)).get()
        assert got.dtype == a.dtype
        np.testing.assert_array_equal(got, a)
    with pytest.raises(TypeError):
        cp.positive(cp.asarray(m))
    with pytest.raises(TypeError):
        cp.gcd(cp.asarray(m), cp.asarray(m))


# ---------------------------------------------------------------------------------------------
# rounding to decimals, clip, where
# ---------------------------------------------------------------------------------------------
def test_around(cp):
    a = uniform(-100, 100, 'float64')
    d = cp.asarray(a)
    for dec in (0, 1, 2, 5, -1):
        np.testing.assert_array_equal(cp.around(d, dec).get(), np.around(a, dec), err_msg=str(dec))
    np.testing.assert_array_equal(d.round(3).get(), a.round(3))
    e = (ints('int32', -800, 800) / 8).astype('float32')                             # exact in float32 after * 10^k
    for dec in (0, 1, 2):
        got = cp.around(cp.asarray(e), dec).get()
        assert got.dtype == np.float32
        check_ulp(got, np.around(e.astype(np.float64), dec).astype(np.float32), 1, 'around f32 %d' % dec)
    h = (ints('int32', -80, 80) / 4).astype('float16')
    np.testing.assert_array_equal(cp.around(cp.asarray(h)).get(), np.around(h))
    i = np.array([15, 25, 35, -15, -25, 149, 151, 250, 1250, 1350, -1250, 99999, 0, 7], dtype='int64')
    di = cp.asarray(i)
    for dec in (-1, -2, -3, 0, 2):
        got = cp.around(di, dec).get()
        assert got.dtype == i.dtype
        np.testing.assert_array_equal(got, np.around(i, dec), err_msg=str(dec))
    for dt in ('int8', 'uint8', 'int16', 'int32', 'uint32'):
        j = ints(dt, -120, 120)
        np.testing.assert_array_equal(cp.around(cp.asarray(j), -1).get(), np.around(j, -1), err_msg=dt)


@pytest.mark.parametrize('dt', ['int8', 'uint8', 'int32', 'int64', 'uint64', 'float16', 'float32', 'float64'])
def test_clip(cp, dt):
    a = ints(dt, -100, 100) if np.dtype(dt).kind != 'f' else uniform(-100, 100, dt)
    d = cp.asarray(a)
    lo, hi = (10, 60) if np.dtype(dt).kind == 'u' else (-30, 40)
    for args in ((lo, hi), (None, hi), (lo, None), (hi, lo)):
        got = cp.clip(d, *args).get()
        want = np.clip(a, *args)
        assert got.dtype == want.dtype
        np.testing.assert_array_equal(got, want, err_msg=str(args))
    np.testing.assert_array_equal(d.clip(lo, hi).get(), a.clip(lo, hi))
    lo_a = np.full(SHAPE[1], lo, dtype=dt)
    np.testing.assert_array_equal(cp.clip(d, cp.asarray(lo_a), hi).get(), np.clip(a, lo_a, hi))
    out = cp.empty_like(d)
    assert cp.clip(d, lo, hi, out=out) is out
    np.testing.assert_array_equal(out.get(), np.clip(a, lo, hi))


@pytest.mark.parametrize('dt', ['?', 'int8', 'int32', 'uint64', 'float16', 'float32', 'float64'])
def test_where(cp, dt):
    m = RS.rand(*SHAPE) > 0.5
    x = ints(dt) if np.dtype(dt).kind in 'iu' else (RS.rand(*SHAPE) > 0.3) if dt == '?' else uniform(-5, 5, dt)
    y = x[::-1].copy()
    got = cp.where(cp.asarray(m), cp.asarray(x), cp.asarray(y)).get()
    assert got.dtype == x.dtype
    np.testing.assert_array_equal(got, np.where(m, x, y))
    np.testing.assert_array_equal(cp.where(cp.asarray(m[:, :1]), cp.asarray(x), cp.asarray(y[:1])).get(),
                                  np.where(m[:, :1], x, y[:1]))
    if dt != '?':
        np.testing.assert_array_equal(cp.where(cp.asarray(x), cp.asarray(x), cp.asarray(y)).get(), np.where(x, x, y))
    with pytest.raises(ValueError):
        cp.where(cp.asarray(m), cp.asarray(x))


@pytest.mark.parametrize('n', [1, 7, 8, 9, 4095, 100003, (1 << 20) + 5])
def test_mixed_item_sizes_any_alignment(cp, n):
    """Loops over operands of different item sizes widen their vector (tunables['flat_mixed_vec']) only while
    every operand keeps the alignment its accesses need: views starting at odd offsets, ragged tails."""
    x = uniform(-4, 4, 'float32', (n + 16,))
    y = uniform(-4, 4, 'float32', (n + 16,))
    m = RS.rand(n + 16) > 0.5
    h = uniform(-4, 4, 'float16', (n + 16,))
    x[::5] = np.nan
    dx, dy, dm, dh = cp.asarray(x), cp.asarray(y), cp.asarray(m), cp.asarray(h)
    for ox, om in ((0, 0), (1, 0), (0, 1), (4, 4), (0, 8), (3, 5), (8, 16), (2, 2)):
        sx, sy, sm, sh = slice(ox, ox + n), slice(ox, ox + n), slice(om, om + n), slice(om, om + n)
        np.testing.assert_array_equal(cp.where(dm[sm], dx[sx], dy[sy]).get(), np.where(m[sm], x[sx], y[sy]))
        np.testing.assert_array_equal(cp.isnan(dx[sx]).get(), np.isnan(x[sx]))
        np.testing.assert_array_equal(cp.greater(dx[sx], dy[sy]).get(), x[sx] > y[sy])
        np.testing.assert_array_equal(cp.add(dx[sx], dh[sh]).get(), x[sx] + h[sh])            # float32 + float16
        out = cp.asarray(np.zeros(n + 16, dtype=bool))
        cp.less(dx[sx], 0, out=out[sm])
        want = np.zeros(n + 16, dtype=bool)
        want[sm] = x[sx] < 0
        np.testing.assert_array_equal(out.get(), want)                                        # nothing outside the view


# ---------------------------------------------------------------------------------------------
# logic
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', ['?', 'int8', 'uint8', 'int32', 'int64', 'uint64', 'float16', 'float32', 'float64'])
def test_logical_ops(cp, dt):
    kind = np.dtype(dt).kind
    a = (RS.rand(*SHAPE) > 0.5) if kind == 'b' else ints(dt, -2, 3) if kind in 'iu' else ints('int32', -2, 3).astype(dt)
    b = a[::-1].copy()
    da, db = cp.asarray(a), cp.asarray(b)
    for name in ('logical_and', 'logical_or', 'logical_xor'):
        got = getattr(cp, name)(da, db).get()
        assert got.dtype == np.bool_
        np.testing.assert_array_equal(got, getattr(np, name)(a, b), err_msg=name)
    np.testing.assert_array_equal(cp.logical_not(da).get(), np.logical_not(a))
    np.testing.assert_array_equal(np.logical_and(da, db).get(), np.logical_and(a, b))


@pytest.mark.parametrize('dt', FLOATS)
def test_isclose_allclose_array_equal(cp, dt):
    a = uniform(1, 10, dt)
    b = a.copy()
    flat = b.ravel()
    flat[::3] *= np.dtype(dt).type(1.01)                 # clearly apart
    flat[5] = np.nan
    a.ravel()[5] = np.nan
    a.ravel()[7] = b.ravel()[7] = np.inf
    a.ravel()[9], b.ravel()[9] = np.inf, -np.inf
    da, db = cp.asarray(a), cp.asarray(b)
    rtol = 1e-3 if dt == 'float16' else 1e-5
    for eq in (False, True):
        got = cp.isclose(da, db, rtol=rtol, equal_nan=eq).get()
        np.testing.assert_array_equal(got, np.isclose(a, b, rtol=rtol, equal_nan=eq))
    assert cp.allclose(da, db).get() == np.allclose(a, b)
    assert bool(cp.allclose(da, da, equal_nan=True).get())
    assert not bool(cp.allclose(da, da).get())           # a NaN is not close to itself
    fin = cp.asarray(uniform(1, 2, dt))
    assert bool(cp.allclose(fin, fin).get()) and bool(cp.array_equal(fin, fin).get())
    assert not bool(cp.array_equal(da, da).get())
    assert bool(cp.array_equal(da, da, equal_nan=True).get())
    assert not bool(cp.array_equal(da, db, equal_nan=True).get())
    assert not bool(cp.array_equal(da, da[:5]).get())
    i = cp.asarray(ints('int32'))
    assert bool(cp.array_equal(i, i).get()) and bool(cp.allclose(i, i).get())
    np.testing.assert_array_equal(cp.isclose(i, i + 1, atol=1).get(), np.ones(SHAPE, bool))


# ---------------------------------------------------------------------------------------------
# NaN-ignoring moments
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dt', FLOATS)
@pytest.mark.parametrize('shape,axis', [((5000,), None), ((300, 257), 0), ((300, 257), 1), ((9, 33, 20), (0, 2)),
                                        ((300, 257), None)])
def test_nan_moments(cp, dt, shape, axis):
    a = uniform(-1, 1, dt, shape)
    a[RS.rand(*shape) < 0.2] = np.nan
    d = cp.asarray(a)
    a64 = a.astype(np.float64)
    tol = {'float16': 2e-3, 'float32': 2e-6, 'float64': 1e-13}[dt]
    got = cp.nanmean(d, axis=axis).get()
    assert got.dtype == np.dtype(dt)
    np.testing.assert_allclose(got.astype(np.float64), np.nanmean(a64, axis=axis), rtol=0, atol=tol)
    for ddof in (0, 1):
        got = cp.nanvar(d, axis=axis, ddof=ddof).get()
        assert got.dtype == np.dtype(dt)
        np.testing.assert_allclose(got.astype(np.float64), np.nanvar(a64, axis=axis, ddof=ddof), rtol=tol * 10, atol=tol)
    got = cp.nanstd(d, axis=axis, keepdims=True).get()
    want = np.nanstd(a64, axis=axis, keepdims=True)
    assert got.shape == want.shape
    np.testing.assert_allclose(got.astype(np.float64), want, rtol=tol * 10, atol=tol)


def test_nan_moments_edge_cases(cp):
    a = np.array([[np.nan, np.nan, np.nan], [1.0, np.nan, 3.0]], dtype='float32')
    d = cp.asarray(a)
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            np.testing.assert_array_equal(cp.nanmean(d, axis=1).get(), np.nanmean(a, axis=1))       # all-NaN row -> NaN
            np.testing.assert_array_equal(cp.nanvar(d, axis=1).get(), np.nanvar(a, axis=1))
    i = np.arange(12, dtype='int32').reshape(3, 4)
    np.testing.assert_array_equal(cp.nanmean(cp.asarray(i), axis=0).get(), np.nanmean(i, axis=0))
    np.testing.assert_allclose(cp.nanvar(cp.asarray(i), axis=1).get(), np.nanvar(i, axis=1), rtol=1e-14)
    out = cp.empty((3,), 'float64')
    f = uniform(-1, 1, 'float32', (40, 3))
    f[::4] = np.nan
    cp.nanvar(cp.asarray(f), axis=0, out=out)
    np.testing.assert_allclose(out.get(), np.nanvar(f.astype(np.float64), axis=0), rtol=1e-5)


@pytest.mark.parametrize('dt', FLOATS)
def test_nan_scans(cp, dt):
    a = uniform(0.5, 1.5, dt, (40, 50))
    a[RS.rand(40, 50) < 0.2] = np.nan
    d = cp.asarray(a)
    a64 = a.astype(np.float64)
    tol = {'float16': 2e-2, 'float32': 1e-5, 'float64': 1e-13}[dt]
    for axis in (None, 0, 1):
        got = cp.nancumsum(d, axis=axis).get()
        assert got.dtype == np.dtype(dt)
        np.testing.assert_allclose(got.astype(np.float64), np.nancumsum(a64, axis=axis), rtol=tol)
    small = cp.asarray(a[:4, :8])
    np.testing.assert_allclose(cp.nancumprod(small, axis=1).get().astype(np.float64), np.nancumprod(a64[:4, :8], axis=1), rtol=tol)
    out = cp.empty_like(d)
    assert cp.nancumsum(d, axis=1, out=out) is out
    np.testing.assert_allclose(out.get().astype(np.float64), np.nancumsum(a64, axis=1), rtol=tol)
    i = ints('int32', -5, 5, (40, 50))
    np.testing.assert_array_equal(cp.nancumsum(cp.asarray(i), axis=0).get(), np.nancumsum(i, axis=0))


def test_average(cp):
    a = uniform(-1, 1, 'float32', (30, 20))
    w1 = uniform(0.1, 2, 'float64', (20,))
    w2 = uniform(0.1, 2, 'float32', (30, 20))
    d = cp.asarray(a)
    np.testing.assert_allclose(cp.average(d).get(), np.average(a), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cp.average(d, axis=0).get(), np.average(a, axis=0), rtol=1e-5, atol=1e-7)
    got, scl = cp.average(d, axis=1, weights=cp.asarray(w1), returned=True)
    want, wscl = np.average(a, axis=1, weights=w1, returned=True)
    assert got.dtype == want.dtype == np.float64 and scl.shape == wscl.shape
    np.testing.assert_allclose(got.get(), want, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(scl.get(), wscl, rtol=1e-12)
    got = cp.average(d, weights=cp.asarray(w2), axis=0, keepdims=True)
    want = np.average(a, weights=w2, axis=0, keepdims=True)
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_allclose(got.get(), want, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(cp.average(d, weights=cp.asarray(w2)).get(), np.average(a, weights=w2), rtol=1e-4, atol=1e-6)
    i = ints('int32', 1, 9, (30, 20))
    got = cp.average(cp.asarray(i), weights=cp.asarray(i), axis=1)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got.get(), np.average(i, weights=i, axis=1), rtol=1e-13)
    got, scl = cp.average(d, axis=0, returned=True)
    want, wscl = np.average(a, axis=0, returned=True)
    np.testing.assert_allclose(scl.get(), np.broadcast_to(wscl, want.shape))
    with pytest.raises(ZeroDivisionError):
        cp.average(d, weights=cp.asarray(np.zeros((30, 20), 'float32')))
    with pytest.raises(TypeError):
        cp.average(d, weights=cp.asarray(w1))


# ---------------------------------------------------------------------------------------------
# user kernels can call the same device helpers the routine strings use
# ---------------------------------------------------------------------------------------------
def test_user_kernel_helpers(cp):
    a, b = ints('int32'), ints('int32', nonzero=True)
    k = cp.ElementwiseKernel('T x, T y', 'T q, float64 c', 'q = _floor_divide(x, y); c = M_PI * x', 'fd_pi')
    q, c = k(cp.asarray(a), cp.asarray(b))
    np.testing.assert_array_equal(q.get(), a // b)
    np.testing.assert_array_equal(c.get(), np.pi * a)


def test_fused_chain_of_new_ufuncs(cp):
    x = uniform(-4, 4, 'float32')

    @cp.fuse()
    def f(v):
        return cp.where(v > 0, cp.floor(v), cp.ceil(v)) + cp.sign(v)

    got = f(cp.asarray(x)).get()
    np.testing.assert_array_equal(got, np.where(x > 0, np.floor(x), np.ceil(x)) + np.sign(x))
