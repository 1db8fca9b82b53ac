"""The rest of the elementwise family around the hot path: every instance here is a `create_ufunc`
table rendered by NVRTC into the FLAT / ROWWISE / TILED tilers of
b200/elementwise.cuh, exactly like `add` or `exp` -- no new device code, only loop tables.

What each table follows in the reference (type loops and routine text are the contract a caller sees:
which dtype comes out, what happens on zero divisors, NaNs, signed zeros):

  trigonometric / hyperbolic   cupy/_math/trigonometric.py:39-105, cupy/_math/hyperbolic.py:33-57
  logaddexp / logaddexp2       cupy/_math/explog.py:73-107
  rint floor ceil trunc fix    cupy/_math/rounding.py:43-118;  around / round  :11-40, cupy/_core/core.pyx:1354-1377, 2728-2814
  reciprocal fmod modf float_power   cupy/_math/arithmetic.py:11-165
  positive floor_divide remainder divmod clip   cupy/_core/_routines_math.pyx:957-966, 1087-1110, 1146-1151; core.pyx:2704-2725
  signbit copysign ldexp frexp nextafter   cupy/_math/floating.py:9-66
  cbrt fabs sign heaviside fmax fmin nan_to_num   cupy/_math/misc.py:149-465
  gcd lcm                      cupy/_math/rational.py:14-63
  logical_and/or/not/xor       cupy/_logic/ops.py:6-46, cupy/_core/_routines_logic.pyx:101-118
  isfinite isinf isnan isneginf isposinf   cupy/_logic/content.py:9-135
  isclose allclose array_equal   cupy/_logic/comparison.py:10-132
  where                        cupy/_sorting/search.py:167-210 (one-argument form: _core/_compaction.py)
"""
from __future__ import annotations

import numpy

from cupy_b200._core import _routines_math as _math
from cupy_b200._core._kernel import create_ufunc
from cupy_b200._core._ndarray import ndarray

_INT = 'bBhHiIlLqQ'
_FLT = 'efd'


def _loops(chars, nin=1, out=None, code=None):
    """('bb->b', ...) for every char; `out` replaces the output char, `code` gives the loops their own routine."""
    sigs = tuple('%s->%s' % (c * nin, out or c) for c in chars)
    return sigs if code is None else tuple((s, code) for s in sigs)


def _float_unary(math_name, name):
    return _math._create_math_ufunc(math_name, 1, name)


# ---- trigonometric / hyperbolic ---------------------------------------------------------------
arcsin = _float_unary('asin', 'cupy_arcsin')
arccos = _float_unary('acos', 'cupy_arccos')
arctan = _float_unary('atan', 'cupy_arctan')
arcsinh = _float_unary('asinh', 'cupy_arcsinh')
arccosh = _float_unary('acosh', 'cupy_arccosh')
arctanh = _float_unary('atanh', 'cupy_arctanh')
deg2rad = create_ufunc('cupy_deg2rad', _loops(_FLT), 'out0 = in0 * (out0_type)(M_PI / 180)')
rad2deg = create_ufunc('cupy_rad2deg', _loops(_FLT), 'out0 = in0 * (out0_type)(180 / M_PI)')
radians, degrees = deg2rad, rad2deg

# ---- exponents and logarithms -----------------------------------------------------------------
logaddexp = create_ufunc(
    'cupy_logaddexp', _loops(_FLT, 2),
    'if (in0 == in1) { out0 = in0 + log(2.0); }'          # equal operands: covers infinities of one sign
    ' else { out0 = fmax(in0, in1) + log1p(exp(-fabs(in0 - in1))); }')
logaddexp2 = create_ufunc(
    'cupy_logaddexp2', _loops(_FLT, 2),
    'if (in0 == in1) { out0 = in0 + 1.0; }'
    ' else { out0 = fmax(in0, in1) + log2(1 + exp2(-fabs(in0 - in1))); }')

# ---- rounding ---------------------------------------------------------------------------------
rint = _float_unary('rint', 'cupy_rint')


def _rounding(name, code):
    # integers and bools round to themselves
    return create_ufunc(name, _loops('?' + _INT) + _loops(_FLT, code=code), 'out0 = in0')


floor = _rounding('cupy_floor', 'out0 = floor(in0)')
ceil = _rounding('cupy_ceil', 'out0 = ceil(in0)')
trunc = _rounding('cupy_trunc', 'out0 = trunc(in0)')
fix = _rounding('cupy_fix', 'out0 = (in0 >= 0.0) ? floor(in0) : ceil(in0)')

_round_preamble = '''
template <typename T> __device__ T pow10(long long n) {
    T x = 1, a = 10;
    for (; n; n >>= 1, a *= a) if (n & 1) x *= a;
    return x;
}
'''
_round_float = ('if (in1 == 0) { out0 = rint(in0); } else {'
                ' double x = pow10<double>(in1 < 0 ? -in1 : in1);'
                ' out0 = in1 < 0 ? rint(in0 / x) * x : rint(in0 * x) / x; }')
_round_ufunc = create_ufunc(
    'cupy_round',
    ('?q->e',) + tuple('%sq->%s' % (c, c) for c in _INT) + tuple(('%sq->%s' % (c, c), _round_float) for c in _FLT),
    'out0 = in0', preamble=_round_preamble)
# integers with negative `decimals`: scale, round the last two digits half-to-even, unscale
_round_ufunc_neg_int = create_ufunc(
    'cupy_round_neg_uint', ('?q->e',) + tuple('%sq->%s' % (c, c) for c in _INT),
    'long long x = pow10<long long>(in1 - 1);'
    ' long long q = in0 / x / 100; int r = in0 - q * x * 100;'
    ' out0 = (q * 100 + __float2ll_rn(r / (x * 10.0f)) * 10) * x;',
    preamble=_round_preamble)


def around(a, decimals=0, out=None):
    """Rounds to `decimals` places, half to even (`numpy.around`; cupy/_core/core.pyx:1365-1377)."""
    a = _math._as_array(a)
    if decimals < 0 and a.dtype.kind in 'iu':
        return _round_ufunc_neg_int(a, -decimals, out=out)
    return _round_ufunc(a, decimals, out=out)


round = round_ = around

# ---- arithmetic -------------------------------------------------------------------------------
_float_recip = 'out0 = 1 / in0'
reciprocal = create_ufunc('cupy_reciprocal', _loops(_INT) + _loops(_FLT, code=_float_recip),
                          'out0 = in0 == 0 ? 0 : (1 / in0)')


def _positive_boolean_error():
    raise TypeError('The cupy boolean positive, the `+` operator, is not supported.')


positive = create_ufunc('cupy_positive', (('?->?', _positive_boolean_error),) + _loops(_INT + _FLT), 'out0 = +in0')

# `_floor_divide(x, y)` is a device helper user kernels may call as well (b200/carray.cuh); integer
# division by zero yields 0, as in the reference
floor_divide = create_ufunc('cupy_floor_divide', _loops(_INT + _FLT, 2), 'out0 = _floor_divide(in0, in1)')
_float_rem = 'out0 = in0 - _floor_divide(in0, in1) * in1'
remainder = create_ufunc('cupy_remainder', _loops(_INT, 2) + _loops(_FLT, 2, code=_float_rem),
                         'out0 = (in0 - _floor_divide(in0, in1) * in1) * (in1 != 0)')
mod = remainder
_divmod_float = 'out0_type a = _floor_divide(in0, in1); out0 = a; out1 = in0 - a * in1'
divmod = create_ufunc(
    'cupy_divmod',
    tuple('%s%s->%s%s' % (c, c, c, c) for c in _INT) + tuple(('%s%s->%s%s' % (c, c, c, c), _divmod_float) for c in _FLT),
    'if (in1 == 0) { out0 = 0; out1 = 0; }'
    ' else { out0_type a = _floor_divide(in0, in1); out0 = a; out1 = in0 - a * in1; }')
fmod = create_ufunc(
    'cupy_fmod',
    _loops(_INT, 2) + (('ee->e', 'out0 = fmodf(in0, in1)'), ('ff->f', 'out0 = fmodf(in0, in1)'),
                       ('dd->d', 'out0 = fmod(in0, in1)')),
    'out0 = in1 == 0 ? 0 : fmod((double)in0, (double)in1)')
modf = create_ufunc(
    'cupy_modf', ('e->ee', 'f->ff', ('d->dd', 'double iptr; out0 = modf(in0, &iptr); out1 = iptr')),
    'float iptr; out0 = modff(in0, &iptr); out1 = iptr')
float_power = create_ufunc('cupy_float_power', ('dd->d',), 'out0 = pow(in0, in1)')

# ---- floating-point pieces --------------------------------------------------------------------
signbit = create_ufunc('cupy_signbit', _loops(_FLT, out='?'), 'out0 = signbit(in0)')
copysign = _math._create_math_ufunc('copysign', 2, 'cupy_copysign')
# float16 steps by one half-precision ulp: through `float` the neighbour would round straight back
_nextafter_half = '''
__device__ float16 nextafter(float16 x, float16 y) {
    const float fx = x, fy = y;
    if (fx != fx || fy != fy) return float16(fx + fy);
    if (fx == fy) return y;
    unsigned short b = __half_as_ushort(x.raw());
    if ((b & 0x7fff) == 0) b = (fy > 0 ? 0x0000 : 0x8000) | 1;            // off zero: the smallest subnormal
    else if ((fx < fy) == (fx > 0)) ++b;                                   // away from zero
    else --b;
    return float16(__ushort_as_half(b));
}
'''
nextafter = create_ufunc('cupy_nextafter', _loops(_FLT, 2), 'out0 = nextafter(in0, in1)', preamble=_nextafter_half)
ldexp = create_ufunc('cupy_ldexp', ('ei->e', 'fi->f', 'el->e', 'fl->f', 'di->d', 'dq->d'), 'out0 = ldexp(in0, in1)')
frexp = create_ufunc('cupy_frexp', ('e->ei', 'f->fi', 'd->di'), 'int nptr; out0 = frexp(in0, &nptr); out1 = nptr')

# ---- miscellany -------------------------------------------------------------------------------
cbrt = create_ufunc('cupy_cbrt', _loops(_FLT), 'out0 = cbrt(in0)')
fabs = create_ufunc('cupy_fabs', _loops(_FLT), 'out0 = fabs(in0)')

_signed_sign = 'out0 = (in0 > 0) - (in0 < 0)'
_unsigned_sign = 'out0 = in0 > 0'
# NaN -> NaN and +-0 -> +-0 through `in0 - in0`
_float_sign = 'if (in0 < 0 || in0 > 0) { out0 = copysign(static_cast<in0_type>(1), in0); } else { out0 = in0 - in0; }'
sign = create_ufunc(
    'cupy_sign',
    tuple(('%s->%s' % (c, c), _unsigned_sign if c.isupper() else _signed_sign) for c in _INT) + _loops(_FLT),
    _float_sign)
heaviside = create_ufunc(
    'cupy_heaviside', _loops(_FLT, 2),
    'if (isnan(in0)) { out0 = in0; } else if (in0 == 0) { out0 = in1; } else { out0 = (in0 > 0); }')
# fmax / fmin: a NaN operand loses (C's fmax / fmin), unlike maximum / minimum
fmax = create_ufunc('cupy_fmax', _loops('?' + _INT, 2) + _loops(_FLT, 2, code='out0 = fmax(in0, in1)'),
                    'out0 = max(in0, in1)')
fmin = create_ufunc('cupy_fmin', _loops('?' + _INT, 2) + _loops(_FLT, 2, code='out0 = fmin(in0, in1)'),
                    'out0 = min(in0, in1)')

clip_ufunc = create_ufunc(
    'cupy_clip', _loops('?' + _INT + _FLT, 3),
    'out0 = in1 > in2 ? in2 : (in0 < in1 ? in1 : (in0 > in2 ? in2 : in0))')


def clip(a, a_min=None, a_max=None, out=None):
    """`maximum(minimum(a, a_max), a_min)` in one kernel.  A missing bound becomes the dtype's own limit;
    when a_min > a_max every element becomes a_max (cupy/_core/_routines_math.pyx:139-151)."""
    a = _math._as_array(a)
    kind = a.dtype.kind
    if a_min is None:
        a_min = a.dtype.type('-inf') if kind == 'f' else numpy.iinfo(a.dtype).min if kind in 'iu' else None
    if a_max is None:
        a_max = a.dtype.type('inf') if kind == 'f' else numpy.iinfo(a.dtype).max if kind in 'iu' else None
    return clip_ufunc(a, a_min, a_max, out=out)


_nan_to_num_preamble = '''
template <class T> __device__ T nan_to_num(T x, T nan, T posinf, T neginf) {
    if (isnan(x)) return nan;
    if (isinf(x)) return x > 0 ? posinf : neginf;
    return x;
}
'''
_nan_to_num = create_ufunc(
    'cupy_nan_to_num_', _loops('?' + _INT, 4) + _loops(_FLT, 4, code='out0 = nan_to_num(in0, in1, in2, in3)'),
    'out0 = in0', preamble=_nan_to_num_preamble)


def nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    """NaN -> `nan`, +-inf -> `posinf` / `neginf` (default: the dtype's largest finite values)."""
    if not isinstance(x, ndarray):
        x = _math._as_array(x)
    if x.dtype.kind != 'f':
        return x.copy() if copy else x
    info = numpy.finfo(x.dtype)
    hi = info.max if posinf is None else posinf
    lo = info.min if neginf is None else neginf
    t = x.dtype.type
    return _nan_to_num(x, t(nan), t(hi), t(lo), out=None if copy else x)


def _bool_gcd_error():
    raise TypeError('gcd cannot be computed with boolean arrays')


def _bool_lcm_error():
    raise TypeError('lcm cannot be computed with boolean arrays')


_gcd_preamble = '''
template <typename T> inline __device__ T gcd(T a, T b) {
    while (b != 0) { T r = a % b; a = b; b = r; }
    return a < 0 ? -a : a;
}
'''
_lcm_preamble = _gcd_preamble + '''
template <typename T> inline __device__ T lcm(T a, T b) {
    T g = gcd(a, b);
    if (g == 0) return 0;
    T r = a / g * b;
    return r < 0 ? -r : r;
}
'''
gcd = create_ufunc('cupy_gcd', (('??->?', _bool_gcd_error),) + _loops(_INT, 2), 'out0 = gcd(in0, in1)',
                   preamble=_gcd_preamble)
lcm = create_ufunc('cupy_lcm', (('??->?', _bool_lcm_error),) + _loops(_INT, 2), 'out0 = lcm(in0, in1)',
                   preamble=_lcm_preamble)

# ---- logic ------------------------------------------------------------------------------------
# the comparison loop table of cupy/_core/_routines_logic.pyx:101-118 (mixed int64 / uint64 loops included)
logical_and = _math._create_comparison('logical_and', '&&')
logical_or = _math._create_comparison('logical_or', '||')
logical_not = create_ufunc('cupy_logical_not', _loops('?' + _INT + _FLT, out='?'), 'out0 = !in0')
logical_xor = create_ufunc('cupy_logical_xor', _loops('?' + _INT + _FLT, 2, out='?'), 'out0 = !in0 != !in1')

isfinite = create_ufunc('cupy_isfinite', _loops(_FLT, out='?'), 'out0 = isfinite(in0)')
isinf = create_ufunc('cupy_isinf', _loops(_FLT, out='?'), 'out0 = isinf(in0)')
isnan = create_ufunc('cupy_isnan', _loops(_FLT, out='?'), 'out0 = isnan(in0)')


def _signed_inf(x, out, negative):
    x = _math._as_array(x)
    inf = isinf(x)
    if x.dtype.kind != 'f':
        # integers hold no infinities
        return inf if out is None else _math.elementwise_copy(inf, out)
    sb = signbit(x)
    if not negative:
        sb = logical_not(sb)
    return logical_and(inf, sb, out=out)


def isneginf(x, out=None):
    """True where x is -inf."""
    return _signed_inf(x, out, True)


def isposinf(x, out=None):
    """True where x is +inf."""
    return _signed_inf(x, out, False)


_is_close = create_ufunc(
    'cupy_is_close', ('eeee?->?', 'ffff?->?', 'dddd?->?'),
    'bool equal_nan = in4;'
    ' if (isfinite(in0) && isfinite(in1)) { out0 = fabs(in0 - in1) <= in3 + in2 * fabs(in1); }'
    ' else if (equal_nan) { out0 = (in0 == in1) || (isnan(in0) && isnan(in1)); }'
    ' else { out0 = (in0 == in1); }')


def isclose(a, b, rtol=1.e-5, atol=1.e-8, equal_nan=False):
    """|a - b| <= atol + rtol * |b| elementwise, infinities equal to themselves (`numpy.isclose`)."""
    a, b = _math._as_array(a), _math._as_array(b)
    if a.dtype.kind not in 'f' or b.dtype.kind not in 'f':
        # integer inputs compare in float64, as in the reference (`astype(result_type(a, b, float))`)
        dt = numpy.result_type(a.dtype, b.dtype, numpy.float64)
        a, b = a.astype(dt), b.astype(dt)
    return _is_close(a, b, rtol, atol, equal_nan)


def allclose(a, b, rtol=1.e-5, atol=1.e-8, equal_nan=False):
    """0-d boolean array (no device synchronisation), like the reference."""
    from cupy_b200._core import _routines_more
    return _routines_more.all(isclose(a, b, rtol, atol, equal_nan))


def array_equal(a1, a2, equal_nan=False):
    """0-d boolean array: same shape and all elements equal."""
    from cupy_b200._core import _routines_more
    from cupy_b200._core._ndarray import asarray
    a1, a2 = _math._as_array(a1), _math._as_array(a2)
    if a1.shape != a2.shape:
        return asarray(numpy.array(False))
    if not equal_nan or (a1.dtype.kind != 'f' and a2.dtype.kind != 'f'):
        return _routines_more.all(_math.equal(a1, a2))
    both_nan = logical_and(_math.not_equal(a1, a1), _math.not_equal(a2, a2))
    return _routines_more.all(logical_or(_math.equal(a1, a2), both_nan))


_where_ufunc = create_ufunc('cupy_where', tuple('?%s%s->%s' % (c, c, c) for c in '?' + _INT + _FLT),
                            'out0 = in0 ? in1 : in2')


def where(condition, x=None, y=None):
    """Elements of x where `condition` holds, of y elsewhere; `where(condition)` alone is `nonzero(condition)`
    (the scan-based compaction of _core/_compaction.py)."""
    missing = (x is None, y is None)
    if missing == (True, True):
        from cupy_b200._core import _compaction
        return _compaction.nonzero(condition)
    if missing != (False, False):
        raise ValueError('Must provide both \'x\' and \'y\' or neither.')
    condition = _math._as_array(condition)
    if condition.dtype != numpy.bool_:
        condition = _math.not_equal(condition, 0)
    return _where_ufunc(condition, x, y)
