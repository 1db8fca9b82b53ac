"""The rest of the reduction family around the hot path:
`all / any / count_nonzero` (cupy/_core/_routines_logic.pyx:43-57, cupy/_sorting/count.py:29-33),
`nansum / nanprod` (cupy/_core/_routines_math.pyx:810-835), `nanmin / nanmax / nanargmin /
nanargmax / ptp` (cupy/_core/_routines_statistics.pyx:48-125, 310-391; cupy/_statistics/order.py),
`nanmean / nanvar / nanstd` (cupy/_core/_routines_statistics.pyx:667-768, cupy/_statistics/meanvar.py:213-291),
`nancumsum / nancumprod` (cupy/_math/sumprod.py:183-220, 351-360), `average` (cupy/_statistics/meanvar.py:72-138).

Same routine-string quadruples as the reference; they are NVRTC-compiled into the structured
FULL / ROWS / COLS skeletons of b200/reduce.cuh like any other `create_reduction_func`.
"""
from __future__ import annotations

from cupy_b200._core._reduction import ReductionKernel, create_reduction_func
from cupy_b200._core import _routines_math as _math
from cupy_b200._core import _routines_statistics as _stat

_LOGIC_TYPES = ('?->?', 'b->?', 'B->?', 'h->?', 'H->?', 'i->?', 'I->?', 'l->?', 'L->?', 'q->?', 'Q->?',
                'e->?', 'f->?', 'd->?')

_all = create_reduction_func('cupy_all', _LOGIC_TYPES, ('in0 != type_in0_raw(0)', 'a & b', 'out0 = a', 'bool'), 'true', '')
_any = create_reduction_func('cupy_any', _LOGIC_TYPES, ('in0 != type_in0_raw(0)', 'a | b', 'out0 = a', 'bool'), 'false', '')

_count_nonzero = create_reduction_func(
    'cupy_count_nonzero',
    ('?->l', 'b->l', 'B->l', 'h->l', 'H->l', 'i->l', 'I->l', 'l->l', 'L->l', 'q->l', 'Q->l', 'e->l', 'f->l', 'd->l'),
    ('in0 != type_in0_raw(0)', 'a + b', 'out0 = a', None), 0)

_nansum_auto = create_reduction_func(
    'cupy_nansum', _math._sumprod_types,
    ('(in0 == in0) ? in0 : type_in0_raw(0)', 'a + b', 'out0 = type_out0_raw(a)', None), 0)
_nansum_keep = create_reduction_func(
    'cupy_nansum_with_dtype', _math._keep_types,
    ('(in0 == in0) ? in0 : type_in0_raw(0)', 'a + b', 'out0 = type_out0_raw(a)', None), 0)
_nanprod_auto = create_reduction_func(
    'cupy_nanprod', _math._sumprod_types,
    ('(in0 == in0) ? in0 : type_in0_raw(1)', 'a * b', 'out0 = type_out0_raw(a)', None), 1)
_nanprod_keep = create_reduction_func(
    'cupy_nanprod_with_dtype', _math._keep_types,
    ('(in0 == in0) ? in0 : type_in0_raw(1)', 'a * b', 'out0 = type_out0_raw(a)', None), 1)

# NaN-ignoring extrema: a NaN operand loses (CUDA's min()/max() semantics in the reference's my_min / my_max)
_nan_preamble = _stat._min_max_preamble + '''
template <typename T>
__device__ min_max_st<T> my_nanmin(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (_b200_isnan(a.value)) return b;
    if (_b200_isnan(b.value)) return a;
    return min_max_st<T>(b.value < a.value ? b.value : a.value);
}
template <typename T>
__device__ min_max_st<T> my_nanmax(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (_b200_isnan(a.value)) return b;
    if (_b200_isnan(b.value)) return a;
    return min_max_st<T>(a.value < b.value ? b.value : a.value);
}
'''
_SAME = _stat._ALL1 + ('e->e', 'f->f', 'd->d')
_nanmin = create_reduction_func(
    'cupy_nanmin', _SAME,
    ('min_max_st<type_in0_raw>(in0)', 'my_nanmin(a, b)', 'out0 = a.value', 'min_max_st<type_in0_raw>'), None, _nan_preamble)
_nanmax = create_reduction_func(
    'cupy_nanmax', _SAME,
    ('min_max_st<type_in0_raw>(in0)', 'my_nanmax(a, b)', 'out0 = a.value', 'min_max_st<type_in0_raw>'), None, _nan_preamble)

_ARG = tuple('%s->q' % c for c in '?bBhHiIlLqQ')
_nanargmin = create_reduction_func(
    'cupy_nanargmin',
    _ARG + (('e->q', (None, 'my_argmin_float(a, b)', None, None)), ('f->q', (None, 'my_argmin_float(a, b)', None, None)),
            ('d->q', (None, 'my_argmin_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0, _b200_isnan(in0) ? -1 : _J)', 'my_argmin(a, b)', 'out0 = a.index',
     'min_max_st<type_in0_raw>'), None, _stat._min_max_preamble, sort_reduce_axis=False)
_nanargmax = create_reduction_func(
    'cupy_nanargmax',
    _ARG + (('e->q', (None, 'my_argmax_float(a, b)', None, None)), ('f->q', (None, 'my_argmax_float(a, b)', None, None)),
            ('d->q', (None, 'my_argmax_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0, _b200_isnan(in0) ? -1 : _J)', 'my_argmax(a, b)', 'out0 = a.index',
     'min_max_st<type_in0_raw>'), None, _stat._min_max_preamble, sort_reduce_axis=False)


def all(a, axis=None, out=None, keepdims=False):
    return _all(_math._as_array(a), axis=axis, out=out, keepdims=keepdims)


def any(a, axis=None, out=None, keepdims=False):
    return _any(_math._as_array(a), axis=axis, out=out, keepdims=keepdims)


def count_nonzero(a, axis=None):
    """Returns a 0-d array for axis=None, like the reference (no device synchronisation)."""
    return _count_nonzero(_math._as_array(a), axis=axis)


def nansum(a, axis=None, dtype=None, out=None, keepdims=False):
    k = _nansum_auto if dtype is None else _nansum_keep
    return k(_math._as_array(a), axis, dtype, out, keepdims)


def nanprod(a, axis=None, dtype=None, out=None, keepdims=False):
    k = _nanprod_auto if dtype is None else _nanprod_keep
    return k(_math._as_array(a), axis, dtype, out, keepdims)


def nanmin(a, axis=None, out=None, keepdims=False):
    """NaN is returned for an all-NaN slice (the reference also warns, which needs a device sync)."""
    return _nanmin(_math._as_array(a), axis=axis, out=out, keepdims=keepdims)


def nanmax(a, axis=None, out=None, keepdims=False):
    return _nanmax(_math._as_array(a), axis=axis, out=out, keepdims=keepdims)


def _all_nan_guard(a, axis):
    a = _math._as_array(a)
    if a.dtype.kind in 'biu':
        return a
    from cupy_b200._core import _routines_math as m
    nan = m.not_equal(a, a)
    if bool(_any(_all(nan, axis=axis)).get()):
        raise ValueError('All-NaN slice encountered')
    return a


def nanargmin(a, axis=None, dtype=None, out=None, keepdims=False):
    """cupy/_sorting/search.py `nanargmin`: raises ValueError for an all-NaN slice."""
    a = _all_nan_guard(a, axis)
    return _nanargmin(a, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def nanargmax(a, axis=None, dtype=None, out=None, keepdims=False):
    a = _all_nan_guard(a, axis)
    return _nanargmax(a, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def ptp(a, axis=None, out=None, keepdims=False):
    """max - min along an axis (cupy/_core/_routines_statistics.pyx:48-66: two reductions and a subtract)."""
    a = _math._as_array(a)
    hi = a.max(axis=axis, keepdims=keepdims)
    lo = a.min(axis=axis, keepdims=keepdims)
    return _math.subtract(hi, lo, out=out) if out is not None else _math.subtract(hi, lo)


# ---- NaN-ignoring moments ----------------------------------------------------------------------
# (sum of the non-NaN values, how many there were) carried as one accumulator; float16 sums in float
_nanmean_preamble = '''
template <typename T>
struct nanmean_st {
    T value;
    long long count;
    __device__ nanmean_st() : value(0), count(0) { }
    __device__ nanmean_st(T v) : value(v == v ? v : T(0)), count(v == v ? 1 : 0) { }
    __device__ nanmean_st(T v, long long c) : value(v), count(c) { }
};
template <typename T>
__device__ nanmean_st<T> my_nanmean(const nanmean_st<T>& a, const nanmean_st<T>& b) {
    return nanmean_st<T>(a.value + b.value, a.count + b.count);
}
'''
_nanmean_func = create_reduction_func(
    'cupy_nanmean',
    (('e->e', ('nanmean_st<float>(in0)', None, 'out0 = a.value / float(a.count)', 'nanmean_st<float>')), 'f->f', 'd->d'),
    ('nanmean_st<type_out0_raw>(in0)', 'my_nanmean(a, b)', 'out0 = a.value / type_out0_raw(a.count)',
     'nanmean_st<type_out0_raw>'), None, _nanmean_preamble)
_count_non_nan = create_reduction_func(
    'cupy_count_non_nan', ('e->q', 'f->q', 'd->q'), ('(in0 == in0) ? 1 : 0', 'a + b', 'out0 = a', None), 0)

_nanvar_preamble = '''
template <typename S, typename T>
__device__ T nanvar_impl(S x, T mean, long long alpha) {
    return (x == x ? T((x - mean) * (x - mean)) : T(0)) / alpha;
}
'''
_nanvar_core = ReductionKernel(
    'S x, T sum, int64 _count, int64 ddof', 'S out',
    'nanvar_impl(x, sum / _count, max(_count - ddof, 0LL))', 'a + b', 'out = a', '0', 'cupy_nanvar_core',
    preamble=_nanvar_preamble)
# float16 results accumulate the squared deviations in float, like `var` (the reference's float16 accumulator
# loses the small terms of a long sum)
_nanvar_core_float16 = ReductionKernel(
    'S x, T sum, int64 _count, int64 ddof', 'float16 out',
    'nanvar_impl(static_cast<float>(x), static_cast<float>(sum) / _count, max(_count - ddof, 0LL))', 'a + b', 'out = a',
    '0', 'cupy_nanvar_core_float16', reduce_type='float', preamble=_nanvar_preamble)
_nanvar_core_out = ReductionKernel(
    'S x, T sum, int64 _count, int64 ddof', 'U out',
    'nanvar_impl(x, sum / static_cast<T>(_count), max(_count - ddof, 0LL))', 'a + b', 'out = a', '0',
    'cupy_nanvar_core_out', preamble=_nanvar_preamble)


def nanmean(a, axis=None, dtype=None, out=None, keepdims=False):
    """Mean of the non-NaN elements; NaN for an all-NaN slice (0 / 0), without a device synchronisation."""
    a = _math._as_array(a)
    if a.dtype.kind in 'biu':
        return a.mean(axis=axis, dtype=dtype, out=out, keepdims=keepdims)
    return _nanmean_func(a, axis=axis, dtype=dtype, out=out, keepdims=keepdims)


def nanvar(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    """Variance of the non-NaN elements: count and sum of the valid ones, then the squared deviations
    from their mean (the reference's three launches, cupy/_core/_routines_statistics.pyx:712-729)."""
    a = _math._as_array(a)
    if a.dtype.kind in 'biu':
        return a.var(axis=axis, dtype=dtype, out=out, ddof=ddof, keepdims=keepdims)
    count = _count_non_nan(a, axis=axis, keepdims=True)
    total = nansum(a, axis=axis, dtype=dtype, keepdims=True)
    if out is None:
        core = _nanvar_core_float16 if a.dtype.char == 'e' and total.dtype.char == 'e' else _nanvar_core
        return core(a, total, count, ddof, axis=axis, keepdims=keepdims)
    _nanvar_core_out(a, total, count, ddof, out, axis=axis, keepdims=keepdims)
    return out


def nanstd(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    a = _math._as_array(a)
    if a.dtype.kind in 'biu':
        return a.std(axis=axis, dtype=dtype, out=out, ddof=ddof, keepdims=keepdims)
    return _math.sqrt(nanvar(a, axis, dtype, None, ddof, keepdims), dtype=dtype, out=out)


# ---- NaN-ignoring scans: NaNs become the identity, then the ordinary scan ------------------------
_replace_nan_memo = []


def _replace_nan(a, val, out=None):
    from cupy_b200._core._kernel import ElementwiseKernel
    from cupy_b200._core._ndarray import empty_like
    if not _replace_nan_memo:
        _replace_nan_memo.append(ElementwiseKernel(
            'T a, T val', 'T out', 'if (a == a) { out = a; } else { out = val; }', 'cupy_replace_nan'))
    if out is None or a.dtype != out.dtype:
        out = empty_like(a)
    _replace_nan_memo[0](a, val, out)
    return out


def nancumsum(a, axis=None, dtype=None, out=None):
    a = _replace_nan(_math._as_array(a), 0, out=out)
    return _math.cumsum(a, axis=axis, dtype=dtype, out=out)


def nancumprod(a, axis=None, dtype=None, out=None):
    a = _replace_nan(_math._as_array(a), 1, out=out)
    return _math.cumprod(a, axis=axis, dtype=dtype, out=out)


def average(a, axis=None, weights=None, returned=False, *, keepdims=False):
    """Weighted mean along an axis.  With weights the zero-weight-sum check reads one flag back from the
    device, as the reference does."""
    import numpy
    a = _math._as_array(a)
    if weights is None:
        avg = a.mean(axis=axis, keepdims=keepdims)
        if not returned:
            return avg
        from cupy_b200._core._ndarray import full
        return avg, full(avg.shape, a.size / max(avg.size, 1), avg.dtype)
    wgt = _math._as_array(weights)
    if a.dtype.kind in 'iub':
        result_dtype = numpy.promote_types(numpy.promote_types(a.dtype, wgt.dtype), 'f8')
    else:
        result_dtype = numpy.promote_types(a.dtype, wgt.dtype)
    if a.shape != wgt.shape:
        if axis is None:
            raise TypeError('Axis must be specified when shapes of a and weights differ.')
        if wgt.ndim != 1:
            raise TypeError('1D weights expected when shapes of a and weights differ.')
        if wgt.shape[0] != a.shape[axis]:
            raise ValueError('Length of weights not compatible with specified axis.')
        wgt = wgt.broadcast_to((a.ndim - 1) * (1,) + wgt.shape).swapaxes(-1, axis)
    scl = wgt.sum(axis=axis, dtype=result_dtype, keepdims=keepdims)
    if bool(any(_math.equal(scl, 0.0)).get()):
        raise ZeroDivisionError('Weights sum to zero, can\'t be normalized')
    avg = _math.true_divide(_math.multiply(a, wgt, dtype=result_dtype).sum(axis, keepdims=keepdims), scl)
    if not returned:
        return avg
    if scl.shape != avg.shape:
        scl = scl.broadcast_to(avg.shape).copy()
    return avg, scl
