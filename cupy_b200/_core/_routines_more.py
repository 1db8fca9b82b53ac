"""The rest of the reduction family around the hot path:
`all / any / count_nonzero` (cupy/_core/_routines_logic.pyx:43-57, cupy/_sorting/count.py:29-33),
`nansum / nanprod` (cupy/_core/_routines_math.pyx:810-835), `nanmin / nanmax / nanargmin /
nanargmax / ptp` (cupy/_core/_routines_statistics.pyx:48-125, 310-391; cupy/_statistics/order.py).

Same routine-string quadruples as the reference; they are NVRTC-compiled into the structured
FULL / ROWS / COLS skeletons of b200/reduce.cuh like any other `create_reduction_func`.
"""
from __future__ import annotations

from cupy_b200._core._reduction import create_reduction_func
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
