"""max / min / argmax / argmin / mean / var / std.

Mirror of the hot-path parts of cupy/_core/_routines_statistics.pyx: routers
`_ndarray_max/_min/_argmax/_argmin/_mean/_var/_std` (:27-181), the `min_max_st`
functors and `_amax/_amin/_argmax/_argmin` tables (:190-353), `_mean_core`
(:647-655), `_var` (:556-600).  The preamble text is the interface user
ReductionKernels can rely on (`min_max_st`, `my_max`, ...), re-stated.
"""
from __future__ import annotations

import math

import numpy

from cupy_b200 import _lib
from cupy_b200._core import _accelerator, _scalar
from cupy_b200._core import _routines_math as _math
from cupy_b200._core._ndarray import ndarray
from cupy_b200._core._reduction import ReductionKernel, create_reduction_func, _get_axis

_min_max_preamble = '''
template <typename T>
struct min_max_st{
    T value;
    IndexT index;
    __device__ min_max_st() : index(-1) { }
    __device__ min_max_st(T v) : value(v), index(0) { }
    __device__ min_max_st(T v, IndexT i) : value(v), index(i) { }
};
template <typename T> __device__ bool _b200_isnan(T) { return false; }
__device__ bool _b200_isnan(float v) { return v != v; }
__device__ bool _b200_isnan(double v) { return v != v; }
__device__ bool _b200_isnan(float16 v) { return b200::isnan(v); }

template <typename T>
__device__ min_max_st<T> my_min(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    return min_max_st<T>(b.value < a.value ? b.value : a.value);
}
template <typename T>
__device__ min_max_st<T> my_min_float(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (_b200_isnan(a.value)) return a;
    if (_b200_isnan(b.value)) return b;
    return min_max_st<T>(b.value < a.value ? b.value : a.value);
}
template <typename T>
__device__ min_max_st<T> my_max(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    return min_max_st<T>(a.value < b.value ? b.value : a.value);
}
template <typename T>
__device__ min_max_st<T> my_max_float(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (_b200_isnan(a.value)) return a;
    if (_b200_isnan(b.value)) return b;
    return min_max_st<T>(a.value < b.value ? b.value : a.value);
}
template <typename T>
__device__ min_max_st<T> my_argmin(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (a.value == b.value) return min_max_st<T>(a.value, a.index < b.index ? a.index : b.index);
    return (a.value <= b.value) ? a : b;
}
template <typename T>
__device__ min_max_st<T> my_argmin_float(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (a.value == b.value) return min_max_st<T>(a.value, a.index < b.index ? a.index : b.index);
    const bool na = _b200_isnan(a.value), nb = _b200_isnan(b.value);
    if (na && nb) return (a.index <= b.index) ? a : b;     // first NaN, as NumPy
    if (na) return a;
    if (nb) return b;
    return (a.value <= b.value) ? a : b;
}
template <typename T>
__device__ min_max_st<T> my_argmax(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (a.value == b.value) return min_max_st<T>(a.value, a.index < b.index ? a.index : b.index);
    return (a.value >= b.value) ? a : b;
}
template <typename T>
__device__ min_max_st<T> my_argmax_float(const min_max_st<T>& a, const min_max_st<T>& b) {
    if (a.index == -1) return b;
    if (b.index == -1) return a;
    if (a.value == b.value) return min_max_st<T>(a.value, a.index < b.index ? a.index : b.index);
    const bool na = _b200_isnan(a.value), nb = _b200_isnan(b.value);
    if (na && nb) return (a.index <= b.index) ? a : b;
    if (na) return a;
    if (nb) return b;
    return (a.value >= b.value) ? a : b;
}
'''

_ALL1 = ('?->?', 'b->b', 'B->B', 'h->h', 'H->H', 'i->i', 'I->I', 'l->l', 'L->L', 'q->q', 'Q->Q')

_amin = create_reduction_func(
    'cupy_min',
    _ALL1 + (('e->e', (None, 'my_min_float(a, b)', None, None)),
             ('f->f', (None, 'my_min_float(a, b)', None, None)),
             ('d->d', (None, 'my_min_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0)', 'my_min(a, b)', 'out0 = a.value', 'min_max_st<type_in0_raw>'),
    None, _min_max_preamble, prebuilt=_lib.OP_MIN)

_amax = create_reduction_func(
    'cupy_max',
    _ALL1 + (('e->e', (None, 'my_max_float(a, b)', None, None)),
             ('f->f', (None, 'my_max_float(a, b)', None, None)),
             ('d->d', (None, 'my_max_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0)', 'my_max(a, b)', 'out0 = a.value', 'min_max_st<type_in0_raw>'),
    None, _min_max_preamble, prebuilt=_lib.OP_MAX)

_arg_int_loops = tuple('{}->{}'.format(d, r) for r in 'qlihb' for d in '?BhHiIlLqQ') + ('b->q',)

_argmin = create_reduction_func(
    'cupy_argmin',
    _arg_int_loops + (('e->q', (None, 'my_argmin_float(a, b)', None, None)),
                      ('f->q', (None, 'my_argmin_float(a, b)', None, None)),
                      ('d->q', (None, 'my_argmin_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0, _J)', 'my_argmin(a, b)', 'out0 = a.index', 'min_max_st<type_in0_raw>'),
    None, _min_max_preamble, sort_reduce_axis=False, prebuilt=_lib.OP_ARGMIN)

_argmax = create_reduction_func(
    'cupy_argmax',
    _arg_int_loops + (('e->q', (None, 'my_argmax_float(a, b)', None, None)),
                      ('f->q', (None, 'my_argmax_float(a, b)', None, None)),
                      ('d->q', (None, 'my_argmax_float(a, b)', None, None))),
    ('min_max_st<type_in0_raw>(in0, _J)', 'my_argmax(a, b)', 'out0 = a.index', 'min_max_st<type_in0_raw>'),
    None, _min_max_preamble, sort_reduce_axis=False, prebuilt=_lib.OP_ARGMAX)

_mean_types = ('?->d', 'B->d', 'b->d', 'h->d', 'H->d', 'i->d', 'I->d', 'l->d', 'L->d', 'q->d', 'Q->d',
               ('e->e', (None, None, None, 'float')), 'f->f', 'd->d')
_mean_core = create_reduction_func(
    'cupy_mean', _mean_types,
    ('in0', 'a + b', 'out0 = a / _type_reduce(_in_ind.size() / _out_ind.size())', None),
    prebuilt=_lib.OP_MEAN)
_mean_core_empty = create_reduction_func(
    'cupy_mean_empty', _mean_types,
    ('in0', 'a + b', 'out0 = a / _type_reduce(_in_ind.size() / _out_ind.size())', None), 0)


def _ndarray_max(self, axis, out, dtype, keepdims):
    return _amax(self, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def _ndarray_min(self, axis, out, dtype, keepdims):
    return _amin(self, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def _ndarray_argmax(self, axis, out, dtype, keepdims):
    return _argmax(self, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def _ndarray_argmin(self, axis, out, dtype, keepdims):
    return _argmin(self, axis=axis, out=out, dtype=dtype, keepdims=keepdims)


def _ndarray_mean(self, axis, dtype, out, keepdims):
    """cupy/_core/_routines_statistics.pyx:128-175 (dtype rules :132-146)."""
    dtype_sum = dtype_out = dtype
    if dtype is None and self.dtype.char not in 'd':
        if self.dtype.kind in 'iub':
            dtype_out = numpy.float64
            dtype_sum = numpy.float64
        elif self.dtype.char == 'e':
            dtype_sum = numpy.float32
            dtype_out = numpy.float16
    elif dtype is not None and numpy.dtype(dtype).kind in 'iub':
        dtype_out = dtype
        dtype_sum = numpy.float64
    if self.size == 0:
        result = _mean_core_empty(self, axis, dtype_sum, out, keepdims)
    elif dtype is None and out is None:
        # natural loop (ints -> float64, float16 accumulated in float): prebuilt MeanOp
        return _mean_core(self, axis, None, None, keepdims)
    else:
        result = _mean_core(self, axis, dtype_sum, out, keepdims)
    if dtype_out is not None and out is None:
        result = result.astype(dtype_out, copy=False)
    return result


# second pass of the reference's two-pass variance; kept because it is the
# documented `_var_core_*` ReductionKernel surface and the `out=` / `dtype=` path
_norm_preamble = '''
template <typename T> __device__ T my_norm(T x) { return x * x; }
'''
_var_core_out = ReductionKernel(
    'S x, T mean, U alpha', 'U out', 'my_norm(x - mean)', 'a + b', 'out = alpha * a', '0',
    'cupy_var_core_out', preamble=_norm_preamble)

# float16 results: the reference's `_var_core_float16` (:611-616) accumulates the squared deviations in float16 and
# overflows to inf past 65504 (as NumPy does); here they are accumulated in float -- the same arithmetic as the
# single-pass functor, so the answer does not depend on which layout route a call takes
_var_core_float16 = ReductionKernel(
    'S x, T mean, float32 alpha', 'float16 out',
    'my_norm(static_cast<float>(x) - static_cast<float>(mean))', 'a + b', 'out = alpha * a', '0',
    'cupy_var_core_float16', reduce_type='float', preamble=_norm_preamble)

_var_types = ('?->d', 'b->d', 'B->d', 'h->d', 'H->d', 'i->d', 'I->d', 'l->d', 'L->d', 'q->d', 'Q->d',
              'e->e', 'f->f', 'd->d')


class _VarKernel:
    """Single-pass variance on the prebuilt Welford/Chan functor (B200_OP_VAR)."""

    def __init__(self):
        # routine strings are only used as a (never taken) fallback description
        self._k = create_reduction_func(
            'cupy_var', _var_types, ('in0', 'a + b', 'out0 = type_out0_raw(a)', None), 0, prebuilt=_lib.OP_VAR)

    def __call__(self, a, axis, ddof, keepdims):
        k = self._k
        out_args = []
        return k._call([a], out_args, a.shape, axis, None, keepdims, True, None, param=float(ddof))


_var_single_pass = _VarKernel()
_var_hot = {}      # (dtype, shape, strides, axis) -> normalised axis tuple of call shapes the single pass takes


def _var(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    """cupy/_core/_routines_statistics.pyx:556-600."""
    hot_key = None
    fast = _accelerator.fast_paths_enabled()     # the single-pass functor is one of the accelerated routes
    if dtype is None and out is None and fast:
        # call shapes already found eligible for the single-pass functor skip the layout analysis below
        hot_key = (a.dtype, a._shape, a._strides, tuple(axis) if isinstance(axis, list) else axis)
        hot_axis = _var_hot.get(hot_key)
        if hot_axis is not None:
            return _var_single_pass(a, hot_axis, ddof, keepdims)
    if axis is None:
        axis = tuple(range(a.ndim))
    if not isinstance(axis, tuple):
        axis = (axis,) if not isinstance(axis, list) else tuple(axis)
    dtype_mean = a.dtype
    if dtype is None:
        if a.dtype.kind in 'biu':
            dtype_mean = numpy.dtype('float64')
            dtype_out = numpy.dtype('float64')
        else:
            dtype_out = a.dtype
    else:
        dtype_out = numpy.dtype(dtype)
    reduce_axis, out_axis = _get_axis(axis, a.ndim)
    items = 1
    for ax in reduce_axis:
        items *= a.shape[ax]

    # ---- hot path: one read of `a` (the reference reads it twice)
    if fast and dtype is None and out is None and a.size > 0 and items > 0:
        from cupy_b200._core import _reduction
        layout = _reduction._classify(a.shape, a.strides, a.dtype.itemsize, reduce_axis, out_axis, False)
        if layout.kind >= 0:
            desc = _lib.ReduceDesc(_lib.OP_VAR, layout.kind, _scalar.dtype_id(a.dtype),
                                   _scalar.dtype_id(dtype_out), layout.batch, layout.n_reduce,
                                   layout.n_out, float(ddof))
            import ctypes
            if _lib.lib.b200_reduce_supported(ctypes.byref(desc)):
                if len(_var_hot) >= 512:
                    _var_hot.clear()
                _var_hot[hot_key] = axis
                return _var_single_pass(a, axis, ddof, keepdims)

    # ---- general path: the reference's algorithm (mean, then sum of squared deviations)
    div = max(items - ddof, 0)
    alpha = 1. / div if div != 0 else math.nan
    arrmean = a.mean(axis=axis, dtype=dtype_mean, out=None, keepdims=True)
    if out is None and dtype_out == numpy.float16:
        return _var_core_float16(a, arrmean, numpy.float32(alpha), axis=axis, keepdims=keepdims)
    if out is None:
        res = ndarray(_out_shape(a.shape, reduce_axis, out_axis, keepdims), dtype_out)
        _var_core_out(a, arrmean, numpy.asarray(alpha, dtype=dtype_out)[()], res, axis=axis, keepdims=keepdims)
        return res
    _var_core_out(a, arrmean, numpy.asarray(alpha, dtype=out.dtype)[()], out, axis=axis, keepdims=keepdims)
    return out.astype(dtype_out, copy=False)


def moments(a, out=None):
    """(n, mean, M2) of all elements of a dense array in ONE pass (B200_OP_MOMENTS): three
    float64 values (accumulated in float32 for float16/float32 inputs, float64 otherwise).
    M2 = sum((x - mean)^2); var = M2 / (n - ddof).  This is what a sharded
    variance exchanges (cupy_b200.distributed.sharded_var); the reference needs two passes
    (cupy/_core/_routines_statistics.pyx:556-600)."""
    import ctypes
    from cupy_b200._core import _reduction, _workspace, _dryrun
    from cupy_b200._core._kernel import current_stream_ptr
    if a.size == 0:
        raise ValueError('moments of an empty array')
    layout = _reduction._classify(a.shape, a.strides, a.dtype.itemsize, tuple(range(a.ndim)), (), False)
    if layout.kind != _lib.RED_FULL:
        a = a.copy()
        layout = _reduction._classify(a.shape, a.strides, a.dtype.itemsize, tuple(range(a.ndim)), (), False)
    ftype = numpy.dtype('float64')
    out = ndarray((3,), ftype) if out is None else out
    desc = _lib.ReduceDesc(_lib.OP_MOMENTS, layout.kind, _scalar.dtype_id(a.dtype), _scalar.dtype_id(ftype),
                           layout.batch, layout.n_reduce, layout.n_out, 0.0)
    if not _lib.lib.b200_reduce_supported(ctypes.byref(desc)):
        raise NotImplementedError('moments: dtype %s has no prebuilt kernel' % a.dtype)
    if _dryrun.enabled:
        _dryrun.record('prebuilt_reduce', name='cupy_moments', layout=layout.kind, batch=layout.batch,
                       n_reduce=layout.n_reduce, n_out=layout.n_out)
        return out
    st = current_stream_ptr()
    need = ctypes.c_size_t()
    _lib.check(_lib.lib.b200_reduce_workspace_bytes(ctypes.byref(desc), ctypes.byref(need)))
    ws_ptr, ws_bytes = _workspace.get(need.value, st)
    _lib.check(_lib.lib.b200_reduce_run(ctypes.byref(desc), a.ptr, out.ptr, ws_ptr, ws_bytes, st))
    return out


def _out_shape(shape, reduce_axis, out_axis, keepdims):
    from cupy_b200._core._reduction import _get_out_shape
    return _get_out_shape(shape, reduce_axis, out_axis, keepdims)


def _ndarray_var(self, axis, dtype, out, ddof, keepdims):
    return _var(self, axis=axis, dtype=dtype, out=out, ddof=ddof, keepdims=keepdims)


def _ndarray_std(self, axis, dtype, out, ddof, keepdims):
    ret = _var(self, axis=axis, dtype=dtype, out=None, ddof=ddof, keepdims=keepdims)
    return _math.sqrt(ret, dtype=dtype, out=out)


def amax(a, axis=None, out=None, keepdims=False):
    return _math._as_array(a).max(axis=axis, out=out, keepdims=keepdims)


def amin(a, axis=None, out=None, keepdims=False):
    return _math._as_array(a).min(axis=axis, out=out, keepdims=keepdims)


def argmax(a, axis=None, dtype=None, out=None, keepdims=False):
    return _math._as_array(a).argmax(axis=axis, dtype=dtype, out=out, keepdims=keepdims)


def argmin(a, axis=None, dtype=None, out=None, keepdims=False):
    return _math._as_array(a).argmin(axis=axis, dtype=dtype, out=out, keepdims=keepdims)


def mean(a, axis=None, dtype=None, out=None, keepdims=False):
    return _math._as_array(a).mean(axis=axis, dtype=dtype, out=out, keepdims=keepdims)


def var(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    return _math._as_array(a).var(axis=axis, dtype=dtype, out=out, ddof=ddof, keepdims=keepdims)


def std(a, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
    return _math._as_array(a).std(axis=axis, dtype=dtype, out=out, ddof=ddof, keepdims=keepdims)
