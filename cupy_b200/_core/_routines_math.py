"""Arithmetic ufunc table, sum / prod, and the cumsum / cumprod scan.

Mirror of the hot-path parts of cupy/_core/_routines_math.pyx: the ufunc
definitions (:878-1178; dtype loop tables and routine strings are the same
interface data, re-stated), `_ndarray_sum/_ndarray_prod` (:88-128),
`_sum_auto_dtype/_sum_keep_dtype` (:762-807) and `scan_core` (:702-751).
"""
from __future__ import annotations

import ctypes

import numpy

from cupy_b200 import _lib
from cupy_b200._core import _accelerator, _dryrun, _kernel, _scalar, _workspace
from cupy_b200._core._kernel import create_ufunc
from cupy_b200._core._ndarray import ndarray, normalize_axis_index, current_stream_ptr
from cupy_b200._core._reduction import create_reduction_func

_INT_LOOPS = ('bb->b', 'BB->B', 'hh->h', 'HH->H', 'ii->i', 'II->I', 'll->l', 'LL->L', 'qq->q', 'QQ->Q')
_INT_LOOPS1 = ('b->b', 'B->B', 'h->h', 'H->H', 'i->i', 'I->I', 'l->l', 'L->L', 'q->q', 'Q->Q')


def _create_arithmetic(name, op, boolop, doc='', scatter_op=None):
    if isinstance(boolop, str):
        boolop = 'out0 = in0 %s in1' % boolop
    return create_ufunc(
        'cupy_' + name,
        (('??->?', boolop),) + _INT_LOOPS + ('ee->e', 'ff->f', 'dd->d'),
        'out0 = in0 %s in1' % op, doc=doc, prebuilt=name, scatter_op=scatter_op)


def _subtract_boolean_error():
    raise TypeError('cupy boolean subtract, the `-` operator, is deprecated, use the '
                    'bitwise_xor, the `^` operator, or the logical_xor function instead.')


def _negative_boolean_error():
    raise TypeError('The cupy boolean negative, the `-` operator, is not supported, '
                    'use the `~` operator or the logical_not function instead.')


add = _create_arithmetic('add', '+', '|', 'Adds two arrays elementwise.', scatter_op='add')
subtract = _create_arithmetic('subtract', '-', _subtract_boolean_error, 'Subtracts arguments elementwise.',
                              scatter_op='sub')
multiply = _create_arithmetic('multiply', '*', '&', 'Multiplies two arrays elementwise.')

true_divide = create_ufunc(
    'cupy_true_divide',
    ('qq->d', 'qQ->d', 'Qq->d', 'QQ->d', 'ee->e', 'ff->f', 'dd->d'),
    'out0 = static_cast<out0_type>(in0) / static_cast<out0_type>(in1)',
    doc='Elementwise true division (i.e. division as floating values).',
    out_ops=('ee->e', 'ff->f', 'dd->d'), prebuilt='true_divide')
divide = true_divide

negative = create_ufunc(
    'cupy_negative', (('?->?', _negative_boolean_error),) + _INT_LOOPS1 + ('e->e', 'f->f', 'd->d'),
    'out0 = -in0', doc='Takes numerical negative elementwise.', prebuilt='negative')

absolute = create_ufunc(
    'cupy_absolute',
    (('?->?', 'out0 = in0'), 'b->b', ('B->B', 'out0 = in0'), 'h->h', ('H->H', 'out0 = in0'),
     'i->i', ('I->I', 'out0 = in0'), 'l->l', ('L->L', 'out0 = in0'), 'q->q', ('Q->Q', 'out0 = in0'),
     ('e->e', 'out0 = fabsf(in0)'), ('f->f', 'out0 = fabsf(in0)'), ('d->d', 'out0 = fabs(in0)')),
    'out0 = in0 > 0 ? in0 : -in0', doc='Elementwise absolute value function.', prebuilt='absolute')

square = create_ufunc(
    'cupy_square', _INT_LOOPS1 + ('e->e', 'f->f', 'd->d'), 'out0 = in0 * in0',
    doc='Elementwise square function.', prebuilt='square')

sqrt = create_ufunc('cupy_sqrt', ('e->e', 'f->f', 'd->d'), 'out0 = sqrt(in0)',
                    doc='Elementwise square root function.', prebuilt='sqrt')


def _create_math_ufunc(math_name, nargs, name, doc='', prebuilt=None):
    """cupy/_math/ufunc.py:7-20."""
    if nargs == 1:
        return create_ufunc(name, ('e->e', 'f->f', 'd->d'), 'out0 = %s(in0)' % math_name, doc=doc, prebuilt=prebuilt)
    return create_ufunc(name, ('ee->e', 'ff->f', 'dd->d'), 'out0 = %s(in0, in1)' % math_name, doc=doc, prebuilt=prebuilt)


exp = _create_math_ufunc('exp', 1, 'cupy_exp', 'Elementwise exponential function.', prebuilt='exp')
log = _create_math_ufunc('log', 1, 'cupy_log', 'Elementwise natural logarithm function.', prebuilt='log')
expm1 = _create_math_ufunc('expm1', 1, 'cupy_expm1')
exp2 = _create_math_ufunc('exp2', 1, 'cupy_exp2')
log2 = _create_math_ufunc('log2', 1, 'cupy_log2')
log10 = _create_math_ufunc('log10', 1, 'cupy_log10')
log1p = _create_math_ufunc('log1p', 1, 'cupy_log1p')
sin = _create_math_ufunc('sin', 1, 'cupy_sin')
cos = _create_math_ufunc('cos', 1, 'cupy_cos')
tan = _create_math_ufunc('tan', 1, 'cupy_tan')
tanh = _create_math_ufunc('tanh', 1, 'cupy_tanh')
sinh = _create_math_ufunc('sinh', 1, 'cupy_sinh')
cosh = _create_math_ufunc('cosh', 1, 'cupy_cosh')
arctan2 = _create_math_ufunc('atan2', 2, 'cupy_arctan2')
hypot = _create_math_ufunc('hypot', 2, 'cupy_hypot')

_float_maximum = 'out0 = (isnan(in0) | isnan(in1)) ? out0_type(NAN) : out0_type(max(in0, in1))'
_float_minimum = 'out0 = (isnan(in0) | isnan(in1)) ? out0_type(NAN) : out0_type(min(in0, in1))'
_float_preamble = '''
#ifndef NAN
#define NAN __int_as_float(0x7fffffff)
#endif
'''
maximum = create_ufunc(
    'cupy_maximum',
    ('??->?',) + _INT_LOOPS + (('ee->e', _float_maximum), ('ff->f', _float_maximum), ('dd->d', _float_maximum)),
    'out0 = max(in0, in1)', preamble=_float_preamble,
    doc='Takes the maximum of two arrays elementwise. If NaN appears, it returns the NaN.', prebuilt='maximum',
    scatter_op='max')
minimum = create_ufunc(
    'cupy_minimum',
    ('??->?',) + _INT_LOOPS + (('ee->e', _float_minimum), ('ff->f', _float_minimum), ('dd->d', _float_minimum)),
    'out0 = min(in0, in1)', preamble=_float_preamble,
    doc='Takes the minimum of two arrays elementwise. If NaN appears, it returns the NaN.', prebuilt='minimum',
    scatter_op='min')

power = create_ufunc(
    'cupy_power',
    ('??->b',) + _INT_LOOPS + (('ee->e', 'out0 = powf(in0, in1)'), ('ff->f', 'out0 = powf(in0, in1)'),
                               ('dd->d', 'out0 = pow(in0, in1)')),
    'out0 = integral_power(in0, in1)',
    preamble='''
template <typename T>
inline __device__ T integral_power(T in0, T in1) {
    if (in1 < 0) {
        if (in0 == -1) {return (in1 & 1) ? -1 : 1;}
        else {return (in0 == 1) ? 1 : 0;}
    }
    T out0 = 1;
    while (in1 > 0) {
        if (in1 & 1) out0 *= in0;
        in0 *= in0;
        in1 >>= 1;
    }
    return out0;
}
''', doc='Computes ``x1 ** x2`` elementwise.')

# fused multiply-add, one rounding (the FFMA the reference's JIT contracts a*x+y into)
fma = create_ufunc('cupy_fma', ('eee->e', 'fff->f', 'ddd->d'), 'out0 = fma(in0, in1, in2)',
                   doc='out = in0 * in1 + in2 with a single rounding.', prebuilt='fma')


def _create_comparison(name, op):
    return create_ufunc(
        'cupy_' + name,
        ('??->?', 'bb->?', 'BB->?', 'hh->?', 'HH->?', 'ii->?', 'II->?', 'll->?', 'LL->?', 'qq->?', 'QQ->?',
         'ee->?', 'ff->?', 'dd->?'),
        'out0 = in0 %s in1' % op)


greater = _create_comparison('greater', '>')
greater_equal = _create_comparison('greater_equal', '>=')
less = _create_comparison('less', '<')
less_equal = _create_comparison('less_equal', '<=')
equal = _create_comparison('equal', '==')
not_equal = _create_comparison('not_equal', '!=')


# ---- sum / prod ------------------------------------------------------------------------
_sumprod_types = (
    '?->l', 'b->l', 'B->L', 'h->l', 'H->L', 'i->l', 'I->L', 'l->l', 'L->L', 'q->q', 'Q->Q',
    ('e->e', (None, None, None, 'float')), 'f->f', 'd->d')
_keep_types = (
    '?->?', 'b->b', 'B->B', 'h->h', 'H->H', 'i->i', 'I->I', 'l->l', 'L->L', 'q->q', 'Q->Q',
    ('e->e', (None, None, None, 'float')), 'f->f', 'd->d')

_sum_auto_dtype = create_reduction_func(
    'cupy_sum', _sumprod_types, ('in0', 'a + b', 'out0 = type_out0_raw(a)', None), 0, prebuilt=_lib.OP_SUM)
_sum_keep_dtype = create_reduction_func(
    'cupy_sum_with_dtype', _keep_types, ('in0', 'a + b', 'out0 = type_out0_raw(a)', None), 0)
_prod_auto_dtype = create_reduction_func(
    'cupy_prod', _sumprod_types, ('in0', 'a * b', 'out0 = type_out0_raw(a)', None), 1, prebuilt=_lib.OP_PROD)
_prod_keep_dtype = create_reduction_func(
    'cupy_prod_with_dtype', _keep_types, ('in0', 'a * b', 'out0 = type_out0_raw(a)', None), 1)


def _ndarray_sum(self, axis, dtype, out, keepdims):
    if dtype is None:
        return _sum_auto_dtype(self, axis, dtype, out, keepdims)
    return _sum_keep_dtype(self, axis, dtype, out, keepdims)


def _ndarray_prod(self, axis, dtype, out, keepdims):
    if dtype is None:
        return _prod_auto_dtype(self, axis, dtype, out, keepdims)
    return _prod_keep_dtype(self, axis, dtype, out, keepdims)


def sum(a, axis=None, dtype=None, out=None, keepdims=False):
    """cupy.sum (cupy/_math/sumprod.py:13-42)."""
    return _as_array(a).sum(axis, dtype, out, keepdims)


def prod(a, axis=None, dtype=None, out=None, keepdims=False):
    return _as_array(a).prod(axis, dtype, out, keepdims)


def _as_array(a):
    if isinstance(a, ndarray) or hasattr(a, '__cupy_override_reduction_kernel__'):   # arrays, fusion variables
        return a
    from cupy_b200._core import _ndarray
    return _ndarray.asarray(a)


# ---- scan ------------------------------------------------------------------------------
def _scan_flat(src, dst, op):
    """Inclusive scan of the C-contiguous 1-D `src` into the C-contiguous `dst`."""
    n = src.size
    if n == 0:
        return
    st = current_stream_ptr()
    in_id, out_id = _scalar.dtype_id(src.dtype), _scalar.dtype_id(dst.dtype)
    if not _lib.lib.b200_scan_supported(op, in_id, out_id):
        # No prebuilt (in, out) pair: scan in the widest accumulator of the result's kind
        # (modular arithmetic commutes with the final truncation) and cast once at the end.
        kind = dst.dtype.kind
        wide = numpy.dtype({'i': 'int64', 'b': 'int64', 'u': 'uint64'}.get(kind, 'float64' if dst.dtype.itemsize == 8 else 'float32'))
        if _lib.lib.b200_scan_supported(op, in_id, _scalar.dtype_id(wide)):
            tmp = ndarray(dst.shape, wide)
            _scan_flat(src, tmp, op)
        else:
            tmp = ndarray(dst.shape, wide)
            _scan_flat(src.astype(wide), tmp, op)
        _kernel.elementwise_copy(tmp, dst)
        return
    if src.ptr % 16 or dst.ptr % 16:
        tmp_in = src.copy() if src.ptr % 16 else src
        if dst.ptr % 16:
            tmp_out = ndarray(dst.shape, dst.dtype)
            _scan_flat(tmp_in, tmp_out, op)
            _kernel.elementwise_copy(tmp_out, dst)
            return
        return _scan_flat(tmp_in, dst, op)
    if _dryrun.enabled:
        _dryrun.record('prebuilt_scan', op=op, n=n, in_dtype=src.dtype.name, out_dtype=dst.dtype.name)
        return
    need = ctypes.c_size_t()
    _lib.check(_lib.lib.b200_scan_workspace_bytes(n, out_id, ctypes.byref(need)))
    ws_ptr, ws_bytes = _workspace.get(16384 + need.value, st)
    _lib.check(_lib.lib.b200_scan_run(op, in_id, out_id, src.ptr, dst.ptr, n, ws_ptr + 16384, ws_bytes - 16384, st))


def scan_core(a, axis, op, dtype=None, out=None):
    """cupy/_core/_routines_math.pyx:702-751 (dtype rules :704-714)."""
    a = _as_array(a)
    if _accelerator.reference_first(routine=True):
        r = _accelerator.try_reference('scan', 'cumsum' if op == _lib.OP_CUMSUM else 'cumprod', a,
                                       axis=axis, dtype=dtype, out=out)
        if r is not None:
            return r
    # ---- memoised call shape (dense input, fresh output, prebuilt pair): what the rest of this function derives
    # depends only on (dtype, shape, strides, alignment, axis, dtype=, op)
    mkey = None
    if out is None and type(a) is ndarray and a._c_contiguous and a.size and _accelerator.fast_paths_enabled():
        mkey = (a.dtype, a._shape, a.ptr & 15, axis, dtype, op, _kernel._memo_epoch)
        memo = _kernel._thread_local.__dict__.setdefault('scan_memo', {})
        e = memo.get(mkey)
        if e is not None and not _dryrun.enabled:
            in_id, out_id, oshape, odtype, ostrides, geom, need = e
            res = ndarray._fresh(oshape, odtype, ostrides, a.size)
            st = current_stream_ptr()
            if geom is None:
                ws_ptr, ws_bytes = _workspace.get(16384 + need, st)
                _lib.check(_lib.lib.b200_scan_run(op, in_id, out_id, a.ptr, res.ptr, a.size, ws_ptr + 16384,
                                                  ws_bytes - 16384, st))
            else:
                ws_ptr = 0
                if need:
                    scratch = ndarray._fresh((need,), _U8, (1,), need)
                    ws_ptr = scratch.ptr
                _lib.check(_lib.lib.b200_scan_axis_run(op, in_id, out_id, a.ptr, res.ptr, geom[0], geom[1], geom[2],
                                                       ws_ptr, need, st))
            return res
    if out is None:
        if dtype is None:
            kind = a.dtype.kind
            if kind in 'bi':
                dtype = numpy.dtype('int64')
            elif kind == 'u':
                dtype = numpy.dtype('uint64')
            else:
                dtype = a.dtype
        dtype = _scalar.get_dtype(dtype)
    else:
        if not isinstance(out, ndarray):
            raise TypeError('Output arguments type must be cupy.ndarray')
        dtype = out.dtype
    if axis is None:
        if out is not None and out.size != a.size:
            raise ValueError('Provided out is the wrong size for the reduction')
        src = a if a._c_contiguous else a.copy()
        src = src.reshape(-1)
        if out is not None and (out._c_contiguous or out._f_contiguous):
            result = out
            _scan_flat(src, result.reshape(-1) if out._c_contiguous else result.T.reshape(-1), op)
            return out
        result = ndarray((a.size,), dtype)
        _scan_flat(src, result, op)
        if mkey is not None and src.ptr % 16 == 0 and result.ptr % 16 == 0 and not _dryrun.enabled:
            in_id, out_id = _scalar.dtype_id(src.dtype), _scalar.dtype_id(result.dtype)
            if _lib.lib.b200_scan_supported(op, in_id, out_id):
                need = ctypes.c_size_t()
                _lib.check(_lib.lib.b200_scan_workspace_bytes(a.size, out_id, ctypes.byref(need)))
                _remember_scan(mkey, (in_id, out_id, result._shape, result.dtype, result._strides, None, need.value))
        if out is not None:
            _kernel.elementwise_copy(result.reshape(out.shape), out)
            return out
        return result
    axis = normalize_axis_index(axis, a.ndim)
    res = _scan_axis(a, axis, op, dtype, out)
    if mkey is not None and not _dryrun.enabled:
        in_id, out_id = _scalar.dtype_id(a.dtype), _scalar.dtype_id(dtype)
        if _lib.lib.b200_scan_supported(op, in_id, out_id) and res._c_contiguous:
            geom = (_kernel._prod(a.shape[:axis]), int(a.shape[axis]), _kernel._prod(a.shape[axis + 1:]))
            need = ctypes.c_size_t()
            _lib.check(_lib.lib.b200_scan_axis_workspace_bytes(geom[0], geom[1], geom[2], ctypes.byref(need)))
            _remember_scan(mkey, (in_id, out_id, res._shape, res.dtype, res._strides, geom, need.value))
    return res


_U8 = numpy.dtype('uint8')


def _remember_scan(mkey, entry):
    memo = _kernel._thread_local.__dict__.setdefault('scan_memo', {})
    if len(memo) >= 512:
        memo.clear()
    memo[mkey] = entry


def _scan_axis(a, axis, op, dtype, out):
    """Scan along one axis: every line is scanned with a per-line restart.

    The line restart is expressed as a segmented scan over the flattened array
    with the scanned axis made innermost (the reference does the same roll +
    reshape, cupy/_core/_routines_math.pyx:692-699)."""
    from cupy_b200._core import _scan_axis as impl
    return impl.scan_axis(a, axis, op, dtype, out)


def cumsum(a, axis=None, dtype=None, out=None):
    """cupy.cumsum (cupy/_math/sumprod.py:145-161)."""
    return scan_core(a, axis, _lib.OP_CUMSUM, dtype, out)


def cumprod(a, axis=None, dtype=None, out=None):
    return scan_core(a, axis, _lib.OP_CUMPROD, dtype, out)
