"""Stream compaction and gather: the callers the scan was built for.

`nonzero / argwhere / flatnonzero / where(condition)`, boolean-mask `a[mask]` and `a[mask] = v`,
`compress / extract`, and integer-array `take` / `a[indices]` / `a[indices] = v`.

Reference: cupy/_core/_routines_indexing.pyx -- `_ndarray_argwhere` / `_ndarray_nonzero` (:78-141: not_equal,
inclusive scan of the flags in int32 (int64 past 2^31 - 1 elements), ONE device-to-host read of the last rank for
the output size, then an ElementwiseKernel that writes each hit's coordinates at `rank - 1`), `_getitem_mask_single`
/ `_prepare_mask_indexing_single` (:755-840), `_scatter_op_mask_single` (:1024-1047), `_take` (:143-222, 562-580),
`_scatter_op_single` 'update' (:899-941); cupy/_indexing/indexing.py (`compress`, `extract`, `take`),
cupy/_sorting/search.py:175-231 (`nonzero`, `flatnonzero`, `argwhere`, one-argument `where`).

Here: the flags are ranked by the TMA-pipelined scan (b200/scan_pipe.cuh, bool -> int32 pair: 5 bytes per element),
and one NVRTC ElementwiseKernel with a `raw` destination does the scatter / gather -- the same two launches and the
same single synchronisation as the reference.
"""
from __future__ import annotations

import numpy

from cupy_b200._core import _routines_math as _math

_kernels = {}


def _k(name, *spec, **kw):
    k = _kernels.get(name)
    if k is None:
        from cupy_b200._core._kernel import ElementwiseKernel
        k = _kernels[name] = ElementwiseKernel(*spec, name, **kw)
    return k


def _flags(a):
    """C-ordered flat boolean flags of `a != 0`."""
    a = _math._as_array(a)
    f = a if a.dtype == numpy.bool_ else _math.not_equal(a, 0)
    return f.ravel()


def _rank(flags):
    """-> (inclusive scan of the flat flags, number of hits).  The one synchronisation of a compaction."""
    if flags.size == 0:
        return None, 0
    dtype = numpy.int32 if flags.size <= 2 ** 31 - 1 else numpy.int64
    rank = _math.cumsum(flags, dtype=dtype)
    return rank, int(rank[-1].item())


def argwhere(a):
    """(hits, ndim) int64 coordinates of the non-zero elements, in C order."""
    from cupy_b200._core._ndarray import empty
    a = _math._as_array(a)
    flags = _flags(a)
    rank, n = _rank(flags)
    dst = empty((n, a.ndim), numpy.int64)
    if dst.size == 0:
        return dst
    if a.ndim == 1:
        _k('cupy_nonzero_1d', 'bool m, S rank', 'raw int64 dst', 'if (m) dst[rank - 1] = i;')(flags, rank, dst)
        return dst
    # `_ind` presents the un-collapsed shape: the coordinates of element i (reduce_dims=False)
    _k('cupy_nonzero_kernel', 'bool m, S rank', 'raw int64 dst',
       'if (m) { for (int j = 0; j < _ind.ndim; j++) { dst[(rank - 1) * _ind.ndim + j] = _ind.get()[j]; } }',
       reduce_dims=False)(flags.reshape(a.shape), rank.reshape(a.shape), dst)
    return dst


def nonzero(a):
    """Tuple of index arrays, one per axis (views of one (hits, ndim) array, like the reference)."""
    a = _math._as_array(a)
    if a.ndim == 0:
        raise ValueError('Calling nonzero on 0d arrays is not allowed. Use cp.atleast_1d(scalar).nonzero() instead.')
    dst = argwhere(a)
    return tuple(dst[:, i] for i in range(a.ndim))


def flatnonzero(a):
    return nonzero(_math._as_array(a).ravel())[0]


# ---- boolean masks -----------------------------------------------------------------------------
def _prepare_mask(a, mask):
    """-> (mask broadcast to the indexed part of a's shape, its rank, shape of a[mask])."""
    if mask.ndim > a.ndim:
        raise IndexError('too many indices for array')
    for i, m in enumerate(mask.shape):
        if m not in (0, a.shape[i]):
            raise IndexError('boolean index did not match indexed array along dimension %d; dimension is %d but '
                             'corresponding boolean dimension is %d' % (i, a.shape[i], m))
    rshape = a.shape[mask.ndim:]
    if mask.size == 0:
        return None, None, (0,) + rshape
    if mask.ndim < a.ndim:
        # a mask over the leading axes selects whole sub-arrays: rank the mask broadcast over the rest
        mask = mask.reshape(mask.shape + (1,) * len(rshape)).broadcast_to(a.shape)
        flat = mask.ravel()
        rank, n = _rank(flat)
        inner = int(numpy.prod(rshape, dtype=numpy.int64))
        return mask, rank.reshape(a.shape), (n // inner if inner else 0,) + rshape
    rank, n = _rank(mask.ravel())
    return mask, rank.reshape(a.shape), (n,)


def getitem_mask(a, mask):
    from cupy_b200._core._ndarray import empty
    mask, rank, shape = _prepare_mask(a, mask)
    out = empty(shape, a.dtype)
    if out.size == 0:
        return out
    _k('cupy_getitem_mask', 'T a, bool mask, S mask_scanned', 'raw T out', 'if (mask) out[mask_scanned - 1] = a')(
        a, mask, rank, out)
    return out


def setitem_mask(a, mask, value):
    """a[mask] = value: a scalar / broadcastable array fills under the mask (one masked copy); an array with one
    element per hit is gathered by rank."""
    from cupy_b200._core._ndarray import ndarray, asarray
    if mask.ndim > a.ndim:
        raise IndexError('too many indices for array')
    rshape = a.shape[mask.ndim:]
    full_mask = mask if mask.ndim == a.ndim else mask.reshape(mask.shape + (1,) * len(rshape)).broadcast_to(a.shape)
    fill = _k('cupy_fill_mask', 'T v, bool mask', 'T a', 'if (mask) a = v')
    if not isinstance(value, ndarray):
        host = numpy.asarray(value)
        if host.ndim == 0:
            # one number travels by value: no device allocation, no copy
            fill(host.astype(a.dtype)[()], full_mask, a)
            return
        value = asarray(host)
    if value.ndim <= len(rshape) or value.size == 1:
        # the value broadcasts against the trailing axes (or is one number): no ranking needed
        v = value if value.dtype == a.dtype else value.astype(a.dtype)
        fill(v, full_mask, a)
        return
    mask_b, rank, shape = _prepare_mask(a, mask)
    if value.shape != shape:
        if value.ndim == 1 + len(rshape) and value.shape[0] == 1:
            value = value.broadcast_to(shape)
        else:
            raise ValueError('NumPy boolean array indexing assignment cannot assign %d input values to the %d output '
                             'values where the mask is true' % (value.size, int(numpy.prod(shape, dtype=numpy.int64))))
    if mask_b is None:
        return
    v = value if value._c_contiguous else value.copy()
    # the cast to a's dtype happens in the kernel: no `astype` pass over the values
    _k('cupy_setitem_mask', 'raw V v, bool mask, S mask_scanned', 'T a', 'if (mask) a = v[mask_scanned - 1]')(
        v, mask_b, rank, a)


def compress(condition, a, axis=None, out=None):
    """Slices of `a` along `axis` where the 1-D `condition` holds."""
    a = _math._as_array(a)
    condition = _math._as_array(condition)
    if condition.ndim != 1:
        raise ValueError('condition must be an 1-d array')
    if condition.dtype != numpy.bool_:
        condition = _math.not_equal(condition, 0)
    res = take(a, flatnonzero(condition), axis=axis)
    if out is None:
        return res
    from cupy_b200._core import _kernel
    _kernel.elementwise_copy(res, out)
    return out


def extract(condition, a):
    """ravel(a)[ravel(condition) != 0]"""
    a = _math._as_array(a)
    condition = _math._as_array(condition)
    if condition.shape != a.shape:
        raise ValueError('Shape mismatch: condition and a must have the same shape')
    if condition.dtype != numpy.bool_:
        condition = _math.not_equal(condition, 0)
    return getitem_mask(a.ravel(), condition.ravel())


# ---- integer arrays ------------------------------------------------------------------------------
def take(a, indices, axis=None, out=None):
    """a[..., indices, ...] along `axis` (the flattened array when None); negative indices wrap once, out-of-range
    ones wrap around like the reference's (`mode='wrap'`-like, no bounds check on the device)."""
    from cupy_b200._core._ndarray import ndarray, asarray, empty, normalize_axis_index
    a = _math._as_array(a)
    if not isinstance(indices, ndarray):
        indices = asarray(numpy.asarray(indices))
    if indices.dtype.kind not in 'iu':
        raise IndexError('arrays used as indices must be of integer (or boolean) type')
    if axis is None:
        a = a.ravel()
        axis = 0
    elif a.ndim == 0:
        normalize_axis_index(axis, 1)
        a = a.reshape(1)
        axis = 0
    else:
        axis = normalize_axis_index(axis, a.ndim)
    lshape, rshape, adim = a.shape[:axis], a.shape[axis + 1:], a.shape[axis]
    shape = lshape + indices.shape + rshape
    if out is None:
        res = empty(shape, a.dtype)
    else:
        if out.dtype != a.dtype:
            raise TypeError('Output dtype mismatch')
        if out.shape != shape:
            raise ValueError('Output shape mismatch')
        res = out
    if res.size == 0:
        return res
    if adim == 0:
        raise IndexError('cannot do a non-empty take from an empty axes.')
    if not a._c_contiguous:
        a = a.copy()
    cdim = indices.size
    rdim = int(numpy.prod(rshape, dtype=numpy.int64))
    idx = indices.reshape((1,) * len(lshape) + indices.shape + (1,) * len(rshape)).broadcast_to(shape)
    _k('cupy_take', 'raw T a, S indices, int64 cdim, int64 rdim, int64 adim', 'T out',
       '''
       ptrdiff_t at = indices % adim;
       if (at < 0) at += adim;
       const ptrdiff_t li = i / (rdim * cdim);
       const ptrdiff_t ri = i % rdim;
       out = a[(li * adim + at) * rdim + ri];
       ''')(a, idx, cdim, rdim, adim, res)
    return res


def _is_index_array(k):
    from cupy_b200._core._ndarray import ndarray
    return isinstance(k, (ndarray, numpy.ndarray, list))


def _as_index(k):
    from cupy_b200._core._ndarray import ndarray, asarray
    return k if isinstance(k, ndarray) else asarray(numpy.asarray(k))


def getitem_advanced(a, key):
    """a[key] where `key` holds index arrays: one boolean mask over the leading axes, or integer arrays for the
    leading axes (broadcast together).  Slices mixed in are not part of this path."""
    from cupy_b200._core import _scatter
    if not isinstance(key, tuple):
        key = (key,)
    if any(not _is_index_array(k) for k in key):
        raise NotImplementedError('index arrays mixed with slices / integers are outside this path: '
                                  'index the leading axes with arrays only, or use take(..., axis=)')
    idx = [_as_index(k) for k in key]
    if len(idx) == 1 and idx[0].dtype == numpy.bool_:
        return getitem_mask(a, idx[0])
    if any(s.dtype == numpy.bool_ for s in idx):
        raise NotImplementedError('boolean arrays mixed with other index arrays are outside this path')
    if len(idx) == 1:
        return take(a, idx[0], axis=0)
    flat, stop = _scatter._normalize_index(a, tuple(idx))
    lead = int(numpy.prod(a.shape[:stop], dtype=numpy.int64))
    src = a if a._c_contiguous else a.copy()
    return take(src.reshape((lead,) + a.shape[stop:]), flat, axis=0)


def setitem_advanced(a, key, value):
    """a[key] = value for the index forms of `getitem_advanced` (repeated indices: one of the values wins)."""
    from cupy_b200._core import _scatter
    from cupy_b200._core._ndarray import ndarray, asarray
    if not isinstance(key, tuple):
        key = (key,)
    if any(not _is_index_array(k) for k in key):
        raise NotImplementedError('index arrays mixed with slices / integers are outside this path')
    idx = [_as_index(k) for k in key]
    if len(idx) == 1 and idx[0].dtype == numpy.bool_:
        return setitem_mask(a, idx[0], value)
    if any(s.dtype == numpy.bool_ for s in idx):
        raise NotImplementedError('boolean arrays mixed with other index arrays are outside this path')
    index, stop = _scatter._normalize_index(a, tuple(idx) if len(idx) > 1 else idx[0])
    v = value.astype(a.dtype, copy=False) if isinstance(value, ndarray) else asarray(numpy.asarray(value, a.dtype))
    rshape = a.shape[stop:]
    adim = int(numpy.prod(a.shape[:stop], dtype=numpy.int64))
    rdim = int(numpy.prod(rshape, dtype=numpy.int64))
    v_shape = index.shape + rshape
    if int(numpy.prod(v_shape, dtype=numpy.int64)) == 0:
        return
    if adim == 0:
        raise IndexError('index out of bounds for an axis of size 0')
    if not a._c_contiguous:
        raise NotImplementedError('assignment through index arrays needs a C-contiguous destination')
    v = v.broadcast_to(v_shape)
    index = index.reshape(index.shape + (1,) * len(rshape)).broadcast_to(v_shape)
    _k('cupy_scatter_update', 'T v, S indices, int64 cdim, int64 rdim, int64 adim', 'raw T a',
       '''
       ptrdiff_t at = indices;
       if (at < 0) at += adim;
       const ptrdiff_t li = i / (rdim * cdim);
       const ptrdiff_t ri = i % rdim;
       a[(li * adim + at) * rdim + ri] = v;
       ''')(v, index, index.size, rdim, adim, a)
