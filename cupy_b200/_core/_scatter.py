"""`ufunc.at` and `add.reduceat`: the two remaining methods of the reference's ufunc class.

Reference: `ufunc.at` -> `ndarray._scatter_op(indices, b, op)` (cupy/_core/_kernel.pyx:1446-1457,
cupy/_core/_routines_indexing.pyx:899-1022, 1064-1084: an ElementwiseKernel over the broadcast values that
applies one atomic per element to `a[(l * adim + index) * rdim + r]`), and `add.reduceat` ->
`_add_reduceat` (cupy/_core/_routines_indexing.pyx:1255-1270: differences of the inclusive scan at the
segment ends, `a[indices[i]]` where a segment is empty).

Built here from the pieces this package already has: the scan kernels for the prefix sums and one NVRTC
ElementwiseKernel per operation with the destination as a `raw` operand; an atomic whose result is unused
compiles to a fire-and-forget `RED` on sm_100a.  Index forms: one integer array (applied along axis 0), a
tuple of integer arrays (one per leading axis, broadcast together), or a boolean mask of `a`'s leading shape.
"""
from __future__ import annotations

import numpy as np

_ATOMICS = r'''
// atomics for the scatter kernels: one overload set per operation, keyed by the destination type
template <typename T> struct b200_bits;
template <> struct b200_bits<float>  { typedef unsigned int type; };
template <> struct b200_bits<double> { typedef unsigned long long type; };

__device__ inline void scat_add(int* p, int v) { atomicAdd(p, v); }
__device__ inline void scat_add(unsigned int* p, unsigned int v) { atomicAdd(p, v); }
__device__ inline void scat_add(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
__device__ inline void scat_add(long long* p, long long v) {      // two's complement: same bits as unsigned add
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}
__device__ inline void scat_add(float* p, float v) { atomicAdd(p, v); }
__device__ inline void scat_add(double* p, double v) { atomicAdd(p, v); }
__device__ inline void scat_add(float16* p, float16 v) { atomicAdd(reinterpret_cast<__half*>(p), v.raw()); }
__device__ inline void scat_sub(int* p, int v) { atomicSub(p, v); }
__device__ inline void scat_sub(unsigned int* p, unsigned int v) { atomicSub(p, v); }

#define B200_SCAT_INT(NAME, FN)                                                                          \
__device__ inline void NAME(int* p, int v) { FN(p, v); }                                                 \
__device__ inline void NAME(unsigned int* p, unsigned int v) { FN(p, v); }                               \
__device__ inline void NAME(long long* p, long long v) { FN(p, v); }                                     \
__device__ inline void NAME(unsigned long long* p, unsigned long long v) { FN(p, v); }
B200_SCAT_INT(scat_max, atomicMax)
B200_SCAT_INT(scat_min, atomicMin)
B200_SCAT_INT(scat_and, atomicAnd)
B200_SCAT_INT(scat_or, atomicOr)
B200_SCAT_INT(scat_xor, atomicXor)

// floating max / min by compare-and-swap; a NaN value wins and then stays (numpy.maximum / minimum)
template <typename T, bool IS_MAX>
__device__ inline void scat_minmax_fp(T* p, T v) {
    typedef typename b200_bits<T>::type U;
    U* up = reinterpret_cast<U*>(p);
    U old = *up;
    for (;;) {
        T cur;
        memcpy(&cur, &old, sizeof(T));
        const bool replace = (cur == cur) && ((v != v) || (IS_MAX ? (v > cur) : (v < cur)));
        if (!replace) return;
        U want;
        memcpy(&want, &v, sizeof(T));
        const U seen = atomicCAS(up, old, want);
        if (seen == old) return;
        old = seen;
    }
}
__device__ inline void scat_max(float* p, float v) { scat_minmax_fp<float, true>(p, v); }
__device__ inline void scat_max(double* p, double v) { scat_minmax_fp<double, true>(p, v); }
__device__ inline void scat_min(float* p, float v) { scat_minmax_fp<float, false>(p, v); }
__device__ inline void scat_min(double* p, double v) { scat_minmax_fp<double, false>(p, v); }
'''

_I32 = (np.int32, np.uint32)
_I3264 = (np.int32, np.int64, np.uint32, np.uint64)
_SUPPORTED = {     # the dtype gates of _scatter_op_single (_routines_indexing.pyx:942-1018)
    'add': _I3264 + (np.float16, np.float32, np.float64),
    'sub': _I32,
    'max': _I3264 + (np.float32, np.float64),
    'min': _I3264 + (np.float32, np.float64),
    'and': _I3264, 'or': _I3264, 'xor': _I3264,
}
_UFUNC_NAME = {'add': 'add', 'sub': 'subtract', 'max': 'maximum', 'min': 'minimum',
               'and': 'bitwise_and', 'or': 'bitwise_or', 'xor': 'bitwise_xor'}
_kernels = {}


def _kernel(op):
    k = _kernels.get(op)
    if k is None:
        from cupy_b200._core._kernel import ElementwiseKernel
        k = _kernels[op] = ElementwiseKernel(
            'T v, S indices, int64 cdim, int64 rdim, int64 adim', 'raw T a',
            '''
            ptrdiff_t at = indices;
            if (at < 0) at += adim;            // negative indices count from the end
            const ptrdiff_t li = i / (rdim * cdim);
            const ptrdiff_t ri = i %% rdim;
            scat_%s(&a[(li * adim + at) * rdim + ri], v);
            ''' % op, 'cupy_scatter_' + op, preamble=_ATOMICS)
    return k


def _normalize_index(a, indices):
    """-> (flat integer index array over the first `stop` axes of `a`, stop)."""
    from cupy_b200._core import _ndarray as nd
    from cupy_b200._core._kernel import _broadcast_core
    if not isinstance(indices, tuple):
        indices = (indices,)
    idx = []
    for s in indices:
        if isinstance(s, (slice, type(Ellipsis))) or s is None:
            raise NotImplementedError('ufunc.at takes integer arrays or one boolean mask here (no slices)')
        idx.append(s if isinstance(s, nd.ndarray) else nd.asarray(np.asarray(s)))
    if len(idx) == 1 and idx[0].dtype == np.bool_:
        mask = idx[0]
        if mask.shape != a.shape[:mask.ndim]:
            raise IndexError('boolean index did not match indexed array')
        # positions of the True entries in C order: rank by inclusive scan, then scatter the position
        from cupy_b200._core._kernel import ElementwiseKernel
        from cupy_b200._core import _routines_math as rm
        flat = mask.ravel()
        rank = rm.cumsum(flat)
        n_true = int(rank[-1].item()) if flat.size else 0
        pos = nd.empty((n_true,), np.int64)
        if n_true:
            k = _kernels.get('_nonzero')
            if k is None:
                k = _kernels['_nonzero'] = ElementwiseKernel(
                    'bool m, int64 rank', 'raw int64 pos', 'if (m) pos[rank - 1] = i;', 'cupy_scatter_mask_positions')
            k(flat, rank, pos)
        return pos, mask.ndim
    if len(idx) > a.ndim:
        raise IndexError('too many indices for array')
    for s in idx:
        if s.dtype.kind not in 'iu':
            raise IndexError('arrays used as indices must be of integer (or boolean) type')
    if len(idx) == 1:
        return idx[0], 1
    # several index arrays: fold them into one index over the flattened leading axes, wrapping negatives
    from cupy_b200._core._kernel import ElementwiseKernel
    arrs = list(idx)
    shape = _broadcast_core(arrs)
    flat = nd.zeros(shape, np.int64)
    k = _kernels.get('_fold')
    if k is None:
        k = _kernels['_fold'] = ElementwiseKernel(
            'S s, int64 dim', 'int64 flat', 'flat = flat * dim + (s < 0 ? s + dim : s);', 'cupy_scatter_fold_index')
    for ax, s in enumerate(arrs):
        k(s, a.shape[ax], flat)
    return flat, len(idx)


def scatter_op(a, indices, value, op):
    """a[indices] op= value with repeated indices accumulated (numpy.ufunc.at)."""
    from cupy_b200._core import _ndarray as nd
    if not isinstance(a, nd.ndarray):
        raise TypeError('ufunc.at needs a cupy_b200.ndarray as its first operand')
    if value is None:
        raise ValueError('second operand needed for ufunc')
    if a.dtype.type not in _SUPPORTED[op]:
        names = sorted({np.dtype(t).name for t in _SUPPORTED[op]})
        raise TypeError('cupy.%s.at only supports %s as data type' % (_UFUNC_NAME[op], ', '.join(names)))
    index, stop = _normalize_index(a, indices)
    v = value.astype(a.dtype, copy=False) if isinstance(value, nd.ndarray) else nd.asarray(np.asarray(value, a.dtype))
    rshape = a.shape[stop:]
    adim = int(np.prod(a.shape[:stop], dtype=np.int64))
    rdim = int(np.prod(rshape, dtype=np.int64))
    v_shape = index.shape + rshape
    if int(np.prod(v_shape, dtype=np.int64)) == 0:
        return
    if adim == 0:
        raise IndexError('index out of bounds for an axis of size 0')
    cdim = index.size
    v = v.broadcast_to(v_shape)
    index = index.reshape(index.shape + (1,) * len(rshape)).broadcast_to(v_shape)
    _kernel(op)(v, index, cdim, rdim, adim, a)


def add_reduceat(array, indices, axis, dtype, out):
    """numpy.add.reduceat: out[i] = sum(array[indices[i]:indices[i+1]]) along `axis` (to the end for the last
    index), array[indices[i]] where indices[i] >= indices[i+1]."""
    from cupy_b200._core import _ndarray as nd
    from cupy_b200._core._kernel import ElementwiseKernel
    from cupy_b200._core import _routines_math as rm
    nd_axis = axis + array.ndim if axis < 0 else axis
    if not 0 <= nd_axis < array.ndim:
        raise np.exceptions.AxisError(axis, array.ndim)
    host_idx = indices.get() if isinstance(indices, nd.ndarray) else np.asarray(indices)
    if host_idx.ndim != 1 or host_idx.dtype.kind not in 'iu':
        raise TypeError('indices must be a 1-d integer array')
    n = array.shape[nd_axis]
    if host_idx.size and (host_idx.min() < 0 or host_idx.max() >= n):
        raise IndexError('index out of bounds for axis %d with size %d' % (nd_axis, n))
    acc = rm.cumsum(array, nd_axis, dtype)
    lo = nd.asarray(host_idx.astype(np.int64))
    hi = nd.asarray(np.append(host_idx[1:], n).astype(np.int64))
    res_shape = array.shape[:nd_axis] + (host_idx.size,) + array.shape[nd_axis + 1:]
    inner = int(np.prod(array.shape[nd_axis + 1:], dtype=np.int64))
    res = nd.empty(res_shape, acc.dtype)
    if res.size:
        k = _kernels.get('_reduceat')
        if k is None:
            k = _kernels['_reduceat'] = ElementwiseKernel(
                'raw T acc, raw X x, raw int64 lo, raw int64 hi, int64 m, int64 n, int64 inner', 'T y',
                '''
                const ptrdiff_t r = i % inner;
                const ptrdiff_t s = (i / inner) % m;
                const ptrdiff_t l = i / (inner * m);
                const ptrdiff_t row = l * n;
                const ptrdiff_t b = lo[s], e = hi[s];
                if (b >= e) {
                    y = static_cast<T>(x[(row + b) * inner + r]);
                } else {
                    const T upper = acc[(row + e - 1) * inner + r];
                    y = b ? static_cast<T>(upper - acc[(row + b - 1) * inner + r]) : upper;
                }
                ''', 'cupy_add_reduceat')
        k(acc, array, lo, hi, host_idx.size, n, inner, res)
    if out is None:
        return res
    from cupy_b200._core._kernel import elementwise_copy
    if out.shape != res.shape:
        raise ValueError('output parameter has the wrong shape')
    elementwise_copy(res, out)
    return out
