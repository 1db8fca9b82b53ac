"""The reduction engine: `_SimpleReductionKernel` (create_reduction_func) and
`ReductionKernel`.

Host-side mirror of cupy/_core/_reduction.pyx -- same classes, signatures and
errors (`_get_axis` :147-161, `_get_out_shape` :164-176,
`_AbstractReductionKernel._call` :309-427, `_SimpleReductionKernel` :565-684,
`ReductionKernel` :716-917) -- over the B200 C ABI.

What differs is the launch.  The reference transposes the reduce axes to the
front and runs one generic shared-memory tree (`_reduction.pyx:44-126`), or hands
C/F-contiguous cases to CUB (`_cub_reduction.pyx`, `cupy/cuda/cub.pyx`).  Here the
input's memory order is analysed once (`_classify`) into FULL / ROWS / COLS /
GENERIC and

* built-in reductions with a prebuilt functor go to `b200_reduce_run`;
* everything else (other dtype loops, `dtype=`, user ReductionKernels) gets its
  (map, reduce, post_map, identity) strings wrapped into a functor and compiled
  by NVRTC against the SAME skeleton kernels (b200/reduce.cuh).
"""
from __future__ import annotations

import ctypes

from cupy_b200 import _lib
from cupy_b200._core import _accelerator, _codegen_reduce, _dryrun, _jit, _kernel, _scalar, _workspace
from cupy_b200._core._kernel import (_broadcast, _decide_params_type_core,
                                     _get_param_info, _preprocess_args, _stream_ptr)
from cupy_b200._core._ndarray import ndarray, normalize_axis_index, current_stream_ptr
from cupy_b200._core._scalar import CScalar, get_dtype, get_typename


def _get_axis(axis, ndim):
    """cupy/_core/_reduction.pyx:147-161."""
    if axis is None:
        return tuple(range(ndim)), ()
    if isinstance(axis, (tuple, list)):
        axis = tuple(axis)
    else:
        axis = (axis,)
    reduce_axis = tuple(sorted(normalize_axis_index(int(d), ndim) for d in axis))
    out_axis = tuple(d for d in range(ndim) if d not in reduce_axis)
    if len(reduce_axis) + len(out_axis) != ndim:
        raise ValueError("duplicate value in 'axis'")
    return reduce_axis, out_axis


def _get_out_shape(shape, reduce_axis, out_axis, keepdims):
    if keepdims:
        out_shape = list(shape)
        for i in reduce_axis:
            out_shape[i] = 1
        return tuple(out_shape)
    return tuple(shape[i] for i in out_axis)


def _prod(seq):
    r = 1
    for s in seq:
        r *= int(s)
    return r


class Layout:
    """Result of `_classify`: how the reduced and kept axes sit in memory."""
    __slots__ = ('kind', 'batch', 'n_reduce', 'n_out')

    def __init__(self, kind, batch, n_reduce, n_out):
        self.kind, self.batch, self.n_reduce, self.n_out = kind, batch, n_reduce, n_out

    def __repr__(self):
        names = {_lib.RED_FULL: 'FULL', _lib.RED_ROWS: 'ROWS', _lib.RED_COLS: 'COLS', -1: 'GENERIC'}
        return 'Layout(%s, batch=%d, n_reduce=%d, n_out=%d)' % (names[self.kind], self.batch, self.n_reduce, self.n_out)


def _classify(shape, strides, itemsize, reduce_axis, out_axis, ordered_index):
    """Memory-order analysis of ONE dense array.

    Walk the axes from the slowest to the fastest varying one (descending stride);
    if the array is dense in that order, label each axis R(educed) or O(ut) and
    merge runs.  [R] -> FULL, [O R] -> ROWS, [R O] / [O R O] -> COLS, else GENERIC.
    Kept axes must appear in their original relative order (the result is written
    C-contiguously); with `ordered_index` (arg-reductions: `_J` is the C-order
    index over the reduced axes) so must the reduced ones.
    """
    generic = Layout(-1, 1, _prod(shape[i] for i in reduce_axis), _prod(shape[i] for i in out_axis))
    axes = [i for i in range(len(shape)) if shape[i] != 1]
    axes.sort(key=lambda i: -abs(strides[i]))
    st = itemsize
    for i in reversed(axes):
        if strides[i] != st:
            return generic
        st *= shape[i]
    rset = set(reduce_axis)
    o_seq = [i for i in axes if i not in rset]
    r_seq = [i for i in axes if i in rset]
    if o_seq != sorted(o_seq):
        return generic
    if ordered_index and r_seq != sorted(r_seq):
        return generic
    runs = []
    for i in axes:
        lab = 'R' if i in rset else 'O'
        if runs and runs[-1][0] == lab:
            runs[-1][1] *= shape[i]
        else:
            runs.append([lab, shape[i]])
    pat = ''.join(r[0] for r in runs)
    n_reduce, n_out = generic.n_reduce, generic.n_out
    if pat in ('R', ''):
        return Layout(_lib.RED_FULL, 1, n_reduce, 1) if n_out == 1 else generic
    if pat == 'O':           # reduced axes all have extent 1
        return Layout(_lib.RED_ROWS, 1, 1, n_out)
    if pat == 'OR':
        return Layout(_lib.RED_ROWS, 1, n_reduce, n_out)
    if pat == 'RO':
        return Layout(_lib.RED_COLS, 1, n_reduce, n_out)
    if pat == 'ORO':
        return Layout(_lib.RED_COLS, runs[0][1], n_reduce, runs[2][1])
    return generic


def _operand_kinds(arrays, x0, layout, reduce_axis, out_axis):
    """Per array: 0 = laid out like x0, 1 = broadcast along the reduced axes and C-dense over the kept
    ones, 2 = broadcast along the kept axes and dense over the reduced ones like x0.  None when some
    operand is none of these (generic kernel)."""
    shape = x0.shape
    e0 = [t // x0.dtype.itemsize for t in x0.strides]
    inner_out = layout.n_out if layout.kind == _lib.RED_COLS else 1      # kept extent inside the reduced run
    dense_out, st = {}, 1
    for i in reversed(out_axis):
        dense_out[i] = st
        st *= shape[i]
    kinds = []
    for a in arrays:
        isz = a.dtype.itemsize
        if a.shape != shape or any(t % isz for t in a.strides):
            return None
        e = [t // isz for t in a.strides]
        live = [i for i in range(len(shape)) if shape[i] > 1]
        if all(e[i] == e0[i] for i in live):
            kinds.append(0)
        elif (all(e[i] == 0 for i in reduce_axis if shape[i] > 1)
              and all(e[i] == dense_out[i] for i in out_axis if shape[i] > 1)):
            kinds.append(1)
        elif (layout.kind != _lib.RED_FULL and all(e[i] == 0 for i in out_axis if shape[i] > 1)
              and all(e[i] * inner_out == e0[i] for i in reduce_axis if shape[i] > 1)):
            kinds.append(2)
        else:
            return None
    return tuple(kinds)


class _AbstractReductionKernel:

    def __init__(self, name, identity, in_params, out_params):
        self.name = name
        self.__name__ = name
        self.identity = identity
        self.in_params = _get_param_info(in_params, True)
        self.out_params = _get_param_info(out_params, False)
        self._cached_codes = {}
        self._memo = {}

    # -- hooks ------------------------------------------------------------------------
    def _get_expressions_and_types(self, in_args, out_args, dtype):
        raise NotImplementedError

    def _get_out_args(self, out_args, out_types, out_shape):
        raise NotImplementedError

    _prebuilt_op = None
    _ordered_index = False
    preamble = ''
    options = ()

    # -- the call ----------------------------------------------------------------------
    def _call(self, in_args, out_args, a_shape, axis, dtype, keepdims, reduce_dims, stream,
              param=0.0):
        # ---- memoised call shape of the prebuilt route (one dense input, fresh output): everything below depends
        # only on (dtype, shape, strides, alignment, axis, dtype=, keepdims, param)
        mkey = None
        given = out_args[0] if out_args else None
        if (len(in_args) == 1 and stream is None and type(in_args[0]) is ndarray
                and (given is None or (len(out_args) == 1 and type(given) is ndarray and given._c_contiguous))
                and self._prebuilt_op is not None and _accelerator.fast_paths_enabled()):
            a = in_args[0]
            ax = tuple(axis) if isinstance(axis, list) else axis
            mkey = (a.dtype, a._shape, a._strides, a.ptr & 15, ax, dtype, bool(keepdims), float(param),
                    None if given is None else (given.dtype, given._shape, given.ptr & 15))
            memo = _kernel._thread_local.__dict__.setdefault('reduce_memo', {}).setdefault(id(self), {})
            e = memo.get(mkey)
            if e is not None and not (_dryrun.enabled and e[6] is None):
                desc, oshape, odtype, ostrides, osize, need, dry = e
                out = ndarray._fresh(oshape, odtype, ostrides, osize) if given is None else given
                if _dryrun.enabled:
                    _dryrun.log.append(dict(dry))
                    return out
                st = current_stream_ptr()
                ws_ptr, ws_bytes = _workspace.get(need, st)
                _lib.check(_lib.lib.b200_reduce_run(ctypes.byref(desc), a.ptr, out.ptr, ws_ptr, ws_bytes, st))
                return out
        if dtype is not None:
            dtype = get_dtype(dtype)
        (map_expr, reduce_expr, post_map_expr, in_types, out_types, reduce_type,
         type_map) = self._get_expressions_and_types(in_args, out_args, dtype)

        reduce_axis, out_axis = _get_axis(axis, len(a_shape))
        out_shape = _get_out_shape(a_shape, reduce_axis, out_axis, keepdims)
        out_args = self._get_out_args(out_args, out_types, out_shape)
        ret = out_args[0]
        if ret.size == 0:
            return ret
        if self.identity == '' and 0 in a_shape:
            raise ValueError('zero-size array to reduction operation %s which has no identity' % self.name)

        for x, t in zip(in_args, in_types):
            if isinstance(x, CScalar):
                x.apply_dtype(t)
        st = current_stream_ptr() if stream is None else _stream_ptr(stream)
        n_reduce = _prod(a_shape[i] for i in reduce_axis)
        n_out = _prod(a_shape[i] for i in out_axis)

        arrays = [a for a in in_args if isinstance(a, ndarray)]
        # structured kernels stream one operand -- or several arrays that share ONE layout (same
        # shape and element strides after broadcasting), as a tuple of their elements
        # structured kernels stream one operand -- or several arrays as a tuple of their elements: those
        # laid out like the first full-size one (kind 0), and those broadcast along the reduced axes
        # (kind 1: dense over the kept axes, e.g. a keepdims mean) or along the kept axes (kind 2: dense over
        # the reduced axes, e.g. weights)
        layout = Layout(-1, 1, n_reduce, n_out)
        kinds = None
        plain = (len(arrays) >= 1 and len(out_args) == 1 and n_reduce > 0
                 and _accelerator.fast_paths_enabled()
                 and not any(p.raw for p in self.in_params + self.out_params))
        if plain:
            x0 = next((a for a in arrays if 0 not in [t for t, n_ in zip(a.strides, a.shape) if n_ > 1]), None)
            if x0 is not None:
                layout = _classify(x0.shape, x0.strides, x0.dtype.itemsize, reduce_axis, out_axis, self._ordered_index)
                if layout.kind >= 0:
                    kinds = _operand_kinds(arrays, x0, layout, reduce_axis, out_axis)
                    if kinds is None:
                        layout = Layout(-1, 1, n_reduce, n_out)
        uniform = kinds is not None
        single = uniform and len(arrays) == 1

        # the fast layouts write a dense C-ordered result of the loop's natural dtype
        out = out_args[0]
        direct = out._c_contiguous
        if layout.kind >= 0 and n_reduce > 0:
            target = out if direct else ndarray(out.shape, out.dtype)
            # ---- prebuilt functor
            if (self._prebuilt_op is not None and single
                    and arrays[0].dtype == in_types[0] and target.dtype == out_types[0]):
                desc = _lib.ReduceDesc(self._prebuilt_op, layout.kind, _scalar.dtype_id(arrays[0].dtype),
                                       _scalar.dtype_id(target.dtype), layout.batch, layout.n_reduce,
                                       layout.n_out, float(param))
                if _lib.lib.b200_reduce_supported(ctypes.byref(desc)):
                    need = ctypes.c_size_t()
                    _lib.check(_lib.lib.b200_reduce_workspace_bytes(ctypes.byref(desc), ctypes.byref(need)))
                    if _dryrun.enabled:
                        _dryrun.record('prebuilt_reduce', name=self.name, layout=layout.kind, batch=layout.batch,
                                       n_reduce=layout.n_reduce, n_out=layout.n_out)
                    else:
                        ws_ptr, ws_bytes = _workspace.get(need.value, st)
                        _lib.check(_lib.lib.b200_reduce_run(ctypes.byref(desc), arrays[0].ptr, target.ptr,
                                                            ws_ptr, ws_bytes, st))
                    if target is not out:
                        _kernel.elementwise_copy(target.reshape(out.shape), out)
                    elif mkey is not None and out._c_contiguous and arrays[0] is in_args[0]:
                        memo = _kernel._thread_local.__dict__.setdefault('reduce_memo', {}).setdefault(id(self), {})
                        if len(memo) >= 512:
                            memo.clear()
                        memo[mkey] = (desc, out._shape, out.dtype, out._strides, out.size, need.value,
                                      dict(_dryrun.log[-1]) if _dryrun.enabled else None)
                    return ret
            # ---- NVRTC functor on the same skeleton
            if uniform:
                _codegen_reduce.launch_structured(
                    self, layout, in_args, target, in_types, out_types, type_map,
                    map_expr, reduce_expr, post_map_expr, reduce_type, st, kinds=kinds)
                if target is not out:
                    _kernel.elementwise_copy(target.reshape(out.shape), out)
                return ret

        # ---- generic strided kernel (any layout, several array operands)
        _codegen_reduce.launch_generic(
            self, in_args, out_args, a_shape, reduce_axis, out_axis, keepdims, in_types, out_types,
            type_map, map_expr, reduce_expr, post_map_expr, reduce_type, st)
        return ret

    @property
    def cached_codes(self):
        if len(self._cached_codes) == 0:
            import warnings
            warnings.warn('No codes are cached because compilation is deferred until the first '
                          'function call or a prebuilt kernel was used.')
        return dict(self._cached_codes.items())

    @property
    def cached_code(self):
        codes = self._cached_codes
        if len(codes) > 1:
            import warnings
            warnings.warn('The input types of the kernel could not be inferred. Please use `.cached_codes` instead.')
        return next(iter(codes.values()))


# -------------------------------------------------------------------------------------
# create_reduction_func
# -------------------------------------------------------------------------------------
class _SimpleReductionKernel(_AbstractReductionKernel):
    """cupy/_core/_reduction.pyx:565-684."""

    def __init__(self, name, ops, identity, preamble, sort_reduce_axis=True, prebuilt=None):
        super().__init__(name, '' if identity is None else str(identity), 'T in0', 'T out0')
        self._ops = ops
        self.preamble = preamble
        self.nin = 1
        self.nout = 1
        self._routine_cache = {}
        self._ordered_index = not sort_reduce_axis
        self._prebuilt_op = prebuilt

    def __call__(self, a, axis=None, dtype=None, out=None, keepdims=False):
        if hasattr(a, '__cupy_override_reduction_kernel__'):
            return a.__cupy_override_reduction_kernel__(self, axis, dtype, out, keepdims)
        arr = _kernel._convert_arg(a)
        if not isinstance(arr, ndarray):
            raise TypeError("Argument 'a' has incorrect type (expected cupy.ndarray, got %s)" % type(a).__name__)
        if out is not None and not isinstance(out, ndarray):
            raise TypeError('Output arguments type must be cupy.ndarray')
        if _accelerator.reference_first():
            r = _accelerator.try_reference('reduction', self.name, arr, axis=axis, dtype=dtype, out=out, keepdims=keepdims)
            if r is not None:
                return r
        out_args = [] if out is None else [out]
        return self._call([arr], out_args, arr.shape, axis, dtype, keepdims, True, None)

    def _get_expressions_and_types(self, in_args, out_args, dtype):
        op = self._ops.guess_routine(self.name, self._routine_cache, in_args, dtype, self._ops)
        map_expr, reduce_expr, post_map_expr, reduce_type = op.routine
        if reduce_type is None:
            reduce_type = get_typename(op.out_types[0])
        out_type = out_args[0].dtype if out_args else op.out_types[0]
        type_map = (('type_in0_raw', in_args[0].dtype), ('type_out0_raw', get_dtype(out_type)))
        return map_expr, reduce_expr, post_map_expr, op.in_types, op.out_types, reduce_type, type_map

    def _get_out_args(self, out_args, out_types, out_shape):
        return _kernel._get_out_args_from_optionals(out_args, out_types, out_shape, 'unsafe')


def create_reduction_func(name, ops, routine=None, identity=None, preamble='',
                          sort_reduce_axis=True, prebuilt=None):
    ops = _kernel._Ops.from_tuples(ops, routine)
    return _SimpleReductionKernel(name, ops, identity, preamble, sort_reduce_axis, prebuilt)


# -------------------------------------------------------------------------------------
# ReductionKernel
# -------------------------------------------------------------------------------------
class ReductionKernel(_AbstractReductionKernel):
    """User-defined reduction kernel (drop-in for cupy.ReductionKernel,
    cupy/_core/_reduction.pyx:716-917)."""

    def __init__(self, in_params, out_params, map_expr, reduce_expr, post_map_expr, identity,
                 name='reduce_kernel', reduce_type=None, reduce_dims=True, preamble='', options=()):
        if not _jit.is_valid_kernel_name(name):
            raise ValueError('Invalid kernel name: "%s"' % name)
        super().__init__(name, '' if identity is None else str(identity), in_params, out_params)
        self.nin = len(self.in_params)
        self.nout = len(self.out_params)
        self.nargs = self.nin + self.nout
        self.reduce_expr = reduce_expr
        self.map_expr = map_expr
        self.post_map_expr = post_map_expr
        self.options = tuple(options)
        self.reduce_dims = reduce_dims
        self.reduce_type = self.out_params[0].ctype if reduce_type is None else reduce_type
        self.preamble = preamble

    def __call__(self, *args, **kwargs):
        out = kwargs.pop('out', None)
        axis = kwargs.pop('axis', None)
        keepdims = kwargs.pop('keepdims', False)
        stream = kwargs.pop('stream', None)
        if kwargs:
            raise TypeError('Wrong arguments %s' % kwargs)
        n_args = len(args)
        if n_args != self.nin and n_args != self.nargs:
            raise TypeError('Wrong number of arguments for %s' % self.name)
        out_args = list(args[self.nin:])
        if out is not None:
            if self.nout != 1:
                raise NotImplementedError('')
            if len(out_args) != 0:
                raise ValueError("cannot specify 'out' as both a positional and keyword argument")
            out_args = [out]
        in_args = _preprocess_args(args[:self.nin])
        out_args = _preprocess_args(out_args)
        in_args, broad_shape = _broadcast(in_args, self.in_params, False)
        return self._call(in_args, out_args, broad_shape, axis, None, keepdims, self.reduce_dims, stream)

    def _get_expressions_and_types(self, in_args, out_args, dtype):
        in_ndarray_types = tuple(a.dtype if isinstance(a, ndarray) else None for a in in_args)
        out_ndarray_types = tuple(a.dtype if isinstance(a, ndarray) else None for a in out_args)
        in_types, out_types, type_map = _decide_params_type_core(
            self.in_params, self.out_params, in_ndarray_types, out_ndarray_types)
        return (self.map_expr, self.reduce_expr, self.post_map_expr, in_types, out_types,
                self.reduce_type, type_map)

    def _get_out_args(self, out_args, out_types, out_shape):
        return _kernel._get_out_args_with_params(out_args, out_types, out_shape, self.out_params, False)
