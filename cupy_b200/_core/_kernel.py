"""The elementwise engine: `ufunc`, `ElementwiseKernel` and the launcher.

Host-side mirror of cupy/_core/_kernel.pyx -- same classes, argument meaning,
defaults and error messages -- over the B200 C ABI:

* `_broadcast_core`              cupy/_core/internal.pyx:320-376
* `ParameterInfo`, `_get_param_info`, `_decide_params_type_core`
                                 cupy/_core/_kernel.pyx:464-604
* `ElementwiseKernel`            cupy/_core/_kernel.pyx:743-1002
* `ufunc`, `_Op`, `_Ops`, `create_ufunc` (NEP-50 weak scalars)
                                 cupy/_core/_kernel.pyx:1103-1766
* overlap guard                  cupy/_core/_kernel.pyx:673-683, _memory_range.pyx

What differs is below the call: instead of rendering a one-element-per-thread
CUPY_FOR loop per (dtype, ndim, contiguity) (`_kernel.pyx:78-107`), the call is
handed to `b200_ew_plan` (collapse dims + classify FLAT / ROWWISE / TILED), and
either a prebuilt kernel is launched (`b200_ufunc_launch`) or the operation
string is wrapped in the tiler glue of `_codegen.py` and compiled by NVRTC.
"""
from __future__ import annotations

import ctypes
import os as _os
import re as _re
import threading

import numpy

from cupy_b200 import _lib
from cupy_b200._core import _codegen, _dryrun, _jit, _scalar
from cupy_b200._core._ndarray import ndarray, current_stream_ptr
from cupy_b200._core._scalar import CScalar, get_dtype, get_typename

_thread_local = threading.local()


# ---------------------------------------------------------------------------
# argument preprocessing
# ---------------------------------------------------------------------------
def _convert_arg(arg):
    if isinstance(arg, ndarray):
        return arg
    if hasattr(arg, '__cuda_array_interface__') or type(arg).__module__.startswith('torch'):
        from cupy_b200._core import _ndarray
        return _ndarray.asarray(arg)
    return CScalar(arg)


def _preprocess_args(args):
    return [_convert_arg(a) for a in args]


def _broadcast_core(arrays):
    """Broadcast the ndarray entries of `arrays` in place; returns the shape."""
    idx = [i for i, a in enumerate(arrays) if isinstance(a, ndarray)]
    if not idx:
        return ()
    nd = max(arrays[i].ndim for i in idx)
    shape = []
    for d in range(nd):
        s = 1
        for i in idx:
            a = arrays[i]
            k = d - (nd - a.ndim)
            if k < 0:
                continue
            a_sh = a.shape[k]
            if a_sh == s or a_sh == 1:
                continue
            if s == 1:
                s = a_sh
                continue
            raise ValueError(
                'operands could not be broadcast together with shapes {}'.format(
                    ' '.join([str(x.shape) if isinstance(x, ndarray) else '()' for x in arrays])))
        shape.append(s)
    shape = tuple(shape)
    for i in idx:
        a = arrays[i]
        if a.shape != shape:
            arrays[i] = a.broadcast_to(shape)
    return shape


def _get_bound(a):
    left, right = 0, a.dtype.itemsize
    for s, t in zip(a.shape, a.strides):
        if s == 0:
            return a.ptr, a.ptr
        tmp = (s - 1) * t
        if tmp > 0:
            right += tmp
        else:
            left += tmp
    return a.ptr + left, a.ptr + right


def may_share_bounds(a, b):
    if a.size == 0 or b.size == 0:
        return False
    al, ar = _get_bound(a)
    bl, br = _get_bound(b)
    return al < br and bl < ar


def _copy_in_args_if_needed(in_args, out_args):
    for i, a in enumerate(in_args):
        if isinstance(a, ndarray):
            for out in out_args:
                if a is not out and may_share_bounds(a, out):
                    in_args[i] = a.copy()
                    break


# ---------------------------------------------------------------------------
# parameters of user kernels
# ---------------------------------------------------------------------------
class ParameterInfo:
    __slots__ = ('name', 'dtype', 'ctype', 'raw', 'is_const')

    def __init__(self, param, is_const):
        self.name = None
        self.dtype = None
        self.ctype = None
        self.raw = False
        self.is_const = is_const
        s = tuple(i for i in param.split() if len(i) != 0)
        if len(s) < 2:
            raise Exception('Syntax error: %s' % param)
        t, self.name = s[-2:]
        if t == 'CIndexer':
            pass
        elif len(t) == 1:
            self.ctype = t
        else:
            self.dtype = get_dtype(t)
            self.ctype = get_typename(self.dtype)
        for i in s[:-2]:
            if i == 'raw':
                self.raw = True
            elif i == '_non_const':
                self.is_const = False
            else:
                raise Exception('Unknown keyword "%s"' % i)

    def _key(self):
        return (self.name, self.dtype, self.ctype, self.raw, self.is_const)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return isinstance(other, ParameterInfo) and self._key() == other._key()

    def __repr__(self):
        return '<ParameterInfo name=%r dtype=%r ctype=%r raw=%r is_const=%r>' % self._key()


_param_info_memo = {}


def _get_param_info(s, is_const):
    key = (s, is_const)
    r = _param_info_memo.get(key)
    if r is None:
        r = () if len(s) == 0 else tuple(ParameterInfo(i, is_const) for i in s.strip().split(','))
        _param_info_memo[key] = r
    return r


def _decide_params_type_core(in_params, out_params, in_args_dtype, out_args_dtype):
    type_dict = {}
    if out_args_dtype:
        assert len(out_params) == len(out_args_dtype)
        for p, a in zip(out_params, out_args_dtype):
            if a is None:
                raise TypeError('Output arguments must be cupy.ndarray')
            if p.dtype is not None:
                if get_dtype(a) != get_dtype(p.dtype):
                    raise TypeError('Type is mismatched. %s %s %s' % (p.name, a, p.dtype))
            elif p.ctype in type_dict:
                t = type_dict[p.ctype]
                if get_dtype(t) != get_dtype(a):
                    raise TypeError('Type is mismatched. %s %s %s %s' % (p.name, a, t, p.ctype))
            else:
                type_dict[p.ctype] = a
    assert len(in_params) == len(in_args_dtype)
    for p, a in zip(in_params, in_args_dtype):
        if a is None:
            continue
        if p.dtype is not None:
            if numpy.dtype(a) != numpy.dtype(p.dtype):
                raise TypeError('Type is mismatched. %s %s %s' % (p.name, a, p.dtype))
        elif p.ctype in type_dict:
            t = type_dict[p.ctype]
            if numpy.dtype(t) != numpy.dtype(a):
                raise TypeError('Type is mismatched. %s %s %s %s' % (p.name, a, t, p.ctype))
        else:
            type_dict[p.ctype] = a
    try:
        in_types = tuple(type_dict[p.ctype] if p.dtype is None else p.dtype for p in in_params)
        out_types = tuple(type_dict[p.ctype] if p.dtype is None else p.dtype for p in out_params)
    except KeyError as e:
        raise TypeError('Type %s of a parameter could not be inferred from the arguments' % e)
    type_map = tuple(sorted((k, get_dtype(v)) for k, v in type_dict.items()))
    return in_types, out_types, type_map


def _broadcast(args, params, use_size):
    value = []
    any_nonraw_array = False
    for a, p in zip(args, params):
        if not p.raw and isinstance(a, ndarray):
            any_nonraw_array = True
            value.append(a)
        else:
            value.append(None)
    if use_size:
        if any_nonraw_array:
            raise ValueError('Specified \'size\' can be used only if all of the ndarray are \'raw\'.')
    else:
        if not any_nonraw_array:
            raise ValueError('Loop size is undecided.')
    shape = _broadcast_core(value)
    for i, a in enumerate(value):
        if a is None:
            value[i] = args[i]
    return value, shape


def _get_out_args_from_optionals(out_args, out_types, out_shape, casting):
    out_args = list(out_args)
    while len(out_args) < len(out_types):
        out_args.append(None)
    for i, a in enumerate(out_args):
        if a is None:
            out_args[i] = ndarray(out_shape, out_types[i])
            continue
        if not isinstance(a, ndarray):
            raise TypeError('Output arguments type must be cupy.ndarray')
        if a.shape != tuple(out_shape):
            raise ValueError('Out shape is mismatched')
        _scalar.raise_if_invalid_cast(out_types[i], a.dtype, casting, 'output operand')
    return out_args


def _get_out_args_with_params(out_args, out_types, out_shape, out_params, is_size_specified):
    if not out_args:
        for p in out_params:
            if p.raw and not is_size_specified:
                raise ValueError('Output array size is Undecided')
        return [ndarray(out_shape, t) for t in out_types]
    for a, p in zip(out_args, out_params):
        if not isinstance(a, ndarray):
            raise TypeError('Output arguments type must be cupy.ndarray')
        if not p.raw and a.shape != tuple(out_shape):
            raise ValueError('Out shape is mismatched')
    return list(out_args)


# ---------------------------------------------------------------------------
# the launcher
# ---------------------------------------------------------------------------
def _make_operands(args, params, n_in, loop_shape):
    """args (ndarray | CScalar) -> C operand array.  Non-raw arrays must already
    be broadcast to `loop_shape`."""
    n = len(args)
    if n > _lib.MAX_ARGS:
        raise ValueError('too many kernel arguments (%d > %d)' % (n, _lib.MAX_ARGS))
    ops = (_lib.Operand * n)()
    for k, (a, p) in enumerate(zip(args, params)):
        o = ops[k]
        if isinstance(a, ndarray):
            if a.ndim > _lib.MAX_NDIM:
                a = _collapse_for_abi(a)
            o.data = a.ptr
            o.kind = _lib.KIND_RAW if p.raw else _lib.KIND_ARRAY
            o.dtype = _scalar.dtype_id(a.dtype)
            o.ndim = a.ndim
            o.is_output = 1 if k >= n_in else 0
            for d in range(a.ndim):
                o.shape[d] = a.shape[d]
                o.strides[d] = a.strides[d]
        else:
            o.kind = _lib.KIND_SCALAR
            o.dtype = _scalar.dtype_id(a.descr)
            raw = a.raw_bytes()
            o.scalar[0] = int.from_bytes(raw[:8], 'little', signed=True)
            o.scalar[1] = int.from_bytes(raw[8:16], 'little', signed=True)
    return ops


def _collapse_for_abi(a):
    raise NotImplementedError('arrays with more than %d dimensions are not supported' % _lib.MAX_NDIM)


def plan_elementwise(args, params, n_in, loop_shape, keep_order=True):
    ops = _make_operands(args, params, n_in, loop_shape)
    plan = _lib.EwPlan()
    if not any(o.kind == _lib.KIND_ARRAY for o in ops):
        # all-raw kernel with explicit size: a 1-D loop of `size` elements
        plan.variant = _lib.EW_FLAT
        plan.ndim = 1
        plan.vec = 1
        plan.idx32 = 1 if _prod(loop_shape) < 2 ** 31 else 0
        plan.nargs = len(args)
        plan.size = _prod(loop_shape)
        plan.shape[0] = plan.size
        return ops, plan
    _lib.check(_lib.lib.b200_ew_plan_ex(len(args), ops, _lib.PLAN_KEEP_ORDER if keep_order else 0, ctypes.byref(plan)))
    return ops, plan


def _prod(shape):
    r = 1
    for s in shape:
        r *= int(s)
    return r


def _prebuilt_fits(plan, ops, args):
    """Prebuilt kernels decide operand kinds at run time, which costs 5-17 % when it matters: by-value
    scalars (a branch per operand in the load phase) and ROWWISE operands that are not unit-stride.
    Those calls are specialised by NVRTC instead (the reference compiles every call shape)."""
    if any(o.kind == _lib.KIND_SCALAR for o in ops):
        return False
    if plan.variant == _lib.EW_ROWWISE:
        last = plan.ndim - 1
        return all(plan.strides[k][last] == a.dtype.itemsize for k, a in enumerate(args))
    return True


# tunables of the generated kernels (bench/tuning scripts override these)
class _Tunables(dict):
    """A/B knobs; changing one invalidates the memoised call shapes (they hold plans made under the old value)."""

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, value)
        clear_call_shape_memo()


def clear_call_shape_memo():
    global _memo_epoch
    _memo_epoch += 1


_memo_epoch = 0
tunables = _Tunables({
    'threads': 256,
    'flat_unroll': 4,
    'flat_mixed_vec': 8,      # FLAT loops over operands of different item sizes: widest vector (elements) so the
                              # narrow operand still moves >= 8 bytes per access; 0 = the planner's 16 / largest item
                              # (B200 sweep profiles/r02_mixed_width_probe.log: where(mask, x, y) 73 -> 102 % of the
                              # measured copy peak, isnan 89 -> 97 %; 16 is level with 8, 32 elements per thread lose)
    'flat_mixed_items': 16,   # ... and elements per thread there (vector x unroll)
    'row_unroll': 2,
    'blocks_per_sm': 0,       # 0 = library default
    'tma_stages': 0,          # TILED_TMA ring depth, 0 = library default
    'reg_unroll': 0,          # TILED_REG blocks per thread, 0 = library default (8 vector loads in flight)
    'reg_min_blocks': 0,      # TILED_REG __launch_bounds__ min blocks/SM, 0 = from the register estimate
})


def _mixed_width_vector(args, ops, vec, unroll):
    """FLAT loop over arrays of different item sizes (a comparison writing bool, `where` reading a mask, a cast):
    the planner's vector is 16 bytes of the LARGEST item, which leaves the 1-byte operand with 4-byte accesses.
    Widen the vector (wide operands then take several 16-byte accesses per step) while every operand keeps
    the alignment its accesses need; the elements per thread stay what the tunables say."""
    sizes = [args[k].dtype.itemsize for k, o in enumerate(ops) if o.kind == _lib.KIND_ARRAY]
    if not sizes or min(sizes) == max(sizes):
        return vec, unroll
    want = min(tunables['flat_mixed_vec'], 16 // min(sizes))
    while want > vec:
        if all(args[k].ptr % min(16, want * args[k].dtype.itemsize) == 0
               for k, o in enumerate(ops) if o.kind == _lib.KIND_ARRAY):
            return want, max(1, tunables['flat_mixed_items'] // want)
        want //= 2
    return vec, unroll


def _launch_jit(name, spec, args, params, n_in, ops, plan, block_size=None, stream=None, ind_shape=None):
    """spec: _codegen.EwSpec describing parameters / operation strings.  `ind_shape`: the un-collapsed loop
    shape `_ind` must present (reduce_dims=False kernels that read `_ind`), else None."""
    variant = plan.variant
    threads = tunables['threads'] if block_size is None else int(block_size)
    if variant == _lib.EW_TILED:
        threads, unroll, vec = 256, 4, 1
    elif variant == _lib.EW_TILED_REG:
        esz = 16 // plan.vec
        # 8 vector loads of the staged operand in flight for unary calls (== b200::reg_tile_unroll());
        # half of that with more operands: occupancy hides the latency better than registers do there
        # (B200 sweep in profiles/r01_tiled_lab.md)
        n_arrays = sum(1 for o in ops if o.kind == _lib.KIND_ARRAY)
        unroll = tunables['reg_unroll'] or ({4: 2, 2: 1, 8: 4}[esz] if n_arrays <= 2 else {4: 1, 2: 1, 8: 2}[esz])
        threads, vec = 256, plan.vec
    elif variant == _lib.EW_TILED_TMA:
        threads, unroll, vec = 288, tunables['tma_stages'], plan.vec   # "unroll" slot = ring depth override (0 = auto)
    elif variant == _lib.EW_FLAT:
        unroll, vec = tunables['flat_unroll'], plan.vec
        if plan.staged_mask:          # periodic operands: the plan assumed 256-thread blocks
            threads = 256
        elif tunables['flat_mixed_vec'] > vec > 1:
            vec, unroll = _mixed_width_vector(args, ops, vec, unroll)
            plan.vec = vec
    else:
        unroll, vec = tunables['row_unroll'], plan.vec
    arginfo = tuple(
        ('raw', a.dtype.char, a.ndim, a._c_contiguous) if (isinstance(a, ndarray) and p.raw)
        else ('arr', a.dtype.char) if isinstance(a, ndarray) else ('scalar', a.descr.char)
        for a, p in zip(args, params))
    access, min_blocks = 0, 1
    if variant == _lib.EW_FLAT and plan.staged_mask:
        access = 1                    # FlatTiler<..., PERIODIC = true>
    if variant == _lib.EW_ROWWISE:
        # innermost-stride kind per operand (RowTiler SPEC, 2 bits each): unit / zero / other
        last = plan.ndim - 1
        for k, o in enumerate(ops):
            if o.kind != _lib.KIND_ARRAY:
                continue
            si = plan.strides[k][last]
            isz = args[k].dtype.itemsize
            access |= (1 if si == isz else 2 if si == 0 else 3) << (2 * k)
    if variant == _lib.EW_TILED_REG:
        # per-operand access, fixed at compile time (RegTileTiler SPEC, 3 bits each): staged /
        # unit along O / other, plus "broadcast along I"; registers a thread's blocks take
        last, ax = plan.ndim - 1, plan.tile_axis
        block_regs = unroll * plan.vec * plan.vec * esz // 4
        regs_in = regs_out = 0
        for k, o in enumerate(ops):
            if o.kind != _lib.KIND_ARRAY:
                continue
            bits = 1 if (plan.staged_mask >> k) & 1 else 2 if plan.strides[k][last] == esz else 3
            regs = block_regs
            if bits != 1 and plan.strides[k][ax] == 0:
                bits |= 4
                regs //= plan.vec
            access |= bits << (3 * k)
            if o.is_output:
                regs_out += regs
            else:
                regs_in += regs
        est = max(regs_in, regs_out) + (64 if esz == 2 else 30)
        min_blocks = tunables['reg_min_blocks'] or max(1, min(3 if n_arrays <= 2 else 5, 65536 // (256 * est)))
    ind_ndim = len(ind_shape) if (ind_shape is not None and spec.uses_ind) else 0
    key = (name, variant, vec, unroll, threads, bool(plan.idx32), plan.ndim if spec.uses_ind else -1, arginfo, access,
           min_blocks, ind_ndim)
    fn = spec.memo.get(key)
    if fn is None:
        source = _codegen.render_elementwise(
            spec, name, args, params, variant=variant, vec=vec, unroll=unroll, threads=threads,
            idx32=bool(plan.idx32), ndim=plan.ndim, access_spec=access, min_blocks=min_blocks, ind_ndim=ind_ndim)
        spec.last_source = source
        fn = _jit.get_function(source, name, spec.options)
        spec.memo[key] = fn
    plan.reserved = (plan.reserved & 0xff000000) | (unroll & 0xff) | ((tunables['blocks_per_sm'] & 0xffff) << 8)
    if _dryrun.enabled:
        _dryrun.record('jit_elementwise', name=name, variant=variant, vec=vec, ndim=plan.ndim,
                       idx32=bool(plan.idx32), staged_mask=plan.staged_mask, tile_axis=plan.tile_axis,
                       shape=tuple(plan.shape[:plan.ndim]), source=spec.last_source)
        return fn.handle, threads
    st = current_stream_ptr() if stream is None else _stream_ptr(stream)
    if ind_ndim:
        shp = (ctypes.c_int64 * ind_ndim)(*ind_shape)
        _lib.check(_lib.lib.b200_jit_ew_launch_ex(fn.handle, ctypes.byref(plan), len(args), ops, threads, ind_ndim, shp, st))
    else:
        _lib.check(_lib.lib.b200_jit_ew_launch_ex(fn.handle, ctypes.byref(plan), len(args), ops, threads, 0, None, st))
    return fn.handle, threads


def _stream_ptr(stream):
    if stream is None:
        return current_stream_ptr()
    if isinstance(stream, int):
        return stream
    for attr in ('cuda_stream', 'ptr'):
        if hasattr(stream, attr):
            return int(getattr(stream, attr))
    raise TypeError('unsupported stream object %r' % (stream,))


# ---------------------------------------------------------------------------
# ElementwiseKernel
# ---------------------------------------------------------------------------
class ElementwiseKernel:
    """User-defined elementwise kernel (drop-in for cupy.ElementwiseKernel,
    cupy/_core/_kernel.pyx:743-1002: same constructor, `__call__(*args, size=,
    stream=, block_size=)`, same errors)."""

    def __init__(self, in_params, out_params, operation, name='kernel', reduce_dims=True,
                 preamble='', no_return=False, return_tuple=False, **kwargs):
        if not _jit.is_valid_kernel_name(name):
            raise ValueError('Invalid kernel name: "%s"' % name)
        self.in_params = _get_param_info(in_params, True)
        self.out_params = _get_param_info(out_params, False)
        self.nin = len(self.in_params)
        self.nout = len(self.out_params)
        self.nargs = self.nin + self.nout
        self.params = self.in_params + self.out_params
        self.operation = operation
        self.name = name
        self.__name__ = name
        self.reduce_dims = reduce_dims
        self.preamble = preamble
        self.no_return = no_return
        self.return_tuple = return_tuple
        write_only = bool(kwargs.pop('_write_only_outputs', False))     # private: text generated by cupy_b200.fuse
        self.kwargs = kwargs
        bad = set(kwargs) - {'options', 'loop_prep', 'after_loop'}
        if bad:
            raise TypeError('Wrong arguments %s' % {k: kwargs[k] for k in bad})
        names = [p.name for p in self.params]
        if 'i' in names:
            raise ValueError('Can not use \'i\' as a parameter name')
        # the loop order is free unless the code can observe the C-order linear index
        text = ' '.join((operation, kwargs.get('loop_prep', ''), kwargs.get('after_loop', '')))
        self._keeps_order = (any(p.raw for p in self.params) or _re.search(r'\b(i|_ind)\b', text) is not None)
        self._has_raw = any(p.raw for p in self.params)
        self._params_type_memo = {}
        self._cached_codes = {}
        self._spec = _codegen.EwSpec(
            mode='elementwise', operation=operation, preamble=preamble,
            loop_prep=kwargs.get('loop_prep', ''), after_loop=kwargs.get('after_loop', ''),
            options=tuple(kwargs.get('options', ())), write_only_outputs=write_only)

    def __call__(self, *args, **kwargs):
        # ---- memoised call shape (see _fast_key): plain arrays / scalars, no keyword arguments
        fkey = None
        if not kwargs and (len(args) == self.nin or len(args) == self.nargs) and not self._has_raw:
            fkey = _fast_key(args, None)
            if fkey is not None:
                memo = _thread_local.__dict__.setdefault('ufunc_memo', {}).setdefault(id(self), {})
                e = memo.get(fkey)
                if e is not None:
                    r = self._fast_launch(e, args)
                    if r is not _MISS:
                        return r
        size = kwargs.pop('size', -1)
        stream = kwargs.pop('stream', None)
        block_size = kwargs.pop('block_size', 128)
        if len(kwargs):
            raise TypeError('Wrong arguments %s' % kwargs)
        if block_size <= 0:
            raise ValueError('block_size must be greater than zero')
        n_args = len(args)
        if n_args != self.nin and n_args != self.nargs:
            raise TypeError(
                'Wrong number of arguments for {!r}. It must be either {} or {} (with outputs), '
                'but given {}.'.format(self.name, self.nin, self.nargs, n_args))
        for arg in args:
            if hasattr(arg, '__cupy_override_elementwise_kernel__'):
                return arg.__cupy_override_elementwise_kernel__(self, *args, **kwargs)
        arg_list = _preprocess_args(args)
        out_args = arg_list[self.nin:]
        bcast, shape = _broadcast(arg_list, self.params, size != -1)
        in_args = bcast[:self.nin]

        in_ndarray_types = tuple(a.dtype if isinstance(a, ndarray) else None for a in in_args)
        out_ndarray_types = tuple(a.dtype if isinstance(a, ndarray) else None for a in out_args)
        in_types, out_types, type_map = self._decide_params_type(in_ndarray_types, out_ndarray_types)

        is_size_specified = False
        if size != -1:
            shape = (int(size),)
            is_size_specified = True
        out_args = _get_out_args_with_params(out_args, out_types, shape, self.out_params, is_size_specified)
        if self.no_return:
            ret = None
        elif not self.return_tuple and self.nout == 1:
            ret = out_args[0]
        else:
            ret = tuple(out_args)
        if 0 in shape:
            return ret

        weak_ts = [x.weak_t if isinstance(x, CScalar) else None for x in in_args]
        for i, x in enumerate(in_args):
            if isinstance(x, CScalar):
                x.apply_dtype(in_types[i])
        inout_args = in_args + out_args
        ops, plan = plan_elementwise(inout_args, self.params, self.nin, shape, keep_order=self._keeps_order)
        # the 128-thread default of the reference is a floor for its scalar loop;
        # the tilers are built for 256 -- a caller's explicit block_size is honoured
        bs = None if block_size == 128 else block_size
        self._spec.type_map = type_map
        # reduce_dims=False: `_ind` presents the un-collapsed loop shape (cupy/_core/_kernel.pyx:926-929);
        # the operands are still collapsed for addressing -- only the indexer keeps the original rank
        ind_shape = None if (self.reduce_dims or len(shape) == 0) else tuple(shape)
        spec = self._spec.bind(type_map)
        handle, threads = _launch_jit(self.name, spec, inout_args, self.params, self.nin, ops, plan,
                                      block_size=bs, stream=stream, ind_shape=ind_shape)
        key = tuple(t for t in in_ndarray_types if t is not None)
        if key not in self._cached_codes:
            self._cached_codes[key] = spec.last_source
        if (fkey is not None and stream is None and bs is None and size == -1
                and not (ind_shape is not None and spec.uses_ind)
                and all(o._c_contiguous or n_args == self.nargs for o in out_args)):
            self._remember(fkey, ops, plan, inout_args, in_types, weak_ts, out_args, n_args == self.nargs, handle, threads)
        return ret

    def _remember(self, fkey, ops, plan, inout_args, in_types, weak_ts, out_args, outs_given, handle, threads):
        memo = _thread_local.__dict__.setdefault('ufunc_memo', {}).setdefault(id(self), {})
        if len(memo) >= _FAST_MAX:
            memo.clear()
        e = _FastEntry()
        e.ops, e.plan, e.nargs = ops, plan, len(inout_args)
        e.prebuilt, e.handle, e.threads = None, handle, threads
        e.array_slots, e.scalar_slots = [], []
        for k in range(self.nin):
            if isinstance(inout_args[k], ndarray):
                e.array_slots.append(k)
            else:
                t = get_dtype(in_types[k])
                lo = hi = None
                if t.kind in 'iu':
                    info = numpy.iinfo(t)
                    lo, hi = int(info.min), int(info.max)
                e.scalar_slots.append((k, t, weak_ts[k], lo, hi, {}))
        e.out_slot = self.nin
        # outputs: given positionally (patched from the call) or created fresh from the remembered metadata
        e.out_dtype = None if outs_given else [(o.dtype, o._shape, o._strides, o.size) for o in out_args]
        e.out_shape = e.out_strides = e.out_size = None
        e.empty = not outs_given
        e.dry = dict(_dryrun.log[-1]) if (_dryrun.enabled and _dryrun.log) else None
        memo[fkey] = e

    def _fast_launch(self, e, args):
        if _dryrun.enabled and e.dry is None:
            return _MISS
        ops = e.ops
        if e.empty:
            outs = [ndarray._fresh(sh, dt, st, sz) for dt, sh, st, sz in e.out_dtype]
        else:
            outs = list(args[self.nin:])
            for k in e.array_slots:
                a = args[k]
                for o in outs:
                    if a is not o and may_share_bounds(a, o):
                        return _MISS
        for k in e.array_slots:
            ops[k].data = args[k].ptr
        for k, t, weak_t, lo, hi, seen in e.scalar_slots:
            v = args[k]
            tv = type(v)
            sk = _scalar_key(tv, v)
            words = seen.get(sk) if sk is not None else None
            if words is None:
                words = _scalar_words(v, t, weak_t if not isinstance(v, numpy.generic) else False, lo, hi)
                if sk is not None:
                    if len(seen) >= 64:
                        seen.clear()
                    seen[sk] = words
            ops[k].scalar[0] = words[0]
            ops[k].scalar[1] = words[1]
        for j, o in enumerate(outs):
            ops[e.out_slot + j].data = o.ptr
        if _dryrun.enabled:
            if e.dry is not None:
                _dryrun.log.append(dict(e.dry))
        else:
            _lib.check(_lib.lib.b200_jit_ew_launch_ex(e.handle, ctypes.byref(e.plan), e.nargs, ops, e.threads, 0, None,
                                                      current_stream_ptr()))
        if self.no_return:
            return None
        if not self.return_tuple and self.nout == 1:
            return outs[0]
        return tuple(outs)

    def _decide_params_type(self, in_args_dtype, out_args_dtype):
        key = (in_args_dtype, out_args_dtype)
        ret = self._params_type_memo.get(key)
        if ret is None:
            ret = _decide_params_type_core(self.in_params, self.out_params, in_args_dtype, out_args_dtype)
            self._params_type_memo[key] = ret
        return ret

    @property
    def cached_codes(self):
        if len(self._cached_codes) == 0:
            import warnings
            warnings.warn('No codes are cached because compilation is deferred until the first function call.')
        return dict(self._cached_codes.items())

    @property
    def cached_code(self):
        codes = self._cached_codes
        if len(codes) > 1:
            import warnings
            warnings.warn('The input types of the kernel could not be inferred. Please use `.cached_codes` instead.')
        return next(iter(codes.values()))


# ---------------------------------------------------------------------------
# ufunc
# ---------------------------------------------------------------------------
def _get_kind_score(kind):
    if issubclass(kind, (numpy.bool_, bool)):
        return 0
    if issubclass(kind, (numpy.integer, int)):
        return 1
    if issubclass(kind, (numpy.inexact, float, complex)):
        return 2
    return 3


def _check_should_use_weak_scalar(in_types, weaks):
    """cupy/_core/_kernel.pyx:1114-1144."""
    if weaks is None:
        return False
    max_array_kind = -1
    max_scalar_kind = -1
    for in_t, w_t in zip(in_types, weaks):
        if w_t:
            max_scalar_kind = max(max_scalar_kind, _get_kind_score(w_t))
        else:
            max_array_kind = max(max_array_kind, _get_kind_score(in_t.type))
    all_scalars_or_arrays = max_scalar_kind == -1 or max_array_kind == -1
    return not all_scalars_or_arrays and max_array_kind >= max_scalar_kind


class _Op:
    def __init__(self, in_types, out_types, routine, error_func):
        self.in_types = tuple(get_dtype(t) for t in in_types)
        self.out_types = tuple(get_dtype(t) for t in out_types)
        self.nin = len(in_types)
        self.nout = len(out_types)
        self.routine = routine
        self.error_func = error_func

    @staticmethod
    def from_type(typ, routine, error_func=None):
        types = typ.split('->')
        if len(types) == 1:
            in_types = out_types = tuple(types)
        else:
            in_types, out_types = map(tuple, types)
        return _Op(in_types, out_types, routine, error_func)

    def check_valid(self):
        if self.error_func is not None:
            self.error_func()

    def __repr__(self):
        return '_Op(%s->%s)' % (''.join(t.char for t in self.in_types), ''.join(t.char for t in self.out_types))


class _Ops:
    def __init__(self, ops):
        assert len(ops) > 0
        self.ops = tuple(ops)
        self.nin = ops[0].nin
        self.nout = ops[0].nout
        for op in ops:
            if op.nin != self.nin or op.nout != self.nout:
                raise ValueError('invalid op %s, wrong nin or nout.' % op)

    @staticmethod
    def from_tuples(ops, routine):
        ops_ = []
        for t in ops:
            if isinstance(t, tuple):
                typ, rt = t
                if rt is None:
                    rt = routine
                elif isinstance(rt, tuple):
                    rt = tuple(r1 or r2 for r1, r2 in zip(rt, routine))
                elif not isinstance(rt, str):
                    assert callable(rt)
                    ops_.append(_Op.from_type(typ, None, rt))
                    continue
            else:
                typ, rt = t, routine
            ops_.append(_Op.from_type(typ, rt))
        return _Ops(ops_)

    def guess_routine(self, name, cache, in_args, dtype, out_ops):
        if dtype is None:
            in_types, weaks, any_weak = [], [], False
            for a in in_args:
                if isinstance(a, CScalar):
                    t, w = a.descr, a.weak_t
                    if w is not False:
                        any_weak = True
                elif isinstance(a, ndarray):
                    t, w = a.dtype, False
                else:
                    raise RuntimeError('Need array or CScalar got %s' % type(a))
                in_types.append(t)
                weaks.append(w)
            in_types = tuple(in_types)
            weaks = tuple(weaks) if any_weak else None
            if not _check_should_use_weak_scalar(in_types, weaks):
                weaks = (False,) * len(in_args)
            op = cache.get((in_types, weaks), ())
            if op == ():
                op = self._guess_routine_from_in_types(in_types, weaks)
                cache[(in_types, weaks)] = op
        else:
            op = cache.get(dtype, ())
            if op == ():
                op = (out_ops or self)._guess_routine_from_dtype(dtype)
                cache[dtype] = op
        if op is not None:
            op.check_valid()
            return op
        if dtype is None:
            dtype = in_types
        raise TypeError('Wrong type (%s) of arguments for %s' % (dtype, name))

    def _guess_routine_from_in_types(self, in_types, weaks=None):
        for op in self.ops:
            for i in range(self.nin):
                it, ot = in_types[i], op.in_types[i]
                weak_t = weaks[i] if weaks is not None else False
                if not numpy.can_cast(it, ot):
                    if not weak_t:
                        break
                    try:
                        if numpy.result_type(weak_t(0), ot) != ot:
                            break
                    except TypeError:
                        break
            else:
                return op
        return None

    def _guess_routine_from_dtype(self, dtype):
        for op in self.ops:
            if all(t == dtype for t in op.out_types):
                return op
        return None


# ---------------------------------------------------------------------------
# Call-shape memo of the ufunc launcher.  Everything `ufunc.__call__` derives from the call's SHAPE --
# loop selection, broadcasting, output dtype / shape, the collapsed plan, prebuilt-vs-NVRTC routing -- depends
# only on (dtype, shape, strides, pointer alignment) of the array arguments and the types of the scalar ones.
# The first call of a shape goes through the full path and leaves the filled C structures behind; later calls
# patch the data pointers / scalar bytes into them and launch: one dict lookup and one ctypes crossing instead
# of ~30 us of Python (the reference spends 12-20 us in Cython here, docs/source/user_guide/performance.rst).
# Entries hold mutable ctypes buffers, so the memo is per thread.
# ---------------------------------------------------------------------------
_MISS = object()
_FAST_MAX = 512


def _fast_key(args, out):
    key = []
    for a in args:
        ta = type(a)
        if ta is ndarray:
            key.append((a.dtype, a._shape, a._strides, a.ptr & 15))
        elif ta is float or ta is bool:
            key.append(ta)
        elif ta is int:
            if not -(1 << 63) <= a < (1 << 63):
                return None
            key.append(ta)
        elif isinstance(a, numpy.generic):
            key.append(a.dtype)
        else:
            return None
    if out is not None:
        if type(out) is not ndarray:
            return None
        key.append((out.dtype, out._shape, out._strides, out.ptr & 15, 'out'))
    # plans depend on the A/B knobs: the tunables (epoch) and the planner's environment switch
    key.append(_memo_epoch)
    key.append(_os.environ.get('B200_EW_TILED_MODE'))
    return tuple(key)


def _scalar_key(tv, v):
    """Hashable identity of a scalar VALUE (bit pattern: -0.0 and 0.0 differ, NaNs with equal bits agree)."""
    if tv is float:
        return (tv, v.hex())
    if tv is int or tv is bool:
        return (tv, v)
    if isinstance(v, numpy.generic):
        return (tv, v.tobytes())
    return None


class _FastEntry:
    __slots__ = ('ops', 'plan', 'nargs', 'prebuilt', 'handle', 'threads', 'array_slots', 'scalar_slots',
                 'out_slot', 'out_dtype', 'out_shape', 'out_strides', 'out_size', 'empty', 'dry')


def _scalar_words(value, dtype, weak_t, lo, hi):
    """The two 64-bit words of a by-value scalar operand cast to `dtype` (CScalar.apply_dtype + raw_bytes)."""
    if weak_t is int and lo is not None and not (lo <= value <= hi):
        raise OverflowError('Python integer %d out of bounds for %s' % (value, dtype))
    if dtype.kind == 'b':
        value = bool(value)
    with numpy.errstate(over='ignore', invalid='ignore'):
        b = numpy.asarray(value).astype(dtype, casting='unsafe').tobytes()
    b = b + b'\0' * (16 - len(b))
    return int.from_bytes(b[:8], 'little', signed=True), int.from_bytes(b[8:16], 'little', signed=True)


class ufunc:
    """Universal function (drop-in for cupy.ufunc, cupy/_core/_kernel.pyx:1147-1493)."""

    def __init__(self, name, nin, nout, ops, preamble='', loop_prep='', doc='',
                 default_casting=None, out_ops=None, prebuilt=None, scatter_op=None):
        self.name = name
        self._scatter_op = scatter_op
        self.__name__ = name
        self.nin = nin
        self.nout = nout
        self.nargs = nin + nout
        self._ops = ops
        self._out_ops = out_ops
        self._preamble = preamble
        self._loop_prep = loop_prep
        self.__doc__ = doc
        self._default_casting = 'same_kind' if default_casting is None else default_casting
        self._prebuilt = _lib.UFUNC_IDS.get(prebuilt) if prebuilt else None
        self._params = tuple(ParameterInfo('T in%d' % i, True) for i in range(nin)) + \
            tuple(ParameterInfo('T out%d' % i, False) for i in range(nout))
        self._params_with_where = self._params[:nin] + (ParameterInfo('T _where', True),) + self._params[nin:]
        self._routine_cache = {}
        self._specs = {}
        # ufunc routines that read the loop index `i` (none of the built-ins) pin the loop order
        self._keeps_order = any(isinstance(op.routine, str) and _re.search(r'\b(i|_ind)\b', op.routine)
                                for op in ops.ops) or bool(_re.search(r'\b(i|_ind)\b', loop_prep or ''))

    def __repr__(self):
        return '<ufunc \'%s\'>' % self.name

    @property
    def types(self):
        return ['%s->%s' % (''.join(t.char for t in op.in_types), ''.join(t.char for t in op.out_types))
                for op in self._ops.ops]

    def __call__(self, *args, **kwargs):
        fusing = getattr(_thread_local, 'fusion', None)
        # ---- memoised call shape (plain arrays / scalars, at most an `out=`)
        fkey = None
        if fusing is None and (not kwargs or (len(kwargs) == 1 and 'out' in kwargs)):
            fout = kwargs.get('out') if kwargs else None
            if len(args) == self.nin or (fout is None and len(args) == self.nargs and self.nout == 1):
                if len(args) == self.nargs and self.nout == 1 and fout is None:
                    fkey = _fast_key(args[:self.nin], args[self.nin])
                    fout = args[self.nin]
                else:
                    fkey = _fast_key(args, fout)
                if fkey is not None:
                    memo = _thread_local.__dict__.setdefault('ufunc_memo', {}).setdefault(id(self), {})
                    entry = memo.get(fkey)
                    if entry is not None:
                        r = self._fast_launch(entry, args, fout)
                        if r is not _MISS:
                            return r
        for arg in args:
            if hasattr(arg, '__cupy_override_elementwise_kernel__'):
                return arg.__cupy_override_elementwise_kernel__(self, *args, **kwargs)
        if fusing is not None:
            return fusing.call_ufunc(self, *args, **kwargs)

        out = kwargs.pop('out', None)
        where = kwargs.pop('_where', None)
        has_where = where is not None
        dtype = kwargs.pop('dtype', None)
        casting = kwargs.pop('casting', self._default_casting)
        if dtype is not None:
            dtype = get_dtype(dtype)
        if kwargs:
            raise TypeError('Wrong arguments %s' % kwargs)
        n_args = len(args)
        if not (self.nin <= n_args <= self.nargs):
            raise TypeError(
                'Wrong number of arguments for {!r}. It must be either {} or {} (with outputs), '
                'but given {}.'.format(self.name, self.nin, self.nargs, n_args))
        in_args = args[:self.nin]
        out_args = args[self.nin:]
        if out is not None:
            if out_args:
                raise ValueError('Cannot specify \'out\' as both a positional and keyword argument')
            if isinstance(out, tuple):
                if len(out) != self.nout:
                    raise ValueError("The 'out' tuple must have exactly one entry per ufunc output")
                out_args = out
            else:
                if 1 != self.nout:
                    raise ValueError("'out' must be a tuple of arrays")
                out_args = (out,)
        in_args = _preprocess_args(in_args)
        out_args = [None if o is None else _convert_arg(o) for o in out_args]
        given_out_args = [o for o in out_args if o is not None]

        if has_where:
            w = _convert_arg(where)
            if isinstance(w, ndarray):
                if w.dtype != numpy.bool_:
                    raise TypeError('Cannot cast array data from %r to %r according to the rule \'safe\''
                                    % (w.dtype, numpy.dtype(bool)))
            else:
                w = CScalar(bool(w.value))
            where_args = [w]
        else:
            where_args = []

        ids_before = [id(a) for a in in_args]
        _copy_in_args_if_needed(in_args, given_out_args)
        _copy_in_args_if_needed(where_args, given_out_args)
        if fkey is not None and ids_before != [id(a) for a in in_args]:
            fkey = None                      # an input was copied because it overlaps the output: not memoised
        inout_args = in_args + where_args + given_out_args
        shape = _broadcast_core(inout_args)
        in_args = inout_args[:self.nin]
        where_args = inout_args[self.nin:self.nin + len(where_args)]

        op = self._ops.guess_routine(self.name, self._routine_cache, in_args, dtype, self._out_ops)
        out_args = _get_out_args_from_optionals(out_args, op.out_types, shape, casting)
        ret = out_args[0] if self.nout == 1 else tuple(out_args)
        if 0 in shape:
            return ret
        weak_ts = [a.weak_t if isinstance(a, CScalar) else None for a in in_args]
        for i, t in enumerate(op.in_types):
            if isinstance(in_args[i], CScalar):
                in_args[i].apply_dtype(t)

        all_args = in_args + where_args + out_args
        params = self._params_with_where if has_where else self._params
        n_in = self.nin + len(where_args)
        ops, plan = plan_elementwise(all_args, params, n_in, shape, keep_order=self._keeps_order)
        st = current_stream_ptr()

        # ---- prebuilt kernel?
        # (register-tiled calls are prebuilt for unary ufuncs only; NVRTC fixes each operand's access otherwise)
        if (self._prebuilt is not None and not has_where and self.nout == 1
                and not (plan.variant == _lib.EW_TILED_REG and self.nin > 1)
                and not (plan.variant == _lib.EW_FLAT and plan.staged_mask)
                and _prebuilt_fits(plan, ops, all_args)
                and all(a.dtype == t if isinstance(a, ndarray) else True for a, t in zip(in_args, op.in_types))
                and (self._prebuilt == 0 or out_args[0].dtype == op.out_types[0])):
            in_ids = (ctypes.c_int32 * self.nin)(*[_scalar.dtype_id(t) for t in op.in_types])
            if self._prebuilt == 0:   # copy: keyed by the memory dtypes
                in_ids[0] = _scalar.dtype_id(in_args[0].dtype if isinstance(in_args[0], ndarray) else op.in_types[0])
            if _lib.lib.b200_ufunc_supported(self._prebuilt, self.nin, in_ids, _scalar.dtype_id(out_args[0].dtype)):
                if _dryrun.enabled:
                    _dryrun.record('prebuilt_ufunc', name=self.name, variant=plan.variant, vec=plan.vec,
                                   ndim=plan.ndim, idx32=bool(plan.idx32), staged_mask=plan.staged_mask,
                                   tile_axis=plan.tile_axis, shape=tuple(plan.shape[:plan.ndim]))
                else:
                    _lib.check(_lib.lib.b200_ufunc_launch(self._prebuilt, ctypes.byref(plan), len(all_args), ops, st))
                if fkey is not None and not has_where and self.nout == 1:
                    self._remember(fkey, ops, plan, all_args, op, weak_ts, out_args[0], self._prebuilt, None, 0)
                return ret

        # ---- NVRTC route: the routine string inside the same tiler glue
        spec = self._get_spec(op, has_where)
        handle, threads = _launch_jit(self._kernel_name(all_args, has_where), spec, all_args, params, n_in, ops, plan)
        if fkey is not None and not has_where and self.nout == 1:
            self._remember(fkey, ops, plan, all_args, op, weak_ts, out_args[0], None, handle, threads)
        return ret

    def _remember(self, fkey, ops, plan, all_args, op, weak_ts, out, prebuilt, handle, threads):
        memo = _thread_local.__dict__.setdefault('ufunc_memo', {}).setdefault(id(self), {})
        if len(memo) >= _FAST_MAX:
            memo.clear()
        e = _FastEntry()
        e.ops, e.plan, e.nargs = ops, plan, len(all_args)
        e.prebuilt, e.handle, e.threads = prebuilt, handle, threads
        e.array_slots, e.scalar_slots = [], []
        for k in range(self.nin):
            a = all_args[k]
            if isinstance(a, ndarray):
                e.array_slots.append(k)
            else:
                t = op.in_types[k]
                lo = hi = None
                if t.kind in 'iu':
                    info = numpy.iinfo(t)
                    lo, hi = int(info.min), int(info.max)
                e.scalar_slots.append((k, t, weak_ts[k], lo, hi, {}))
        e.out_slot = len(all_args) - 1
        e.out_dtype, e.out_shape, e.out_strides, e.out_size = out.dtype, out._shape, out._strides, out.size
        e.empty = out._c_contiguous
        e.dry = dict(_dryrun.log[-1]) if (_dryrun.enabled and _dryrun.log) else None
        memo[fkey] = e

    def _fast_launch(self, e, args, out):
        """A call whose shape has been seen: patch pointers and scalar bytes into the remembered operand block."""
        if _dryrun.enabled and e.dry is None:
            return _MISS                         # remembered on a live run: the dry run records through the full path
        if out is not None:
            for k in e.array_slots:
                a = args[k]
                if a is not out and may_share_bounds(a, out):
                    return _MISS                     # the slow path copies overlapping inputs
        ops = e.ops
        for k in e.array_slots:
            ops[k].data = args[k].ptr
        for k, t, weak_t, lo, hi, seen in e.scalar_slots:
            v = args[k]
            tv = type(v)
            sk = _scalar_key(tv, v)
            words = seen.get(sk) if sk is not None else None
            if words is None:
                words = _scalar_words(v, t, weak_t if not isinstance(v, numpy.generic) else False, lo, hi)
                if sk is not None:
                    if len(seen) >= 64:
                        seen.clear()
                    seen[sk] = words
            ops[k].scalar[0] = words[0]
            ops[k].scalar[1] = words[1]
        if out is None:
            out = ndarray._fresh(e.out_shape, e.out_dtype, e.out_strides, e.out_size)
        ops[e.out_slot].data = out.ptr
        if _dryrun.enabled:
            if e.dry is not None:
                _dryrun.log.append(dict(e.dry))
            return out
        st = current_stream_ptr()
        if e.prebuilt is not None:
            _lib.check(_lib.lib.b200_ufunc_launch(e.prebuilt, ctypes.byref(e.plan), e.nargs, ops, st))
        else:
            _lib.check(_lib.lib.b200_jit_ew_launch_ex(e.handle, ctypes.byref(e.plan), e.nargs, ops, e.threads, 0, None, st))
        return out

    def _get_spec(self, op, has_where):
        key = (op, has_where)
        spec = self._specs.get(key)
        if spec is None:
            spec = _codegen.EwSpec(
                mode='ufunc', operation=op.routine, preamble=self._preamble, loop_prep=self._loop_prep,
                after_loop='', options=(), in_types=op.in_types, out_types=op.out_types, has_where=has_where)
            self._specs[key] = spec
        return spec

    def _kernel_name(self, args, has_where):
        name = self.name + ('_where' if has_where else '')
        words = []
        for a in args:
            if isinstance(a, ndarray):
                words.append(a.dtype.name)
            else:
                words.append(a.descr.name.rstrip('0123456789'))
        return '{}__{}'.format(name, '_'.join(words))

    def outer(self, A, B, **kwargs):
        from cupy_b200._core import _ndarray
        A = _ndarray.asarray(A)
        B = _ndarray.asarray(B)
        A = A.reshape(A.shape + (1,) * B.ndim)
        B = B.reshape((1,) * (A.ndim - B.ndim) + B.shape) if B.ndim else B
        return self(A, B, **kwargs)

    def at(self, a, indices, b=None):
        """In-place `a[indices] = ufunc(a[indices], b)` with repeated indices accumulated (_kernel.pyx:1446-1457)."""
        if self._scatter_op is None:
            raise NotImplementedError('`%s.at` is not supported yet' % self.name)
        from cupy_b200._core import _scatter
        _scatter.scatter_op(a, indices, b, self._scatter_op)

    def reduce(self, array, axis=0, dtype=None, out=None, keepdims=False):
        if self.name == 'cupy_add':
            return array.sum(axis, dtype, out, keepdims)
        if self.name == 'cupy_multiply':
            return array.prod(axis, dtype, out, keepdims)
        if self.name in ('cupy_maximum', 'cupy_minimum') and dtype is None:
            return (array.max if self.name == 'cupy_maximum' else array.min)(axis, out, keepdims)
        raise NotImplementedError('`%s.reduce` is not supported yet' % self.name)

    def accumulate(self, array, axis=0, dtype=None, out=None):
        if self.name == 'cupy_add':
            return array.cumsum(axis, dtype, out)
        if self.name == 'cupy_multiply':
            return array.cumprod(axis, dtype, out)
        raise NotImplementedError('`%s.accumulate` is not supported yet' % self.name)

    def reduceat(self, array, indices, axis=0, dtype=None, out=None):
        if self.name == 'cupy_add':
            from cupy_b200._core import _scatter
            return _scatter.add_reduceat(array, indices, axis, dtype, out)
        raise NotImplementedError('`%s.reduceat` is not supported yet' % self.name)


def create_ufunc(name, ops, routine=None, preamble='', doc='', default_casting=None,
                 loop_prep='', out_ops=None, prebuilt=None, scatter_op=None):
    ops_ = _Ops.from_tuples(ops, routine)
    _out_ops = None if out_ops is None else _Ops.from_tuples(out_ops, routine)
    return ufunc(name, ops_.nin, ops_.nout, ops_, preamble, loop_prep, doc,
                 default_casting=default_casting, out_ops=_out_ops, prebuilt=prebuilt, scatter_op=scatter_op)


# ---------------------------------------------------------------------------
# elementwise_copy (cupy/_core/_ufuncs.py:7-13) and helpers built on the engine
# ---------------------------------------------------------------------------
_copy_ufunc = create_ufunc(
    'cupy_copy',
    ('?->?', 'b->b', 'B->B', 'h->h', 'H->H', 'i->i', 'I->I', 'l->l', 'L->L',
     'q->q', 'Q->Q', 'e->e', 'f->f', 'd->d'),
    'out0 = in0', default_casting='unsafe', prebuilt='copy')


def elementwise_copy(src, dst):
    """dst[...] = src with dtype cast (`astype`, `copy`, `fill`, `out=` handling)."""
    return _copy_ufunc(src, dst)


_arange_memo = []


def _arange_kernel():
    if not _arange_memo:
        _arange_memo.append(ElementwiseKernel('T start, T step', 'T y', 'y = start + step * i', 'cupy_arange'))
    return _arange_memo[0]
