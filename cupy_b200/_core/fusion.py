"""cupy_b200.fuse -- kernel fusion of elementwise chains (+ one trailing reduction).

Drop-in for `cupy.fuse` (cupy/_core/fusion.pyx, new_fusion.pyx; entry cupy/__init__.py:899;
hooks cupy/_core/_kernel.pyx:1261-1262 and cupy/_math/sumprod.py:30-39): the decorated
function is run ONCE per argument signature (dtype / ndim / scalar kind) on symbolic
variables; every ufunc call and operator it makes is recorded; the recording becomes the
body of ONE ElementwiseKernel (or the pre-map of ONE ReductionKernel when the function ends
in `sum/prod/max/min`), which then runs on the same tilers as every other kernel -- so
`exp(x) + v` on a transposed `x` is one pass over HBM instead of two.

What differs from the reference: there is no separate fusion code generator -- the trace is
rendered into the operation string of the public kernel classes, so fused kernels get the
FLAT / ROWWISE / TILED_REG classification, vector widths and NVRTC cache for free.
"""
from __future__ import annotations

import functools

import numpy

from cupy_b200._core import _kernel
from cupy_b200._core._ndarray import ndarray
from cupy_b200._core._scalar import CScalar, get_dtype, get_typename

_thread_local = _kernel._thread_local


class _Var:
    """A value inside a fused function: a kernel parameter or a temporary."""

    __array_priority__ = 200      # NumPy scalars defer to our reflected operators

    def __init__(self, trace, name, dtype, is_array, ndim=0, weak_t=False, param_index=None):
        self._trace = trace
        self.name = name
        self.dtype = get_dtype(dtype)
        self.is_array = is_array
        self.ndim = ndim
        self.weak_t = weak_t
        self.param_index = param_index
        self.param_name = name if param_index is not None else None   # `name` follows in-place updates
        self.reduced = False
        self.used = False             # read by some recorded operation, returned, or reduced

    # ---- the subset of the ndarray surface a fused function may touch
    def _bin(self, uf, other, swap=False):
        return uf(other, self) if swap else uf(self, other)

    def __add__(self, o): return self._bin(_m().add, o)
    def __radd__(self, o): return self._bin(_m().add, o, True)
    def __sub__(self, o): return self._bin(_m().subtract, o)
    def __rsub__(self, o): return self._bin(_m().subtract, o, True)
    def __mul__(self, o): return self._bin(_m().multiply, o)
    def __rmul__(self, o): return self._bin(_m().multiply, o, True)
    def __truediv__(self, o): return self._bin(_m().true_divide, o)
    def __rtruediv__(self, o): return self._bin(_m().true_divide, o, True)
    def __pow__(self, o): return self._bin(_m().power, o)
    def __rpow__(self, o): return self._bin(_m().power, o, True)
    def __neg__(self): return _m().negative(self)
    def __abs__(self): return _m().absolute(self)
    def __lt__(self, o): return self._bin(_m().less, o)
    def __le__(self, o): return self._bin(_m().less_equal, o)
    def __gt__(self, o): return self._bin(_m().greater, o)
    def __ge__(self, o): return self._bin(_m().greater_equal, o)
    def __eq__(self, o): return self._bin(_m().equal, o)
    def __ne__(self, o): return self._bin(_m().not_equal, o)
    __hash__ = object.__hash__

    def __iadd__(self, o): return self._trace.assign(self, self._bin(_m().add, o))
    def __isub__(self, o): return self._trace.assign(self, self._bin(_m().subtract, o))
    def __imul__(self, o): return self._trace.assign(self, self._bin(_m().multiply, o))
    def __itruediv__(self, o): return self._trace.assign(self, self._bin(_m().true_divide, o))

    def astype(self, dtype, copy=True):
        return self._trace.cast(self, get_dtype(dtype))

    def copy(self):
        return self._trace.cast(self, self.dtype)

    def sum(self, axis=None, dtype=None, out=None, keepdims=False):
        return _m()._ndarray_sum(self, axis, dtype, out, keepdims)

    def prod(self, axis=None, dtype=None, out=None, keepdims=False):
        return _m()._ndarray_prod(self, axis, dtype, out, keepdims)

    def max(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as st
        return st._amax(self, axis=axis, out=out, dtype=None, keepdims=keepdims)

    def min(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as st
        return st._amin(self, axis=axis, out=out, dtype=None, keepdims=keepdims)

    def __bool__(self):
        raise TypeError('the truth value of a value inside a fused function is not available while tracing '
                        '(data-dependent Python control flow cannot be fused)')

    # ---- hooks the kernels look for (cupy/_core/_kernel.pyx:1261-1262, _reduction.pyx:609-611)
    def __cupy_override_elementwise_kernel__(self, kernel, *args, **kwargs):
        return self._trace.call_ufunc(kernel, *args, **kwargs)

    def __cupy_override_reduction_kernel__(self, kernel, axis, dtype, out, keepdims):
        return self._trace.call_reduction(kernel, self, axis, dtype, out, keepdims)


def _m():
    from cupy_b200._core import _routines_math
    return _routines_math


class _Trace:
    def __init__(self, name):
        self.name = name
        self.params = []          # _Var, in argument order
        self.steps = []           # text of each recorded operation
        self.preambles = []
        self.n_tmp = 0
        self.assigned = {}        # param index -> _Var written back into that array (in-place ops / out=)
        self.reduction = None

    # ---- building blocks
    def new_param(self, arg, k):
        if isinstance(arg, ndarray):
            v = _Var(self, '_p%d' % k, arg.dtype, True, arg.ndim, param_index=k)
        else:
            s = CScalar(arg)
            v = _Var(self, '_p%d' % k, s.descr, False, 0, weak_t=s.weak_t, param_index=k)
        self.params.append(v)
        return v

    def new_tmp(self, dtype, is_array, ndim):
        v = _Var(self, '_t%d' % self.n_tmp, dtype, is_array, ndim)
        self.n_tmp += 1
        return v

    def _check_open(self):
        if self.reduction is not None:
            raise NotImplementedError('a fused function may only end in a reduction: no operation can follow it')

    def _operand(self, a):
        """-> (C expression, dtype, weak type, is_array, ndim)"""
        if isinstance(a, _Var):
            if a._trace is not self:
                raise ValueError('value belongs to another fused function')
            if a.reduced:
                raise NotImplementedError('the result of the reduction cannot be used inside the fused function')
            a.used = True
            return a.name, a.dtype, a.weak_t, a.is_array, a.ndim
        if isinstance(a, ndarray):
            raise TypeError('arrays must be passed to a fused function as arguments, not captured from outside')
        s = CScalar(a)
        return _literal(s), s.descr, s.weak_t, False, 0

    def cast(self, v, dtype):
        self._check_open()
        expr, _, _, is_array, ndim = self._operand(v)
        out = self.new_tmp(dtype, is_array, ndim)
        self.steps.append('%s %s = static_cast<%s>(%s);' % (get_typename(dtype), out.name, get_typename(dtype), expr))
        return out

    def assign(self, target, value):
        """In-place operators and `out=`: the value is written back into a parameter array."""
        if not (isinstance(target, _Var) and target.param_index is not None and target.is_array):
            raise NotImplementedError('only argument arrays can be updated in place inside a fused function')
        if value.dtype != target.dtype:
            if not numpy.can_cast(value.dtype, target.dtype, 'same_kind'):
                raise TypeError("Cannot cast ufunc output from %r to %r with casting rule 'same_kind'"
                                % (value.dtype, target.dtype))
            value = self.cast(value, target.dtype)
        self.assigned[target.param_index] = value
        # every reference to the parameter (the user's own variable included: `cp.add(a, b, out=a); a * 2`)
        # reads the new value from here on, as with a real array
        target.name = value.name
        return target

    # ---- recorded calls
    def call_ufunc(self, uf, *args, **kwargs):
        self._check_open()
        if not isinstance(uf, _kernel.ufunc):
            raise NotImplementedError('only ufuncs can be fused (got %r)' % (uf,))
        out = kwargs.pop('out', None)
        dtype = kwargs.pop('dtype', None)
        kwargs.pop('casting', None)
        if kwargs:
            raise TypeError('Wrong arguments %s' % kwargs)
        if uf.nout != 1:
            raise NotImplementedError('ufuncs with several outputs cannot be fused')
        if len(args) == uf.nin + 1:
            args, out = args[:uf.nin], args[uf.nin]
        if len(args) != uf.nin:
            raise TypeError('Wrong number of arguments for %r' % uf.name)
        ops = [self._operand(a) for a in args]
        in_types = tuple(o[1] for o in ops)
        if dtype is None:
            any_weak = any(o[2] is not False for o in ops)
            weaks = tuple(o[2] for o in ops) if any_weak else None
            if not _kernel._check_should_use_weak_scalar(in_types, weaks):
                weaks = None
            op = uf._ops._guess_routine_from_in_types(in_types, weaks)
        else:
            op = (uf._out_ops or uf._ops)._guess_routine_from_dtype(get_dtype(dtype))
        if op is None:
            raise TypeError('Wrong type (%s) of arguments for %s' % (in_types if dtype is None else dtype, uf.name))
        op.check_valid()
        is_array = any(o[3] for o in ops)
        ndim = max(o[4] for o in ops)
        res = self.new_tmp(op.out_types[0], is_array, ndim)
        if uf._preamble and uf._preamble not in self.preambles:
            self.preambles.append(uf._preamble)
        lines = ['%s %s;' % (get_typename(op.out_types[0]), res.name), '{']
        for k, (o, t) in enumerate(zip(ops, op.in_types)):
            lines.append('  typedef %s in%d_type;' % (get_typename(t), k))
            expr = o[0]
            if o[1].kind == 'b' and t.kind != 'b':
                expr = '(%s) ? 1 : 0' % expr
            lines.append('  const in%d_type in%d = static_cast<in%d_type>(%s);' % (k, k, k, expr))
        lines.append('  typedef %s out0_type;' % get_typename(op.out_types[0]))
        lines.append('  out0_type out0;')
        lines.append('  ' + op.routine + ';')
        lines.append('  %s = out0;' % res.name)
        lines.append('}')
        self.steps.append('\n'.join(lines))
        if out is not None:
            return self.assign(out, res)
        return res

    def call_reduction(self, kernel, a, axis, dtype, out, keepdims):
        self._check_open()
        if out is not None:
            raise NotImplementedError('out= of a reduction inside a fused function')
        if not a.is_array:
            raise TypeError('reduction of a scalar inside a fused function')
        a.used = True
        self.reduction = (kernel, a, axis, dtype, keepdims)
        r = _Var(self, '_reduced', a.dtype, True, 0)
        r.reduced = True
        return r


def _literal(s):
    """C literal of a Python / NumPy scalar constant."""
    v, dt = s.value, s.descr
    if dt.kind == 'b':
        return 'true' if v else 'false'
    if dt.kind in 'iu':
        return ('%dull' % int(v)) if dt.kind == 'u' and int(v) >= (1 << 63) else '%dll' % int(v)
    f = float(v)
    if f != f:
        return '(0.0 / 0.0)' if dt.itemsize == 8 else '(0.0f / 0.0f)'
    if f in (float('inf'), float('-inf')):
        return ('-' if f < 0 else '') + ('(1.0 / 0.0)' if dt.itemsize == 8 else '(1.0f / 0.0f)')
    return repr(f) if dt.itemsize == 8 else '%sf' % repr(float(numpy.float32(f)))


class _FusedKernel:
    """One traced signature: the kernel object and how to call / unpack it."""

    def __init__(self, func, name, args):
        trace = _Trace(name)
        params = [trace.new_param(a, k) for k, a in enumerate(args)]
        prev = getattr(_thread_local, 'fusion', None)
        _thread_local.fusion = trace
        try:
            ret = func(*params)
        finally:
            _thread_local.fusion = prev
        self.trace = trace
        self.n_args = len(args)
        self.ret_none = ret is None
        self.ret_tuple = isinstance(ret, tuple)
        rets = () if ret is None else (ret if isinstance(ret, tuple) else (ret,))
        for r in rets:
            if not isinstance(r, _Var):
                raise TypeError('a fused function must return values computed from its arguments (got %r)' % (r,))
            r.used = True
        # arguments the function never reads or updates stay out of the kernel: they must not take part in the
        # broadcast that shapes the loop (`lambda x, y: y * y` has y's shape)
        self.used = [k for k, p in enumerate(trace.params) if p.used or k in trace.assigned]
        self.used_params = [trace.params[k] for k in self.used]
        in_decl = ', '.join('%s %s' % (p.dtype.name, p.param_name) for p in self.used_params)   # NumPy type names
        preamble = '\n'.join(trace.preambles)
        body = '\n'.join(trace.steps)
        self.inplace = sorted(trace.assigned.items())
        if trace.reduction is not None:
            self._build_reduction(rets, in_decl, preamble, body, name)
            return
        # ---- elementwise: outputs = returned values, then arrays updated in place
        self.reduction = False
        outs, self.ret_spec = [], []
        for r in rets:
            if not r.is_array:
                raise NotImplementedError('a fused function must return arrays')
            outs.append(r)
        out_decl, stores = [], []
        for k, r in enumerate(outs):
            out_decl.append('%s _o%d' % (r.dtype.name, k))
            stores.append('_o%d = %s;' % (k, r.name))
        for k, (pidx, v) in enumerate(self.inplace):
            out_decl.append('%s _w%d' % (trace.params[pidx].dtype.name, k))
            stores.append('_w%d = %s;' % (k, trace.params[pidx].name))
        if not out_decl:
            raise ValueError('the fused function neither returns nor updates an array')
        self.n_ret = len(outs)
        self.ret_dtypes = [r.dtype for r in outs]
        # every output is assigned exactly once, at the end: the kernel never has to load them
        op = body + '\n' + '\n'.join(stores)
        self.kernel = _kernel.ElementwiseKernel(in_decl, ', '.join(out_decl), op, name, preamble=preamble,
                                                return_tuple=True, _write_only_outputs=True)

    def _build_reduction(self, rets, in_decl, preamble, body, name):
        from cupy_b200._core._reduction import ReductionKernel
        trace = self.trace
        kernel, a, axis, dtype, keepdims = trace.reduction
        if len(rets) != 1 or not rets[0].reduced or self.inplace:
            raise NotImplementedError('a fused function that reduces must return exactly the reduction result')
        if dtype is None:
            op = kernel._ops._guess_routine_from_in_types((a.dtype,), None)
        else:
            op = kernel._ops._guess_routine_from_dtype(get_dtype(dtype))
        if op is None:
            raise TypeError('Wrong type (%s) of arguments for %s' % (a.dtype if dtype is None else dtype, kernel.name))
        map_expr, reduce_expr, post_map_expr, reduce_type = op.routine
        out_t = op.out_types[0]
        if reduce_type is None:
            reduce_type = get_typename(out_t)
        args_decl = ', '.join('const %s& %s' % (get_typename(p.dtype), p.param_name) for p in self.used_params)
        args_call = ', '.join(p.param_name for p in self.used_params)
        pre = [preamble, kernel.preamble,
               'typedef %s type_in0_raw;' % get_typename(a.dtype),
               'typedef %s type_out0_raw;' % get_typename(out_t),
               '__device__ __forceinline__ %s _fused_chain(%s) {\n%s\n  return %s;\n}'
               % (get_typename(a.dtype), args_decl, body, a.name),
               '__device__ __forceinline__ %s _fused_premap(const %s in0) { return %s; }'
               % (reduce_type, get_typename(op.in_types[0]), map_expr)]
        import re
        post = re.sub(r'\bout0\b', '_o0', post_map_expr)
        self.kernel = ReductionKernel(in_decl, '%s _o0' % out_t.name,
                                      '_fused_premap(_fused_chain(%s))' % args_call, reduce_expr, post,
                                      kernel.identity if kernel.identity != '' else None, name,
                                      reduce_type=reduce_type, preamble='\n'.join(p for p in pre if p))
        self.reduction = True
        self.red_axis, self.red_keepdims = axis, keepdims


class Fusion:
    """The object `fuse` returns: callable like the function it wraps."""

    def __init__(self, func, name=None):
        self.func = func
        self.name = name or getattr(func, '__name__', 'fused')
        if not self.name.isidentifier():
            self.name = 'fused'
        self._cache = {}
        functools.update_wrapper(self, func)

    def __repr__(self):
        return '<Fusion %r>' % self.name

    @staticmethod
    def _signature(args):
        key = []
        for a in args:
            if isinstance(a, ndarray):
                key.append(('a', a.dtype.char, a.ndim))
            elif isinstance(a, numpy.generic):
                key.append(('n', a.dtype.char))
            elif isinstance(a, (bool, int, float)):
                key.append(('p', type(a).__name__))
            else:
                return None
        return tuple(key)

    def __call__(self, *args, **kwargs):
        if kwargs:
            raise TypeError('keyword arguments are not supported by fused functions')
        # nested fusion / plain NumPy call: run the Python function as it is
        if getattr(_thread_local, 'fusion', None) is not None or any(isinstance(a, _Var) for a in args):
            return self.func(*args)
        if not any(isinstance(a, ndarray) for a in args):
            return self.func(*args)
        key = self._signature(args)
        if key is None:
            raise TypeError('unsupported argument type for a fused function: %s' % ', '.join(type(a).__name__ for a in args))
        fk = self._cache.get(key)
        if fk is None:
            fk = _FusedKernel(self.func, self.name, args)
            self._cache[key] = fk
        return _run(fk, args)


def _run(fk, args):
    targets = [args[pidx] for pidx, _ in fk.inplace]
    args = [args[k] for k in fk.used]
    if fk.reduction:
        return fk.kernel(*args, axis=fk.red_axis, keepdims=fk.red_keepdims)
    if fk.n_ret == 0:
        fk.kernel(*args, *targets)
        return None
    if targets:
        # returned arrays are allocated with the broadcast shape of the call, like the kernel would
        shape = numpy.broadcast_shapes(*[a.shape for a in args if isinstance(a, ndarray)])
        rets = [ndarray(shape, dt) for dt in fk.ret_dtypes]
        fk.kernel(*args, *rets, *targets)
    else:
        rets = list(fk.kernel(*args))
    if fk.ret_tuple:
        return tuple(rets)
    return rets[0]


def fuse(*args, **kwargs):
    """Decorator: `@fuse()` / `@fuse(kernel_name='name')` / `@fuse` (cupy/_core/fusion.pyx `fuse`)."""
    def wrap(f, kernel_name=None):
        return Fusion(f, kernel_name)
    if len(args) == 1 and len(kwargs) == 0 and callable(args[0]):
        return wrap(args[0])
    kernel_name = kwargs.pop('kernel_name', None)
    if args or kwargs:
        raise TypeError('fuse() takes only the keyword argument kernel_name')
    return lambda f: wrap(f, kernel_name)
