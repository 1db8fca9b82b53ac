"""`ndarray`: a strided view over device memory -- the array adapter the kernel
engine launches on.

The reference's container (cupy/_core/core.pyx, 3378 lines: memory pool,
indexing, hundreds of methods) is OUT of scope.
What the hot path needs from it is small and is what this class provides:
shape / byte strides / dtype / base pointer, the contiguity flags the launcher
keys on (core.pyx `_c_contiguous`, `_f_contiguous`), zero-copy views
(transpose, reshape, basic slicing, broadcast), `__cuda_array_interface__`
interop, and the method surface `sum/max/min/argmax/argmin/mean/var/std/
cumsum/cumprod/prod` (core.pyx:1288-1465) plus arithmetic operators
(core.pyx:1560-1700) that delegate to the engine.

Device memory comes from PyTorch's caching allocator (a flat uint8 tensor per
allocation): PyTorch is the plumbing for memory and streams, not the compute.
"""
from __future__ import annotations

import numpy
import torch

from cupy_b200._core import _dryrun, _scalar

_TORCH_DTYPES = {
    numpy.dtype('bool'): torch.bool, numpy.dtype('int8'): torch.int8,
    numpy.dtype('uint8'): torch.uint8, numpy.dtype('int16'): torch.int16,
    numpy.dtype('int32'): torch.int32, numpy.dtype('int64'): torch.int64,
    numpy.dtype('float16'): torch.float16, numpy.dtype('float32'): torch.float32,
    numpy.dtype('float64'): torch.float64,
    numpy.dtype('uint16'): torch.uint16, numpy.dtype('uint32'): torch.uint32,
    numpy.dtype('uint64'): torch.uint64,
}
_NUMPY_DTYPES = {v: k for k, v in _TORCH_DTYPES.items()}


class AxisError(ValueError, IndexError):
    pass


try:  # share NumPy's class so `except numpy.exceptions.AxisError` works
    from numpy.exceptions import AxisError  # noqa: F811
except Exception:  # pragma: no cover
    pass


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _stream_ctx(stream):
    """Context that makes `stream` (cupy_b200.cuda.Stream, torch stream or raw pointer) current."""
    if stream is None:
        return _NullCtx()
    if hasattr(stream, '_st'):
        return torch.cuda.stream(stream._st)
    if isinstance(stream, torch.cuda.Stream):
        return torch.cuda.stream(stream)
    return torch.cuda.stream(torch.cuda.ExternalStream(int(getattr(stream, 'ptr', stream))))


def _cai_stream():
    if _dryrun.enabled or not torch.cuda.is_available():
        return 1
    p = torch.cuda.current_stream().cuda_stream
    return p if p else 1


# torch's raw accessors: the cudaStream_t of the calling thread's current stream without building a
# torch.cuda.Stream object around it (5 us on the GPU box's host, a third of a small call; profiles/r02_host_cost.log)
_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_raw_device = getattr(torch._C, '_cuda_getDevice', None)


def current_stream_ptr():
    if _dryrun.enabled:
        return 0
    if _raw_stream is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _prod(seq):
    r = 1
    for s in seq:
        r *= int(s)
    return r


def _c_strides(shape, itemsize):
    strides = []
    st = itemsize
    for s in reversed(shape):
        strides.append(st)
        st *= max(int(s), 1)
    return tuple(reversed(strides))


def _f_strides(shape, itemsize):
    strides = []
    st = itemsize
    for s in shape:
        strides.append(st)
        st *= max(int(s), 1)
    return tuple(strides)


def _is_c_contiguous(shape, strides, itemsize):
    st = itemsize
    for s, t in zip(reversed(shape), reversed(strides)):
        if s == 0:
            return True
        if s != 1:
            if t != st:
                return False
            st *= s
    return True


def _is_f_contiguous(shape, strides, itemsize):
    st = itemsize
    for s, t in zip(shape, strides):
        if s == 0:
            return True
        if s != 1:
            if t != st:
                return False
            st *= s
    return True


def normalize_axis_index(axis, ndim):
    if not -ndim <= axis < ndim:
        raise AxisError('axis %d is out of bounds for array of dimension %d' % (axis, ndim))
    return axis + ndim if axis < 0 else axis


class _Flags:
    __slots__ = ('c_contiguous', 'f_contiguous', 'owndata')

    def __init__(self, c, f, own):
        self.c_contiguous, self.f_contiguous, self.owndata = c, f, own

    def __getitem__(self, name):
        return getattr(self, name.lower())

    def __repr__(self):
        return '  C_CONTIGUOUS : %s\n  F_CONTIGUOUS : %s\n  OWNDATA : %s' % (
            self.c_contiguous, self.f_contiguous, self.owndata)


_UFUNCS = {}      # operator name -> ufunc, filled on first use (the routines module imports this one)


class ndarray:
    """N-dimensional strided device array (see module docstring)."""

    __array_priority__ = 100
    __slots__ = ('_mem', 'ptr', '_shape', '_strides', 'dtype', 'base',
                 '_c_contiguous', '_f_contiguous', 'size', '__weakref__')

    def __init__(self, shape, dtype=float, memptr=None, strides=None, order='C', _mem=None, _base=None):
        if isinstance(shape, (int, numpy.integer)):
            shape = (shape,)
        shape = tuple(int(s) for s in shape)
        if any(s < 0 for s in shape):
            raise ValueError('negative dimensions are not allowed')
        self.dtype = _scalar.get_dtype(dtype)
        _scalar.dtype_id(self.dtype)
        itemsize = self.dtype.itemsize
        self._shape = shape
        self.size = _prod(shape)
        if strides is None:
            strides = _f_strides(shape, itemsize) if order in ('F', 'f') else _c_strides(shape, itemsize)
        self._strides = tuple(int(s) for s in strides)
        if memptr is None:
            nbytes = max(self.size * itemsize, 1)
            if _dryrun.enabled:
                self._mem = None
                self.ptr = _dryrun.fake_alloc(nbytes)
            else:
                self._mem = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
                self.ptr = self._mem.data_ptr()
        else:
            self._mem = _mem
            self.ptr = int(memptr)
        self.base = _base
        self._update_contiguity()

    # ---- construction helpers -------------------------------------------------
    @classmethod
    def _fresh(cls, shape, dtype, strides, size):
        """A new C-contiguous array whose metadata the caller already holds (the launcher's memoised
        call shapes): skips the normalisation of __init__."""
        self = object.__new__(cls)
        self.dtype = dtype
        self._shape = shape
        self._strides = strides
        self.size = size
        nbytes = size * dtype.itemsize
        if _dryrun.enabled:
            self._mem = None
            self.ptr = _dryrun.fake_alloc(nbytes)
        else:
            self._mem = torch.empty(nbytes if nbytes > 0 else 1, dtype=torch.uint8, device='cuda')
            self.ptr = self._mem.data_ptr()
        self.base = None
        self._c_contiguous = True
        self._f_contiguous = _is_f_contiguous(shape, strides, dtype.itemsize)
        return self

    def _update_contiguity(self):
        isz = self.dtype.itemsize
        self._c_contiguous = _is_c_contiguous(self._shape, self._strides, isz)
        self._f_contiguous = _is_f_contiguous(self._shape, self._strides, isz)

    def _view(self, shape, strides, ptr=None, dtype=None):
        base = self if self.base is None else self.base
        return ndarray(shape, self.dtype if dtype is None else dtype,
                       memptr=self.ptr if ptr is None else ptr, strides=strides,
                       _mem=self._mem, _base=base)

    @classmethod
    def _from_pointer(cls, ptr, shape, dtype, strides=None, owner=None):
        """Wrap foreign device memory (no ownership): used for CAI / torch interop
        and by the CPU-side launcher tests (fake pointers, never dereferenced)."""
        return cls(shape, dtype, memptr=ptr, strides=strides, _mem=owner)

    # ---- basic properties -------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, newshape):
        v = self.reshape(newshape)
        if v.ptr != self.ptr or (v.base is None and v is not self):
            raise AttributeError('incompatible shape for a non-contiguous array')
        self._shape, self._strides = v._shape, v._strides
        self._update_contiguity()

    @property
    def strides(self):
        return self._strides

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def flags(self):
        return _Flags(self._c_contiguous, self._f_contiguous, self.base is None)

    @property
    def data(self):
        return self

    @property
    def device(self):
        return self._mem.device if self._mem is not None else torch.device('cuda', torch.cuda.current_device())

    @property
    def T(self):
        return self.transpose()

    @property
    def __cuda_array_interface__(self):
        desc = {
            'shape': self._shape,
            'typestr': self.dtype.str,
            'descr': self.dtype.descr,
            'data': (self.ptr, False),
            'version': 3,
            # consumers must order their work after what is enqueued on the producing stream: torch's current one
            # (1 = the legacy default stream, CUDA Array Interface v3)
            'stream': _cai_stream(),
        }
        if not self._c_contiguous:
            desc['strides'] = self._strides
        return desc

    def __len__(self):
        if not self._shape:
            raise TypeError('len() of unsized object')
        return self._shape[0]

    # ---- views ----------------------------------------------------------------
    def transpose(self, *axes):
        if len(axes) == 1 and (axes[0] is None or isinstance(axes[0], (tuple, list))):
            axes = axes[0]
        if not axes:
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(normalize_axis_index(int(a), self.ndim) for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError('axes don\'t match array')
        return self._view(tuple(self._shape[a] for a in axes), tuple(self._strides[a] for a in axes))

    def swapaxes(self, a, b):
        axes = list(range(self.ndim))
        a, b = normalize_axis_index(a, self.ndim), normalize_axis_index(b, self.ndim)
        axes[a], axes[b] = axes[b], axes[a]
        return self.transpose(axes)

    def _reshape_strides(self, newshape):
        """Strides of a no-copy reshape, or None (NumPy's _attempt_nocopy_reshape)."""
        isz = self.dtype.itemsize
        if self.size == 0:
            return _c_strides(newshape, isz)
        old = [(s, t) for s, t in zip(self._shape, self._strides) if s != 1]
        newstrides = [0] * len(newshape)
        oi, ni = 0, 0
        oldn, newn = len(old), len(newshape)
        while oi < oldn and ni < newn:
            np_, op_ = newshape[ni], old[oi][0]
            oj, nj = oi + 1, ni + 1
            while np_ != op_:
                if np_ < op_:
                    np_ *= newshape[nj]
                    nj += 1
                else:
                    op_ *= old[oj][0]
                    oj += 1
            for k in range(oi, oj - 1):
                if old[k][1] != old[k + 1][0] * old[k + 1][1]:
                    return None
            st = old[oj - 1][1]
            for k in range(nj - 1, ni - 1, -1):
                newstrides[k] = st
                st *= newshape[k]
            oi, ni = oj, nj
        last = newstrides[ni - 1] if ni > 0 else isz
        for k in range(ni, newn):
            newstrides[k] = last
        return tuple(newstrides)

    def reshape(self, *shape, order='C'):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if shape.count(-1) > 1:
            raise ValueError('can only specify one unknown dimension')
        if -1 in shape:
            known = _prod(s for s in shape if s != -1)
            if known == 0 or self.size % known:
                raise ValueError('cannot reshape array of size %d into shape %s' % (self.size, tuple(shape)))
            shape[shape.index(-1)] = self.size // known
        shape = tuple(shape)
        if _prod(shape) != self.size:
            raise ValueError('cannot reshape array of size %d into shape %s' % (self.size, shape))
        strides = self._reshape_strides(shape)
        if strides is not None:
            return self._view(shape, strides)
        return self.copy()._view_owned(shape)

    def _view_owned(self, shape):
        return self._view(shape, _c_strides(shape, self.dtype.itemsize))

    def ravel(self, order='C'):
        return self.reshape(-1)

    def flatten(self):
        return self.copy().reshape(-1)

    def view(self, dtype=None):
        if dtype is None:
            return self._view(self._shape, self._strides)
        dtype = _scalar.get_dtype(dtype)
        if dtype.itemsize != self.dtype.itemsize:
            raise ValueError('view() only supports dtypes of the same itemsize')
        return self._view(self._shape, self._strides, dtype=dtype)

    def squeeze(self, axis=None):
        if axis is None:
            keep = [i for i, s in enumerate(self._shape) if s != 1]
        else:
            axes = (axis,) if isinstance(axis, int) else tuple(axis)
            axes = {normalize_axis_index(a, self.ndim) for a in axes}
            for a in axes:
                if self._shape[a] != 1:
                    raise ValueError('cannot select an axis to squeeze out which has size not equal to one')
            keep = [i for i in range(self.ndim) if i not in axes]
        return self._view(tuple(self._shape[i] for i in keep), tuple(self._strides[i] for i in keep))

    def broadcast_to(self, shape):
        shape = tuple(int(s) for s in shape)
        nd = len(shape)
        if nd < self.ndim:
            raise ValueError('input operand has more dimensions than allowed by the axis remapping')
        strides = [0] * nd
        off = nd - self.ndim
        for i, (s, t) in enumerate(zip(self._shape, self._strides)):
            if s == shape[off + i]:
                strides[off + i] = t
            elif s != 1:
                raise ValueError('operands could not be broadcast together with shapes %s %s' % (self._shape, shape))
        return self._view(shape, tuple(strides))

    def __getitem__(self, key):
        """Basic indexing (ints, slices, None, Ellipsis) gives views.  Index ARRAYS -- one boolean mask, or integer
        arrays for the leading axes -- copy through the compaction / gather kernels (_core/_compaction.py)."""
        if not isinstance(key, tuple):
            key = (key,)
        if any(isinstance(k, (list, numpy.ndarray, ndarray)) for k in key):
            from cupy_b200._core import _compaction
            return _compaction.getitem_advanced(self, key)
        n_real = sum(1 for k in key if k is not None and k is not Ellipsis)
        if n_real > self.ndim:
            raise IndexError('too many indices for array')
        if key.count(Ellipsis) > 1:
            raise IndexError('an index can only have a single ellipsis')
        if Ellipsis in key:
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - n_real) + key[i + 1:]
        else:
            key = key + (slice(None),) * (self.ndim - n_real)
        shape, strides, ptr, dim = [], [], self.ptr, 0
        for k in key:
            if k is None:
                shape.append(1)
                strides.append(0)
                continue
            s, t = self._shape[dim], self._strides[dim]
            if isinstance(k, slice):
                start, stop, step = k.indices(s)
                n = len(range(start, stop, step))
                ptr += start * t if n > 0 else 0
                shape.append(n)
                strides.append(t * step)
            else:
                k = int(k)
                if not -s <= k < s:
                    raise IndexError('index %d is out of bounds for axis %d with size %d' % (k, dim, s))
                ptr += (k + s if k < 0 else k) * t
            dim += 1
        return self._view(tuple(shape), tuple(strides), ptr=ptr)

    def __setitem__(self, key, value):
        from cupy_b200._core import _kernel
        if any(isinstance(k, (list, numpy.ndarray, ndarray)) for k in (key if isinstance(key, tuple) else (key,))):
            from cupy_b200._core import _compaction
            return _compaction.setitem_advanced(self, key, value)
        dst = self[key]
        _kernel.elementwise_copy(value, dst)

    # ---- data movement ------------------------------------------------------------
    def copy(self, order='C'):
        from cupy_b200._core import _kernel
        if order in ('K', 'A'):
            order = 'F' if (self._f_contiguous and not self._c_contiguous) else 'C'
        out = ndarray(self._shape, self.dtype, order=order)
        if self.size:
            _kernel.elementwise_copy(self, out)
        return out

    def astype(self, dtype, order='K', casting=None, subok=None, copy=True):
        from cupy_b200._core import _kernel
        dtype = _scalar.get_dtype(dtype)
        if order in ('K', 'A'):
            order = 'F' if (self._f_contiguous and not self._c_contiguous) else 'C'
        if not copy and dtype == self.dtype and (
                (order == 'C' and self._c_contiguous) or (order == 'F' and self._f_contiguous)):
            return self
        out = ndarray(self._shape, dtype, order=order)
        if self.size:
            _kernel.elementwise_copy(self, out)
        return out

    def fill(self, value):
        from cupy_b200._core import _kernel
        _kernel.elementwise_copy(value, self)

    def _bytes_view(self):
        """uint8 torch tensor over this (contiguous) array's bytes."""
        nbytes = self.nbytes
        if self._mem is not None and isinstance(self._mem, torch.Tensor):
            off = self.ptr - self._mem.data_ptr()
            return self._mem[off:off + nbytes]
        if isinstance(self._mem, _Keep):
            # foreign device memory (imported through __cuda_array_interface__): hand torch a byte view of it
            return torch.as_tensor(_ByteSpan(self.ptr, nbytes, self._mem), device='cuda')
        raise ValueError('array does not own torch-visible memory')

    def get(self, stream=None, order='C', out=None, blocking=True):
        """Device -> host copy (cupy/_core/core.pyx `ndarray.get`): enqueued on `stream`
        (default: the current stream).  With `out=` a page-locked C-contiguous array and
        `blocking=False` the copy is asynchronous; otherwise the call returns when the data
        is on the host."""
        a = self
        if not ((order == 'C' and a._c_contiguous) or (order == 'F' and a._f_contiguous)):
            a = a.copy(order=order if order in ('C', 'F') else 'C')
        if a.size == 0:
            return numpy.empty(a._shape, a.dtype, order=order if order in ('C', 'F') else 'C') if out is None else out
        ctx = _stream_ctx(stream)
        with ctx:
            if out is not None:
                if not isinstance(out, numpy.ndarray):
                    raise TypeError('Only numpy.ndarray can be obtained from cupy_b200.ndarray')
                if out.dtype != a.dtype:
                    raise TypeError('{} array cannot be obtained from {} array'.format(out.dtype, a.dtype))
                if out.shape != a._shape:
                    raise ValueError('Shape mismatch. Expected shape: {}, actual shape: {}'.format(a._shape, out.shape))
                ok = out.flags.c_contiguous if order == 'C' else out.flags.f_contiguous
                if not ok:
                    raise RuntimeError('`out` cannot be specified when copying to non-contiguous ndarray')
                flat = out.reshape(-1, order='C' if order == 'C' else 'F') if out.ndim > 1 else out
                dst = torch.from_numpy(flat.view(numpy.uint8) if flat.ndim == 1 else flat.reshape(-1).view(numpy.uint8))
                dst.copy_(a._bytes_view(), non_blocking=not blocking)
                if blocking:
                    torch.cuda.current_stream().synchronize()
                return out
            host = a._bytes_view().cpu().numpy().view(a.dtype)
        if order == 'F' and a.ndim > 1:
            host = host.reshape(a._shape[::-1]).T
        else:
            host = host.reshape(a._shape)
        return host

    def set(self, arr, stream=None):
        """Host -> device copy (cupy/_core/core.pyx `ndarray.set`), enqueued on `stream`
        (default: the current stream); asynchronous when `arr` is page-locked and laid out
        like this array."""
        if not isinstance(arr, numpy.ndarray):
            raise TypeError('Only numpy.ndarray can be set to cupy_b200.ndarray')
        if arr.dtype != self.dtype:
            raise TypeError('{} array cannot be set to {} array'.format(arr.dtype, self.dtype))
        if arr.shape != self._shape:
            raise ValueError('Shape mismatch. Old shape: %s, new shape: %s' % (self._shape, arr.shape))
        with _stream_ctx(stream):
            if self._c_contiguous and arr.flags.c_contiguous and self.size:
                src = torch.from_numpy(arr.reshape(-1).view(numpy.uint8))
                self._bytes_view().copy_(src, non_blocking=True)
                return
            if self.size == 0:
                return
            tmp = asarray(arr)
            from cupy_b200._core import _kernel
            _kernel.elementwise_copy(tmp, self)

    def item(self):
        return self.get().item()

    def tolist(self):
        return self.get().tolist()

    def to_torch(self):
        """Zero-copy torch.Tensor over the same memory (element-aligned strides only)."""
        isz = self.dtype.itemsize
        if any(t % isz for t in self._strides):
            raise ValueError('strides are not multiples of the itemsize')
        tdt = _TORCH_DTYPES[self.dtype]
        off = self.ptr - self._mem.data_ptr()
        if off % isz:
            raise ValueError('data pointer is not element-aligned inside its allocation')
        typed = self._mem.view(tdt) if self._mem.numel() % isz == 0 else self._mem[:self._mem.numel() // isz * isz].view(tdt)
        if any(t < 0 for t in self._strides):
            raise ValueError('negative strides cannot be expressed as a torch tensor')
        return torch.as_strided(typed, self._shape, tuple(t // isz for t in self._strides), off // isz)

    # ---- python protocol ------------------------------------------------------------
    def __repr__(self):
        return 'array(' + repr(self.get())[6:] if self.size < 1000 else '<cupy_b200.ndarray shape=%s dtype=%s>' % (self._shape, self.dtype)

    def __float__(self):
        return float(self.get())

    def __int__(self):
        return int(self.get())

    def __bool__(self):
        if self.size != 1:
            raise ValueError('The truth value of an array with more than one element is ambiguous.')
        return bool(self.get())

    def __array__(self, dtype=None, copy=None):
        raise TypeError('Implicit conversion to a NumPy array is not allowed. '
                        'Please use `.get()` to construct a NumPy array explicitly.')

    # ---- NumPy dispatch protocols (cupy/_core/core.pyx:1969-2037) ------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """`numpy.multiply(x, 2)` on a device array runs the engine's ufunc of the same name.  Host
        numpy.ndarray operands are not converted silently: the ufunc raises TypeError, as in the
        reference."""
        import cupy_b200
        inout = inputs
        if 'out' in kwargs:
            out = kwargs['out']
            if not isinstance(out, tuple):
                out = (out,)
            if len(out) != 1:
                raise ValueError('The \'out\' parameter must have exactly one array value')
            inout += out
            kwargs['out'] = out[0]
        if method not in ('__call__', 'outer', 'at', 'reduce', 'accumulate', 'reduceat'):
            return NotImplemented
        name = ufunc.__name__
        func = getattr(cupy_b200, name, None)
        from cupy_b200._core._kernel import ufunc as _ufunc_class
        if not isinstance(func, _ufunc_class):
            return NotImplemented
        if method != '__call__':
            func = getattr(func, method)
        for x in inout:
            if not (isinstance(x, (ndarray, numpy.ndarray, numpy.generic, int, float, bool, complex))
                    or hasattr(x, '__cuda_array_interface__')):
                return NotImplemented
        if name in ('greater', 'greater_equal', 'less', 'less_equal', 'equal', 'not_equal'):
            # workaround for numpy/numpy#12142 (0-d host arrays in comparisons)
            inputs = tuple(x.item() if isinstance(x, numpy.ndarray) and x.ndim == 0 else x for x in inputs)
        return func(*inputs, **kwargs)

    def __array_function__(self, func, types, args, kwargs):
        """`numpy.sum(x, axis=0)`, `numpy.cumsum(x)` ... -> the cupy_b200 function of the same name."""
        import cupy_b200
        if (func.__module__ or '').split('.')[0] != 'numpy' or len((func.__module__ or '').split('.')) > 2:
            return NotImplemented
        mine = getattr(cupy_b200, func.__name__, None)
        if mine is None or mine is func or not callable(mine):
            return NotImplemented
        if not all(issubclass(t, (ndarray, numpy.ndarray, numpy.generic, int, float, bool)) for t in types):
            return NotImplemented
        return mine(*args, **kwargs)

    # ---- arithmetic (cupy/_core/core.pyx:1560-1700) ----------------------------------
    def _binop(self, name, other, reflected=False):
        f = _UFUNCS.get(name)
        if f is None:
            from cupy_b200._core import _routines_math as m, _routines_binary as b, _routines_elementwise as e
            f = _UFUNCS[name] = getattr(m, name, None) or getattr(b, name, None) or getattr(e, name)
        to = type(other)
        if (to is ndarray or to is float or to is int or to is bool
                or isinstance(other, (ndarray, int, float, bool, numpy.generic))
                or (isinstance(other, numpy.ndarray) and other.ndim == 0)):
            return f(other, self) if reflected else f(self, other)
        return NotImplemented

    def __add__(self, o): return self._binop('add', o)
    def __radd__(self, o): return self._binop('add', o, True)
    def __sub__(self, o): return self._binop('subtract', o)
    def __rsub__(self, o): return self._binop('subtract', o, True)
    def __mul__(self, o): return self._binop('multiply', o)
    def __rmul__(self, o): return self._binop('multiply', o, True)
    def __truediv__(self, o): return self._binop('true_divide', o)
    def __rtruediv__(self, o): return self._binop('true_divide', o, True)

    def _ibinop(self, name, other):
        from cupy_b200._core import _routines_math as m, _routines_binary as b, _routines_elementwise as e
        (getattr(m, name, None) or getattr(b, name, None) or getattr(e, name))(self, other, out=self)
        return self

    def __iadd__(self, o): return self._ibinop('add', o)
    def __isub__(self, o): return self._ibinop('subtract', o)
    def __imul__(self, o): return self._ibinop('multiply', o)
    def __itruediv__(self, o): return self._ibinop('true_divide', o)
    def __floordiv__(self, o): return self._binop('floor_divide', o)
    def __rfloordiv__(self, o): return self._binop('floor_divide', o, True)
    def __ifloordiv__(self, o): return self._ibinop('floor_divide', o)
    def __mod__(self, o): return self._binop('remainder', o)
    def __rmod__(self, o): return self._binop('remainder', o, True)
    def __imod__(self, o): return self._ibinop('remainder', o)
    def __divmod__(self, o): return self._binop('divmod', o)
    def __rdivmod__(self, o): return self._binop('divmod', o, True)

    def __and__(self, o): return self._binop('bitwise_and', o)
    def __rand__(self, o): return self._binop('bitwise_and', o, True)
    def __or__(self, o): return self._binop('bitwise_or', o)
    def __ror__(self, o): return self._binop('bitwise_or', o, True)
    def __xor__(self, o): return self._binop('bitwise_xor', o)
    def __rxor__(self, o): return self._binop('bitwise_xor', o, True)
    def __lshift__(self, o): return self._binop('left_shift', o)
    def __rlshift__(self, o): return self._binop('left_shift', o, True)
    def __rshift__(self, o): return self._binop('right_shift', o)
    def __rrshift__(self, o): return self._binop('right_shift', o, True)
    def __iand__(self, o): return self._ibinop('bitwise_and', o)
    def __ior__(self, o): return self._ibinop('bitwise_or', o)
    def __ixor__(self, o): return self._ibinop('bitwise_xor', o)

    def __invert__(self):
        from cupy_b200._core import _routines_binary as b
        return b.invert(self)

    def __lt__(self, o): return self._binop('less', o)
    def __le__(self, o): return self._binop('less_equal', o)
    def __gt__(self, o): return self._binop('greater', o)
    def __ge__(self, o): return self._binop('greater_equal', o)
    def __eq__(self, o): return self._binop('equal', o)
    def __ne__(self, o): return self._binop('not_equal', o)
    __hash__ = None
    def __pow__(self, o): return self._binop('power', o)

    def __neg__(self):
        from cupy_b200._core import _routines_math as m
        return m.negative(self)

    def __pos__(self):
        return self.copy()

    def __abs__(self):
        from cupy_b200._core import _routines_math as m
        return m.absolute(self)

    # ---- reductions / scans (cupy/_core/core.pyx:1288-1465) ---------------------------
    def sum(self, axis=None, dtype=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_math as m
        return m._ndarray_sum(self, axis, dtype, out, keepdims)

    def prod(self, axis=None, dtype=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_math as m
        return m._ndarray_prod(self, axis, dtype, out, keepdims)

    def cumsum(self, axis=None, dtype=None, out=None):
        from cupy_b200._core import _routines_math as m
        return m.cumsum(self, axis, dtype, out)

    def cumprod(self, axis=None, dtype=None, out=None):
        from cupy_b200._core import _routines_math as m
        return m.cumprod(self, axis, dtype, out)

    def max(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_max(self, axis, out, None, keepdims)

    def min(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_min(self, axis, out, None, keepdims)

    def argmax(self, axis=None, out=None, dtype=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_argmax(self, axis, out, dtype, keepdims)

    def argmin(self, axis=None, out=None, dtype=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_argmin(self, axis, out, dtype, keepdims)

    def mean(self, axis=None, dtype=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_mean(self, axis, dtype, out, keepdims)

    def var(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_var(self, axis, dtype, out, ddof, keepdims)

    def std(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
        from cupy_b200._core import _routines_statistics as s
        return s._ndarray_std(self, axis, dtype, out, ddof, keepdims)

    def all(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_more as r
        return r.all(self, axis, out, keepdims)

    def any(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_more as r
        return r.any(self, axis, out, keepdims)

    def nonzero(self):
        from cupy_b200._core import _compaction
        return _compaction.nonzero(self)

    def take(self, indices, axis=None, out=None):
        from cupy_b200._core import _compaction
        return _compaction.take(self, indices, axis=axis, out=out)

    def compress(self, condition, axis=None, out=None):
        from cupy_b200._core import _compaction
        return _compaction.compress(condition, self, axis=axis, out=out)

    def clip(self, min=None, max=None, out=None):
        from cupy_b200._core import _routines_elementwise as e
        return e.clip(self, min, max, out=out)

    def round(self, decimals=0, out=None):
        from cupy_b200._core import _routines_elementwise as e
        return e.around(self, decimals, out=out)

    def ptp(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_more as r
        return r.ptp(self, axis, out, keepdims)


# ---- creation ------------------------------------------------------------------------
def empty(shape, dtype=float, order='C'):
    return ndarray(shape, dtype, order=order)


def empty_like(a, dtype=None, order='K', shape=None):
    dtype = a.dtype if dtype is None else dtype
    if order in ('K', 'A'):
        order = 'F' if (a._f_contiguous and not a._c_contiguous) else 'C'
    return ndarray(a.shape if shape is None else shape, dtype, order=order)


def _filled(shape, dtype, value, order='C'):
    a = ndarray(shape, dtype, order=order)
    if a.size:
        a.fill(value)
    return a


def zeros(shape, dtype=float, order='C'):
    a = ndarray(shape, dtype, order=order)
    if a.size and not _dryrun.enabled:
        a._bytes_view().zero_()
    return a


def ones(shape, dtype=float, order='C'):
    return _filled(shape, dtype, 1, order)


def full(shape, fill_value, dtype=None, order='C'):
    if dtype is None:
        dtype = numpy.asarray(fill_value).dtype
    return _filled(shape, dtype, fill_value, order)


def zeros_like(a, dtype=None):
    return zeros(a.shape, a.dtype if dtype is None else dtype)


def ones_like(a, dtype=None):
    return ones(a.shape, a.dtype if dtype is None else dtype)


def asarray(a, dtype=None, order=None):
    """Host (NumPy / scalar / list) or device (ndarray, torch.Tensor, CAI) -> ndarray."""
    if isinstance(a, ndarray):
        if dtype is None or _scalar.get_dtype(dtype) == a.dtype:
            return a
        return a.astype(dtype)
    if isinstance(a, torch.Tensor):
        return from_torch(a) if dtype is None else from_torch(a).astype(dtype)
    if hasattr(a, '__cuda_array_interface__'):
        r = from_cuda_array_interface(a)
        return r if dtype is None or _scalar.get_dtype(dtype) == r.dtype else r.astype(dtype)
    h = numpy.asarray(a, dtype=dtype)
    # order 'K' (default): keep the memory order of the host array -- upload the bytes of the
    # axis permutation that is C-contiguous and view them with the permuted strides
    if order in ('C', 'c'):
        perm = tuple(range(h.ndim))
    elif order in ('F', 'f'):
        perm = tuple(reversed(range(h.ndim)))
    else:
        perm = tuple(sorted(range(h.ndim), key=lambda i: (-abs(h.strides[i]), i)))
    hp = numpy.asarray(h.transpose(perm), order='C')     # (ascontiguousarray would turn 0-d into 1-d)
    dev = ndarray(hp.shape, hp.dtype)
    if dev.size and not _dryrun.enabled:
        src = torch.from_numpy(hp.reshape(-1).view(numpy.uint8))
        dev._bytes_view().copy_(src, non_blocking=src.is_pinned())
    inv = [0] * h.ndim
    for k, ax in enumerate(perm):
        inv[ax] = k
    return dev.transpose(inv) if h.ndim > 1 else dev


def array(a, dtype=None, copy=True, order='K'):
    r = asarray(a, dtype, order)
    if copy and r is a:
        r = r.copy()
    return r


def asnumpy(a, stream=None, order='C', out=None):
    if isinstance(a, ndarray):
        return a.get(order=order, out=out)
    return numpy.asarray(a, order=order)


def from_torch(t):
    """Zero-copy view of a CUDA torch.Tensor (the tensor's storage is kept alive)."""
    if not t.is_cuda:
        raise ValueError('from_torch needs a CUDA tensor (cupy_b200 has no CPU arrays)')
    dt = _NUMPY_DTYPES.get(t.dtype)
    if dt is None:
        raise TypeError('Unsupported torch dtype %s' % t.dtype)
    isz = dt.itemsize
    st = t.untyped_storage()
    mem = torch.empty(0, dtype=torch.uint8, device=t.device).set_(st, 0, (st.nbytes(),), (1,))
    return ndarray(tuple(t.shape), dt, memptr=t.data_ptr(),
                   strides=tuple(s * isz for s in t.stride()), _mem=mem)


def from_cuda_array_interface(obj):
    d = obj.__cuda_array_interface__
    dt = numpy.dtype(d['typestr'])
    return ndarray(tuple(d['shape']), dt, memptr=d['data'][0], strides=d.get('strides'), _mem=_Keep(obj))


class _Keep:
    def __init__(self, obj):
        self.obj = obj


class _ByteSpan:
    """`nbytes` of foreign device memory as a CUDA-array-interface object (keeps the owner alive)."""

    def __init__(self, ptr, nbytes, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


def arange(start, stop=None, step=1, dtype=None):
    from cupy_b200._core import _kernel
    if stop is None:
        start, stop = 0, start
    if dtype is None:
        dtype = numpy.result_type(*[numpy.asarray(v).dtype for v in (start, stop, step)])
        if dtype.kind in 'iu':
            dtype = numpy.dtype('int64')
    dtype = _scalar.get_dtype(dtype)
    n = int(numpy.ceil((stop - start) / step))
    out = ndarray((max(n, 0),), dtype)
    if out.size:
        _kernel._arange_kernel()(dtype.type(start), dtype.type(step), out)
    return out


def empty_pinned(shape, dtype=float):
    """Page-locked host array (numpy) for asynchronous H2D / D2H in benchmarks."""
    dtype = _scalar.get_dtype(dtype)
    n = _prod(shape if not isinstance(shape, int) else (shape,))
    t = torch.empty(max(n * dtype.itemsize, 1), dtype=torch.uint8, pin_memory=True)
    return t.numpy()[:n * dtype.itemsize].view(dtype).reshape(shape)
