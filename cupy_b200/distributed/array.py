"""`DistributedArray` -- an N-d array cut into chunks that live on several GPUs, with the index-map and mode
semantics of the reference's `cupyx.distributed.array` (cupyx/distributed/array/_array.py:65-133, 223-240,
310-340; _modes.py:45-68; _reduction.py:15-92; _elementwise.py:77-289; _chunk.py; _index_arith.py), re-cast
for this build's process model: ONE PROCESS PER GPU.

  * `index_map`: {rank: [index, ...]} -- which slices of the global array each rank owns (a rank may hold
    several chunks; chunks may overlap).  Every rank constructs the array collectively with the SAME
    index_map and keeps only its own chunks; the index arithmetic (normalisation, intersection of strided
    slices through the Chinese remainder theorem, sub-indexing) is the reference's.
  * modes: `REPLICA` (overlapping chunks hold identical copies) and the op modes `SUM / PROD / MAX / MIN`
    (the true value of an element is the op-reduction over all chunks that cover it).  `change_mode` moves
    data only where chunks intersect: a forward pass folds every chunk into the later ones that overlap it,
    a backward pass copies the folded values back (`_all_reduce_intersections`, _chunk.py:203-231).
  * reductions along an axis run the engine's kernel on every local chunk and return an array in the matching
    op mode WITHOUT any exchange; the partial results meet only when a REPLICA view is asked for
    (`change_mode(REPLICA)`, `get()`, or an elementwise operation) -- the reference's lazy SUM mode.
  * elementwise kernels (`cupy_b200.ufunc`, `ElementwiseKernel`) run chunk by chunk through the
    `__cupy_override_elementwise_kernel__` hook; operands with different index maps are resharded first.

Differences from the reference, which is single-process with one thread driving all devices: chunk transfers
are `NCCLBackend.send / recv` between the owning ranks (all ranks walk the chunk pairs in one global order, so
the blocking pairs cannot deadlock); transfers are applied eagerly (the reference buffers them as "partial
updates" and flushes lazily -- an optimisation of its single-threaded scheduling, not part of the semantics);
`all_chunks()` returns this rank's chunks; `matmul` is out of scope.
"""
from __future__ import annotations

import numpy

# -------------------------------------------------------------------------------------------------
# index arithmetic (cupyx/distributed/array/_index_arith.py)
# -------------------------------------------------------------------------------------------------


def _extgcd(a, b):
    """(g, x) with g = gcd(a, b) and a*x == g (mod b)."""
    c, d = a, b
    x, u = 1, 0
    while d:
        r = c // d
        c, d = d, c - d * r
        x, u = u, x - u * r
    return c, x


def _crt(a1, n1, a2, n2):
    """Smallest x >= max(a1, a2) with x == a1 (mod n1), x == a2 (mod n2), and lcm(n1, n2); None if none."""
    g, m1 = _extgcd(n1, n2)
    if (a2 - a1) % g != 0:
        return None
    n = n1 * (n2 // g)
    x = a1 + (a2 - a1) // g * m1 % (n // n1) * n1
    if x < a2:
        x += ((a2 - x - 1) // n + 1) * n
    return x, n


def _slice_intersection(a, b, length):
    a_start, a_stop, a_step = a.indices(length)
    b_start, b_stop, b_step = b.indices(length)
    r = _crt(a_start, a_step, b_start, b_step)
    if r is None:
        return None
    c_start, c_step = r
    c_stop = min(a_stop, b_stop)
    if c_start >= c_stop:
        return None
    return slice(c_start, c_stop, c_step)


def _index_for_subslice(a, sub, length):
    """slice c with array[a][c] == array[sub] (sub contained in a)."""
    a_start, _, a_step = a.indices(length)
    sub_start, sub_stop, sub_step = sub.indices(length)
    return slice((sub_start - a_start) // a_step, (sub_stop - a_start - 1) // a_step + 1, sub_step // a_step)


def _index_intersection(a_idx, b_idx, shape):
    res = tuple(_slice_intersection(a, b, n) for a, b, n in zip(a_idx, b_idx, shape))
    return None if None in res else res


def _index_for_subindex(a_idx, sub_idx, shape):
    return tuple(_index_for_subslice(a, s, n) for a, s, n in zip(a_idx, sub_idx, shape))


def _shape_after_indexing(outer_shape, idx):
    shape = list(outer_shape)
    for i in range(len(idx)):
        start, stop, step = idx[i].indices(shape[i])
        shape[i] = (stop - start - 1) // step + 1
    return tuple(shape)


def _normalize_index(shape, idx):
    if not isinstance(idx, tuple):
        idx = (idx,)
    ndim = len(shape)
    if len(idx) > ndim:
        raise IndexError('too many indices for array: array is %d-dimensional, but %d were indexed' % (ndim, len(idx)))
    idx = idx + (slice(None),) * (ndim - len(idx))
    new_idx = []
    for i in range(ndim):
        if isinstance(idx[i], (int, numpy.integer)):
            if idx[i] >= shape[i]:
                raise IndexError('Index %d is out of bounds for axis %d with size %d' % (idx[i], i, shape[i]))
            new_idx.append(slice(int(idx[i]), int(idx[i]) + 1, 1))
        elif isinstance(idx[i], slice):
            start, stop, step = idx[i].indices(shape[i])
            if step <= 0:
                raise ValueError('Slice step must be positive.')
            if start == stop:
                raise ValueError('The index is empty on axis %d' % i)
            new_idx.append(slice(start, stop, step))
        else:
            raise ValueError('Invalid index on axis %d' % i)
    return tuple(new_idx)


def _slice_key(idx):
    return tuple((s.start, s.stop, s.step) for s in idx)


def _normalize_index_map(shape, index_map):
    new = {}
    for dev, idxs in index_map.items():
        if not isinstance(idxs, list):
            idxs = [idxs]
        idxs = [_normalize_index(shape, idx) for idx in idxs]
        idxs.sort(key=_slice_key)
        new[int(dev)] = idxs
    return new


def _same_index_map(a, b):
    return (list(a.keys()) == list(b.keys())
            and all([_slice_key(i) for i in a[k]] == [_slice_key(i) for i in b[k]] for k in a))


def make_2d_index_map(i_partitions, j_partitions, devices):
    """`index_map` of a 2-D matrix cut at the given row / column boundaries, `devices[i][j]` = the set of ranks
    owning block (i, j) (cupyx/distributed/array/_linalg.py:346-395)."""
    assert i_partitions[0] == 0 and sorted(set(i_partitions)) == list(i_partitions)
    assert j_partitions[0] == 0 and sorted(set(j_partitions)) == list(j_partitions)
    index_map = {}
    assert len(devices) == len(i_partitions) - 1
    for i in range(len(devices)):
        assert len(devices[i]) == len(j_partitions) - 1
        for j in range(len(devices[i])):
            idx = (slice(i_partitions[i], i_partitions[i + 1]), slice(j_partitions[j], j_partitions[j + 1]))
            for dev in sorted(devices[i][j]):
                index_map.setdefault(dev, []).append(idx)
    return index_map


# -------------------------------------------------------------------------------------------------
# modes (cupyx/distributed/array/_modes.py)
# -------------------------------------------------------------------------------------------------
def _min_value_of(dtype):
    dtype = numpy.dtype(dtype)
    if dtype.kind == 'b':
        return dtype.type(False)
    if dtype.kind in 'iu':
        return dtype.type(numpy.iinfo(dtype).min)
    if dtype.kind == 'f':
        return dtype.type(-numpy.inf)
    raise RuntimeError('Unsupported type: %s' % dtype)


def _max_value_of(dtype):
    dtype = numpy.dtype(dtype)
    if dtype.kind == 'b':
        return dtype.type(True)
    if dtype.kind in 'iu':
        return dtype.type(numpy.iinfo(dtype).max)
    if dtype.kind == 'f':
        return dtype.type(numpy.inf)
    raise RuntimeError('Unsupported type: %s' % dtype)


class _OpMode:
    """An op mode: the binary function that combines overlapping chunks, whether it is idempotent, and its
    identity for a dtype."""

    def __init__(self, name, func_name, idempotent, identity_of):
        self.name = name
        self.func_name = func_name
        self.numpy_func = getattr(numpy, func_name)
        self.idempotent = idempotent
        self.identity_of = identity_of

    def __repr__(self):
        return repr(self.name)


REPLICA = None
MIN = _OpMode('min', 'minimum', True, _max_value_of)
MAX = _OpMode('max', 'maximum', True, _min_value_of)
SUM = _OpMode('sum', 'add', False, lambda dt: numpy.dtype(dt).type(0))
PROD = _OpMode('prod', 'multiply', False, lambda dt: numpy.dtype(dt).type(1))


# -------------------------------------------------------------------------------------------------
# array backend: what a chunk is made of.  The product backend is the engine itself (cupy_b200 arrays, its
# ufuncs and kernels); the CPU test tier injects a NumPy look-alike to check the host logic without a GPU.
# -------------------------------------------------------------------------------------------------
class _EngineBackend:
    name = 'cupy_b200'

    def __init__(self):
        import cupy_b200
        self.cp = cupy_b200
        self.ndarray = cupy_b200.ndarray

    def empty(self, shape, dtype):
        return self.cp.empty(shape, dtype)

    def full(self, shape, value, dtype):
        return self.cp.full(shape, value, dtype)

    def from_host(self, a):
        return self.cp.asarray(numpy.asarray(a, order='C'))

    def to_host(self, a):
        return a.get()

    def contiguous(self, a):
        return a if a.flags.c_contiguous else a.copy()

    def copy(self, a):
        return a.copy()

    def assign(self, a, idx, value):
        a[idx] = value

    def combine(self, func_name, a, idx, value):
        """a[idx] = func(a[idx], value)"""
        view = a[idx]
        getattr(self.cp, func_name)(view, value, out=view)

    def wire(self, a):
        """the object the communicator sends / receives (contiguous)"""
        return a

    def run_elementwise(self, kernel, arrays, kwargs):
        return kernel(*arrays, **kwargs)

    def run_reduction(self, kernel, array, axis, dtype):
        return kernel(array, axis=axis, dtype=dtype)


_backend = None


def _get_backend():
    global _backend
    if _backend is None:
        _backend = _EngineBackend()
    return _backend


def _set_backend(backend):
    """Test seam (tests/ only): inject the NumPy look-alike backend of the CPU tier; None restores the engine."""
    global _backend
    _backend = backend


class _SoloComm:
    """World of one rank (no process group): every transfer is local."""
    rank, _n_devices = 0, 1

    def send(self, array, peer, stream=None):
        raise RuntimeError('no peer in a single-rank world')

    recv = broadcast = send


# -------------------------------------------------------------------------------------------------
class _Chunk:
    __slots__ = ('array', 'index')

    def __init__(self, array, index):
        self.array = array
        self.index = index


def _atleast_1d(xp, a):
    return a.reshape((1,)) if a.ndim == 0 else a


class DistributedArray:
    """N-d array distributed over the ranks of a communicator (see the module docstring).  Create it with
    `distributed_array`; every method that moves data is COLLECTIVE: all ranks call it in the same order."""

    def __init__(self, shape, dtype, index_map, chunks, mode=REPLICA, comm=None):
        self._shape = tuple(shape)
        self.dtype = numpy.dtype(dtype)
        self._index_map = index_map          # normalised, identical on every rank
        self._chunks = chunks                # this rank's chunks, in index_map[rank] order
        self._mode = mode
        self._comm = comm if comm is not None else _SoloComm()
        self._xp = _get_backend()

    # ---- metadata -------------------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, newshape):
        raise NotImplementedError('DistributedArray currently does not support assignment to shape.')

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def size(self):
        n = 1
        for s in self._shape:
            n *= s
        return n

    @property
    def mode(self):
        """How overlaps of the chunks are interpreted: REPLICA (None) or an op mode (SUM / PROD / MAX / MIN)."""
        return self._mode

    @property
    def devices(self):
        """The ranks holding part of the data."""
        return self._index_map.keys()

    @property
    def index_map(self):
        return {dev: list(idxs) for dev, idxs in self._index_map.items()}

    @property
    def rank(self):
        return self._comm.rank

    def all_chunks(self):
        """{this rank: [its chunk arrays]} (the reference, driving every device from one process, returns all)."""
        return {self.rank: [c.array for c in self._chunks]}

    def _pairs(self):
        """Every chunk of the array as (owner rank, position in the owner's list, index), in ONE global order."""
        return [(dev, i, idx) for dev, idxs in self._index_map.items() for i, idx in enumerate(idxs)]

    def _like(self, chunks, mode=None, shape=None, dtype=None, index_map=None):
        return DistributedArray(self._shape if shape is None else shape, self.dtype if dtype is None else dtype,
                                self._index_map if index_map is None else index_map, chunks,
                                self._mode if mode is None else mode, self._comm)

    # ---- moving data between chunks ---------------------------------------------------------------
    def _move(self, src_rank, src_view, dst_rank, shape):
        """The contents of `src_view` (on src_rank) as a fresh contiguous array on dst_rank (None elsewhere)."""
        xp, me = self._xp, self.rank
        if src_rank == dst_rank:
            return xp.copy(src_view) if me == src_rank else None
        if me == src_rank:
            self._comm.send(xp.wire(xp.contiguous(src_view)), dst_rank)
            return None
        if me == dst_rank:
            buf = xp.empty(shape, self.dtype)
            self._comm.recv(xp.wire(buf), src_rank)
            return buf
        return None

    def _apply(self, chunks, src, dst, mode):
        """Fold chunk `src` into chunk `dst` on their overlap (`_Chunk.apply_to`, _chunk.py:132-181): with an
        op mode dst = func(dst, src) there (and, unless the op is idempotent, src is reset to the identity so
        the value is not counted twice); in REPLICA mode dst is overwritten."""
        xp, me = self._xp, self.rank
        (src_rank, src_i, src_idx), (dst_rank, dst_i, dst_idx) = src, dst
        inter = _index_intersection(src_idx, dst_idx, self._shape)
        if inter is None:
            return
        src_new = _index_for_subindex(src_idx, inter, self._shape)
        dst_new = _index_for_subindex(dst_idx, inter, self._shape)
        shape = _shape_after_indexing(self._shape, inter)
        src_view = chunks[src_i].array[src_new] if me == src_rank else None
        data = self._move(src_rank, src_view, dst_rank, shape)
        if me == dst_rank:
            if mode is REPLICA:
                xp.assign(chunks[dst_i].array, dst_new, data)
            else:
                xp.combine(mode.func_name, chunks[dst_i].array, dst_new, data)
        if me == src_rank and mode is not REPLICA and not mode.idempotent:
            xp.assign(chunks[src_i].array, src_new, mode.identity_of(self.dtype))

    def _copy_chunks(self):
        return [_Chunk(self._xp.copy(c.array), c.index) for c in self._chunks]

    def _chunks_in_replica_mode(self):
        chunks = self._copy_chunks()
        if self._mode is not REPLICA:
            pairs = self._pairs()
            for i in range(len(pairs)):                      # fold forward: the last chunk covering an element
                for j in range(i + 1, len(pairs)):           # ends up with its full value ...
                    self._apply(chunks, pairs[i], pairs[j], self._mode)
            for j in range(len(pairs) - 1, -1, -1):          # ... which is then copied back to the earlier ones
                for i in range(j):
                    self._apply(chunks, pairs[j], pairs[i], REPLICA)
        return chunks

    def _chunks_in_op_mode(self, op_mode):
        chunks = self._chunks_in_replica_mode()
        pairs = self._pairs()
        identity = op_mode.identity_of(self.dtype)
        me = self.rank
        for i in range(len(pairs)):                          # keep every element in exactly one chunk (the last)
            a_rank, a_i, a_idx = pairs[i]
            if a_rank != me:
                continue
            for j in range(i + 1, len(pairs)):
                inter = _index_intersection(a_idx, pairs[j][2], self._shape)
                if inter is not None:
                    self._xp.assign(chunks[a_i].array, _index_for_subindex(a_idx, inter, self._shape), identity)
        return chunks

    def _to_op_mode(self, op_mode):
        if self._mode is op_mode:
            return self
        if len(self._pairs()) == 1:
            return self._like(self._chunks, mode=op_mode)
        chunks = self._chunks_in_replica_mode() if op_mode is REPLICA else self._chunks_in_op_mode(op_mode)
        return DistributedArray(self._shape, self.dtype, self._index_map, chunks, op_mode, self._comm)

    def change_mode(self, mode):
        """A view or a copy of the array in the given mode (collective when chunks overlap)."""
        return self._to_op_mode(mode)

    def reshard(self, index_map):
        """A view or a copy of the array with the given index_map (collective)."""
        new_map = _normalize_index_map(self._shape, index_map)
        if _same_index_map(new_map, self._index_map):
            return self
        xp, me = self._xp, self.rank
        src = self._copy_chunks() if self._mode is not REPLICA else list(self._chunks)
        new_chunks = []
        for idx in new_map.get(me, []):
            shape = _shape_after_indexing(self._shape, idx)
            if self._mode is REPLICA:
                arr = xp.empty(shape, self.dtype)
            else:
                arr = xp.full(shape, self._mode.identity_of(self.dtype), self.dtype)
            new_chunks.append(_Chunk(_atleast_1d(xp, arr), idx))
        target = DistributedArray(self._shape, self.dtype, new_map, new_chunks, self._mode, self._comm)
        for s in self._pairs():
            for d in target._pairs():
                self._apply_between(src, s, new_chunks, d)
        return target

    def _apply_between(self, src_chunks, src, dst_chunks, dst):
        """`_apply` with source and destination chunk lists of two different arrays (resharding)."""
        xp, me, mode = self._xp, self.rank, self._mode
        (src_rank, src_i, src_idx), (dst_rank, dst_i, dst_idx) = src, dst
        inter = _index_intersection(src_idx, dst_idx, self._shape)
        if inter is None:
            return
        src_new = _index_for_subindex(src_idx, inter, self._shape)
        dst_new = _index_for_subindex(dst_idx, inter, self._shape)
        shape = _shape_after_indexing(self._shape, inter)
        src_view = src_chunks[src_i].array[src_new] if me == src_rank else None
        data = self._move(src_rank, src_view, dst_rank, shape)
        if me == dst_rank:
            if mode is REPLICA:
                xp.assign(dst_chunks[dst_i].array, dst_new, data)
            else:
                xp.combine(mode.func_name, dst_chunks[dst_i].array, dst_new, data)
        if me == src_rank and mode is not REPLICA and not mode.idempotent:
            xp.assign(src_chunks[src_i].array, src_new, mode.identity_of(self.dtype))

    def get(self, stream=None, order='C', out=None, blocking=True):
        """The whole array as a NumPy array on EVERY rank (collective: each chunk is broadcast by its owner)."""
        if stream is not None:
            raise RuntimeError('Argument `stream` not supported')
        if order != 'C':
            raise RuntimeError('Argument `order` not supported')
        if out is not None:
            raise RuntimeError('Argument `out` not supported')
        xp, me = self._xp, self.rank
        if self._mode is REPLICA:
            res = numpy.empty(self._shape, dtype=self.dtype)
        else:
            res = numpy.full(self._shape, self._mode.identity_of(self.dtype), self.dtype)
        res = numpy.atleast_1d(res)
        world = getattr(self._comm, '_n_devices', 1)
        for rank, i, idx in self._pairs():
            shape = _shape_after_indexing(self._shape, idx) if self._shape else (1,)
            if world > 1:
                buf = xp.contiguous(self._chunks[i].array) if me == rank else xp.empty(shape, self.dtype)
                self._comm.broadcast(xp.wire(buf), root=rank)
            else:
                buf = self._chunks[i].array
            host = xp.to_host(buf).reshape(shape)
            if self._mode is REPLICA:
                res[idx] = host
            else:
                self._mode.numpy_func(res[idx], host, out=res[idx])
        return res.reshape(self._shape)

    # ---- kernel hooks (cupy/_core/_kernel.pyx:1261-1262, _reduction.pyx:609-611) -------------------------
    def __cupy_override_elementwise_kernel__(self, kernel, *args, **kwargs):
        return _execute_elementwise(kernel, args, kwargs)

    def __cupy_override_reduction_kernel__(self, kernel, axis, dtype, out, keepdims):
        if axis is None:
            raise RuntimeError('axis must be specified')
        if out is not None:
            raise RuntimeError('Argument `out` is not supported')
        if keepdims:
            raise RuntimeError('Argument `keepdims` is not supported')
        return _execute_reduction(self, kernel, axis, dtype)

    def _ufunc(self, name, *others, reflected=False):
        import cupy_b200
        f = getattr(cupy_b200, name)
        args = (others + (self,)) if reflected else ((self,) + others)
        return f(*args)

    def __add__(self, o): return self._ufunc('add', o)
    def __sub__(self, o): return self._ufunc('subtract', o)
    def __mul__(self, o): return self._ufunc('multiply', o)
    def __truediv__(self, o): return self._ufunc('true_divide', o)
    def __neg__(self): return self._ufunc('negative')

    def sum(self, axis=None, dtype=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_math
        return _routines_math._ndarray_sum(self, axis, dtype, out, keepdims)

    def prod(self, axis=None, dtype=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_math
        return _routines_math._ndarray_prod(self, axis, dtype, out, keepdims)

    def max(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics
        return _routines_statistics._amax(self, axis=axis, out=out, dtype=None, keepdims=keepdims)

    def min(self, axis=None, out=None, keepdims=False):
        from cupy_b200._core import _routines_statistics
        return _routines_statistics._amin(self, axis=axis, out=out, dtype=None, keepdims=keepdims)

    def __repr__(self):
        return '<DistributedArray shape=%s dtype=%s mode=%r ranks=%s>' % (
            self._shape, self.dtype, self._mode, sorted(self._index_map))

    def _unsupported(name):      # noqa: N805  (the reference overrides these to raise, _array.py:417-817)
        def f(self, *args, **kwargs):
            raise NotImplementedError('DistributedArray currently does not support %s.' % name)
        f.__name__ = name
        return f

    for _n in ('__getitem__', '__setitem__', '__len__', '__iter__', '__copy__', 'all', 'any', 'argmax', 'argmin',
               'astype', 'copy', 'cumprod', 'cumsum', 'dot', 'fill', 'flatten', 'item', 'mean', 'ravel', 'reshape',
               'squeeze', 'std', 'swapaxes', 'tolist', 'transpose', 'var', 'view'):
        locals()[_n] = _unsupported(_n)
    del _n, _unsupported


# -------------------------------------------------------------------------------------------------
def _execute_reduction(arr, kernel, axis, dtype):
    """cupyx/distributed/array/_reduction.py:15-92: the kernel runs on every local chunk; the result stays in the
    op mode of the reduction (no exchange here)."""
    mode_overrides = {'cupy_max': MAX, 'cupy_min': MIN, 'cupy_sum': SUM, 'cupy_prod': PROD}
    if kernel.name not in mode_overrides:
        raise RuntimeError('Unsupported kernel: %s' % kernel.name)
    mode = mode_overrides[kernel.name]
    if not isinstance(axis, (int, numpy.integer)):
        raise RuntimeError('axis must be an integer')
    axis = int(axis)
    if axis < 0:
        axis += arr.ndim
    if mode in (MAX, MIN):
        if arr._mode is not mode:
            arr = arr._to_op_mode(REPLICA)
    else:
        arr = arr._to_op_mode(mode)
    xp = arr._xp
    shape = arr._shape[:axis] + arr._shape[axis + 1:]
    out_dtype = None
    out_chunks = []
    for c in arr._chunks:
        res = xp.run_reduction(kernel, c.array, axis, dtype)
        res = _atleast_1d(xp, res)
        out_dtype = res.dtype
        out_chunks.append(_Chunk(res, c.index[:axis] + c.index[axis + 1:]))
    if out_dtype is None:        # this rank holds no chunk: the dtype follows the kernel's loop table
        probe = xp.run_reduction(kernel, xp.full((1,) * arr.ndim, 0, arr.dtype), axis, dtype)
        out_dtype = probe.dtype
    new_map = {dev: [idx[:axis] + idx[axis + 1:] for idx in idxs] for dev, idxs in arr._index_map.items()}
    return DistributedArray(shape, out_dtype, new_map, out_chunks, mode, arr._comm)


def _execute_elementwise(kernel, args, kwargs):
    """cupyx/distributed/array/_elementwise.py:77-289."""
    for a in list(args) + list(kwargs.values()):
        if not isinstance(a, DistributedArray):
            raise RuntimeError('Mixing a distributed array with a non-distributed one is not supported')
    args = list(args)
    first = args[0] if args else next(iter(kwargs.values()))
    others = [a for a in args[1:] + list(kwargs.values())]
    if any(not _same_index_map(a._index_map, first._index_map) for a in others):
        # The reference lets one device read the peer's chunks directly; with a process per GPU the second
        # operand is resharded onto the first one's index_map instead (same result, same layout of the output).
        if len(args) > 2:
            raise RuntimeError('Element-wise operation over more than two distributed arrays is not supported '
                               'unless they share the same index_map.')
        if kwargs:
            raise RuntimeError('Keyword argument is not supported unless arguments share the same index_map.')
        args = [args[0]] + [a._to_op_mode(REPLICA).reshard(first._index_map) for a in args[1:]]
    args = [a._to_op_mode(REPLICA) for a in args]
    kwargs = {k: a._to_op_mode(REPLICA) for k, a in kwargs.items()}
    first = args[0] if args else next(iter(kwargs.values()))
    xp = first._xp
    out_chunks = []
    out_dtype = None
    for i, c in enumerate(first._chunks):
        res = xp.run_elementwise(kernel, [a._chunks[i].array for a in args],
                                 {k: a._chunks[i].array for k, a in kwargs.items()})
        if not isinstance(res, xp.ndarray):
            raise RuntimeError('Kernels returning other than single array are not supported')
        out_dtype = res.dtype
        out_chunks.append(_Chunk(res, c.index))
    if out_dtype is None:
        probe = xp.run_elementwise(kernel, [xp.full((1,), 1, a.dtype) for a in args],
                                   {k: xp.full((1,), 1, a.dtype) for k, a in kwargs.items()})
        out_dtype = probe.dtype
    return DistributedArray(first._shape, out_dtype, first._index_map, out_chunks, REPLICA, first._comm)


def distributed_array(array, index_map, mode=REPLICA, comm=None):
    """Create a distributed array from data every rank holds (cupyx/distributed/array/_array.py:820-915):
    `array` is a DistributedArray, a cupy_b200.ndarray, or anything numpy.array accepts; `index_map` maps each
    rank to the index (or list of indices) of the chunks it owns.  Collective; `comm` is the NCCLBackend of
    `init_process_group` (omit it in a single-process world).  Does not check that the chunks cover the array."""
    xp = _get_backend()
    if isinstance(array, DistributedArray):
        if array.mode is not mode:
            array = array.change_mode(mode)
        new_map = _normalize_index_map(array.shape, index_map)
        if not _same_index_map(array._index_map, new_map):
            array = array.reshard(index_map)
        return array._like(array._chunks)
    if isinstance(array, xp.ndarray):
        host = xp.to_host(array)
    else:
        host = numpy.array(array)
    if mode is not REPLICA:
        host = host.copy()
    index_map = _normalize_index_map(host.shape, index_map)
    comm = comm if comm is not None else _SoloComm()
    me = comm.rank
    view = numpy.atleast_1d(host)
    chunks = []
    for dev, idxs in index_map.items():
        for idx in idxs:
            if dev == me:
                chunks.append(_Chunk(_atleast_1d(xp, xp.from_host(view[idx])), idx))
            if mode is not REPLICA and not mode.idempotent:
                view[idx] = mode.identity_of(host.dtype)      # later chunks must not count these elements again
    return DistributedArray(host.shape, host.dtype, index_map, chunks, mode, comm)
