"""NVRTC route of the reduction engine: turn (map, reduce, post_map, identity,
reduce_type) strings into a functor for the skeleton of b200/reduce.cuh, choose
the launch geometry, pack the by-value parameter block and launch.

The generated text is the functor plus a one-line `extern "C"` wrapper around
`reduce_full_body` / `reduce_rows_body` / `reduce_cols_body` /
`reduce_generic_body`; the reduction algorithms themselves are the hand-written
header code.  Reference counterpart: `_create_reduction_function_code`
(cupy/_core/_reduction.pyx:44-126) and its launch (`:481-508`).  User-visible
names are kept: parameter names, `in0/out0`, `type_in0_raw/type_out0_raw`, `a`,
`b`, `_J`, `_type_reduce`, `_in_ind.size()`, `_out_ind.size()`, `IndexT`.
"""
from __future__ import annotations

import struct

import numpy

from cupy_b200 import _lib
from cupy_b200._core import _dryrun, _jit, _workspace
from cupy_b200._core._ndarray import ndarray
from cupy_b200._core._scalar import CScalar, get_dtype, get_typename

_THREADS = 256
_TICKET_BYTES = 16384
_MAX_ACC_BYTES = 64
_KNOWN_SIZES = {'bool': 1, 'signed char': 1, 'unsigned char': 1, 'char': 1, 'short': 2, 'unsigned short': 2,
                'int': 4, 'unsigned int': 4, 'long long': 8, 'unsigned long long': 8, 'float16': 2,
                'float': 4, 'double': 8, 'size_t': 8, 'ptrdiff_t': 8, 'long': 8, 'unsigned long': 8}

_PROLOGUE = '''#include <b200/reduce.cuh>
#include <b200/carray.cuh>
using b200::float16;
'''


def _acc_size(reduce_type, type_map):
    t = reduce_type.strip()
    for ctype, dt in type_map:
        if t == ctype:
            return get_dtype(dt).itemsize
    return _KNOWN_SIZES.get(t)


def _functor_source(kernel, in_args, in_params, out_args, out_params, type_map, map_expr, reduce_expr,
                    post_map_expr, reduce_type, index64, structured, kinds=None, layout_kind=None):
    lines = [_PROLOGUE]
    lines.append('typedef %s IndexT;' % ('long long' if index64 else 'int'))
    for ctype, dt in type_map:
        lines.append('typedef %s %s;' % (get_typename(dt), ctype))
    lines.append(kernel.preamble)
    lines.append('typedef %s _type_reduce;' % reduce_type)
    lines.append('static_assert(sizeof(_type_reduce) <= %d, "reduce_type too large");' % _MAX_ACC_BYTES)
    multi = [a for a in in_args if isinstance(a, ndarray)] if structured else []
    if len(multi) > 1:
        # kinds: 0 = laid out like x, 1 = broadcast along the reduced axes (dense over the kept ones),
        # 2 = broadcast along the kept axes (dense over the reduced ones); see _reduction._operand_kinds
        ks = tuple(kinds) if kinds is not None else (0,) * len(multi)
        cols = layout_kind == _lib.RED_COLS
        ts = [get_typename(a.dtype) for a in multi]
        k_all = range(len(multi))

        def step(fmt_by_kind):
            return ' '.join(fmt_by_kind[ks[k]].format(k=k) for k in k_all)

        lines.append('struct __align__(16) _In { %s };' % ' '.join('%s m%d;' % (t, k) for k, t in enumerate(ts)))
        lines.append('struct _InPtr {')
        lines.append('  ' + ' '.join('const %s* p%d;' % (t, k) for k, t in enumerate(ts)))
        # + / []: a step along the contiguous reduced axis (FULL and ROWS skeletons)
        lines.append('  __device__ __forceinline__ _InPtr operator+(long long _o) const { _InPtr _r; %s return _r; }'
                     % step({0: '_r.p{k} = p{k} + _o;', 1: '_r.p{k} = p{k};', 2: '_r.p{k} = p{k} + _o;'}))
        lines.append('  __device__ __forceinline__ _In operator[](long long _i) const { _In _r; %s return _r; }'
                     % step({0: '_r.m{k} = p{k}[_i];', 1: '_r.m{k} = p{k}[0];', 2: '_r.m{k} = p{k}[_i];'}))
        lines.append('};')
        lines.append('__device__ __forceinline__ _InPtr row_ptr(const _InPtr& _x, long long _row, long long _n) { _InPtr _r; %s return _r; }'
                     % step({0: '_r.p{k} = _x.p{k} + _row * _n;', 1: '_r.p{k} = _x.p{k} + _row;', 2: '_r.p{k} = _x.p{k};'}))
        lines.append('__device__ __forceinline__ _InPtr cols_ptr(const _InPtr& _x, long long _b, long long _n, long long _c, long long _c0) '
                     '{ _InPtr _r; %s return _r; }'
                     % step({0: '_r.p{k} = _x.p{k} + _b * _n * _c + _c0;', 1: '_r.p{k} = _x.p{k} + _b * _c + _c0;',
                             2: '_r.p{k} = _x.p{k};'}))
        lines.append('__device__ __forceinline__ _InPtr cols_row(const _InPtr& _x, long long _rw, long long _c) { _InPtr _r; %s return _r; }'
                     % step({0: '_r.p{k} = _x.p{k} + _rw * _c;', 1: '_r.p{k} = _x.p{k};', 2: '_r.p{k} = _x.p{k} + _rw;'}))
        lines.append('template <int _N> __device__ __forceinline__ void load_pack(b200::Pack<_In, _N>& _d, const _InPtr& _p) {')
        scalar_kind = 2 if cols else 1       # one value serves the whole pack
        for k, t in enumerate(ts):
            if ks[k] == scalar_kind:
                lines.append('  { const %s _s = *_p.p%d;' % (t, k))
                lines.append('#pragma unroll')
                lines.append('    for (int _i = 0; _i < _N; ++_i) _d[_i].m%d = _s; }' % k)
            else:
                lines.append('  { b200::Pack<%s, _N> _t; b200::load_pack(_t, _p.p%d);' % (t, k))
                lines.append('#pragma unroll')
                lines.append('    for (int _i = 0; _i < _N; ++_i) _d[_i].m%d = _t[_i]; }' % k)
        lines.append('}')
    lines.append('struct _Op {')
    lines.append('  typedef _type_reduce acc_t; typedef IndexT index_t; struct ctx_t {};')
    lines.append('  static constexpr bool kWideIndex = false;')
    in_bind, out_bind, members, raw_bind = [], [], [], []
    k_arr = 0
    for a, p in zip(in_args, in_params):
        if isinstance(a, ndarray):
            mem_t = get_typename(a.dtype)
            if p.raw:
                # not broadcast, indexed by the user code with _i / _j / _J (cupy/_core/_reduction.pyx:186-225)
                members.append('  b200::RawView _rv_%s;' % p.name)
                raw_bind.append('    const CArray<%s, %d, %s, false> %s(_rv_%s);'
                                % (mem_t, a.ndim, 'true' if a._c_contiguous else 'false', p.name, p.name))
                continue
            ctype = p.ctype if p.ctype else mem_t
            if kernel_is_simple(kernel):
                in_bind.append('    const type_in0_raw in0 = *reinterpret_cast<const type_in0_raw*>(_ptrs[%d]);' % k_arr)
            else:
                in_bind.append('    const %s %s = *reinterpret_cast<const %s*>(_ptrs[%d]);' % (ctype, p.name, mem_t, k_arr))
            k_arr += 1
        else:
            members.append('  alignas(8) %s %s;' % (p.ctype, p.name))
    for k, (a, p) in enumerate(zip(out_args, out_params)):
        if p.raw:
            raise NotImplementedError('raw output arguments of reduction kernels are not supported')
        mem_t = get_typename(a.dtype)
        if kernel_is_simple(kernel):
            out_bind.append('    type_out0_raw& out0 = *reinterpret_cast<type_out0_raw*>(_optrs[%d]);' % k)
        else:
            out_bind.append('    %s& %s = *reinterpret_cast<%s*>(_optrs[%d]);' % (p.ctype, p.name, mem_t, k))
    lines.extend(members)
    lines.append('  long long _in_size, _out_size;')
    ident = kernel.identity
    lines.append('  __device__ acc_t identity() const { return _type_reduce(%s); }' % ident)
    lines.append('  __device__ acc_t combine(const acc_t& a, const acc_t& b) const { return (%s); }' % reduce_expr)
    # _i: output index, _J: index along the reduced axes, _j = _i + _J * out_size: the reference's linear
    # input index when the reduced axes lead (cupy/_core/_reduction.pyx:80-89); generic skeleton only
    lines.append('  __device__ acc_t map_at(const char* const* _ptrs, index_t _J, ptrdiff_t _i = 0) const {')
    lines.append('    const CSizeIndexer _in_ind = {(ptrdiff_t)_in_size}, _out_ind = {(ptrdiff_t)_out_size};')
    if raw_bind:
        lines.append('    const ptrdiff_t _j = _i + (ptrdiff_t)_J * (ptrdiff_t)_out_size; (void)_j;')
    lines.extend(raw_bind)
    lines.extend(in_bind)
    lines.append('    return static_cast<_type_reduce>(%s);' % map_expr)
    lines.append('  }')
    lines.append('  __device__ void post_at(char* const* _optrs, const acc_t& a, ptrdiff_t _i = 0) const {')
    lines.append('    const CSizeIndexer _in_ind = {(ptrdiff_t)_in_size}, _out_ind = {(ptrdiff_t)_out_size};')
    lines.extend(raw_bind)
    lines.extend(out_bind)
    lines.append('    %s;' % post_map_expr)
    lines.append('  }')
    if structured:
        in_arrs = [a for a in in_args if isinstance(a, ndarray)]
        if len(in_arrs) == 1:
            lines.append('  typedef %s in_t; typedef %s out_t;'
                         % (get_typename(in_arrs[0].dtype), get_typename(out_args[0].dtype)))
            lines.append('  __device__ ctx_t step(int) const { return ctx_t(); }')
            lines.append('  __device__ acc_t single(const in_t& _v, index_t _j) const {'
                         ' const char* _q = reinterpret_cast<const char*>(&_v); return map_at(&_q, _j); }')
        else:
            # several arrays of one layout: the skeleton streams a tuple of their elements (_In) through
            # a struct of pointers (_InPtr), see b200::in_ptr in reduce.cuh
            lines.append('  typedef _In in_t; typedef _InPtr ptr_t; typedef %s out_t;' % get_typename(out_args[0].dtype))
            lines.append('  __device__ ctx_t step(int) const { return ctx_t(); }')
            lines.append('  __device__ acc_t single(const in_t& _v, index_t _j) const { const char* _q[%d] = {%s};'
                         ' return map_at(_q, _j); }'
                         % (len(in_arrs), ', '.join('reinterpret_cast<const char*>(&_v.m%d)' % k
                                                    for k in range(len(in_arrs)))))
        lines.append('  __device__ void accumulate(acc_t& _acc, const ctx_t&, const in_t& _v, index_t _j) const {'
                     ' _acc = combine(_acc, single(_v, _j)); }')
        lines.append('  __device__ out_t post(const acc_t& _a, long long) const {'
                     ' out_t _o; char* _q = reinterpret_cast<char*>(&_o); post_at(&_q, _a); return _o; }')
    lines.append('};')
    return '\n'.join(lines)


def kernel_is_simple(kernel):
    return hasattr(kernel, '_ops')


def _pack_raw_view(a):
    """b200::RawView (csrc/include/b200/carray.cuh): data, size, ndim, pad, shape[10], strides[10]."""
    pad = [0] * (_lib.MAX_NDIM - a.ndim)
    return struct.pack('<Qqii%dq%dq' % (_lib.MAX_NDIM, _lib.MAX_NDIM), a.ptr, a.size, a.ndim, 0,
                       *(list(a.shape) + pad), *(list(a.strides) + pad))


def _pack_op(in_args, n_in, n_out, in_params=None):
    b = b''
    for k, a in enumerate(in_args):
        if isinstance(a, ndarray) and in_params is not None and in_params[k].raw:
            b += _pack_raw_view(a)
        if isinstance(a, CScalar):
            raw = numpy.asarray(a.value, dtype=a.descr).tobytes()
            if len(raw) > 8:
                raise NotImplementedError('scalar parameters wider than 8 bytes')
            b += raw + b'\0' * (8 - len(raw))
    b += struct.pack('<qq', int(n_in), int(n_out))
    return b


def _sm_count():
    if _dryrun.enabled:
        return 148
    return _lib.device_info()[0]


def _pick_vec(ptr, inner, itemsize, full):
    """`ptr` / `itemsize` may be lists (several operands of one layout): every one must be aligned."""
    ptrs = ptr if isinstance(ptr, (list, tuple)) else [ptr]
    sizes = itemsize if isinstance(itemsize, (list, tuple)) else [itemsize] * len(ptrs)
    vec = full
    while vec > 1:
        if inner % vec == 0 and all(q % min(vec * sz, 16) == 0 for q, sz in zip(ptrs, sizes)):
            break
        vec >>= 1
    return vec


def launch_structured(kernel, layout, in_args, out, in_types, out_types, type_map,
                      map_expr, reduce_expr, post_map_expr, reduce_type, stream, kinds=None):
    xs = [a for a in in_args if isinstance(a, ndarray)]
    kinds = tuple(kinds) if kinds is not None else (0,) * len(xs)
    x = xs[kinds.index(0)]
    kind = layout.kind
    # operands read with vector loads in this layout (the others are one scalar per row / per column pack)
    vec_kinds = (0, 2) if kind != _lib.RED_COLS else (0, 1)
    vxs = [a for a, k in zip(xs, kinds) if k in vec_kinds]
    isz = max(a.dtype.itemsize for a in vxs)         # the widest vector operand sets the vector width
    acc_size = _acc_size(reduce_type, type_map)
    known = acc_size is not None
    acc_bytes = acc_size if known else _MAX_ACC_BYTES
    index64 = layout.n_reduce >= 2 ** 31
    sm = _sm_count()
    full_vec = min(16 // isz, 8) if known and acc_bytes <= 16 else 1
    unroll = 2 if full_vec >= 8 else 4
    a0 = a1 = a2 = 0
    grid = (1, 1, 1)
    ws_need = 0
    if kind == _lib.RED_FULL:
        vec = _pick_vec([a.ptr for a in vxs], layout.n_reduce, [a.dtype.itemsize for a in vxs], full_vec)
        vec = vec if vec == full_vec else 1
        tile = _THREADS * vec * unroll
        g = max(1, min((layout.n_reduce + tile - 1) // tile, sm * 8))
        grid = (g, 1, 1)
        a0 = layout.n_reduce
        ws_need = _TICKET_BYTES + g * acc_bytes
        body = ('b200::reduce_full_body<_Op, %d, %d, %d>(p.op, p.x, p.y, p.a0, '
                'reinterpret_cast<_Op::acc_t*>(p.ws0), reinterpret_cast<uint32_t*>(p.ws1));' % (vec, unroll, _THREADS))
        tag = 'full_v%d' % vec
    elif kind == _lib.RED_ROWS:
        vec = _pick_vec([a.ptr for a in vxs], layout.n_reduce, [a.dtype.itemsize for a in vxs], full_vec)
        vec = vec if vec == full_vec else 1
        n = layout.n_reduce
        # == rows_group() in csrc/reduce_impl.cuh: the largest group whose unrolled batch fits the row
        group = _THREADS if n >= 2048 else 32 if n >= 32 * vec * unroll else 8 if n >= 8 * vec else 1
        rpb = _THREADS // group
        g = max(1, min((layout.n_out + rpb - 1) // rpb, sm * 64))
        grid = (g, 1, 1)
        a0, a1 = layout.n_out, n
        body = 'b200::reduce_rows_body<_Op, %d, %d, %d, %d>(p.op, p.x, p.y, p.a0, p.a1);' % (vec, unroll, _THREADS, group)
        tag = 'rows_v%d_g%d' % (vec, group)
    else:
        cv = full_vec
        while cv > 1 and 8 * 32 * cv * acc_bytes > 32768:
            cv >>= 1
        vec = _pick_vec([a.ptr for a in vxs], layout.n_out, [a.dtype.itemsize for a in vxs], cv)
        vec = vec if vec == cv else 1
        ru = 2 if vec >= 8 else 4
        wc = 8 if layout.n_out >= 8 * 32 * vec * 2 else 1      # == cols_wc() in csrc/reduce_impl.cuh (plain functors)
        block_cols = 32 * vec * wc
        tiles = (layout.n_out + block_cols - 1) // block_cols
        want = (sm * 8 + tiles * layout.batch - 1) // (tiles * layout.batch)
        nsplit = max(1, min(want, layout.n_reduce // 64, 65535))
        grid = (tiles, nsplit, layout.batch)
        a0, a1 = layout.n_reduce, layout.n_out
        if nsplit > 1:
            ws_need = _TICKET_BYTES + layout.batch * nsplit * layout.n_out * acc_bytes
        body = ('b200::reduce_cols_body<_Op, %d, %d, %d>(p.op, p.x, p.y, p.a0, p.a1, '
                'reinterpret_cast<_Op::acc_t*>(p.ws0), reinterpret_cast<uint32_t*>(p.ws1));' % (vec, ru, wc))
        tag = 'cols_v%d_w%d' % (vec, wc)

    key = ('s', tag, index64, tuple(a.dtype.char for a in xs), kinds, out.dtype.char, type_map, reduce_type,
           tuple(a.descr.char for a in in_args if isinstance(a, CScalar)))
    fn = kernel._memo.get(key)
    name = kernel.name + '_' + tag
    if fn is None:
        src = _functor_source(kernel, in_args, kernel.in_params, [out], kernel.out_params, type_map,
                              map_expr, reduce_expr, post_map_expr, reduce_type, index64, True,
                              kinds=kinds, layout_kind=kind)
        src += '''
struct _Params { _Op op; b200::in_ptr<_Op>::type x; _Op::out_t* y; long long a0, a1, a2; void* ws0; void* ws1; };
extern "C" __global__ void __launch_bounds__(%d) %s(const __grid_constant__ _Params p) {
  %s
}
''' % (_THREADS, name, body)
        kernel._cached_codes.setdefault((x.dtype.char,), src)
        fn = _jit.get_function(src, name, tuple(kernel.options))
        kernel._memo[key] = fn
    ws_ptr, ws_bytes = _workspace.get(ws_need, stream)
    params = _pack_op(in_args, layout.n_reduce * layout.n_out * layout.batch, layout.n_out * layout.batch)
    params += b''.join(struct.pack('<Q', a.ptr) for a in xs)
    params += struct.pack('<QqqqQQ', out.ptr, a0, a1, a2, ws_ptr + _TICKET_BYTES, ws_ptr)
    _launch(fn, grid, _THREADS, params, stream)


def _launch(fn, grid, threads, params, stream):
    if _dryrun.enabled:
        _dryrun.record('jit_reduce', name=fn.name, grid=grid, params_bytes=len(params))
        return
    import ctypes
    buf = ctypes.create_string_buffer(params, len(params))
    _lib.check(_lib.lib.b200_jit_launch(fn.handle, grid[0], grid[1], grid[2], threads, 0,
                                        ctypes.cast(buf, ctypes.c_void_p), len(params), stream))


def launch_generic(kernel, in_args, out_args, a_shape, reduce_axis, out_axis, keepdims, in_types, out_types,
                   type_map, map_expr, reduce_expr, post_map_expr, reduce_type, stream):
    raws = [(a, p) for a, p in zip(in_args, kernel.in_params) if isinstance(a, ndarray) and p.raw]
    arrays = [a for a, p in zip(in_args, kernel.in_params) if isinstance(a, ndarray) and not p.raw]
    if raws and tuple(reduce_axis) + tuple(out_axis) != tuple(range(len(a_shape))):
        # same restriction as the reference (_set_permuted_args, cupy/_core/_reduction.pyx:186-203)
        raise NotImplementedError('Illegal conditions')
    if not arrays:
        raise ValueError('Loop size is undecided.')
    nin, nout = len(arrays), len(out_args)
    red_shape = [a_shape[i] for i in reduce_axis if a_shape[i] != 1]
    out_dims = [i for i in out_axis if a_shape[i] != 1]
    red_dims = [i for i in reduce_axis if a_shape[i] != 1]
    out_shape = [a_shape[i] for i in out_dims]
    if len(red_dims) > _lib.MAX_NDIM or len(out_dims) > _lib.MAX_NDIM:
        raise NotImplementedError('more than %d kept or reduced dimensions' % _lib.MAX_NDIM)
    red_size = 1
    for s in [a_shape[i] for i in reduce_axis]:
        red_size *= s
    out_size = 1
    for s in [a_shape[i] for i in out_axis]:
        out_size *= s
    index64 = red_size >= 2 ** 31

    def pad(seq):
        seq = list(seq)
        return seq + [0] * (_lib.MAX_NDIM - len(seq))

    blob = struct.pack('<iiqq', len(out_dims), len(red_dims), out_size, red_size)
    blob += struct.pack('<%dq' % _lib.MAX_NDIM, *pad(out_shape))
    blob += struct.pack('<%dq' % _lib.MAX_NDIM, *pad(red_shape))
    for a in arrays:
        blob += struct.pack('<%dq' % _lib.MAX_NDIM, *pad(a.strides[i] for i in out_dims))
    for a in arrays:
        blob += struct.pack('<%dq' % _lib.MAX_NDIM, *pad(a.strides[i] for i in red_dims))
    for o in out_args:
        if keepdims:
            ostr = [o.strides[i] for i in out_dims]
        else:
            pos = {ax: k for k, ax in enumerate(out_axis)}
            ostr = [o.strides[pos[i]] for i in out_dims]
        blob += struct.pack('<%dq' % _lib.MAX_NDIM, *pad(ostr))
    for a in arrays:
        blob += struct.pack('<Q', a.ptr)
    for o in out_args:
        blob += struct.pack('<Q', o.ptr)

    group = _THREADS if red_size >= 1024 else 32 if red_size >= 32 else 8 if red_size >= 8 else 1
    per_block = _THREADS // group
    grid = (max(1, min((out_size + per_block - 1) // per_block, _sm_count() * 32)), 1, 1)
    key = ('g', nin, nout, group, index64, tuple(a.dtype.char for a in arrays),
           tuple(o.dtype.char for o in out_args), type_map, reduce_type,
           tuple(a.descr.char for a in in_args if isinstance(a, CScalar)),
           tuple((a.dtype.char, a.ndim, a._c_contiguous) for a, _ in raws))
    fn = kernel._memo.get(key)
    name = kernel.name + '_generic_g%d' % group
    if fn is None:
        src = _functor_source(kernel, in_args, kernel.in_params, out_args, kernel.out_params, type_map,
                              map_expr, reduce_expr, post_map_expr, reduce_type, index64, False)
        src += '''
struct _Params { _Op op; b200::GenericReduceParams<%d, %d> g; };
extern "C" __global__ void __launch_bounds__(%d) %s(const __grid_constant__ _Params p) {
  b200::reduce_generic_body<_Op, %d, %d, %d, %d>(p.op, p.g);
}
''' % (nin, nout, _THREADS, name, nin, nout, _THREADS, group)
        kernel._cached_codes.setdefault(tuple(a.dtype.char for a in arrays), src)
        fn = _jit.get_function(src, name, tuple(kernel.options))
        kernel._memo[key] = fn
    params = _pack_op(in_args, red_size * out_size, out_size, kernel.in_params) + blob
    _launch(fn, grid, _THREADS, params, stream)
