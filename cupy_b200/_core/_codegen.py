"""CUDA source generation for NVRTC: wraps operation strings of ufunc loops,
ElementwiseKernels and ReductionKernels in the hand-written skeleton headers
(csrc/include/b200/*.cuh).  Only glue is generated here -- the loops, tilers,
vector access, shared-memory staging, warp/block/grid combines are the header
code that the prebuilt kernels use too.

Reference counterparts: `_get_simple_elementwise_kernel_code`
(cupy/_core/_kernel.pyx:78-107), `_get_elementwise_kernel_code` (:709-728),
`_get_ufunc_kernel` (:1024-1100), `_create_reduction_function_code`
(cupy/_core/_reduction.pyx:44-126).  User-visible names are kept: parameters by
name, `i`, `_ind`, `in0/out0`, `in0_type/out0_type`, `_raw_<name>` is NOT
provided for non-raw arrays (values live in registers here).
"""
from __future__ import annotations

import re

from cupy_b200 import _lib
from cupy_b200._core._scalar import get_dtype, get_typename

_PROLOGUE = '''#include <b200/elementwise.cuh>
#include <b200/carray.cuh>
using b200::float16;
'''


class EwSpec:
    """Everything that defines the text of an elementwise kernel except the
    per-call layout (variant / vector width / dims)."""

    def __init__(self, mode, operation, preamble='', loop_prep='', after_loop='', options=(),
                 in_types=None, out_types=None, has_where=False, type_map=(), write_only_outputs=False):
        self.mode = mode                               # 'elementwise' (user kernel) | 'ufunc'
        self.write_only_outputs = write_only_outputs   # generated text (cupy_b200.fuse): outputs are only assigned
        self.operation = operation
        self.preamble = preamble
        self.loop_prep = loop_prep
        self.after_loop = after_loop
        self.options = tuple(options)
        self.in_types = in_types
        self.out_types = out_types
        self.has_where = has_where
        self.type_map = type_map
        text = ' '.join((operation, loop_prep, after_loop))
        self.uses_ind = re.search(r'\b_ind\b', text) is not None
        self.memo = {}
        self.last_source = None
        self._bound = {}

    def bind(self, type_map):
        s = self._bound.get(type_map)
        if s is None:
            s = EwSpec(self.mode, self.operation, self.preamble, self.loop_prep, self.after_loop,
                       self.options, self.in_types, self.out_types, self.has_where, type_map,
                       self.write_only_outputs)
            self._bound[type_map] = s
        return s


def _fix_cast_expr(src_type, dst_type, expr):
    """cupy/_core/_kernel.pyx:1005-1021."""
    src_kind = get_dtype(src_type).kind
    dst_kind = get_dtype(dst_type).kind
    if src_kind == dst_kind:
        return expr
    if src_kind == 'b':
        return '(%s) ? 1 : 0' % expr
    return expr


_CONTROL_FLOW = re.compile(r'\b(if|else|for|while|do|switch|case|return|goto|continue|break)\b|[{}]')


def writes_first(operation, name):
    """True when every statement of `operation` runs unconditionally and the first mention
    of output `name` is the left side of a plain assignment at the start of a statement:
    its previous value can never be observed, so the kernel does not have to load it.
    (The reference binds outputs as references into memory, cupy/_core/_kernel.pyx:86-97,
    so an unread output costs it nothing; here operands travel through register packs and
    a load the compiler cannot prove dead doubles the write traffic.)"""
    if _CONTROL_FLOW.search(operation):
        return False
    m = re.search(r'\b%s\b' % re.escape(name), operation)
    if m is None:
        return False
    before = operation[:m.start()].rstrip()
    after = operation[m.end():].lstrip()
    if not ((before == '' or before.endswith(';')) and after.startswith('=') and not after.startswith('==')):
        return False
    # the right-hand side of that first assignment must not read the output (`y = y + x`,
    # `y = m ? x : y`): the statement ends at the next `;` (no braces / control flow here)
    rhs = after[1:].split(';', 1)[0]
    return re.search(r'\b%s\b' % re.escape(name), rhs) is None


def _tiler_type(variant, nargs, vec, unroll, threads, idx32, spec=0):
    if variant == _lib.EW_FLAT:
        return 'b200::FlatTiler<%d, %d, %d, %d, %s>' % (nargs, vec, unroll, threads, 'true' if spec else 'false')
    if variant == _lib.EW_ROWWISE:
        return 'b200::RowTiler<%d, %d, %d, %d, %s, %dull>' % (nargs, vec, unroll, threads, 'true' if idx32 else 'false', spec)
    if variant == _lib.EW_TILED_TMA:
        return 'b200::TmaTileTiler<%d, %d>' % (nargs, 16 // vec)
    if variant == _lib.EW_TILED_REG:
        return 'b200::RegTileTiler<%d, %d, %d, %dull>' % (nargs, 16 // vec, unroll, spec)
    return 'b200::TileTiler<%d>' % nargs


def render_elementwise(spec, name, args, params, variant, vec, unroll, threads, idx32, ndim, access_spec=0,
                       min_blocks=1, ind_ndim=0):
    """`ind_ndim` > 0: `_ind` is built from the un-collapsed loop shape of that rank (reduce_dims=False
    kernels), which arrives as one extra shape-only view behind the raw operands."""
    from cupy_b200._core._ndarray import ndarray
    nargs = len(args)
    if variant == _lib.EW_TILED:
        vec, unroll, threads = 1, 4, 256
    tma = variant == _lib.EW_TILED_TMA
    if tma:
        unroll, threads = vec, 288    # 8 consumer warps + the TMA producer warp
    lines = [_PROLOGUE]
    for ctype, dt in spec.type_map:
        lines.append('typedef %s %s;' % (get_typename(dt), ctype))
    typedefs = []
    if spec.mode == 'ufunc':
        for k, t in enumerate(spec.in_types):
            typedefs.append('typedef %s in%d_type;' % (get_typename(t), k))
        for k, t in enumerate(spec.out_types):
            typedefs.append('typedef %s out%d_type;' % (get_typename(t), k))
    lines.extend(typedefs)
    lines.append(spec.preamble)
    min_blocks = ', %d' % min_blocks if variant == _lib.EW_TILED_REG else ''
    n_views = sum(1 for a, p in zip(args, params) if isinstance(a, ndarray) and p.raw) + (1 if ind_ndim else 0)
    lines.append('extern "C" __global__ void __launch_bounds__(%d%s) %s('
                 'const __grid_constant__ b200::EwParams _p, const __grid_constant__ b200::RawPackN<%d> _rv%s) {'
                 % (threads, min_blocks, name, n_views, ', const __grid_constant__ b200::TileMaps _tm' if tma else ''))
    lines.append('  typedef %s _Tiler;' % _tiler_type(variant, nargs, vec, unroll, threads, idx32, access_spec))
    lines.append('  constexpr int _V = _Tiler::kV, _U = _Tiler::kU;')

    decl, loads, binds, stores = [], [], [], []
    n_raw = 0
    n_in = sum(1 for p in params if p.is_const or p.name == '_where')  # informational only
    for k, (a, p) in enumerate(zip(args, params)):
        if isinstance(a, ndarray):
            mem_t = get_typename(a.dtype)
            if p.raw:
                const = 'const ' if p.is_const else ''
                decl.append('  %sCArray<%s, %d, %s, false> %s(_rv.v[%d]);'
                            % (const, mem_t, a.ndim, 'true' if a._c_contiguous else 'false', p.name, n_raw))
                n_raw += 1
                continue
            reg = '_v_%s' % p.name
            loads.append('    b200::Pack<%s, _V> %s[_U];' % (mem_t, reg))
            is_out = not p.is_const
            if spec.mode == 'ufunc':
                if p.name == '_where':
                    loads.append('    _t.template load<_FULL>(%d, %s);' % (k, reg))
                    binds.insert(0, '        if (!%s[_u][_k]) continue;' % reg)
                elif p.name.startswith('in'):
                    idx = int(p.name[2:])
                    loads.append('    _t.template load<_FULL>(%d, %s);' % (k, reg))
                    binds.append('        const in%d_type in%d(%s);' % (
                        idx, idx, _fix_cast_expr(a.dtype, spec.in_types[idx], '%s[_u][_k]' % reg)))
                else:
                    idx = int(p.name[3:])
                    if spec.has_where:
                        loads.append('    _t.template load<_FULL>(%d, %s);' % (k, reg))   # keep unselected elements
                    binds.append('        out%d_type out%d;' % (idx, idx))
                    stores.append(('        %s[_u][_k] = %s;' % (
                        reg, _fix_cast_expr(spec.out_types[idx], a.dtype, 'out%d' % idx)),
                        '    _t.template store<_FULL>(%d, %s);' % (k, reg)))
            else:
                # user kernel: outputs are read-modify-write capable unless the operation
                # provably writes them first
                if not (is_out and (spec.write_only_outputs or writes_first(spec.operation, p.name))):
                    loads.append('    _t.template load<_FULL>(%d, %s);' % (k, reg))
                if is_out:
                    binds.append('        %s& %s = %s[_u][_k];' % (p.ctype, p.name, reg))
                    stores.append((None, '    _t.template store<_FULL>(%d, %s);' % (k, reg)))
                else:
                    binds.append('        const %s& %s = %s[_u][_k];' % (p.ctype, p.name, reg))
        else:
            if spec.mode == 'ufunc' and p.name == '_where':
                decl.append('  if (!b200::scalar_arg<bool>(_p, %d)) return;' % k)
            elif spec.mode == 'ufunc':
                idx = int(p.name[2:])
                decl.append('  const in%d_type in%d = b200::scalar_arg<in%d_type>(_p, %d);' % (idx, idx, idx, k))
            else:
                t = p.ctype
                decl.append('  const %s %s = b200::scalar_arg<%s>(_p, %d);' % (t, p.name, t, k))
    lines.extend(decl)
    if ind_ndim:
        lines.append('  CIndexer<%d> _ind(_p.size, _rv.v[%d].shape);' % (ind_ndim, n_views - 1))
    else:
        lines.append('  CIndexer<%d> _ind(_p.size, _p.shape);' % (ndim if spec.uses_ind else 1))
    lines.append('  ' + spec.loop_prep + ';')
    lines.append('  _Tiler _t(_p, _tm);' if tma else '  _Tiler _t(_p);')
    lines.append('  auto _tile = [&](auto _full_tag) {')
    lines.append('    constexpr bool _FULL = decltype(_full_tag)::value;')
    lines.extend(loads)
    lines.append('#pragma unroll')
    lines.append('    for (int _u = 0; _u < _U; ++_u) {')
    lines.append('#pragma unroll')
    lines.append('      for (int _k = 0; _k < _V; ++_k) {')
    lines.append('        if (!_t.template in_range<_FULL>(_u, _k)) continue;')
    lines.append('        const ptrdiff_t i = _t.index(_u, _k);')
    if spec.uses_ind:
        lines.append('        _ind.set(i);')
    lines.extend(binds)
    lines.append('        ' + spec.operation + ';')
    for assign, _ in stores:
        if assign:
            lines.append(assign)
    lines.append('      }')
    lines.append('    }')
    for _, st in stores:
        lines.append(st)
    lines.append('  };')
    lines.append('  for (; _t.valid(); _t.next()) {')
    lines.append('    if (_t.is_full()) _tile(b200::true_t()); else _tile(b200::false_t());')
    lines.append('  }')
    lines.append('  ' + spec.after_loop + ';')
    lines.append('}')
    return '\n'.join(lines)
