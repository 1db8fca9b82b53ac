#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- never imported by the product (cupy_b200/).

Renders the REFERENCE's own JIT kernel templates into cubins so that bench.py's `reference_gpu`
leg can time the reference's GPU kernels beside ours on the same B200 (SURVEY.md section 8d
"reference GPU baseline (same device, same run)").

Nothing of the reference is copied into this repository: the kernel templates and routine
preambles are read at BUILD time from the reference tree where it lies (`$REF`, default
/root/reference), rendered exactly as the reference's host code would render them for the
BASELINE configs, and compiled with nvcc against the reference's own headers
(`-I $REF/cupy/_core/include`) with the reference's options (`-ftz=true --std=c++17`,
cupy/cuda/compiler.py:667, cupy/_core/core.pyx:2581-2583) for sm_100 -- the arch the reference
would pick on a B200 (cupy/cuda/compiler.py:192-236).  Outputs go ONLY to oracle/_ref/jit/
(git-ignored; travels to the GPU box): one cubin per kernel + manifest.json holding, per kernel,
the by-value parameter layout and the reference's launch geometry, so the GPU box needs no
reference tree.

What is rendered (reference file:line of each template / routine):
  elementwise template        cupy/_core/_kernel.pyx:86-97  (+ `const T &x = _raw_x[_ind.get()]`
                              binding of ElementwiseKernel :709-728, ufunc glue :1024-1100)
  generic reduction template  cupy/_core/_reduction.pyx:59-112 (+ input/output exprs :565-590,
                              :886-900), geometry `_get_block_specs` :239-253 (restated in
                              `block_specs` below), `linear_launch` cupy/cuda/function.pyx:153-171
  min/max/argmax preamble     cupy/_core/_routines_statistics.pyx:190-276
  var second pass             cupy/_core/_routines_statistics.pyx:603-625 (`_var_core_float32/16`)
  CUB-block JIT reduction     cupy/_core/_cub_reduction.pyx:33-242 (argmax along the contiguous axis)
"""
from __future__ import annotations

import json
import os
import re
import string
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('REF', '/root/reference')
OUT = os.path.join(HERE, '_ref', 'jit')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')

HEADERS = ''.join('#include <%s>\n' % h for h in (
    'cupy/complex.cuh', 'cupy/carray.cuh', 'cupy/atomics.cuh', 'cupy/math_constants.h'))   # core.pyx:2455-2468


def _read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def _template_after(text, marker):
    """The first string.Template('''...''') literal after `marker`."""
    at = text.index(marker)
    m = re.compile(r"string\.Template\('''(.*?)'''\)", re.S).search(text, at)
    return m.group(1)


def _named_literal(text, name):
    m = re.search(r"^cdef\s+%s\s*=\s*'''(.*?)'''" % re.escape(name), text, re.S | re.M)
    return m.group(1)


CTYPE = {'f': 'float', 'e': 'float16', 'q': 'long long', 'd': 'double', 'i': 'int'}


def carray(t, ndim, c_contig, idx32, const=False):
    return {'kind': 'carray', 'ctype': ('const ' if const else '') + 'CArray<%s, %d, %d, %d>' % (
        CTYPE[t], ndim, int(c_contig), int(idx32)), 'ndim': ndim}


def cindexer(ndim, idx32=True):
    return {'kind': 'cindexer', 'ctype': 'CIndexer<%d, %d>' % (ndim, int(idx32)), 'ndim': ndim}


def pointer(const=False):
    return {'kind': 'pointer', 'ctype': 'const void*' if const else 'void*'}


def scalar(t):
    return {'kind': 'scalar', 'ctype': CTYPE[t], 'np': {'f': 'float32', 'i': 'int32', 'd': 'float64'}[t]}


def block_specs(in_size, out_size, contiguous_size, block_size=512):
    """Restates _get_block_specs (cupy/_core/_reduction.pyx:239-253)."""
    reduce_block_size = max(1, in_size // out_size)
    contiguous_size = min(contiguous_size, 32)
    block_stride = max(contiguous_size, block_size // reduce_block_size)
    x = block_stride // 2 + 1            # internal.clp2: next power of two >= x
    block_stride = 1 << (x - 1).bit_length()
    return block_size, block_stride, (out_size + block_stride - 1) // block_stride


def _type_decls(typedefs):
    # float16 needs its header through the type_decls slot (cupy/_core/_scalar.pyx:17, :62-83)
    return '#include "cupy/float16.cuh"\n\n' if any(c == 'e' for _, c in typedefs) else ''


def elementwise_source(tpl, name, typedefs, params, operation, preamble=''):
    plist = ', '.join('%s %s' % (p['ctype'], n) for n, p in params)
    return HEADERS + string.Template(tpl).substitute(
        type_decls=_type_decls(typedefs), typedef_preamble=''.join('typedef %s %s;\n' % (CTYPE[c], t) for t, c in typedefs),
        preamble=preamble, name=name, params=plist, loop_prep='', operation=operation, after_loop='')


def reduction_source(tpl, name, typedefs, params, block_size, reduce_type, identity, pre_map, reduce_expr,
                     post_map, input_expr, output_expr, preamble=''):
    plist = ', '.join('%s %s' % (p['ctype'], n) for n, p in params)
    return HEADERS + string.Template(tpl).substitute(
        type_decls=_type_decls(typedefs), type_preamble=''.join('typedef %s %s;\n' % (CTYPE[c], t) for t, c in typedefs),
        preamble=preamble, name=name, params=plist, block_size=block_size, reduce_type=reduce_type,
        identity=identity, reduce_expr=reduce_expr, pre_map_expr=pre_map, post_map_expr=post_map,
        input_expr=input_expr, output_expr=output_expr)


def cub_block_source(cpyx, name, typedefs, params, block_size, items_per_thread, reduce_type, identity, pre_map,
                     reduce_expr, post_map, preamble=''):
    """_create_cub_reduction_function (cupy/_core/_cub_reduction.pyx:33-242): the module text is assembled from
    the literals of that function, in order, with the `pre_map_expr == 'in0'` alternatives chosen as it does."""
    at = cpyx.index('cdef function.Function _create_cub_reduction_function')
    end = cpyx.index('type_decls = set()', at)
    lits = re.findall(r"'''(.*?)'''", cpyx[at:end], re.S)
    assert len(lits) == 6, len(lits)
    head, load_t, body, load_in0, load_map, tail = lits
    hdr = re.search(r"_cub_path == '<bundle>':\s*_cub_header = '''(.*?)'''", cpyx, re.S).group(1)
    code = hdr + head + (load_t if pre_map == 'in0' else '') + body + (load_in0 if pre_map == 'in0' else load_map) + tail
    plist = ', '.join('%s %s' % (p['ctype'], n) for n, p in params)
    return HEADERS + string.Template(code).substitute(
        name=name, block_size=block_size, items_per_thread=items_per_thread, reduce_type=reduce_type, params=plist,
        type_decls=_type_decls(typedefs), identity=identity, reduce_expr=reduce_expr, pre_map_expr=pre_map,
        post_map_expr=post_map, type_preamble=''.join('typedef %s %s;\n' % (CTYPE[c], t) for t, c in typedefs),
        preamble=preamble)


def kernels():
    kpyx = _read('cupy/_core/_kernel.pyx')
    cpyx = _read('cupy/_core/_cub_reduction.pyx')
    rpyx = _read('cupy/_core/_reduction.pyx')
    spyx = _read('cupy/_core/_routines_statistics.pyx')
    ew_tpl = _template_after(kpyx, 'cdef str _get_simple_elementwise_kernel_code')
    red_tpl = _template_after(rpyx, 'cpdef str _create_reduction_function_code')
    minmax = _named_literal(spyx, '_min_max_preamble')
    norm = _named_literal(spyx, '_norm_preamble')
    out = []

    # ---- C2: ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy'), 1-D contiguous, 32-bit index
    params = [('a', scalar('f')), ('_raw_x', carray('f', 1, 1, 1, True)), ('_raw_y', carray('f', 1, 1, 1, True)),
              ('_raw_z', carray('f', 1, 1, 1)), ('_ind', cindexer(1))]
    op = ('const T &x = _raw_x[_ind.get()];\nconst T &y = _raw_y[_ind.get()];\n'
          'T &z = _raw_z[_ind.get()];\nz = a * x + y')
    out.append(dict(name='ref_axpy', family='elementwise', params=params,
                    source=elementwise_source(ew_tpl, 'ref_axpy', [('T', 'f')], params, op)))

    # ---- ufunc kernels (glue of _get_ufunc_kernel): exp on a transposed 3-D view, add with a broadcast row
    def ufunc(name, nin, routine, ins, outp, ind, tin='f', tout='f'):
        params, op, typedefs = [], [], []
        for i in range(nin):
            params.append(('_raw_in%d' % i, ins[i]))
            typedefs.append(('in%d_type' % i, tin))
            op.append('const in%d_type in%d(_raw_in%d[_ind.get()]);' % (i, i, i))
        params.append(('_raw_out0', outp))
        typedefs.append(('out0_type', tout))
        op.append('out0_type out0;')
        op += [routine, ';', '_raw_out0[_ind.get()] = out0;']
        params.append(('_ind', ind))
        out.append(dict(name=name, family='elementwise', params=params,
                        source=elementwise_source(ew_tpl, name, typedefs, params, '\n'.join(op))))

    ufunc('ref_exp_t3', 1, 'out0 = exp(in0)', [carray('f', 3, 0, 1, True)], carray('f', 3, 1, 1), cindexer(3))
    ufunc('ref_add_b2', 2, 'out0 = in0 + in1', [carray('f', 2, 1, 1, True), carray('f', 2, 0, 1, True)],
          carray('f', 2, 1, 1), cindexer(2))
    # the astype copy scan_core makes before the in-place CUB scan (_routines_math.pyx:726-727): int64 2 GiB
    # arrays -> 64-bit CArray indexing (core.pyx:301), 32-bit indexer (2^28 items)
    ufunc('ref_copy_i64', 1, 'out0 = in0', [carray('q', 1, 1, 0, True)], carray('q', 1, 1, 0), cindexer(1), 'q', 'q')

    # ---- generic reductions (_SimpleReductionKernel): 1-D collapsed input (64-bit CArray indexing for 4 GiB,
    #      32-bit for the 2 GiB fp16 input), 1-D output, block 512
    simple_in = 'const type_in0_raw in0 = _raw_in0[_in_ind.get()];'
    simple_out = 'type_out0_raw &out0 = _raw_out0[_out_ind.get()];'

    def simple(name, tin, tout, idx32_in, pre, red, post, rtype, identity, preamble=''):
        params = [('_raw_in0', carray(tin, 1, 1, idx32_in, True)), ('_raw_out0', carray(tout, 1, 1, 1)),
                  ('_in_ind', cindexer(1)), ('_out_ind', cindexer(1)), ('_block_stride', scalar('i'))]
        typedefs = [('type_in0_raw', tin), ('type_out0_raw', tout), ('IndexT', 'i')]
        out.append(dict(name=name, family='reduction', block_size=512, params=params,
                        source=reduction_source(red_tpl, name, typedefs, params, 512, rtype, identity, pre, red,
                                                post, simple_in, simple_out, preamble)))

    mm = 'min_max_st<type_in0_raw>'
    for t, idx32, acc in (('f', 0, 'float'), ('e', 1, 'float')):
        sfx = {'f': 'f32', 'e': 'f16'}[t]
        rt = CTYPE[t] if t == 'f' else acc          # 'e->e' uses reduce_type float (_routines_math.pyx:782)
        simple('ref_sum_' + sfx, t, t, idx32, 'in0', 'a + b', 'out0 = type_out0_raw(a)', rt, '0')
        simple('ref_mean_' + sfx, t, t, idx32, 'in0', 'a + b',
               'out0 = a / _type_reduce(_in_ind.size() / _out_ind.size())', rt, '0')
        simple('ref_max_' + sfx, t, t, idx32, mm + '(in0)', 'my_max_float(a, b)', 'out0 = a.value', mm, '', minmax)
        simple('ref_argmax_' + sfx, t, 'q', idx32, mm + '(in0, _J)', 'my_argmax_float(a, b)', 'out0 = a.index',
               mm, '', minmax)

    # ---- CUB-block JIT path (_cub_reduction.pyx): argmax along the contiguous axis, one 512-thread block per row,
    #      4 items per thread (_get_cub_block_specs :391-410); args: raw pointers + int32 segment / array size (:657-664)
    for t, sfx in (('f', 'f32'), ('e', 'f16')):
        name = 'ref_cub_argmax_' + sfx
        params = [('_raw_in0', pointer(True)), ('_raw_out0', pointer()),
                  ('_segment_size', dict(scalar('i'), ctype='const int')), ('_array_size', dict(scalar('i'), ctype='const int'))]
        typedefs = [('type_in0_raw', t), ('type_out0_raw', 'q'), ('IndexT', 'i'), ('sizeT', 'i')]
        out.append(dict(name=name, family='cub_block', block_size=512, params=params,
                        source=cub_block_source(cpyx, name, typedefs, params, 512, 4, mm, '', mm + '(in0, _J)',
                                                'my_argmax_float(a, b)', 'out0 = a.index', minmax)))

    # ---- var second pass: ReductionKernel('S x, T mean, float32 alpha', 'float32 out', 'my_norm(x - mean)', ...)
    #      x and the broadcast keepdims mean stay 2-D (_reduced_view_core leaves 2-D non-contiguous sets alone,
    #      _kernel.pyx:404-405).  axis=0: x C-contiguous; axis=1: both transposed to (reduce, out).
    for t, sfx, x32 in (('f', 'f32', 0), ('e', 'f16', 1)):
        for ax, xc in ((0, 1), (1, 0)):
            name = 'ref_var_core_%s_ax%d' % (sfx, ax)
            params = [('_raw_x', carray(t, 2, xc, x32, True)), ('_raw_mean', carray(t, 2, 0, 1, True)),
                      ('alpha', scalar('f')), ('_raw_out', carray(t, 1, 1, 1)),
                      ('_in_ind', cindexer(2)), ('_out_ind', cindexer(1)), ('_block_stride', scalar('i'))]
            typedefs = [('S', t), ('T', t), ('IndexT', 'i')]
            inp = 'const S x = _raw_x[_in_ind.get()];\nconst T mean = _raw_mean[_in_ind.get()];'
            outp = '%s &out = _raw_out[_out_ind.get()];' % ('float' if t == 'f' else 'float16')
            out.append(dict(name=name, family='reduction', block_size=512, params=params,
                            source=reduction_source(red_tpl, name, typedefs, params, 512,
                                                    'float' if t == 'f' else 'float16', '0', 'my_norm(x - mean)',
                                                    'a + b', 'out = alpha * a', inp, outp, norm)))
    return out


def main():
    if not os.path.isdir(os.path.join(REF, 'cupy', '_core')):
        print('no reference tree at %s: skipping oracle/_ref/jit' % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(REF, 'cupy', '_core', 'include')
    flags = ['-cubin', '-arch=sm_100', '-ftz=true', '--std=c++17', '-DCUB_DISABLE_BF16_SUPPORT',
             '-I', inc, '-I', os.path.join(inc, 'cupy', '_cccl', 'libcudacxx'),
             '-I', os.path.join(inc, 'cupy', '_cccl', 'cub'), '-I', os.path.join(inc, 'cupy', '_cccl', 'thrust')]
    manifest = {}
    procs = []
    for k in kernels():
        cu = os.path.join('/tmp', 'refjit_%s.cu' % k['name'])
        with open(cu, 'w') as f:
            f.write(k['source'])
        cubin = os.path.join(OUT, k['name'] + '.cubin')
        procs.append((k, cu, subprocess.Popen([NVCC] + flags + ['-o', cubin, cu], stderr=subprocess.PIPE)))
        manifest[k['name']] = {'family': k['family'], 'block_size': k.get('block_size', 128),
                               'params': [dict(p, name=n) for n, p in k['params']]}
    rc = 0
    for k, cu, p in procs:
        err = p.communicate()[1].decode()
        if p.returncode:
            sys.stderr.write('== %s failed\n%s\n' % (k['name'], err))
            rc = 1
        os.unlink(cu)
    with open(os.path.join(OUT, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    print('rendered %d reference JIT kernels into %s' % (len(manifest), OUT))
    return rc


if __name__ == '__main__':
    sys.exit(main())
