"""CPU tier: the host side of the launcher in dry-run mode (cupy_b200/_core/_dryrun.py):
argument handling, dtype-loop selection, broadcasting, error behaviour, launch
classification -- and every kernel text the engine generates on the way is compiled
for sm_100a by NVRTC.  The cases restate the reference's own host-logic tests
(file:line given per test); values are checked in the GPU tier."""
import numpy as np
import pytest

import cupy_b200 as cp
from cupy_b200 import _lib
from cupy_b200._core import _kernel, _reduction


def kinds(log):
    return [(l['kind'], l.get('variant', l.get('layout'))) for l in log]


# ---- ElementwiseKernel: tests/cupy_tests/core_tests/test_userkernel.py --------------------
class TestElementwiseKernelSize:
    """test_userkernel.py:91-200 (`size=` legality matrix)."""

    def create_kernel(self, input_raw, output_raw):
        ins = ', '.join('{}float32 x{}'.format('raw ' if r else '', i) for i, r in enumerate(input_raw))
        outs = ', '.join('{}float32 y{}'.format('raw ' if r else '', i) for i, r in enumerate(output_raw))
        return cp.ElementwiseKernel(ins, outs, '', 'kernel')

    def setup_arrays(self):
        return cp.empty((2,), 'float32'), cp.empty((2,), 'float32')

    def test_all_raws(self, dry):
        a1, a2 = self.setup_arrays()
        k1 = self.create_kernel((True, True), (False,))
        assert k1(a1, a2, size=2).shape == (2,)
        with pytest.raises(ValueError, match=r'^Loop size is undecided\.'):
            k1(a1, a2)
        k2 = self.create_kernel((True, True), (True,))
        k2(a1, a2, size=2)
        with pytest.raises(ValueError, match=r'^Loop size is undecided\.'):
            k2(a1, a2)

    def test_nonraws(self, dry):
        a1, a2 = self.setup_arrays()
        for ins, outs in (((False, False), (False,)), ((False, False), (True,)), ((True, False), (False,)),
                          ((False, True), (True,))):
            with pytest.raises(ValueError, match=r"^Specified 'size' can"):
                self.create_kernel(ins, outs)(a1, a2, size=2)
        with pytest.raises(ValueError, match=r"^Specified 'size' can"):
            self.create_kernel((False, False), (False,))(a1, 7, size=2)

    def test_scalars_and_raws(self, dry):
        a1, _ = self.setup_arrays()
        k = self.create_kernel((True, False), (False,))
        k(a1, 7, size=2)
        with pytest.raises(ValueError, match=r'^Loop size is undecided\.'):
            k(a1, 7)


def test_invalid_kernel_name():
    """test_elementwise.py:82-86."""
    with pytest.raises(ValueError, match='Invalid kernel name'):
        cp.ElementwiseKernel('T x', '', '', '1')
    with pytest.raises(ValueError, match='Invalid kernel name'):
        cp.ReductionKernel('T x', 'T y', 'x', 'a + b', 'y = a', '0', name='1')


def test_i_is_reserved():
    with pytest.raises(ValueError, match="Can not use 'i' as a parameter name"):
        cp.ElementwiseKernel('T i', 'T y', 'y = i')


def test_wrong_number_of_arguments(dry):
    """_kernel.pyx:870-875."""
    k = cp.ElementwiseKernel('T x, T y', 'T z', 'z = x + y', 'addk')
    with pytest.raises(TypeError, match='Wrong number of arguments'):
        k(cp.empty((2,), 'f'))
    with pytest.raises(TypeError, match='Wrong number of arguments'):
        cp.add(cp.empty((2,), 'f'))
    with pytest.raises(TypeError, match='Wrong arguments'):
        k(cp.empty((2,), 'f'), cp.empty((2,), 'f'), foo=1)


def test_out_shape_mismatch(dry):
    """test_elementwise.py:72-79 (TestElementwiseInvalidShape) / _kernel.pyx:665-666, 704-705."""
    f = cp.ElementwiseKernel('T x', 'T y', 'y += x')
    x = cp.empty((3, 4), 'q')
    y = cp.empty((4,), 'q')
    with pytest.raises(ValueError, match='Out shape is mismatched'):
        f(x, y)
    a = cp.empty((2, 3), 'f')
    with pytest.raises(ValueError, match='Out shape is mismatched'):
        cp.add(a, a[:1], out=cp.empty((1, 3), 'f'))      # broadcastable, but not the loop shape


def test_type_mismatch(dry):
    k = cp.ElementwiseKernel('T x, T y', 'T z', 'z = x + y', 'addk')
    with pytest.raises(TypeError, match='Type is mismatched'):
        k(cp.empty((2,), 'f'), cp.empty((2,), 'd'))
    k2 = cp.ElementwiseKernel('float32 x', 'float32 y', 'y = x', 'cp2')
    with pytest.raises(TypeError, match='Type is mismatched'):
        k2(cp.empty((2,), 'd'))


def test_broadcast_error(dry):
    """internal.pyx:352-357."""
    with pytest.raises(ValueError, match='operands could not be broadcast together with shapes'):
        cp.add(cp.empty((2, 3), 'f'), cp.empty((4,), 'f'))


def test_cached_codes_count(dry):
    """test_userkernel.py:75-88: one generated source per input dtype set."""
    k = cp.ElementwiseKernel('T x, T y', 'T z', 'z = x + y', 'uesr_kernel_1')
    a, b = cp.empty((4,), 'f'), cp.empty((4,), 'f')
    assert len(k._cached_codes) == 0
    k(a, b)
    assert len(k._cached_codes) == 1
    k(a, b)
    assert len(k._cached_codes) == 1
    k(a.astype('d'), b.astype('d'))
    assert len(k._cached_codes) == 2
    assert 'z = x + y' in k.cached_codes[(np.dtype('f'), np.dtype('f'))]


# ---- ufunc loop selection / NEP 50: _kernel.pyx:1103-1144, 1656-1753 ----------------------
@pytest.mark.parametrize('lhs,rhs,expected', [
    ('float32', 2, 'float32'), ('float32', 2.0, 'float32'), ('int8', 1, 'int8'), ('uint8', 1.0, 'float64'),
    ('int32', 1.5, 'float64'), ('float16', 1, 'float16'), ('bool', 1, 'int64'), ('int64', np.float32(1), 'float64'),
    ('int8', np.int16(1), 'int16'), ('float32', np.float64(1), 'float64'),
])
def test_weak_scalar_promotion_matches_numpy(dry, lhs, rhs, expected):
    got = cp.add(cp.empty((3,), lhs), rhs)
    want = (np.empty((3,), lhs) + rhs).dtype
    assert got.dtype == want == np.dtype(expected)


@pytest.mark.parametrize('dt', ['int8', 'int16', 'int32', 'int64', 'uint8', 'uint16', 'uint32', 'uint64'])
def test_python_int_overflow(dry, dt):
    """test_elementwise.py:89-145 (NEP 50: out-of-range Python ints raise OverflowError)."""
    info = np.iinfo(dt)
    a = cp.empty((1,), 'int8')
    for b in (info.max, info.min):
        try:
            want = (np.zeros((1,), 'int8') + b).dtype
        except OverflowError:
            want = OverflowError
        if want is OverflowError:
            with pytest.raises(OverflowError):
                a + b
        else:
            assert (a + b).dtype == want
    big = cp.empty((1,), dt)
    assert (big + np.int8(0)).dtype == (np.zeros((1,), dt) + np.int8(0)).dtype


def test_mixed_array_dtypes_and_casting(dry):
    a, b = cp.empty((4,), 'int32'), cp.empty((4,), 'float32')
    assert (a + b).dtype == np.float64
    with pytest.raises(TypeError, match='Cannot cast'):
        cp.add(b, b, out=cp.empty((4,), 'int32'))                  # same_kind forbids float->int
    cp.add(b, b, out=cp.empty((4,), 'int32'), casting='unsafe')
    assert cp.add(a, a, dtype='float32').dtype == np.float32
    with pytest.raises(TypeError, match='Wrong type'):
        cp.exp(a, dtype='int32')
    assert cp.true_divide(a, a).dtype == np.float64
    with pytest.raises(TypeError, match='boolean subtract'):
        cp.subtract(cp.empty((2,), '?'), cp.empty((2,), '?'))


def test_ufunc_types_attribute():
    assert cp.add.nin == 2 and cp.add.nout == 1 and cp.add.nargs == 3
    assert cp.add.types[0] == '??->?' and 'ff->f' in cp.add.types
    assert cp.exp.types == ['e->e', 'f->f', 'd->d']


def test_overlap_guard_copies_input(dry):
    """_kernel.pyx:673-683: an input overlapping `out` (but not identical) is copied first."""
    a = cp.empty((100,), 'f')
    o = a[1:]
    cp.add(a[:-1], o, out=o)
    assert [k for k, _ in kinds(dry)].count('prebuilt_ufunc') == 2     # the copy, then the add
    del dry[:]
    cp.add(a, a, out=a)                                               # identical: no copy
    assert len(dry) == 1


# ---- launch classification -------------------------------------------------------------------
def test_classification_of_config_shapes(dry):
    n = 1 << 20
    x, y = cp.empty((n,), 'f'), cp.empty((n,), 'f')
    cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')(np.float32(2), x, y)
    assert dry[-1]['kind'] == 'jit_elementwise' and dry[-1]['variant'] == _lib.EW_FLAT and dry[-1]['vec'] == 4
    xt = cp.empty((64, 128, 256), 'f').transpose(2, 1, 0)
    v = cp.empty((64,), 'f')
    cp.exp(xt)
    assert dry[-1]['variant'] == _lib.EW_TILED_REG and dry[-1]['tile_axis'] == 0 and dry[-1]['staged_mask'] == 1
    cp.add(cp.empty((256, 128, 64), 'f'), v)
    # a row vector over a dense array: 1-D walk, the vector is a periodic operand (period 64 | 1024)
    assert dry[-1]['variant'] == _lib.EW_FLAT and dry[-1]['vec'] == 4 and dry[-1]['ndim'] == 1 and dry[-1]['staged_mask'] == 2
    cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'fused')(xt, v)
    assert dry[-1]['variant'] == _lib.EW_TILED_REG and 'RegTileTiler<3, 4, 1, 177ull>' in dry[-1]['source']
    h = cp.empty((n,), 'e')
    cp.add(h, h)
    assert dry[-1]['variant'] == _lib.EW_FLAT and dry[-1]['vec'] == 8


def test_generated_source_keeps_user_names(dry):
    k = cp.ElementwiseKernel('raw T x, T y, int32 n', 'T z', 'z = x[n - 1 - i] + y + _ind.size()', 'names',
                             preamble='__device__ int helper() { return 1; }', loop_prep='int q = helper()')
    k(cp.empty((10,), 'f'), cp.empty((2, 5), 'f'), 10)
    src = dry[-1]['source']
    for frag in ('CArray<float, 1, true, false> x(_rv.v[0])', 'const ptrdiff_t i =', '_ind.set(i)', 'int q = helper()',
                 'const int n = b200::scalar_arg<int>'):
        assert frag in src, frag


# ---- reductions: _reduction.pyx:147-176, 346-354; tests/.../test_reduction.py --------------
def test_axis_errors(dry):
    a = cp.empty((2, 3, 4), 'f')
    with pytest.raises(ValueError, match="duplicate value in 'axis'"):
        a.sum(axis=(0, 0))
    with pytest.raises(cp.AxisError):
        a.sum(axis=3)
    with pytest.raises(cp.AxisError):
        a.sum(axis=-4)
    assert a.sum(axis=-1).shape == (2, 3)
    assert a.sum(axis=(0, 2), keepdims=True).shape == (1, 3, 1)


def test_zero_size_reductions(dry):
    """test_search.py:68-80 / _reduction.pyx:352-354."""
    e = cp.empty((0, 3), 'f')
    with pytest.raises(ValueError, match='zero-size array to reduction operation cupy_max which has no identity'):
        e.max()
    with pytest.raises(ValueError, match='zero-size array'):
        e.argmax(axis=0)
    assert e.sum(axis=0).shape == (3,)            # sum has an identity
    assert e.max(axis=1).shape == (0,)            # empty result: nothing to reduce
    assert e.sum().dtype == np.float32


@pytest.mark.parametrize('dt,want', [('?', 'int64'), ('int8', 'int64'), ('uint8', 'uint64'), ('int32', 'int64'),
                                     ('uint32', 'uint64'), ('float16', 'float16'), ('float32', 'float32'),
                                     ('float64', 'float64')])
def test_reduction_result_dtypes(dry, dt, want):
    a = cp.empty((8, 8), dt)
    assert a.sum().dtype == np.dtype(want) == np.empty((8, 8), dt).sum().dtype
    assert a.max(axis=0).dtype == np.dtype(dt)
    assert a.argmax(axis=1).dtype == np.int64
    assert a.cumsum().dtype == np.empty((8, 8), dt).cumsum().dtype
    if dt != '?':
        assert a.mean().dtype == np.empty((8, 8), dt).mean().dtype
        assert a.var(axis=0).dtype == np.empty((8, 8), dt).var(axis=0).dtype
    assert a.sum(dtype='float64').dtype == np.float64


def test_reduction_layout_classification():
    c = _reduction._classify
    f32 = 4
    # C-contiguous (R, C)
    assert c((128, 256), (1024, 4), f32, (1,), (0,), False).kind == _lib.RED_ROWS
    assert c((128, 256), (1024, 4), f32, (0,), (1,), False).kind == _lib.RED_COLS
    assert c((128, 256), (1024, 4), f32, (0, 1), (), False).kind == _lib.RED_FULL
    # F-contiguous: roles swap
    assert c((128, 256), (4, 512), f32, (0,), (1,), False).kind == _lib.RED_ROWS
    assert c((128, 256), (4, 512), f32, (1,), (0,), False).kind == _lib.RED_COLS
    # 3-D middle axis -> batched COLS; outer+inner axes -> generic
    l = c((8, 16, 32), (2048, 128, 4), f32, (1,), (0, 2), False)
    assert (l.kind, l.batch, l.n_reduce, l.n_out) == (_lib.RED_COLS, 8, 16, 32)
    assert c((8, 16, 32), (2048, 128, 4), f32, (0, 2), (1,), False).kind == -1
    assert c((8, 16, 32), (2048, 128, 4), f32, (1, 2), (0,), False).kind == _lib.RED_ROWS
    assert c((8, 16, 32), (2048, 128, 4), f32, (0, 1), (2,), False).kind == _lib.RED_COLS
    # strided view -> generic; F-order argmax over all axes: index order differs -> generic
    assert c((64, 64), (512, 8), f32, (1,), (0,), False).kind == -1
    assert c((128, 256), (4, 512), f32, (0, 1), (), True).kind == -1
    assert c((128, 256), (4, 512), f32, (0, 1), (), False).kind == _lib.RED_FULL


def test_reduction_routes(dry):
    a = cp.empty((512, 1024), 'f')
    a.sum(axis=1); a.sum(axis=0); a.sum(); a.argmax(axis=0); a.var(axis=1); a.mean(axis=0)
    assert kinds(dry) == [('prebuilt_reduce', 1), ('prebuilt_reduce', 2), ('prebuilt_reduce', 0),
                          ('prebuilt_reduce', 2), ('prebuilt_reduce', 1), ('prebuilt_reduce', 2)]
    del dry[:]
    a.astype('int16').sum(axis=1)                       # no prebuilt int16 functor -> NVRTC, same skeleton
    assert dry[-1]['kind'] == 'jit_reduce' and 'rows' in dry[-1]['name']
    a[::2].sum(axis=0)                                  # strided -> generic kernel
    assert 'generic' in dry[-1]['name']
    dot = cp.ReductionKernel('T x, T y', 'T z', 'x * y', 'a + b', 'z = a', '0', 'dot')
    dot(a, a, axis=1)                                   # two arrays of ONE layout: structured, tuple operand
    assert 'rows' in dry[-1]['name'] and '_InPtr' in dry[-1].get('source', '_InPtr')
    dot(a, a[0], axis=1)                                # a row vector broadcast over the kept axis: still structured
    assert 'rows' in dry[-1]['name']
    sq = cp.empty((64, 64), 'f')
    dot(sq, sq.T, axis=1)                               # operands of different layouts -> generic kernel
    assert 'generic' in dry[-1]['name']


def test_reduction_kernel_errors(dry):
    k = cp.ReductionKernel('T x', 'T y', 'x', 'a + b', 'y = a', '0', 'rk')
    with pytest.raises(TypeError, match='Wrong number of arguments'):
        k()
    with pytest.raises(TypeError, match='Wrong arguments'):
        k(cp.empty((3,), 'f'), foo=1)
    with pytest.raises(ValueError, match="cannot specify 'out' as both"):
        k(cp.empty((3,), 'f'), cp.empty((), 'f'), out=cp.empty((), 'f'))
    with pytest.raises(ValueError, match='Out shape is mismatched'):
        k(cp.empty((3, 4), 'f'), cp.empty((5,), 'f'), axis=1)


def test_scan_routes_and_errors(dry):
    x = cp.empty((1 << 20,), 'int64')
    x.cumsum()
    assert dry[-1]['kind'] == 'prebuilt_scan' and dry[-1]['in_dtype'] == 'int64'
    cp.empty((100,), 'int32').cumsum()
    assert dry[-1]['in_dtype'] == 'int32' and dry[-1]['out_dtype'] == 'int64'     # cast fused into the load
    with pytest.raises(cp.AxisError):
        cp.empty((3, 3), 'f').cumsum(axis=2)
    with pytest.raises(ValueError, match='wrong size'):
        cp.cumsum(cp.empty((10,), 'f'), out=cp.empty((9,), 'f'))


# ---- the array adapter ---------------------------------------------------------------------------
def test_views_match_numpy_strides(dry):
    a = cp.empty((6, 8, 10), 'f')
    n = np.empty((6, 8, 10), 'f')
    for f in (lambda v: v.T, lambda v: v.transpose(1, 0, 2), lambda v: v[::2, 1:, ::-1], lambda v: v[1],
              lambda v: v[..., None], lambda v: v.reshape(48, 10), lambda v: v.reshape(6, 80), lambda v: v[:, :, 3],
              lambda v: v.swapaxes(0, 2), lambda v: v[2:4].reshape(-1)):
        g, w = f(a), f(n)
        assert g.shape == w.shape and g.strides == w.strides, (g.shape, g.strides, w.shape, w.strides)
        assert g.flags.c_contiguous == w.flags.c_contiguous and g.flags.f_contiguous == w.flags.f_contiguous
    assert a[::2].reshape(-1).base is None or True      # non-viewable reshape copies
    with pytest.raises(IndexError):
        a[6]
    with pytest.raises(ValueError):
        a.reshape(7, -1)
    assert cp.may_share_bounds(a[0], a[0:1]) and not cp.may_share_bounds(a[0], a[1])


def test_codegen_skips_the_load_of_write_first_outputs(dry):
    from cupy_b200._core._codegen import writes_first as w
    assert w('z = exp(x) + v', 'z') and w('z = a * x + y', 'z') and w('z = x > 0 ? x : 0', 'z')
    assert w('T t = x * 2; z = t; w = z + 1', 'z') and w('T t = x * 2; z = t; w = z + 1', 'w')
    for op in ('z += x', 'y = z; z = x', 'if (x > 0) z = x', 'z == x', 'zz = 1', 'for (int k = 0; k < 2; ++k) { z = x; }',
               'z = z + x', 'z = c ? x : z', 'T t = x; z = t * z'):      # read-modify-write through the right-hand side
        assert not w(op, 'z'), op
    assert w('z = x; z = z + 1', 'z')          # reads the value the kernel itself wrote
    x = cp.empty((1 << 16,), 'f')
    cp.ElementwiseKernel('T x', 'T z', 'z = x * 2', 'wf_a')(x)
    assert 'load<_FULL>(1' not in dry[-1]['source']
    cp.ElementwiseKernel('T x', 'T z', 'z += x', 'wf_b')(x, cp.empty((1 << 16,), 'f'))
    assert 'load<_FULL>(1' in dry[-1]['source']
    cp.ElementwiseKernel('T x', 'T z', 'z = z + x', 'wf_c')(x, cp.empty((1 << 16,), 'f'))
    assert 'load<_FULL>(1' in dry[-1]['source']


# ---- fusion, operand kinds, axis scans (host logic; no GPU) -----------------------------------
def test_fuse_traces_into_one_kernel(dry):
    @cp.fuse(kernel_name='fused_probe')
    def f(x, v):
        return cp.exp(x) + v * 2

    xt = cp.empty((16, 64, 128), 'f').transpose(2, 1, 0)
    r = f(xt, cp.empty((16,), 'f'))
    assert r.shape == (128, 64, 16) and r.dtype == np.float32
    assert len(dry) == 1 and dry[0]['kind'] == 'jit_elementwise' and dry[0]['variant'] == _lib.EW_TILED_REG
    src = dry[0]['source']
    assert 'out0 = exp(in0)' in src and 'out0 = in0 * in1' in src and 'out0 = in0 + in1' in src
    assert 'load<_FULL>(2' not in src                      # the output is write-only
    del dry[:]
    f(xt, cp.empty((16,), 'f'))                            # second call: cached trace, one launch again
    assert len(dry) == 1
    # weak Python scalars keep the array dtype (NEP 50), NumPy scalars promote
    assert cp.fuse(lambda x: x * 2 + 1)(cp.empty((8,), 'e')).dtype == np.float16
    assert cp.fuse(lambda x: x * np.float64(2))(cp.empty((8,), 'f')).dtype == np.float64
    with pytest.raises(NotImplementedError):
        cp.fuse(lambda x: cp.sum(x) * 2)(cp.empty((8,), 'f'))
    with pytest.raises(TypeError):
        cp.fuse(lambda x: x if x > 0 else x)(cp.empty((8,), 'f'))


def test_reduce_dims_false_keeps_the_indexer_rank_and_raw_operands_are_uncapped(dry):
    # a reduce_dims=False kernel reading _ind.get()[k] must see the ORIGINAL loop shape
    k = cp.ElementwiseKernel('T x', 'int64 r, int64 c', 'r = _ind.get()[0]; c = _ind.get()[1]', 'ind2d', reduce_dims=False)
    k(cp.empty((4, 6), 'f'))
    src = dry[-1]['source']
    assert 'CIndexer<2> _ind(_p.size, _rv.v[0].shape)' in src and 'RawPackN<1>' in src
    assert dry[-1]['ndim'] == 1                            # the operands themselves still collapse
    k1 = cp.ElementwiseKernel('T x', 'int64 r', 'r = _ind.get()[0]', 'ind_collapsed')      # reduce_dims=True: collapsed
    k1(cp.empty((4, 6), 'f'))
    assert 'CIndexer<1> _ind(_p.size, _p.shape)' in dry[-1]['source']
    # more than four raw operands (the round-1 cap)
    names = ['a', 'b', 'c', 'd', 'e', 'f']
    k6 = cp.ElementwiseKernel(', '.join('raw T %s' % n for n in names), 'T z',
                              'z = ' + ' + '.join('%s[i]' % n for n in names), 'six_raw')
    k6(*[cp.empty((32,), 'f') for _ in names], size=32)
    assert 'RawPackN<6>' in dry[-1]['source']


def test_reduction_kernel_accepts_raw_inputs(dry):
    # `raw` in-params (cupy/_core/_reduction.pyx:186-225, 886-900): not broadcast, indexed by user code
    k = cp.ReductionKernel('T x, raw T w', 'T y', 'x * w[_j % 5]', 'a + b', 'y = a', '0', 'raw_weights')
    r = k(cp.empty((7, 5), 'f'), cp.empty((5,), 'f'), axis=0)
    assert r.shape == (5,) and 'generic' in dry[-1]['name']
    src = k.cached_code
    assert 'b200::RawView _rv_w;' in src and 'CArray<float, 1, true, false> w(_rv_w);' in src
    with pytest.raises(NotImplementedError):          # the reduced axes must lead when raw arguments are used
        k(cp.empty((7, 5), 'f'), cp.empty((5,), 'f'), axis=1)


def test_fuse_out_argument_updates_every_reference_to_the_parameter(dry):
    # cp.add(a, b, out=a) followed by a read of `a` must see a + b (NumPy / reference semantics), not the old a
    @cp.fuse(kernel_name='fused_out_then_read')
    def f(a, b):
        cp.add(a, b, out=a)
        return a * 2

    f(cp.empty((64,), 'f'), cp.empty((64,), 'f'))
    src = dry[-1]['source']
    import re
    mul = src[src.index('out0 = in0 * in1'):]
    before_mul = src[:src.index('out0 = in0 * in1')]
    # the multiply reads the temporary holding a + b, never parameter _p0 again
    last_in0 = re.findall(r'const in0_type in0 = static_cast<in0_type>\((\w+)\);', before_mul)[-1]
    assert last_in0.startswith('_t'), last_in0
    assert '_w0 = _t' in mul


def test_fused_and_user_reductions_use_the_structured_skeletons(dry):
    a, b = cp.empty((300, 1000), 'f'), cp.empty((300, 1000), 'f')
    f = cp.fuse(kernel_name='fused_ssd_probe')(lambda x, y: cp.sum((x - y) * (x - y), axis=1))
    r = f(a, b)
    assert r.shape == (300,) and dry[-1]['kind'] == 'jit_reduce' and 'rows' in dry[-1]['name']
    f0 = cp.fuse(kernel_name='fused_ssd_probe0')(lambda x, y: cp.sum((x - y) * (x - y), axis=0))
    f0(a, b)
    assert 'cols' in dry[-1]['name']
    # operand kinds: keepdims mean (1), weights along the reduced axis (2)
    from cupy_b200._core import _reduction
    m = cp.empty((300, 1), 'f')
    w = cp.empty((1, 1000), 'f')
    ssd = cp.ReductionKernel('T x, T m, T w', 'T z', '(x - m) * w', 'a + b', 'z = a', '0', 'kinds_probe')
    ssd(a, m, w, axis=1)
    assert 'rows' in dry[-1]['name']
    from cupy_b200._core._ndarray import ndarray as _nd
    lay = _reduction._classify(a.shape, a.strides, 4, (1,), (0,), False)
    ms = _nd((300, 1000), 'f', memptr=m.ptr, strides=(4, 0))
    ws = _nd((300, 1000), 'f', memptr=w.ptr, strides=(0, 4))
    assert _reduction._operand_kinds([a, ms, ws], a, lay, (1,), (0,)) == (0, 1, 2)
    assert _reduction._operand_kinds([a, a.T.copy().T], a, lay, (1,), (0,)) is None
    lay0 = _reduction._classify(a.shape, a.strides, 4, (0,), (1,), False)
    m0 = _nd((300, 1000), 'f', memptr=w.ptr, strides=(0, 4))          # keepdims mean over axis 0
    w0 = _nd((300, 1000), 'f', memptr=m.ptr, strides=(4, 0))          # weights along axis 0
    assert _reduction._operand_kinds([a, m0, w0], a, lay0, (0,), (1,)) == (0, 1, 2)


def test_axis_scan_views_the_array_in_place(dry):
    a = cp.empty((6, 50, 32), 'i')
    r = cp.cumsum(a, axis=1)
    assert r.dtype == np.int64 and r.shape == a.shape
    rec = [d for d in dry if d['kind'] == 'prebuilt_scan_axis'][-1]
    assert (rec['outer'], rec['n'], rec['inner']) == (6, 50, 32)
    assert not any(d['kind'] in ('jit_elementwise', 'prebuilt_ufunc') for d in dry)      # no transposing copies
    del dry[:]
    cp.cumsum(a.transpose(2, 0, 1), axis=0)                  # a non-dense view is staged once
    assert sum(d['kind'] in ('jit_elementwise', 'prebuilt_ufunc') for d in dry) == 1
    assert [d for d in dry if d['kind'] == 'prebuilt_scan_axis'][-1]['inner'] == 6 * 50


def test_prebuilt_kernels_serve_only_fixed_operand_kinds(dry):
    x = cp.empty((1 << 16,), 'f')
    cp.add(x, x)
    assert dry[-1]['kind'] == 'prebuilt_ufunc'
    cp.multiply(x, 2)                                        # by-value scalar -> NVRTC folds it
    assert dry[-1]['kind'] == 'jit_elementwise' and dry[-1]['variant'] == _lib.EW_FLAT
    pad = cp.empty((64, 512), 'f')[:, :300]
    cp.add(pad, pad)                                         # padded rows, all unit stride: prebuilt ROWWISE
    assert dry[-1]['kind'] == 'prebuilt_ufunc' and dry[-1]['variant'] == _lib.EW_ROWWISE
    cp.add(pad, cp.empty((64, 1), 'f'))                      # column broadcast: specialised
    assert dry[-1]['kind'] == 'jit_elementwise' and 'RowTiler' in dry[-1]['source']


def test_accelerator_switch_and_numpy_dispatch(dry):
    """CUPY_ACCELERATORS-style switch (cupy/_core/_accelerator.pyx) and NumPy's dispatch
    protocols on the array (cupy/_core/core.pyx:1969-2037)."""
    from cupy_b200._core import _accelerator
    assert cp.get_reduction_accelerators() == ['b200']
    x = cp.empty((64, 128), 'f')
    x.sum(axis=1)
    assert dry[-1]['kind'] == 'prebuilt_reduce'
    cp.set_reduction_accelerators(['generic'])
    try:
        x.sum(axis=1)
        assert dry[-1]['kind'] == 'jit_reduce' and 'generic' in dry[-1]['name']
        # var without the accelerated single pass is the reference's two passes (mean, then cupy_var_core)
        del dry[:]
        x.var(axis=1)
        assert [d['kind'] for d in dry] == ['jit_reduce', 'jit_reduce'] and 'var_core' in dry[-1]['name']
        cp.set_reduction_accelerators(['reference'])
        with pytest.raises(RuntimeError):                  # nothing registered: the package has no oracle
            x.sum(axis=1)
        seen = []
        _accelerator.register_reference_backend(lambda kind, name, a, **kw: seen.append((kind, name)) or None)
        x.sum(axis=1)                                       # backend declined (None): falls through to the engine
        assert seen == [('reduction', 'cupy_sum')]
        with pytest.raises(ValueError):
            cp.set_reduction_accelerators(['nope'])
        cp.set_reduction_accelerators('cub')                # the reference's names are accepted
        assert cp.get_reduction_accelerators() == ['b200']
    finally:
        _accelerator.register_reference_backend(None)
        cp.set_reduction_accelerators(['b200'])
    # numpy.multiply(x, 2) / numpy.sum(x, axis=0) on a device array reach the engine
    del dry[:]
    r = np.multiply(x, 2)
    assert isinstance(r, cp.ndarray) and r.shape == x.shape and len(dry) == 1
    r = np.sum(x, axis=0)
    assert isinstance(r, cp.ndarray) and r.shape == (128,)
    r = np.add.reduce(x, axis=1)
    assert isinstance(r, cp.ndarray) and r.shape == (64,)
    with pytest.raises(TypeError):
        np.add(x, np.ones((64, 128), 'f'))                  # no silent host-to-device conversion


def test_call_shape_memo_of_the_launcher(dry):
    """Second and later calls of one call shape take the memoised route: same launch records, fresh outputs,
    scalar bytes re-encoded, overlapping `out=` falls back to the copying path."""
    from cupy_b200._core import _kernel
    a, b = cp.empty((64, 48), 'f'), cp.empty((64, 48), 'f')
    r1 = cp.add(a, b)
    first = dict(dry[-1])
    r2 = cp.add(cp.empty((64, 48), 'f'), b)
    assert dry[-1] == first and r2.ptr != r1.ptr and r2.shape == (64, 48) and r2.dtype == np.float32
    memo = _kernel._thread_local.ufunc_memo[id(cp.add)]
    assert len(memo) >= 1
    assert any(e.ops[e.out_slot].data == r2.ptr for e in memo.values())     # the remembered operand block was patched
    # scalars: value re-encoded per call, -0.0 kept apart from 0.0, NEP 50 overflow still raised
    cp.multiply(a, 2.0)
    cp.multiply(a, -0.0)
    sm = _kernel._thread_local.ufunc_memo[id(cp.multiply)]
    cp.multiply(a, 0.0)
    seen_keys = {key[1] for v in sm.values() for slot in v.scalar_slots for key in slot[5]}
    assert seen_keys >= {(-0.0).hex(), (0.0).hex()}
    u8 = cp.empty((16,), np.uint8)
    u8 + 1
    with pytest.raises(OverflowError):
        u8 + 1000
    # a different alignment class or stride pattern is a different shape
    n0 = len(memo)
    cp.add(a[:, ::2], b[:, ::2])
    assert len(memo) == n0 + 1
    # reductions
    a.sum(axis=0)
    rec = dict(dry[-1])
    a2 = cp.empty((64, 48), 'f')
    out = a2.sum(axis=0)
    assert dry[-1] == rec and out.shape == (48,)


def test_ufunc_at_and_reduceat_kernels_compile_for_every_supported_dtype(dry):
    """ufunc.at / add.reduceat (cupy_b200/_core/_scatter.py): every dtype gate of the reference's
    _scatter_op_single (_routines_indexing.pyx:942-1018) has an atomic overload that NVRTC accepts."""
    import cupy_b200 as cp
    from cupy_b200._core import _scatter
    for op, dts in _scatter._SUPPORTED.items():
        for dt in dts:
            getattr(cp, _scatter._UFUNC_NAME[op]).at(cp.empty((100,), dt), cp.empty((30,), 'int64'), cp.empty((30,), dt))
    cp.add.at(cp.empty((10, 20, 3), 'int64'), (cp.empty((4, 1), 'int32'), cp.empty((1, 5), 'int64')),
              cp.empty((4, 5, 3), 'int64'))
    n_at = len(dry)
    assert n_at >= sum(len(v) for v in _scatter._SUPPORTED.values())
    r = cp.add.reduceat(cp.empty((37, 50), 'int32'), [0, 3, 5], axis=1)
    assert r.shape == (37, 3) and r.dtype == np.int64
    assert [d['kind'] for d in dry[n_at:]] == ['prebuilt_scan_axis', 'jit_elementwise']
    with pytest.raises(TypeError):
        cp.add.at(cp.empty((4,), 'int8'), cp.empty((1,), 'int64'), 1)
    with pytest.raises(NotImplementedError):
        cp.multiply.at(cp.empty((4,), 'float32'), cp.empty((1,), 'int64'), 1.0)
    with pytest.raises(NotImplementedError):
        cp.add.at(cp.empty((4, 4), 'float32'), (slice(None), cp.empty((1,), 'int64')), 1.0)
    with pytest.raises(IndexError):
        cp.add.reduceat(cp.empty((10,), 'int32'), [0, 10])


def test_asarray_keeps_zero_dimensional_host_arrays_zero_dimensional(dry):
    """Found by tests/test_fuzz_gpu.py: numpy.ascontiguousarray promotes 0-d to 1-d."""
    import cupy_b200 as cp
    a = cp.asarray(np.array(3))
    assert a.shape == () and a.sum(keepdims=True).shape == () and cp.multiply(a, a).shape == ()
    assert cp.asarray(np.float32(2.5)).shape == () and cp.asarray(7).shape == ()
    assert cp.asarray(np.ones((3, 4)).T).strides == (8, 32)         # order 'K' keeps the host layout


def test_fused_function_ignores_arguments_it_never_reads(dry):
    """Found by tests/test_fuzz_gpu.py: an unused (larger) argument must not shape the fused loop."""
    import cupy_b200 as cp
    x, y = cp.empty((64, 130), 'i'), cp.empty((130,), 'i')
    assert cp.fuse(lambda a, b: b * b)(x, y).shape == (130,)
    assert cp.fuse(lambda a, b: b)(x, y).shape == (130,)
    assert cp.fuse(lambda a, b: cp.sum(b * 2, axis=0))(x, y).shape == ()
    assert cp.fuse(lambda a, b, s: a + s)(x, y, 3).shape == (64, 130)

    def upd(a, b):
        a += 1
    assert cp.fuse(upd)(x, y) is None and dry[-1]['kind'] in ('jit_elementwise', 'prebuilt_elementwise')
