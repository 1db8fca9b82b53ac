"""GPU tier: seeded differential fuzz against NumPy over the whole hot path -- random shapes (0..4-d, with empty and
unit extents), dtypes, views (steps, negative steps, transposes, broadcasts), axes, keepdims, `out=` -- the way the
reference's `testing.numpy_cupy_*` decorators compare every routine (tests/cupy_tests/math_tests/test_sumprod.py,
core_tests/test_ndarray_reduction.py, core_tests/test_ufunc_methods.py), but generated rather than enumerated.
Integer / bool / index results are bit-exact; floating results within a bound scaled by the reduced magnitude."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_MORE = int(os.environ.get('B200_FUZZ_MORE', '1'))     # one-off deeper runs: multiplies the number of seeds
DTYPES = ['bool', 'int8', 'uint8', 'int16', 'int32', 'uint32', 'int64', 'uint64', 'float16', 'float32', 'float64']


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


def _rand_shape(rs, max_ndim=4, big=False):
    nd = rs.randint(0, max_ndim + 1)
    choices = [0, 1, 2, 3, 5, 8, 17, 33, 64, 100, 257] + ([1024, 4099] if big else [])
    shape = tuple(int(rs.choice(choices)) for _ in range(nd))
    while np.prod(shape, dtype=np.int64) > (1 << 22):
        shape = shape[:-1]
    return shape


def _rand_data(rs, shape, dt):
    dt = np.dtype(dt)
    if dt.kind == 'b':
        return np.asarray(rs.rand(*shape) < 0.5)
    if dt.kind == 'f':
        return np.asarray(rs.rand(*shape) * 4 - 2).astype(dt)
    if dt.kind == 'u':
        return np.asarray(rs.randint(0, 7, size=shape)).astype(dt)
    return np.asarray(rs.randint(-3, 4, size=shape)).astype(dt)


def _rand_view(rs, host, dev):
    """The same random view (slices with steps, a transpose, maybe a broadcast axis) of a host array and its device twin."""
    for _ in range(rs.randint(0, 3)):
        if host.ndim == 0:
            break
        kind = rs.randint(0, 3)
        if kind == 0:
            key = []
            for n in host.shape:
                step = int(rs.choice([1, 1, 2, 3, -1, -2]))
                lo = int(rs.randint(0, max(n // 2, 1)))
                key.append(slice(None, None, step) if rs.rand() < 0.5 else
                           (slice(lo, None, step) if step > 0 else slice(None, lo if lo else None, step)))
            key = tuple(key)
            host, dev = host[key], dev[key]
        elif kind == 1 and host.ndim >= 2:
            perm = tuple(int(p) for p in rs.permutation(host.ndim))
            host, dev = host.transpose(perm), dev.transpose(perm)
        elif kind == 2 and host.ndim >= 1:
            ax = int(rs.randint(0, host.ndim))
            key = tuple(slice(0, 1) if d == ax else slice(None) for d in range(host.ndim))
            if host.shape[ax] >= 1:
                host, dev = host[key], dev[key]
                shape = tuple(4 if d == ax else n for d, n in enumerate(host.shape))
                host, dev = np.broadcast_to(host, shape), dev.broadcast_to(shape)
    return host, dev


def _tol(dt, terms):
    dt = np.dtype(dt)
    eps = {2: 1e-3, 4: 1.2e-7, 8: 2.3e-16}[dt.itemsize]
    return eps * max(terms, 1) * 8


def _check(got, want, terms=1, scale=2.0, what=''):
    got = got.get() if hasattr(got, 'get') else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, '%s: shape %s vs %s' % (what, got.shape, want.shape)
    assert got.dtype == want.dtype, '%s: dtype %s vs %s' % (what, got.dtype, want.dtype)
    if want.dtype.kind == 'f':
        np.testing.assert_allclose(got.astype('f8'), want.astype('f8'), rtol=_tol(want.dtype, 1),
                                   atol=_tol(want.dtype, terms) * scale, err_msg=what, equal_nan=True)
    else:
        np.testing.assert_array_equal(got, want, err_msg=what)


@pytest.mark.parametrize('seed', range(6 * _MORE))
def test_fuzz_reductions(cp, seed):
    rs = np.random.RandomState(1000 + seed)
    for case in range(60):
        dt = DTYPES[rs.randint(len(DTYPES))]
        base = _rand_data(rs, _rand_shape(rs, big=case % 5 == 0), dt)
        h, d = _rand_view(rs, base, cp.asarray(base))
        if h.ndim == 0:
            axis = None
        else:
            k = rs.randint(0, h.ndim + 1)
            axis = None if k == 0 else tuple(sorted(int(a) for a in rs.choice(h.ndim, size=k, replace=False)))
            if axis is not None and len(axis) == 1 and rs.rand() < 0.5:
                axis = axis[0] - (h.ndim if rs.rand() < 0.3 else 0)
        keep = bool(rs.rand() < 0.3)
        red = h.shape if axis is None else [h.shape[a] for a in (axis if isinstance(axis, tuple) else (axis,))]
        terms = int(np.prod(red, dtype=np.int64))
        what = 'seed %d case %d: %s %s strides %s axis=%r keepdims=%r' % (seed, case, dt, h.shape, h.strides, axis, keep)
        for name in ('sum', 'prod', 'max', 'min', 'mean', 'var', 'any', 'all', 'argmax', 'argmin'):
            w = what + ' ' + name
            kw = {'axis': axis, 'keepdims': keep}
            if name in ('argmax', 'argmin'):
                if isinstance(axis, tuple):
                    continue
                if h.dtype.kind == 'f' and terms and rs.rand() < 0.3:
                    pass
            if name in ('max', 'min', 'argmax', 'argmin') and terms == 0:
                # no identity: ValueError -- except that the reference returns an empty result first when the
                # OUTPUT is empty too (cupy/_core/_reduction.pyx:349-354), where NumPy still raises
                try:
                    want = getattr(h, name)(**kw)
                except ValueError:
                    want = None
                if want is None and 0 in [n for i, n in enumerate(h.shape) if not (
                        axis is None or i in [a % h.ndim for a in (axis if isinstance(axis, tuple) else (axis,))])]:
                    assert getattr(d, name)(**kw).size == 0, w
                elif want is None:
                    with pytest.raises(ValueError):
                        getattr(d, name)(**kw)
                else:
                    _check(getattr(d, name)(**kw), want, 1, 0.0, w)
                continue
            if name in ('mean', 'var') and (terms == 0 or h.size == 0):
                continue                                   # NumPy warns and returns NaN; covered by the enumerated tests
            if name == 'prod' and h.dtype.kind == 'f' and terms > 64:
                continue                                   # products of many U[-2,2) values under/overflow: not a parity case
            want = getattr(h, name)(**kw)
            if name in ('var', 'mean') and h.dtype == np.float16:
                # NumPy's float16 intermediates overflow past 65504; the reference accumulates float16 in float
                # (cupy/_core/_routines_statistics.pyx:611-616, 647-655), as the kernels here do
                want = np.asarray(getattr(h.astype(np.float32), name)(**kw)).astype(np.float16)
                if not keep and want.ndim == 0:
                    want = want[()]
            got = getattr(d, name)(**kw)
            scale = 2.0 if name != 'prod' else float(2.0 ** min(terms, 64))
            if name == 'var':
                scale = 8.0
            _check(got, want, terms, scale, w)


@pytest.mark.parametrize('seed', range(4 * _MORE))
def test_fuzz_scans(cp, seed):
    rs = np.random.RandomState(2000 + seed)
    for case in range(60):
        dt = DTYPES[rs.randint(len(DTYPES))]
        base = _rand_data(rs, _rand_shape(rs, big=case % 4 == 0), dt)
        h, d = _rand_view(rs, base, cp.asarray(base))
        axis = None if (h.ndim == 0 or rs.rand() < 0.3) else int(rs.randint(-h.ndim, h.ndim))
        what = 'seed %d case %d: %s %s strides %s axis=%r' % (seed, case, dt, h.shape, h.strides, axis)
        n = h.size if axis is None else h.shape[axis]
        _check(cp.cumsum(d, axis=axis), np.cumsum(h, axis=axis), n, 2.0, what + ' cumsum')
        if h.dtype.kind != 'f' or n <= 32:
            small = h if h.dtype.kind != 'f' else h
            _check(cp.cumprod(d, axis=axis), np.cumprod(small, axis=axis), n, float(2.0 ** min(n, 32)), what + ' cumprod')
        if rs.rand() < 0.3 and h.dtype.kind in 'iu':
            odt = np.dtype(rs.choice(['int32', 'int64', 'float64']))
            _check(cp.cumsum(d, axis=axis, dtype=odt), np.cumsum(h, axis=axis, dtype=odt), n, 2.0, what + ' cumsum dtype=%s' % odt)
        if rs.rand() < 0.3:
            want = np.cumsum(h, axis=axis)
            out = cp.empty(want.shape, want.dtype)
            r = cp.cumsum(d, axis=axis, out=out)
            assert r is out
            _check(out, want, n, 2.0, what + ' cumsum out=')


_BINARY = ['add', 'subtract', 'multiply', 'maximum', 'minimum', 'greater', 'less_equal', 'equal', 'true_divide']


@pytest.mark.parametrize('seed', range(4 * _MORE))
def test_fuzz_elementwise(cp, seed):
    rs = np.random.RandomState(3000 + seed)
    for case in range(80):
        dta, dtb = DTYPES[rs.randint(len(DTYPES))], DTYPES[rs.randint(len(DTYPES))]
        shape = _rand_shape(rs, big=case % 6 == 0)
        a = _rand_data(rs, shape, dta)
        # b: same shape, a trailing-suffix shape (broadcast), or a python scalar
        pick = rs.randint(0, 4)
        if pick == 0 and shape:
            b = _rand_data(rs, shape[int(rs.randint(0, len(shape))):], dtb)
        elif pick == 1:
            b = None
        else:
            b = _rand_data(rs, shape, dtb)
        ha, da = _rand_view(rs, a, cp.asarray(a))
        if b is None:
            hb = db = [2, 3.5, True, -1][rs.randint(4)]
        else:
            hb, db = b, cp.asarray(b)
            if hb.shape == a.shape and ha.shape != a.shape:
                hb, db = _rand_data(rs, ha.shape, dtb), None
                db = cp.asarray(hb)
        name = _BINARY[rs.randint(len(_BINARY))]
        what = 'seed %d case %d: %s(%s %s strides %s, %s)' % (
            seed, case, name, dta, ha.shape, ha.strides, ('%s %s' % (dtb, hb.shape)) if b is not None else repr(hb))
        f_np, f_cp = getattr(np, name), getattr(cp, name)
        try:
            with np.errstate(all='ignore'):
                want = f_np(ha, hb)
        except (TypeError, OverflowError, ValueError) as e:        # bool subtract, out-of-range python int, bad broadcast
            if isinstance(e, OverflowError) and 0 in ha.shape:
                f_cp(da, db)          # the reference returns the empty result before it packs the scalar (_kernel.pyx:1385-1392)
                continue
            with pytest.raises(type(e)):
                f_cp(da, db)
            continue
        try:
            got = f_cp(da, db)
        except OverflowError:
            # NumPy 2 answers comparisons with an out-of-range Python int; the reference (PyArray_Pack of the scalar
            # into the loop's type, cupy/_core/_scalar.pyx:375-379) and this package raise for every ufunc
            assert b is None and name in ('greater', 'less_equal', 'equal') and ha.dtype.kind in 'ub' and hb == -1, what
            continue
        if name == 'true_divide':
            g, w = got.get(), np.asarray(want)
            assert g.dtype == w.dtype and g.shape == w.shape, what
            np.testing.assert_allclose(g.astype('f8'), w.astype('f8'), rtol=_tol(w.dtype, 1), err_msg=what, equal_nan=True)
        else:
            _check(got, want, 1, 0.0, what)                   # IEEE-exact ops: bit-for-bit
        # unary + copy / astype on the same view
        if rs.rand() < 0.4:
            _check(da.copy(), ha.copy(), 1, 0.0, what + ' copy')
            tdt = DTYPES[rs.randint(len(DTYPES))]
            if not (ha.dtype.kind == 'f' and np.dtype(tdt).kind in 'iub'):   # float -> int of negatives/NaN is UB in C
                with np.errstate(all='ignore'):
                    _check(da.astype(tdt), ha.astype(tdt), 1, 0.0, what + ' astype ' + tdt)
            if ha.dtype.kind != 'b':
                _check(cp.negative(da), np.negative(ha), 1, 0.0, what + ' negative')


def _rand_expr(rs, depth, n_in):
    """A random expression tree over inputs x0..x{n-1} and small constants, as source text valid for both modules."""
    if depth == 0 or rs.rand() < 0.2:
        return 'x%d' % rs.randint(n_in) if rs.rand() < 0.8 else str(int(rs.randint(1, 4)))
    kind = rs.randint(0, 8)
    a, b = _rand_expr(rs, depth - 1, n_in), _rand_expr(rs, depth - 1, n_in)
    if kind <= 2:
        return '(%s %s %s)' % (a, '+-*'[kind], b)
    if kind == 3:
        return 'xp.maximum(%s, %s)' % (a, b)
    if kind == 4:
        return 'xp.minimum(%s, %s)' % (a, b)
    if kind == 5:
        return 'xp.negative(%s)' % a if not a.isdigit() else a
    if kind == 6:
        return 'xp.absolute(%s)' % a if not a.isdigit() else a
    return 'xp.square(%s)' % a if not a.isdigit() else a


@pytest.mark.parametrize('seed', range(3 * _MORE))
def test_fuzz_fusion_expression_trees(cp, seed):
    """cupy_b200.fuse of random expression trees (optionally closed by a reduction) against the same Python function
    on NumPy arrays -- fusion_utils.check_fusion in the reference's fusion tests.  Integer trees are bit-exact;
    floating trees may contract a*b+c into one FMA inside the fused kernel, so they get a relative tolerance."""
    rs = np.random.RandomState(4000 + seed)
    for case in range(25):
        n_in = int(rs.randint(1, 4))
        dt = ['int32', 'int64', 'float32', 'float64', 'int16'][rs.randint(5)]
        shape = tuple(int(rs.choice([1, 3, 17, 64, 130])) for _ in range(rs.randint(1, 4)))
        expr = _rand_expr(rs, 4, n_in)
        if not re.search(r'x\d', expr):
            expr = '(x0 + %s)' % expr
        reduce_axis = None
        tail = rs.randint(0, 4)
        if tail == 0:
            reduce_axis = int(rs.randint(0, len(shape)))
            expr = 'xp.sum(%s, axis=%d)' % (expr, reduce_axis)
        elif tail == 1:
            expr = 'xp.sum(%s)' % expr
        src = 'def f(%s):\n    return %s\n' % (', '.join('x%d' % i for i in range(n_in)), expr)
        env_np, env_cp = {'xp': np}, {'xp': cp}
        exec(src, env_np)
        exec(src, env_cp)
        hs = []
        for i in range(n_in):
            sh = shape if (i == 0 or rs.rand() < 0.6) else shape[int(rs.randint(0, len(shape))):]
            hs.append(_rand_data(rs, sh, dt))
        ds = [cp.asarray(h) for h in hs]
        what = 'seed %d case %d: %s %s  %s' % (seed, case, dt, [h.shape for h in hs], expr)
        try:
            with np.errstate(all='ignore'):
                want = np.asarray(env_np['f'](*hs))
        except np.exceptions.AxisError:                         # the tree dropped the full-rank input: not a case
            continue
        fused = cp.fuse(kernel_name='fuzz_%d_%d' % (seed, case))(env_cp['f'])
        got = fused(*ds)
        plain = env_cp['f'](*ds)                                # the same function launch by launch
        g, p = got.get(), plain.get()
        assert g.shape == want.shape and g.dtype == want.dtype, what
        if want.dtype.kind == 'f':
            mag = float(np.abs(want.astype('f8')).max()) if want.size else 1.0
            np.testing.assert_allclose(g.astype('f8'), want.astype('f8'), rtol=1e-4, atol=1e-5 * max(mag, 1.0), err_msg=what)
            np.testing.assert_allclose(p.astype('f8'), want.astype('f8'), rtol=1e-4, atol=1e-5 * max(mag, 1.0), err_msg=what)
        else:
            np.testing.assert_array_equal(g, want, err_msg=what)
            np.testing.assert_array_equal(p, want, err_msg=what)
        got2 = fused(*ds)                                        # second call: memoised kernel, same answer
        np.testing.assert_array_equal(got2.get(), g, err_msg=what)


@pytest.mark.parametrize('seed', range(2 * _MORE))
def test_fuzz_user_kernels(cp, seed):
    """ElementwiseKernel / ReductionKernel with generic types over random views and broadcasts: an axpy-like
    elementwise body and an L1-distance reduction, against NumPy."""
    rs = np.random.RandomState(5000 + seed)
    ew = cp.ElementwiseKernel('T x, T y, T a', 'T z', 'z = a * x + y', 'fuzz_axpy')
    l1 = cp.ReductionKernel('T x, T y', 'T z', 'abs(x - y)', 'a + b', 'z = a', '0', 'fuzz_l1')
    for case in range(40):
        dt = ['int32', 'int64', 'float32', 'float64'][rs.randint(4)]
        shape = _rand_shape(rs, max_ndim=3, big=case % 5 == 0)
        a, b = _rand_data(rs, shape, dt), _rand_data(rs, shape, dt)
        ha, da = _rand_view(rs, a, cp.asarray(a))
        hb = _rand_data(rs, ha.shape, dt)
        if ha.ndim and ha.size and rs.rand() < 0.3:
            hb = hb[(0,) * int(rs.randint(1, ha.ndim + 1))]      # a trailing-suffix shape: broadcast
        db = cp.asarray(hb)
        what = 'seed %d case %d: %s %s strides %s with %s' % (seed, case, dt, ha.shape, ha.strides, hb.shape)
        s = np.dtype(dt).type(2)
        want = s * ha + hb
        got = ew(da, db, s)
        if np.dtype(dt).kind == 'f':
            np.testing.assert_allclose(got.get(), want, rtol=1e-6 if dt == 'float32' else 1e-14, atol=1e-6, err_msg=what)
        else:
            np.testing.assert_array_equal(got.get(), want, err_msg=what)
        if ha.ndim == 0 or ha.size == 0:
            continue
        axis = int(rs.randint(0, ha.ndim))
        hbb = np.broadcast_to(hb, ha.shape)
        want = np.abs(ha - hbb).sum(axis=axis).astype(dt)
        got = l1(da, db.broadcast_to(ha.shape) if db.shape != ha.shape else db, axis=axis)
        terms = ha.shape[axis]
        _check(got, want, terms, 4.0, what + ' l1 axis=%d' % axis)
