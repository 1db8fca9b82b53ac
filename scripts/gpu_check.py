"""Quick end-to-end check of every kernel family against NumPy on a real GPU.
Prints one line per case; used while developing (the real parity tests are tests/)."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, '.')
import torch
import cupy_b200 as cp

rs = np.random.RandomState(0)
fails = 0


def check(title, got, want, rtol=0, atol=0, exact=False):
    global fails
    try:
        g = got.get() if isinstance(got, cp.ndarray) else np.asarray(got)
        w = np.asarray(want)
        assert g.shape == w.shape, (g.shape, w.shape)
        assert g.dtype == w.dtype, (g.dtype, w.dtype)
        if exact:
            np.testing.assert_array_equal(g, w)
        else:
            np.testing.assert_allclose(g, w, rtol=rtol, atol=atol)
        print('PASS', title)
    except Exception as e:
        fails += 1
        print('FAIL', title, type(e).__name__, str(e)[:600].replace('\n', ' | '))


def run(title, f):
    global fails
    try:
        f()
    except Exception as e:
        fails += 1
        print('ERR ', title, type(e).__name__, str(e)[:1200])
        traceback.print_exc(limit=4)


def rnd(*shape, dt=np.float32):
    if np.dtype(dt).kind == 'f':
        return (rs.rand(*shape) * 2 - 1).astype(dt)
    if np.dtype(dt).kind == 'b':
        return rs.rand(*shape) > 0.5
    return rs.randint(-100, 100, size=shape).astype(dt)


def t_elementwise():
    a = rnd(1000, 1000); b = rnd(1000, 1000)
    da, db = cp.asarray(a), cp.asarray(b)
    check('add f32 flat', da + db, a + b, exact=True)
    check('mul scalar', da * 2, a * 2, exact=True)
    check('x*2+1', da * 2 + 1, a * 2 + 1, exact=True)
    check('sub', da - db, a - b, exact=True)
    check('div', da / db, a / b, rtol=1e-6)
    check('exp', cp.exp(da), np.exp(a), rtol=3e-7)
    check('neg', -da, -a, exact=True)
    check('maximum', cp.maximum(da, db), np.maximum(a, b), exact=True)
    v = rnd(1000); dv = cp.asarray(v)
    check('bcast row', da + dv, a + v, exact=True)
    c = rnd(1000, 1); dc = cp.asarray(c)
    check('bcast col', da + dc, a + c, exact=True)
    check('transposed in', cp.exp(da.T), np.exp(a.T), rtol=3e-7)
    check('transposed add', da.T + db, a.T + b, exact=True)
    x3 = rnd(64, 48, 40); d3 = cp.asarray(x3)
    check('3d transpose(2,1,0)', cp.exp(d3.transpose(2, 1, 0)), np.exp(x3.transpose(2, 1, 0)), rtol=3e-7)
    check('3d transpose(1,0,2)', d3.transpose(1, 0, 2) * 3, x3.transpose(1, 0, 2) * 3, exact=True)
    check('strided view', da[::2, 1::3] + 1, a[::2, 1::3] + 1, exact=True)
    i8a, i8b = rnd(333, dt=np.int8), rnd(333, dt=np.int8)
    check('add int8 (jit)', cp.asarray(i8a) + cp.asarray(i8b), i8a + i8b, exact=True)
    ia = rnd(100, 100, dt=np.int32)
    check('mixed i32+f32 (jit)', cp.asarray(ia) + da[:100, :100], ia + a[:100, :100], exact=True)
    h = rnd(513, 257, dt=np.float16); dh = cp.asarray(h)
    check('f16 add', dh + dh, h + h, exact=True)
    check('f16 exp', cp.exp(dh), np.exp(h), rtol=1e-3)
    check('astype f32->f16', da.astype(np.float16), a.astype(np.float16), exact=True)
    check('astype f32->i32', (da * 50).astype(np.int32), (a * 50).astype(np.int32), exact=True)
    out = cp.zeros((1000, 1000), np.float32)
    m = rnd(1000, 1000, dt=np.bool_)
    cp.add(da, db, out=out, _where=cp.asarray(m))
    check('where', out, np.where(m, a + b, 0).astype(np.float32), exact=True)
    k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')
    z = k(np.float32(1.5), da, db)
    want = (np.float64(1.5) * a.astype(np.float64) + b.astype(np.float64)).astype(np.float32)
    check('axpy user kernel (<=1ulp vs fp64)', z, want, rtol=1.2e-7)
    rev = cp.ElementwiseKernel('raw T x, int32 n', 'T y', 'y = x[n - 1 - i]', 'rev')
    check('raw arg', rev(dv, np.int32(1000), size=1000), v[::-1], exact=True)
    fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'expadd')
    check('fused transposed+bcast', fused(d3.transpose(2, 1, 0), cp.asarray(x3[:, 0, 0].copy())),
          np.exp(x3.transpose(2, 1, 0)) + x3[:, 0, 0], rtol=3e-7)
    check('arange', cp.arange(1000), np.arange(1000), exact=True)
    check('inplace iadd', _iadd(da.copy(), db), a + b, exact=True)


def _iadd(x, y):
    x += y
    return x


def t_reduce():
    for dt, rtol in ((np.float32, 2e-6), (np.float16, 2e-3), (np.float64, 1e-12), (np.int32, 0), (np.int64, 0),
                     (np.int8, 0), (np.uint8, 0), (np.bool_, 0), (np.int16, 0)):
        a = rnd(517, 1031, dt=dt) if dt != np.uint8 else rs.randint(0, 255, (517, 1031)).astype(np.uint8)
        d = cp.asarray(a)
        name = np.dtype(dt).name
        ex = rtol == 0
        acc = dict(dtype=np.float32) if dt == np.float16 else {}
        def want_sum(ax):
            w = a.sum(axis=ax, **acc)
            return w.astype(np.float16) if dt == np.float16 else w
        check('sum all ' + name, d.sum(), want_sum(None), rtol=rtol * 10, exact=ex)
        check('sum ax0 ' + name, d.sum(axis=0), want_sum(0), rtol=rtol * 10, atol=1e-4 if not ex else 0, exact=ex)
        check('sum ax1 ' + name, d.sum(axis=1), want_sum(1), rtol=rtol * 10, atol=1e-4 if not ex else 0, exact=ex)
        check('max ax0 ' + name, d.max(axis=0), a.max(axis=0), exact=True)
        check('max ax1 ' + name, d.max(axis=1), a.max(axis=1), exact=True)
        check('min all ' + name, d.min(), a.min(), exact=True)
        check('argmax ax0 ' + name, d.argmax(axis=0), a.argmax(axis=0), exact=True)
        check('argmax ax1 ' + name, d.argmax(axis=1), a.argmax(axis=1), exact=True)
        check('argmin all ' + name, d.argmin(), np.asarray(a.argmin()), exact=True)
        if dt != np.bool_:
            mw = a.mean(axis=1, dtype=np.float32).astype(np.float16) if dt == np.float16 else a.mean(axis=1)
            check('mean ax1 ' + name, d.mean(axis=1), mw, rtol=max(rtol, 1e-12) * 10, atol=1e-6)
            vw0 = a.astype(np.float64).var(axis=0)
            vw1 = a.astype(np.float64).var(axis=1)
            vdt = np.float16 if dt == np.float16 else (np.float32 if dt == np.float32 else np.float64)
            check('var ax0 ' + name, d.var(axis=0), vw0.astype(vdt), rtol=max(rtol, 1e-10) * 10)
            check('var ax1 ' + name, d.var(axis=1), vw1.astype(vdt), rtol=max(rtol, 1e-10) * 10)
            check('var all ddof1 ' + name, d.var(ddof=1), np.asarray(a.astype(np.float64).var(ddof=1)).astype(vdt), rtol=max(rtol, 1e-10) * 10)
    a = rnd(37, 53, 61); d = cp.asarray(a)
    check('3d sum ax1', d.sum(axis=1), a.sum(axis=1), rtol=1e-5, atol=1e-5)
    check('3d sum ax(0,2) generic', d.sum(axis=(0, 2)), a.sum(axis=(0, 2)), rtol=1e-5, atol=1e-5)
    check('3d sum ax(1,2)', d.sum(axis=(1, 2)), a.sum(axis=(1, 2)), rtol=1e-5, atol=1e-5)
    check('3d sum ax(0,1)', d.sum(axis=(0, 1)), a.sum(axis=(0, 1)), rtol=1e-5, atol=1e-5)
    check('3d argmax ax1', d.argmax(axis=1), a.argmax(axis=1), exact=True)
    check('3d keepdims', d.sum(axis=1, keepdims=True), a.sum(axis=1, keepdims=True), rtol=1e-5, atol=1e-5)
    check('noncontig sum', d[::2, :, 1::2].sum(axis=0), a[::2, :, 1::2].sum(axis=0), rtol=1e-5, atol=1e-5)
    check('transposed sum', d.T.sum(axis=0), a.T.sum(axis=0), rtol=1e-5, atol=1e-5)
    check('sum dtype=f64', d.sum(axis=0, dtype=np.float64), a.sum(axis=0, dtype=np.float64), rtol=1e-12)
    nan = a.copy(); nan[3, 5, 7] = np.nan; nan[3, 9, 7] = np.nan
    dn = cp.asarray(nan)
    check('max nan', dn.max(axis=1), nan.max(axis=1), exact=True)
    check('argmax nan', dn.argmax(axis=1), nan.argmax(axis=1), exact=True)
    check('argmax tie', cp.asarray(np.array([0, 5, 2, 3, 4, 5])).argmax(), np.asarray(1), exact=True)
    l2 = cp.ReductionKernel('T x', 'T y', 'x * x', 'a + b', 'y = sqrt(a)', '0', 'l2norm')
    check('ReductionKernel l2', l2(d, axis=1), np.sqrt((a * a).sum(axis=1)), rtol=1e-5)
    check('var ddof out=', d.var(axis=2, ddof=1, out=cp.empty((37, 53), np.float32)), a.var(axis=2, ddof=1), rtol=1e-4)
    check('std', d.std(axis=0), a.std(axis=0), rtol=1e-4)
    big = rnd(1 << 22); dbig = cp.asarray(big)
    check('sum 4M', dbig.sum(), np.asarray(big.sum(dtype=np.float64)).astype(np.float32), rtol=1e-5, atol=1e-3)
    check('tall cols 2^20x3', cp.asarray(big[:3 << 20].reshape(-1, 3)).sum(axis=0), big[:3 << 20].reshape(-1, 3).sum(axis=0, dtype=np.float64).astype(np.float32), rtol=1e-4, atol=1e-2)
    check('wide rows 3x2^20', cp.asarray(big[:3 << 20].reshape(3, -1)).sum(axis=1), big[:3 << 20].reshape(3, -1).sum(axis=1, dtype=np.float64).astype(np.float32), rtol=1e-4, atol=1e-2)


def t_scan():
    for dt in (np.int64, np.int32, np.int8, np.float32, np.float64, np.uint8, np.bool_, np.float16):
        for n in (1, 100, 4096, 4097, 100000, (1 << 20) + 3):
            a = rnd(n, dt=dt) if dt != np.uint8 else rs.randint(0, 255, n).astype(np.uint8)
            d = cp.asarray(a)
            w = np.cumsum(a)
            ex = np.dtype(dt).kind in 'iub'
            check('cumsum %s n=%d' % (np.dtype(dt).name, n), d.cumsum(), w, rtol=1e-3 if dt == np.float16 else 1e-4, atol=1e-2, exact=ex)
    a = np.ones(10000, np.int64); check('ones->arange', cp.asarray(a).cumsum(), np.arange(1, 10001), exact=True)
    a = rnd(50, 60, dt=np.int32); d = cp.asarray(a)
    check('cumsum axis0', d.cumsum(axis=0), a.cumsum(axis=0), exact=True)
    check('cumsum axis1', d.cumsum(axis=1), a.cumsum(axis=1), exact=True)
    check('cumsum 2d flat', d.cumsum(), a.cumsum(), exact=True)
    check('cumprod', cp.asarray(np.full(30, 2, np.int64)).cumprod(), np.full(30, 2, np.int64).cumprod(), exact=True)
    o = cp.empty((3000,), np.int64)
    check('cumsum out=', cp.cumsum(d, out=o), a.cumsum(), exact=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    for name, f in (('elementwise', t_elementwise), ('reduce', t_reduce), ('scan', t_scan)):
        t0 = time.time()
        run(name, f)
        torch.cuda.synchronize()
        print('---- %s done in %.1fs' % (name, time.time() - t0))
    print('FAILS', fails)
    sys.exit(1 if fails else 0)
