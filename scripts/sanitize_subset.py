"""Reduced-size pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
the TMA-pipelined flat scans (cooperative launch, slot ring), the ticket combines of full and split-COLS
reductions, rows / cols reductions, axis scans, elementwise tilers (FLAT / ROWWISE / TILED_REG) and the NVRTC
skeletons.  Results are checked against NumPy so a sanitizer-clean run is also a correct one.

    compute-sanitizer --tool racecheck python scripts/sanitize_subset.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402


def main():
    rs = np.random.RandomState(0)
    # ---- flat scans on the pipelined kernel (n >= 2^20), ragged sizes, casting pairs
    n = (1 << 20) + 4096 * 3 + 5
    xi = rs.randint(-1000, 1000, n).astype(np.int64)
    np.testing.assert_array_equal(cp.asarray(xi).cumsum().get(), np.cumsum(xi))
    x32 = xi.astype(np.int32)
    np.testing.assert_array_equal(cp.asarray(x32).cumsum().get(), np.cumsum(x32, dtype=np.int64))
    xb = (xi & 1).astype(np.bool_)
    np.testing.assert_array_equal(cp.asarray(xb).cumsum().get(), np.cumsum(xb))
    xf = (rs.rand(n) - 0.5).astype(np.float32)
    np.testing.assert_allclose(cp.asarray(xf).cumsum().get(), np.cumsum(xf.astype(np.float64)), atol=1e-2)
    xh = (rs.randint(-2, 3, n)).astype(np.float16)
    got = cp.asarray(xh).cumsum().get().astype(np.float64)
    np.testing.assert_allclose(got, np.cumsum(xh.astype(np.float64)).astype(np.float16).astype(np.float64), atol=1.0)
    np.testing.assert_array_equal(cp.cumprod(cp.asarray(np.ones(n, np.int64))).get(), np.ones(n, np.int64))
    # ---- the look-back scan (small n) and axis scans
    small = xi[:100003]
    np.testing.assert_array_equal(cp.asarray(small).cumsum().get(), np.cumsum(small))
    m = xi[:512 * 384].reshape(512, 384)
    for ax in (0, 1):
        np.testing.assert_array_equal(cp.asarray(m).cumsum(axis=ax).get(), np.cumsum(m, axis=ax))
    tall = xi[:40000 * 24].reshape(40000, 24)          # few columns: split along the scanned axis (totals workspace)
    np.testing.assert_array_equal(cp.asarray(tall).cumsum(axis=0).get(), np.cumsum(tall, axis=0))
    # ---- reductions: FULL (ticket combine), ROWS, COLS, split COLS, arg-reductions, var
    a = (rs.rand(1 << 20) * 2 - 1).astype(np.float32)
    da = cp.asarray(a)
    np.testing.assert_allclose(da.sum().get(), a.astype(np.float64).sum(), rtol=1e-5, atol=1e-3)
    assert int(da.argmax().get()) == int(a.argmax())
    np.testing.assert_allclose(da.var().get(), a.astype(np.float64).var(), rtol=1e-5)
    m2 = a.reshape(1024, 1024)
    dm = cp.asarray(m2)
    for ax in (0, 1):
        np.testing.assert_allclose(dm.sum(axis=ax).get(), m2.astype(np.float64).sum(axis=ax), rtol=1e-4, atol=1e-3)
        np.testing.assert_array_equal(dm.max(axis=ax).get(), m2.max(axis=ax))
        np.testing.assert_array_equal(dm.argmax(axis=ax).get(), m2.argmax(axis=ax))
        np.testing.assert_allclose(dm.var(axis=ax).get(), m2.astype(np.float64).var(axis=ax), rtol=1e-4)
    thin = a[:65536 * 8].reshape(65536, 8)             # split COLS: partials + tickets
    np.testing.assert_allclose(cp.asarray(thin).sum(axis=0).get(), thin.astype(np.float64).sum(axis=0), rtol=1e-4, atol=1e-2)
    h = a[:512 * 640].astype(np.float16).reshape(512, 640)
    np.testing.assert_allclose(cp.asarray(h).sum(axis=1).get().astype(np.float64), h.astype(np.float64).sum(axis=1), atol=0.5)
    # ---- elementwise: FLAT, ROWWISE (broadcast), TILED_REG (transpose), user kernels, fuse, user reduction
    b = (rs.rand(1 << 20)).astype(np.float32)
    db = cp.asarray(b)
    np.testing.assert_array_equal((da * 2 + 1).get(), a * 2 + 1)
    k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'san_axpy')
    z = k(np.float32(1.5), da, db).get()
    np.testing.assert_allclose(z, 1.5 * a + b, rtol=1e-6, atol=1e-6)
    t3 = cp.asarray(a[:64 * 96 * 32].reshape(32, 96, 64)).transpose(2, 1, 0)
    v = cp.asarray(b[:32])
    f = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'san_expadd')
    np.testing.assert_allclose(f(t3, v).get(), np.exp(a[:64 * 96 * 32].reshape(32, 96, 64).transpose(2, 1, 0)) + b[:32], rtol=1e-5)
    np.testing.assert_allclose((dm + cp.asarray(b[:1024])).get(), m2 + b[:1024], rtol=1e-6)
    ff = cp.fuse(kernel_name='san_fused')(lambda p, q: cp.sum((p - q) * (p - q), axis=1))
    np.testing.assert_allclose(ff(dm, cp.asarray(b.reshape(1024, 1024))).get(),
                               ((m2.astype(np.float64) - b.reshape(1024, 1024)) ** 2).sum(axis=1), rtol=1e-4)
    dot = cp.ReductionKernel('T p, T q', 'T r', 'p * q', 'a + b', 'r = a', '0', 'san_dot')
    np.testing.assert_allclose(dot(da, db).get(), (a.astype(np.float64) * b).sum(), rtol=1e-4)
    # ---- narrow / short-row kernels (flat-stream tiles): reductions and scans along both axes, ufunc.at
    for cols, dt in ((3, np.float32), (4, np.float32), (7, np.int32), (16, np.float16), (33, np.float64), (64, np.int8)):
        rows = 70000 // cols * 3 + 5
        p = (rs.rand(rows, cols) * 8 - 4).astype(dt)
        dp = cp.asarray(p)
        tol = 5e-2 if dt == np.float16 else 1e-4
        np.testing.assert_allclose(dp.sum(axis=0).get().astype(np.float64), p.astype(np.float64).sum(axis=0), rtol=tol, atol=rows * tol)
        np.testing.assert_array_equal(dp.argmax(axis=0).get(), p.argmax(axis=0))
        np.testing.assert_array_equal(dp.argmax(axis=1).get(), p.argmax(axis=1))
        np.testing.assert_array_equal(dp.max(axis=1).get(), p.max(axis=1))
        np.testing.assert_allclose(dp.var(axis=0).get().astype(np.float64), p.astype(np.float64).var(axis=0), rtol=1e-2)
        np.testing.assert_allclose(dp.var(axis=1).get().astype(np.float64), p.astype(np.float64).var(axis=1), rtol=1e-2, atol=1e-2)
        if np.dtype(dt).kind == 'i':
            np.testing.assert_array_equal(cp.cumsum(dp, axis=0).get(), np.cumsum(p, axis=0))
            np.testing.assert_array_equal(cp.cumsum(dp, axis=1).get(), np.cumsum(p, axis=1))
        else:
            np.testing.assert_allclose(cp.cumsum(dp, axis=0).get().astype(np.float64), np.cumsum(p.astype(np.float64), axis=0),
                                       rtol=tol, atol=rows * tol)
            np.testing.assert_allclose(cp.cumsum(dp, axis=1).get().astype(np.float64), np.cumsum(p.astype(np.float64), axis=1),
                                       rtol=tol, atol=cols * tol)
    acc = np.zeros(1000, np.int64)
    idx = rs.randint(0, 1000, 50000)
    dacc = cp.asarray(acc)
    cp.add.at(dacc, cp.asarray(idx), 1)
    np.add.at(acc, idx, 1)
    np.testing.assert_array_equal(dacc.get(), acc)
    torch.cuda.synchronize()
    print('sanitize subset ok')


if __name__ == '__main__':
    main()
