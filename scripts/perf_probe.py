"""Device-timed GB/s of the configs in BASELINE.json (development probe; bench.py is the contract)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import cupy_b200 as cp
from cupy_b200._core import _kernel

PEAK = 6546.9


def timeit(f, iters=20, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        f()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def report(name, nbytes, f, **kw):
    try:
        med, mn = timeit(f, **kw)
        print('%-44s %9.3f ms (min %8.3f)  %8.1f GB/s  %5.1f%% of measured peak' % (
            name, med, mn, nbytes / med / 1e6, 100 * nbytes / med / 1e6 / PEAK), flush=True)
    except Exception as e:
        print('%-44s ERROR %s: %s' % (name, type(e).__name__, str(e)[:300]), flush=True)


def randn(shape, dtype):
    t = torch.rand(shape, device='cuda', dtype=torch.float32) * 2 - 1
    return t.to(dtype)


def main():
    which = sys.argv[1:] or ['copy', 'axpy', 'sum', 'c3', 'c4', 'scan', 'axis', 'multi', 'cast', 'rows']
    n = 1 << 28
    if 'copy' in which:
        a = torch.empty(n, device='cuda', dtype=torch.float32); b = torch.empty_like(a)
        report('torch copy_ f32 2^28 (R+W)', 8 * n, lambda: b.copy_(a))
        report('torch sum f32 2^28', 4 * n, lambda: a.sum())
        del a, b
    if 'axpy' in which:
        x = cp.from_torch(randn((n,), torch.float32)); y = cp.from_torch(randn((n,), torch.float32))
        z = cp.empty((n,), np.float32)
        k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')
        a = np.float32(1.5)
        report('axpy ElementwiseKernel 2^28 (JIT)', 12 * n, lambda: k(a, x, y, z))
        for unroll in (1, 2, 8):
            _kernel.tunables['flat_unroll'] = unroll
            report('  axpy unroll=%d' % unroll, 12 * n, lambda: k(a, x, y, z))
        _kernel.tunables['flat_unroll'] = 4
        for bps in (8, 16, 64, 4096):
            _kernel.tunables['blocks_per_sm'] = bps
            report('  axpy blocks/SM=%d' % bps, 12 * n, lambda: k(a, x, y, z))
        _kernel.tunables['blocks_per_sm'] = 0
        report('x*2 prebuilt 2^28', 8 * n, lambda: cp.multiply(x, 2, out=z))
        report('x+y prebuilt 2^28', 12 * n, lambda: cp.add(x, y, out=z))
        report('fma prebuilt 2^28', 16 * n, lambda: cp.fma(x, y, z, out=z))
        report('exp prebuilt 2^28', 8 * n, lambda: cp.exp(x, out=z))
        del x, y, z
    if 'sum' in which:
        x = cp.from_torch(randn((n,), torch.float32))
        report('sum f32 2^28 (full)', 4 * n, lambda: x.sum())
        report('max f32 2^28 (full)', 4 * n, lambda: x.max())
        report('argmax f32 2^28 (full)', 4 * n, lambda: x.argmax())
        report('var f32 2^28 (full)', 4 * n, lambda: x.var())
        del x
    if 'c3' in which:
        for dt, tdt in ((np.float32, torch.float32), (np.float16, torch.float16)):
            m = 32768
            x = cp.from_torch(randn((m, m), tdt))
            nb = m * m * np.dtype(dt).itemsize
            for op in ('sum', 'max', 'argmax', 'var'):
                for ax in (0, 1):
                    report('%s axis=%d %s 32768^2' % (op, ax, np.dtype(dt).name), nb,
                           lambda: getattr(x, op)(axis=ax), iters=10)
            del x
            torch.cuda.empty_cache()
    if 'c4' in which:
        base = cp.from_torch(randn((256, 1024, 1024), torch.float32))
        xt = base.transpose(2, 1, 0)       # shape (1024,1024,256), strides (4, 4096, 4194304)
        v = cp.from_torch(randn((256,), torch.float32))
        out = cp.empty((1024, 1024, 256), np.float32)
        tmp = cp.empty((1024, 1024, 256), np.float32)
        nel = 1 << 28
        report('exp(x^T) transposed -> contiguous', 8 * nel, lambda: cp.exp(xt, out=tmp), iters=10)
        report('tmp + v (broadcast row)', 8 * nel, lambda: cp.add(tmp, v, out=out), iters=10)
        fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'expadd')
        report('fused exp(x^T)+v ElementwiseKernel', 8 * nel, lambda: fused(xt, v, out), iters=10)
        report('copy transposed (ascontiguous)', 8 * nel, lambda: cp.elementwise_copy(xt, out), iters=10)
        ff = cp.fuse(kernel_name='fuse_expadd')(lambda x, v: cp.exp(x) + v)
        report('cupy_b200.fuse: exp(x^T)+v', 8 * nel, lambda: ff(xt, v), iters=10)
        fr = cp.fuse(kernel_name='fuse_sqsum')(lambda x, y: cp.sum((x - y) * (x - y), axis=2))
        report('cupy_b200.fuse: sum((x-tmp)^2, axis=2) contiguous', 8 * nel, lambda: fr(out, tmp), iters=10)
        f1 = cp.fuse(kernel_name='fuse_sumsq')(lambda x: cp.sum(x * x))
        report('cupy_b200.fuse: sum(x*x) one input, full', 4 * nel, lambda: f1(out), iters=10)
        f2 = cp.fuse(kernel_name='fuse_sumsq_ax')(lambda x: cp.sum(x * x, axis=2))
        report('cupy_b200.fuse: sum(x*x, axis=2) one input, rows', 4 * nel, lambda: f2(out), iters=10)
        fa = cp.fuse(kernel_name='fuse_x2p1')(lambda x: x * 2 + 1)
        report('cupy_b200.fuse: x*2+1 (config 1 chain, 2^28)', 8 * nel, lambda: fa(out), iters=10)
        k1 = cp.ElementwiseKernel('T x', 'T z', 'z = exp(x)', 'jit_exp')
        k2 = cp.ElementwiseKernel('T x, T v', 'T z', 'z = x + v', 'jit_addv')
        k3 = cp.ElementwiseKernel('T x, T v', 'T z', 'z = __expf(x) + v', 'jit_fastexp_addv')
        for un, mb in ((0, 0), (1, 4), (1, 5), (2, 3)):
            _kernel.tunables['reg_min_blocks'] = mb
            _kernel.tunables['reg_unroll'] = un
            report('  JIT exp(x^T) unroll=%d min_blocks=%d' % (un, mb), 8 * nel, lambda: k1(xt, out), iters=10)
            report('  JIT x^T + v unroll=%d min_blocks=%d' % (un, mb), 8 * nel, lambda: k2(xt, v, out), iters=10)
            report('  JIT __expf(x^T) + v unroll=%d min_blocks=%d' % (un, mb), 8 * nel, lambda: k3(xt, v, out), iters=10)
            report('  JIT exp(x^T) + v unroll=%d min_blocks=%d' % (un, mb), 8 * nel, lambda: fused(xt, v, out), iters=10)
        _kernel.tunables['reg_min_blocks'] = 0
        _kernel.tunables['reg_unroll'] = 0
        import os
        os.environ['B200_EW_TILED_MODE'] = 'reg'
        for mode in ():
            os.environ['B200_EW_TILED_MODE'] = mode
            report('  [%s] exp(x^T)' % mode, 8 * nel, lambda: cp.exp(xt, out=tmp), iters=10)
            report('  [%s] fused exp(x^T)+v' % mode, 8 * nel, lambda: fused(xt, v, out), iters=10)
            report('  [%s] copy transposed' % mode, 8 * nel, lambda: cp.elementwise_copy(xt, out), iters=10)
        del os.environ['B200_EW_TILED_MODE']
        if 'sweep2' in which:
            for bps in (4, 8, 16, 32, 64):
                _kernel.tunables['blocks_per_sm'] = bps
                report('  [reg] fused blocks/SM=%d' % bps, 8 * nel, lambda: fused(xt, v, out), iters=10)
            _kernel.tunables['blocks_per_sm'] = 0
            for un in (1, 2, 4):
                for mb in (1, 2, 3, 4):
                    _kernel.tunables['reg_min_blocks'] = mb
                    _kernel.tunables['reg_unroll'] = un
                    report('  [reg] fused unroll=%d min_blocks=%d' % (un, mb), 8 * nel, lambda: fused(xt, v, out), iters=10)
            _kernel.tunables['reg_min_blocks'] = 0
            _kernel.tunables['reg_unroll'] = 0
        h = cp.from_torch(randn((256, 1024, 1024), torch.float16)).transpose(2, 1, 0)
        ho = cp.empty((1024, 1024, 256), np.float16)
        report('copy transposed float16', 4 * nel, lambda: cp.elementwise_copy(h, ho), iters=10)
        del h, ho
        sq = cp.from_torch(randn((16384, 16384), torch.float32))
        sqo = cp.empty((16384, 16384), np.float32)
        report('2-D transpose 16384^2 f32', 8 * (1 << 28), lambda: cp.elementwise_copy(sq.T, sqo), iters=10)
        report('2-D x^T + y 16384^2 f32', 12 * (1 << 28), lambda: cp.add(sq.T, sqo, out=sqo), iters=10)
        for un in (1, 2):
            for mb in (1, 2, 3, 4):
                _kernel.tunables['reg_min_blocks'] = mb
                _kernel.tunables['reg_unroll'] = un
                report('  2-D x^T + y unroll=%d min_blocks=%d' % (un, mb), 12 * (1 << 28), lambda: cp.add(sq.T, sqo, out=sqo), iters=10)
        _kernel.tunables['reg_min_blocks'] = 0
        _kernel.tunables['reg_unroll'] = 0
        del sq, sqo
        del base, xt, out, tmp
        torch.cuda.empty_cache()
    if 'cast' in which:
        x32 = cp.from_torch(randn((n,), torch.float32))
        x16 = cp.from_torch(randn((n,), torch.float16))
        xi32 = cp.from_torch(torch.randint(-100, 100, (n,), device='cuda', dtype=torch.int32))
        report('astype f32 -> f16 2^28', 6 * n, lambda: x32.astype(np.float16), iters=10)
        report('astype f16 -> f32 2^28', 6 * n, lambda: x16.astype(np.float32), iters=10)
        report('astype i32 -> f64 2^28', 12 * n, lambda: xi32.astype(np.float64), iters=10)
        report('astype f32 -> i8 2^28', 5 * n, lambda: x32.astype(np.int8), iters=10)
        m2 = x32.reshape(16384, 16384)
        report('ascontiguous of [:, ::2] view f32', 4 * n, lambda: m2[:, ::2].copy(), iters=10)
        report('copy into strided out[:, ::2]', 4 * n, lambda: cp.elementwise_copy(m2[:, :8192], m2[:, ::2]), iters=10)
        report('x.T.astype(f16) 16384^2 (transposed + cast)', 6 * n, lambda: m2.T.astype(np.float16), iters=10)
        del x32, x16, xi32, m2
    if 'rows' in which:
        xs = cp.from_torch(randn((n,), torch.float32))
        for cols in (16, 64, 128, 256, 512, 1024, 4096):
            v2 = xs.reshape(n // cols, cols)
            report('sum axis=1 f32 rows of %d' % cols, 4 * n, lambda: v2.sum(axis=1), iters=10)
        v2 = xs.reshape(n // 256, 256)
        report('max axis=1 f32 rows of 256', 4 * n, lambda: v2.max(axis=1), iters=10)
        report('argmax axis=1 f32 rows of 256', 4 * n, lambda: v2.argmax(axis=1), iters=10)
        report('var axis=1 f32 rows of 256', 4 * n, lambda: v2.var(axis=1), iters=10)
        del xs, v2
    if 'multi' in which:
        m = 16384
        xa = cp.from_torch(randn((m, m), torch.float32)); xb = cp.from_torch(randn((m, m), torch.float32))
        dot = cp.ReductionKernel('T x, T y', 'T z', 'x * y', 'a + b', 'z = a', '0', 'dot')
        for ax in (None, 1, 0):
            report('ReductionKernel dot(x,y) axis=%s f32 16384^2' % ax, 8 * m * m, lambda: dot(xa, xb, axis=ax), iters=10)
        fr = cp.fuse(kernel_name='fuse_sqdiff')(lambda x, y: cp.sum((x - y) * (x - y), axis=1))
        report('cupy_b200.fuse sum((x-y)^2, axis=1) 16384^2', 8 * m * m, lambda: fr(xa, xb), iters=10)
        ssd = cp.ReductionKernel('T x, T m', 'T z', '(x - m) * (x - m)', 'a + b', 'z = a', '0', 'sum_sq_dev')
        for ax in (1, 0):
            mk = xa.mean(axis=ax, keepdims=True)
            report('ReductionKernel sum((x-mean)^2) axis=%d (broadcast mean) 16384^2' % ax, 4 * m * m,
                   lambda: ssd(xa, mk, axis=ax), iters=10)
            report('var(axis=%d, dtype=float32 given: two-pass reference algorithm)' % ax, 8 * m * m,
                   lambda: xa.var(axis=ax, dtype=np.float32), iters=10)
        report('torch (x*y).sum(1) 16384^2 (2 kernels)', 8 * m * m, lambda: (xa.to_torch() * xb.to_torch()).sum(1), iters=5)
        del xa, xb
    if 'axis' in which:
        m = 16384
        xf = cp.from_torch(randn((m, m), torch.float32)); yf = cp.empty((m, m), np.float32)
        report('cumsum axis=1 f32 16384^2', 8 * m * m, lambda: cp.cumsum(xf, axis=1, out=yf), iters=10)
        report('cumsum axis=0 f32 16384^2', 8 * m * m, lambda: cp.cumsum(xf, axis=0, out=yf), iters=10)
        x3 = xf.reshape(1024, 1024, 256); y3 = yf.reshape(1024, 1024, 256)
        report('cumsum axis=1 f32 1024x1024x256', 8 * m * m, lambda: cp.cumsum(x3, axis=1, out=y3), iters=10)
        report('cumsum axis=2 f32 1024x1024x256', 8 * m * m, lambda: cp.cumsum(x3, axis=2, out=y3), iters=10)
        xi = cp.from_torch(torch.randint(-100, 100, (m, m), device='cuda', dtype=torch.int32)); yi = cp.empty((m, m), np.int64)
        report('cumsum axis=1 int32->int64 16384^2', 12 * m * m, lambda: cp.cumsum(xi, axis=1, out=yi), iters=10)
        report('torch cumsum dim=1 f32 16384^2', 8 * m * m, lambda: torch.cumsum(xf.to_torch(), 1), iters=5)
        report('torch cumsum dim=0 f32 16384^2', 8 * m * m, lambda: torch.cumsum(xf.to_torch(), 0), iters=5)
        del xf, yf, xi, yi
    if 'scan' in which:
        xi = cp.from_torch(torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', dtype=torch.int64))
        yo = cp.empty((n,), np.int64)
        report('cumsum int64 2^28', 16 * n, lambda: cp.cumsum(xi, out=yo), iters=10)
        xf = cp.from_torch(randn((n,), torch.float32)); yf = cp.empty((n,), np.float32)
        report('cumsum f32 2^28', 8 * n, lambda: cp.cumsum(xf, out=yf), iters=10)
        report('torch cumsum int64 2^28', 16 * n, lambda: torch.cumsum(xi.to_torch(), 0), iters=5)
        x32 = cp.from_torch(torch.randint(-100, 100, (n,), device='cuda', dtype=torch.int32))
        report('cumsum int32 -> int64 2^28 (casting, flat)', 12 * n, lambda: cp.cumsum(x32, out=yo), iters=10)
        xh = cp.from_torch((torch.rand(n, device='cuda') / 1024).to(torch.float16)); yh = cp.empty((n,), np.float16)
        report('cumsum float16 2^28 (float accumulator, flat)', 4 * n, lambda: cp.cumsum(xh, out=yh), iters=10)
        xb = cp.from_torch(torch.rand(n, device='cuda') > 0.5)
        report('cumsum bool -> int64 2^28', 9 * n, lambda: cp.cumsum(xb, out=yo), iters=10)


if __name__ == '__main__':
    main()
