"""The `configs` table of bench.py: every BASELINE.json config (C1 .. C5) device-timed through the public
API, each entry with its algorithmic bytes, GB/s, fraction of the measured copy
peak, an in-bench parity check (float64 / bit-exact, computed with torch on the same buffers) and -- on
rank 0 at N = 1 -- the REFERENCE's own GPU kernels timed on the same buffers in the same run
(`ref_gpu`: oracle/_ref CUB path + the reference's rendered JIT templates with the reference's launch
geometry, see oracle/ref_gpu.py; baseline-only, like cpu_baseline).

Timing: >= 3 warm-up calls, then `iters` calls each bracketed by CUDA events on the launching (current)
stream; the median is reported.  Working sets are >= 1 GiB (>> 126 MB L2) except where an entry says
`l2_resident`.
"""
from __future__ import annotations

import os
import time

import numpy as np


def _median_ms(f, iters=10, warm=3):
    import torch
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        f()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return float(ts[len(ts) // 2]), float(ts[0])


def _graph_us(f, inner=20, replays=10):
    """Per-launch time of `f` inside a replayed CUDA graph: the kernel (plus launch gaps) with the per-call host cost
    taken out -- for the L2-resident / tiny configs whose per-call time is the host's."""
    import torch
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                f()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(inner):
                f()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / (inner * replays) * 1e3
    except Exception:            # capture is evidence beside the bench line, never a reason to lose it
        torch.cuda.synchronize()
        return None


def _with_graph(e, f, nbytes, peak):
    us = _graph_us(f)
    if us is not None:
        e['graph_replay_us'] = round(us, 2)
        if nbytes:
            e['graph_replay_gbs'] = round(nbytes / us / 1e3, 1)
            e['graph_replay_frac'] = round(nbytes / us / 1e3 / peak, 4)
    return e


def _entry(name, nbytes, ms, peak, check, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    e = {'name': name, 'ms': round(ms, 5), 'bytes': int(nbytes), 'gbs': round(gbs, 1), 'frac': round(gbs / peak, 4),
         'check': check}
    e.update(kw)
    return e


def _ref(e, nbytes, f, what, iters=10):
    """Attach the reference-GPU timing of the same config to entry `e`."""
    try:
        ms, _ = _median_ms(f, iters=iters)
        e['ref_gpu'] = {'ms': round(ms, 5), 'gbs': round(nbytes / (ms * 1e-3) / 1e9, 1), 'what': what}
        e['speedup_vs_ref_gpu'] = round(ms / e['ms'], 3)
    except Exception as ex:      # the baseline leg must never take the bench down
        e['ref_gpu'] = {'error': '%s: %s' % (type(ex).__name__, str(ex)[:200])}


def _ulp16(v):
    import torch
    return torch.from_numpy(np.spacing(np.abs(v.cpu().numpy()).astype(np.float16)).astype(np.float64)).to(v.device)


# ------------------------------------------------------------------------------------------------
# C1: the reference's CPU-runnable case: NumPy on one host core
# ------------------------------------------------------------------------------------------------
def c1_numpy(repeats=25):
    rs = np.random.RandomState(0)
    x = (rs.rand(4096, 4096) * 2 - 1).astype(np.float32)
    out = {}
    for name, f, nbytes in (('x*2+1', lambda: x * 2 + 1, 2 * 4 * 4096 * 4096),
                            ('x.sum(axis=1)', lambda: x.sum(axis=1), 4 * 4096 * 4096 + 4 * 4096)):
        f()
        ts = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            f()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        out[name] = {'best_ms': round(1e3 * ts[0], 3), 'median_ms': round(1e3 * ts[len(ts) // 2], 3),
                     'bytes': nbytes, 'gbs_best': round(nbytes / ts[0] / 1e9, 2), 'repeats': repeats}
    out['cores_used'] = 1
    out['host_cpu_count'] = os.cpu_count()
    out['numpy'] = np.__version__
    return out


# ------------------------------------------------------------------------------------------------
def run_configs(peak, with_ref_gpu=True, quick=False):
    import torch
    import cupy_b200 as cp
    refjit = refcub = None
    ref_note = None
    if with_ref_gpu:
        try:
            from oracle import ref_gpu
            if ref_gpu.available():
                refjit, refcub = ref_gpu.RefJit(), ref_gpu.RefCub()
            else:
                ref_note = 'oracle/_ref not built (needs the reference tree at build time)'
        except Exception as ex:
            ref_note = '%s: %s' % (type(ex).__name__, ex)
    if refjit is not None:
        from oracle.ref_gpu import carray, CUB_SUM, CUB_MAX, CUB_ARGMAX, CUB_CUMSUM
    import ctypes
    keep = []

    def alloc(nbytes):
        t = torch.empty(int(nbytes), dtype=torch.uint8, device='cuda')
        keep.append(t)
        return t.data_ptr()

    entries = []
    it = 5 if quick else 10
    g = torch.Generator(device='cuda')
    g.manual_seed(0)

    # ---------------- C1 on the GPU (64 MiB arrays: L2-resident, launch-bound) ------------------
    x1 = cp.from_torch(torch.rand(4096, 4096, device='cuda', generator=g) * 2 - 1)
    t1 = x1.to_torch()
    fz = cp.fuse(kernel_name='c1_x2p1')(lambda a: a * 2 + 1)
    ms, _ = _median_ms(lambda: fz(x1), iters=50)
    ok = bool(torch.equal(fz(x1).to_torch(), t1 * 2 + 1))
    entries.append(_with_graph(_entry('C1 x*2+1 f32 4096^2 (cupy_b200.fuse, one kernel)', 2 * 4 * 4096 * 4096, ms, peak,
                                      'bit-exact' if ok else 'MISMATCH', l2_resident=True),
                               lambda: fz(x1), 2 * 4 * 4096 * 4096, peak))
    ms, _ = _median_ms(lambda: x1 * 2 + 1, iters=50)
    entries.append(_entry('C1 x*2+1 f32 4096^2 (two ufunc launches; bytes = fused-equivalent)', 2 * 4 * 4096 * 4096, ms,
                          peak, 'bit-exact' if bool(torch.equal((x1 * 2 + 1).to_torch(), t1 * 2 + 1)) else 'MISMATCH',
                          l2_resident=True))
    ms, _ = _median_ms(lambda: x1.sum(axis=1), iters=50)
    err = float((x1.sum(axis=1).to_torch().double() - t1.double().sum(1)).abs().max())
    entries.append(_with_graph(_entry('C1 x.sum(axis=1) f32 4096^2', 4 * 4096 * 4096 + 4 * 4096, ms, peak,
                                      'ok max_abs_err=%.2e' % err if err < 1e-3 else 'MISMATCH %.3e' % err,
                                      l2_resident=True),
                               lambda: x1.sum(axis=1), 4 * 4096 * 4096 + 4 * 4096, peak))
    del x1, t1

    # ---------------- small-array floor (performance.rst:33-34: arange(1000).sum()) --------------
    xs = cp.arange(1000)
    for _ in range(20):
        xs.sum()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        xs.sum()
    host_us = (time.perf_counter() - t0) / 2000 * 1e6
    torch.cuda.synchronize()
    gpu_ms, _ = _median_ms(lambda: xs.sum(), iters=200)
    a_, b_ = cp.arange(1000, dtype=np.float32), cp.arange(1000, dtype=np.float32)
    for _ in range(20):
        cp.add(a_, b_)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        cp.add(a_, b_)
    add_us = (time.perf_counter() - t0) / 2000 * 1e6
    torch.cuda.synchronize()
    entries.append({'name': 'small-array floor: arange(1000).sum() / add(a,b) on 1000 f32',
                    'graph_replay_us_sum': (lambda v: None if v is None else round(v, 2))(_graph_us(lambda: xs.sum())),
                    'host_us_per_call_sum': round(host_us, 2), 'host_us_per_call_add': round(add_us, 2),
                    'gpu_us_per_call_sum': round(gpu_ms * 1e3, 2),
                    'check': 'ok' if int(xs.sum().get()) == 499500 else 'MISMATCH',
                    'reference_documented': 'cupy docs/source/user_guide/performance.rst:33-34,200-205: '
                                            '~20 us CPU-side, 53 us GPU for the same call (other hardware)'})

    # ---------------- C2 (the headline kernels again, for the ref_gpu columns) -------------------
    n = 1 << 28
    tx = torch.rand(n, device='cuda', generator=g) * 2 - 1
    ty = torch.rand(n, device='cuda', generator=g) * 2 - 1
    x, y = cp.from_torch(tx), cp.from_torch(ty)
    z = cp.empty((n,), np.float32)
    axpy = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')
    a = np.float32(1.5)
    ms, _ = _median_ms(lambda: axpy(a, x, y, z), iters=it)
    want = torch.addcmul(ty, tx, torch.tensor(1.5, device='cuda'))       # fused multiply-add: one rounding
    ok = bool(torch.equal(z.to_torch(), want))
    if not ok:   # torch.addcmul may not contract; fall back to the float64 statement of fma
        ok = bool(torch.equal(z.to_torch(), (tx.double() * 1.5 + ty.double()).float()))
    e = _entry('C2 axpy z=a*x+y f32 2^28 (ElementwiseKernel, NVRTC)', 12 * n, ms, peak,
               'bit-exact vs fma' if ok else 'MISMATCH')
    if refjit is not None:
        zr = torch.empty(n, device='cuda')
        sh, st = (n,), (4,)
        f = refjit.elementwise('ref_axpy', [ctypes.c_float(1.5), carray(tx.data_ptr(), sh, st),
                                            carray(ty.data_ptr(), sh, st), carray(zr.data_ptr(), sh, st),
                                            oracle_cindexer(sh)], n)
        _ref(e, 12 * n, f, 'reference ElementwiseKernel template (_kernel.pyx:86-97), linear_launch 128-thread blocks', it)
        e['ref_gpu']['same_bits'] = bool(torch.equal(zr, z.to_torch()))
        del zr
    entries.append(e)
    del want, y, ty, z

    s = cp.empty((), np.float32)
    ms, _ = _median_ms(lambda: x.sum(out=s), iters=it)
    want = float(tx.double().sum())
    e = _entry('C2 sum f32 2^28 (full)', 4 * n, ms, peak,
               'ok rel_err=%.1e' % (abs(float(s.get()) - want) / max(1, abs(want))))
    if refcub is not None:
        yr = torch.empty(1, device='cuda')
        _ref(e, 4 * n, refcub.reduce(tx.data_ptr(), yr.data_ptr(), n, CUB_SUM, 'float32', alloc),
             'reference cub_device_reduce (cupy_cub.cu, CCCL 3.1.2 DeviceReduce::Sum)', it)
    entries.append(e)
    for op, tfun in (('max', torch.amax), ('argmax', torch.argmax), ('var', None)):
        ms, _ = _median_ms(lambda: getattr(x, op)(), iters=it)
        got = getattr(x, op)().get()
        if op == 'var':
            want = float(tx.double().var(unbiased=False))
            chk = 'ok rel_err=%.1e' % (abs(float(got) - want) / want)
        elif op == 'max':
            chk = 'bit-exact' if float(got) == float(tx.max()) else 'MISMATCH'
        else:
            mx = tx.max()
            first = int(torch.nonzero(tx == mx)[0])
            chk = 'exact (first occurrence)' if int(got) == first else 'MISMATCH'
        e = _entry('full %s f32 2^28' % op, 4 * n, ms, peak, chk)
        if refcub is not None and op != 'var':
            yr = torch.empty(4, device='cuda')
            _ref(e, 4 * n, refcub.reduce(tx.data_ptr(), yr.data_ptr(), n, CUB_MAX if op == 'max' else CUB_ARGMAX,
                                         'float32', alloc), 'reference cub_device_reduce', it)
        entries.append(e)
    del x, tx, s
    torch.cuda.empty_cache()

    # ---------------- C3: axis reductions over 32768^2, fp32 and fp16 ----------------------------
    m = 32768
    for tdt, ndt, sfx in ((torch.float32, np.float32, 'f32'), (torch.float16, np.float16, 'f16')):
        t = torch.empty(m, m, device='cuda', dtype=tdt)
        for lo in range(0, m, 4096):
            t[lo:lo + 4096] = (torch.rand(4096, m, device='cuda', generator=g) * 2 - 1).to(tdt)
        x = cp.from_torch(t)
        isz = np.dtype(ndt).itemsize
        nb = m * m * isz
        # float64 references, one axis at a time, in row slabs (no 8 GiB temporaries)
        s64 = [torch.zeros(m, device='cuda', dtype=torch.float64) for _ in range(2)]
        q64 = [torch.zeros(m, device='cuda', dtype=torch.float64) for _ in range(2)]
        a64 = [torch.zeros(m, device='cuda', dtype=torch.float64) for _ in range(2)]      # sum |x|
        for lo in range(0, m, 2048):
            c = t[lo:lo + 2048].double()
            s64[0] += c.sum(0); q64[0] += (c * c).sum(0); a64[0] += c.abs().sum(0)
            s64[1][lo:lo + 2048] = c.sum(1); q64[1][lo:lo + 2048] = (c * c).sum(1); a64[1][lo:lo + 2048] = c.abs().sum(1)
        for op in ('sum', 'max', 'argmax', 'var'):
            for ax in (0, 1):
                ms, _ = _median_ms(lambda: getattr(x, op)(axis=ax), iters=it)
                got = getattr(x, op)(axis=ax).to_torch()
                if op == 'sum':
                    # a sum's error scales with sum|x| (the values cancel: |sum| ~ 100, sum|x| = 16384), which is what
                    # the reference's own CUB-vs-NumPy tests allow for (rtol on shaped_random, test_sumprod.py:213-296)
                    want = s64[ax]
                    if sfx == 'f32':
                        err = float(((got.double() - want).abs() / a64[ax]).max())
                        chk = ('ok' if err <= 1e-5 else 'MISMATCH') + ' max_err=%.1e of sum|x| (tol 1e-5)' % err
                    else:
                        # fp32 accumulate, one rounding to fp16 at the end: half an fp16 ulp + the fp32 accumulation error
                        over = float(((got.double() - want).abs() - 0.5 * _ulp16(want) - 1e-6 * a64[ax]).max())
                        chk = ('ok' if over <= 0 else 'MISMATCH') + ' |err| <= 0.5 fp16 ulp + 1e-6 sum|x| (fp32 accumulate): margin %.1e' % -over
                elif op == 'var':
                    mean = s64[ax] / m
                    want = q64[ax] / m - mean * mean
                    if sfx == 'f32':
                        err = float(((got.double() - want).abs() / want).max())
                        chk = ('ok' if err <= 1e-5 else 'MISMATCH') + ' max_rel_err=%.1e (tol 1e-5)' % err
                    else:
                        ulps = float(((got.double() - want).abs() / _ulp16(want)).max())
                        chk = ('ok' if ulps <= 1.0 else 'MISMATCH') + ' max_err=%.2f fp16 ulp (fp32 accumulate; tol 1)' % ulps
                elif op == 'max':
                    chk = 'bit-exact' if bool(torch.equal(got, torch.amax(t, dim=ax))) else 'MISMATCH'
                else:
                    mx = torch.amax(t, dim=ax, keepdim=True)
                    first = torch.full((m,), m, device='cuda', dtype=torch.int64)
                    ar = torch.arange(m, device='cuda')
                    for lo in range(0, m, 2048):          # first index holding the maximum, slab by slab
                        if ax == 0:
                            hit = t[lo:lo + 2048] == mx
                            cand = torch.where(hit, ar[lo:lo + 2048, None], m).amin(0)
                            first = torch.minimum(first, cand)
                        else:
                            hit = t[lo:lo + 2048] == mx[lo:lo + 2048]
                            first[lo:lo + 2048] = torch.where(hit, ar[None, :], m).amin(1)
                    chk = 'exact (first occurrence)' if bool(torch.equal(got, first)) else 'MISMATCH'
                    del first, mx
                e = _entry('C3 %s axis=%d %s 32768^2' % (op, ax, sfx), nb, ms, peak, chk)
                if refjit is not None:
                    try:
                        _c3_ref(e, refjit, refcub, t, op, ax, sfx, m, isz, nb, alloc, it)
                    except Exception as ex:
                        e['ref_gpu'] = {'error': '%s: %s' % (type(ex).__name__, str(ex)[:200])}
                entries.append(e)
        del x, t, s64, q64, a64
        torch.cuda.empty_cache()

    # ---------------- C4a: exp(x^T) + row vector, 1024 x 1024 x 256 f32 ----------------------------
    base = torch.rand(256, 1024, 1024, device='cuda', generator=g) * 2 - 1
    tv = torch.rand(256, device='cuda', generator=g) * 2 - 1
    xt = cp.from_torch(base).transpose(2, 1, 0)         # shape (1024,1024,256), strides (4, 4096, 4194304)
    v = cp.from_torch(tv)
    out = cp.empty((1024, 1024, 256), np.float32)
    tmp = cp.empty((1024, 1024, 256), np.float32)
    nel = 1 << 28
    nb4 = 8 * nel + 1024
    want = torch.exp(base.permute(2, 1, 0).double()) + tv.double()
    fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'expadd')

    ex64 = torch.exp(base.permute(2, 1, 0).double())
    big = torch.maximum(ex64, want.abs()).float()
    ulp = (torch.nextafter(big, torch.tensor(float('inf'), device='cuda')) - big).double()
    del ex64, big

    def ulp_err(o):
        # error budget: <= 2 ulp of exp(x) + the half-ulp rounding of the add, in ulps of max(exp(x), |result|)
        return float(((o.double() - want).abs() / ulp).max())

    ms, _ = _median_ms(lambda: fused(xt, v, out), iters=it)
    u = ulp_err(out.to_torch())
    e_fused = _entry('C4a fused exp(x^T)+v f32 1024x1024x256 (ElementwiseKernel, one kernel)', nb4, ms, peak,
                     ('ok' if u <= 2.5 else 'MISMATCH') + ' max_err=%.2f ulp (tol 2.5 = 2 ulp exp + add rounding)' % u)
    ff = cp.fuse(kernel_name='fuse_expadd')(lambda x_, v_: cp.exp(x_) + v_)
    ms, _ = _median_ms(lambda: ff(xt, v), iters=it)
    same = bool(torch.equal(ff(xt, v).to_torch(), out.to_torch()))
    e_fuse = _entry('C4a cupy_b200.fuse(exp(x^T)+v) (one kernel)', nb4, ms, peak,
                    'identical bits to the user kernel' if same else 'MISMATCH')

    def two():
        cp.exp(xt, out=tmp)
        cp.add(tmp, v, out=out)
    ms, _ = _median_ms(two, iters=it)
    u = ulp_err(out.to_torch())
    e_two = _entry('C4a two launches exp(x^T) then +v (bytes = fused-equivalent)', nb4, ms, peak,
                   ('ok' if u <= 2.5 else 'MISMATCH') + ' max_err=%.2f ulp (tol 2.5 = 2 ulp exp + add rounding)' % u)
    if refjit is not None:
        rt = torch.empty(1024, 1024, 256, device='cuda')
        ro = torch.empty(1024, 1024, 256, device='cuda')
        sh3, st_in, st_c = (1024, 1024, 256), (4, 4096, 4194304), (1024 * 256 * 4, 256 * 4, 4)
        f1 = refjit.elementwise('ref_exp_t3', [carray(base.data_ptr(), sh3, st_in), carray(rt.data_ptr(), sh3, st_c),
                                               oracle_cindexer(sh3)], nel)
        sh2 = (1024 * 1024, 256)
        f2 = refjit.elementwise('ref_add_b2', [carray(rt.data_ptr(), sh2, (1024, 4)), carray(tv.data_ptr(), sh2, (0, 4)),
                                               carray(ro.data_ptr(), sh2, (1024, 4)), oracle_cindexer(sh2)], nel)

        def ref_two():
            f1()
            f2()
        _ref(e_two, nb4, ref_two, 'reference: exp ufunc on the 3-D transposed view (CArray<float,3,0,1>, CIndexer<3>) '
                                  'then add with the broadcast row collapsed to 2-D; two launches, 128-thread blocks', it)
        e_fused['ref_gpu'] = e_fuse['ref_gpu'] = e_two.get('ref_gpu')
        if 'ms' in e_two.get('ref_gpu', {}):
            for e in (e_fused, e_fuse):
                e['speedup_vs_ref_gpu'] = round(e_two['ref_gpu']['ms'] / e['ms'], 3)
            u = ulp_err(ro)
            e_two['ref_gpu']['max_err_ulp'] = round(u, 2)
        del rt, ro
    entries += [e_fused, e_fuse, e_two]
    del base, want, out, tmp, xt, ulp
    torch.cuda.empty_cache()

    # ---------------- C4b: int64 cumsum 2^28 (bit-exact) + the casting / axis scans -----------------
    ti = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', dtype=torch.int64, generator=g)
    xi = cp.from_torch(ti)
    yo = cp.empty((n,), np.int64)
    ms, _ = _median_ms(lambda: cp.cumsum(xi, out=yo), iters=it)
    want = torch.cumsum(ti, 0)
    e = _entry('C4b cumsum int64 2^28', 16 * n, ms, peak, 'bit-exact' if bool(torch.equal(yo.to_torch(), want)) else 'MISMATCH')
    if refjit is not None:
        rc = torch.empty(n, device='cuda', dtype=torch.int64)
        fcopy = refjit.elementwise('ref_copy_i64', [carray(ti.data_ptr(), (n,), (8,)), carray(rc.data_ptr(), (n,), (8,)),
                                                    oracle_cindexer((n,))], n)
        fscan = refcub.scan(rc.data_ptr(), rc.data_ptr(), n, CUB_CUMSUM, 'int64', alloc)

        def ref_scan():
            fcopy()
            fscan()
        _ref(e, 16 * n, ref_scan, 'reference scan_core: astype copy kernel (_routines_math.pyx:726-727) + in-place '
                                  'cub_device_scan (CCCL DeviceScan::InclusiveSum)', it)
        ref_scan()
        e['ref_gpu']['same_bits'] = bool(torch.equal(rc, want))
        ms2, _ = _median_ms(fscan, iters=it)
        e['ref_gpu']['cub_scan_alone_ms'] = round(ms2, 5)
        e['ref_gpu']['cub_scan_alone_gbs'] = round(16 * n / ms2 / 1e6, 1)
        del rc
    entries.append(e)
    del want, yo

    tf = torch.rand(n, device='cuda', generator=g) * 2 - 1
    xf = cp.from_torch(tf)
    ms, _ = _median_ms(lambda: cp.cumsum(xf), iters=it)
    err = float((cp.cumsum(xf).to_torch().double() - torch.cumsum(tf.double(), 0)).abs().max())
    e = _entry('cumsum f32 2^28', 8 * n, ms, peak, ('ok' if err < 0.5 else 'MISMATCH') + ' max_abs_err=%.2e' % err)
    if refcub is not None:
        rc = torch.empty(n, device='cuda')
        fcopy = None
        fscan = refcub.scan(tf.data_ptr(), rc.data_ptr(), n, CUB_CUMSUM, 'float32', alloc)
        _ref(e, 8 * n, fscan, 'reference cub_device_scan alone (no astype copy charged)', it)
        del rc
    entries.append(e)
    del xf, tf

    for src_dt, tsrc, label, bpe in ((np.int32, torch.int32, 'int32->int64', 12), (np.bool_, torch.bool, 'bool->int64', 9)):
        if tsrc is torch.bool:
            tt = torch.rand(n, device='cuda', generator=g) < 0.5
        else:
            tt = torch.randint(-1000, 1000, (n,), device='cuda', dtype=tsrc, generator=g)
        xx = cp.from_torch(tt)
        ms, _ = _median_ms(lambda: cp.cumsum(xx), iters=it)
        ok = bool(torch.equal(cp.cumsum(xx).to_torch(), torch.cumsum(tt.to(torch.int64), 0)))
        entries.append(_entry('casting cumsum %s 2^28' % label, bpe * n, ms, peak, 'bit-exact' if ok else 'MISMATCH'))
        del xx, tt
    th = (torch.rand(n, device='cuda', generator=g) * 2 - 1).half()
    xh = cp.from_torch(th)
    ms, _ = _median_ms(lambda: cp.cumsum(xh), iters=it)
    goth = cp.cumsum(xh).to_torch()
    wanth = torch.cumsum(th.double(), 0)
    # float accumulator, one rounding to fp16 per output: half an fp16 ulp + fp32 accumulation error (of cumsum|x|)
    over = float(((goth.double() - wanth).abs() - 0.5 * _ulp16(wanth) - 1e-6 * torch.cumsum(th.double().abs(), 0)).max())
    entries.append(_entry('casting cumsum float16 (float accumulate) 2^28', 4 * n, ms, peak,
                          ('ok' if over <= 0 else 'MISMATCH') + ' |err| <= 0.5 fp16 ulp + 1e-6 cumsum|x|: margin %.1e' % -over))
    del xh, th, goth, wanth
    torch.cuda.empty_cache()

    t2 = torch.rand(16384, 16384, device='cuda', generator=g) * 2 - 1
    x2 = cp.from_torch(t2)
    for ax in (0, 1):
        ms, _ = _median_ms(lambda: cp.cumsum(x2, axis=ax), iters=it)
        err = float((cp.cumsum(x2, axis=ax).to_torch().double() - torch.cumsum(t2.double(), ax)).abs().max())
        entries.append(_entry('cumsum axis=%d f32 16384^2' % ax, 8 * 16384 * 16384, ms, peak,
                              ('ok' if err < 0.05 else 'MISMATCH') + ' max_abs_err=%.2e' % err))
    # short rows and strided copies
    for cols in (64, 96):
        rows = (1 << 28) // cols
        tr = torch.rand(rows, cols, device='cuda', generator=g)
        xr = cp.from_torch(tr)
        ms, _ = _median_ms(lambda: xr.sum(axis=1), iters=it)
        err = float((xr.sum(axis=1).to_torch().double() - tr.double().sum(1)).abs().max())
        entries.append(_entry('sum axis=1 f32 rows of %d (%d rows)' % (cols, rows), 4 * rows * cols + 4 * rows, ms, peak,
                              ('ok' if err < 1e-3 else 'MISMATCH') + ' max_abs_err=%.1e' % err))
        del xr, tr
    # tall narrow matrices (point clouds, feature tables, class scores): the flat-stream tile kernels
    for rows, cols in ((1 << 26, 3), (1 << 23, 32)):
        tn = torch.rand(rows, cols, device='cuda', generator=g) * 2 - 1
        xn = cp.from_torch(tn)
        nb = 4 * rows * cols
        tag = 'f32 (%d, %d)' % (rows, cols)
        ms, _ = _median_ms(lambda: xn.sum(axis=0), iters=it)
        err = float((xn.sum(axis=0).to_torch().double() - tn.double().sum(0)).abs().max())
        entries.append(_entry('sum axis=0 ' + tag, nb + 4 * cols, ms, peak,
                              ('ok' if err < 1e-5 * rows else 'MISMATCH') + ' max_abs_err=%.1e' % err))
        ms, _ = _median_ms(lambda: xn.var(axis=0), iters=it)
        rel = float(((xn.var(axis=0).to_torch().double() - tn.double().var(0, unbiased=False)).abs() /
                     tn.double().var(0, unbiased=False)).max())
        entries.append(_entry('var axis=0 ' + tag, nb + 4 * cols, ms, peak, ('ok' if rel < 1e-5 else 'MISMATCH') + ' rel_err=%.1e' % rel))
        ms, _ = _median_ms(lambda: xn.argmax(axis=1), iters=it)
        same = bool(torch.equal(xn.argmax(axis=1).to_torch(), tn.argmax(1)))
        entries.append(_entry('argmax axis=1 ' + tag, nb + 8 * rows, ms, peak, 'same indices' if same else 'MISMATCH'))
        for ax in (0, 1):
            ms, _ = _median_ms(lambda: cp.cumsum(xn, axis=ax), iters=it)
            got = cp.cumsum(xn, axis=ax).to_torch()
            ref = torch.cumsum(tn.double(), ax)
            err = float((got.double() - ref).abs().max())
            bound = 1e-6 * float(torch.cumsum(tn.double().abs(), ax).max())
            entries.append(_entry('cumsum axis=%d %s' % (ax, tag), 2 * nb, ms, peak,
                                  ('ok' if err <= bound else 'MISMATCH') + ' max_abs_err=%.1e (bound %.1e)' % (err, bound)))
            del got, ref
        del xn, tn
        torch.cuda.empty_cache()
    src = cp.from_torch(t2)
    dst = cp.empty((16384, 32768), np.float32)
    dst_t = dst.to_torch()
    dst_t.zero_()
    ms, _ = _median_ms(lambda: cp.elementwise_copy(src, dst[:, ::2]), iters=it)
    ok = bool(torch.equal(dst_t[:, ::2], t2)) and float(dst_t[:, 1::2].abs().max()) == 0.0
    entries.append(_entry('strided scatter copy out[:, ::2] = x, f32 16384^2 (algorithmic 8 B/elem; DRAM moves whole '
                          '32-byte sectors: 12 B/elem floor)', 8 * 16384 * 16384, ms, peak, 'bit-exact' if ok else 'MISMATCH'))
    wide = cp.from_torch(dst_t)
    dense = cp.empty((16384, 16384), np.float32)
    ms, _ = _median_ms(lambda: cp.elementwise_copy(wide[:, ::2], dense), iters=it)
    ok = bool(torch.equal(dense.to_torch(), dst_t[:, ::2]))
    entries.append(_entry('strided gather copy y = x[:, ::2], f32 16384^2', 8 * 16384 * 16384, ms, peak,
                          'bit-exact' if ok else 'MISMATCH'))
    del t2, x2, src, dst, dst_t, wide, dense
    torch.cuda.empty_cache()

    # ---------------- the widened elementwise family (cupy_b200/_core/_routines_elementwise.py) ----------------
    # operands of different item sizes (bool masks) and the division / rounding / NaN-aware members, 2^28 float32
    n = 1 << 28
    tx = torch.rand(n, device='cuda', generator=g) * 8 - 4
    ty = torch.rand(n, device='cuda', generator=g) * 3 + 1
    tm = torch.rand(n, device='cuda', generator=g) > 0.5
    tx[::97] = float('nan')
    fx, fy, fm = cp.from_torch(tx), cp.from_torch(ty), cp.from_torch(tm)
    fo = cp.empty((n,), np.float32)
    fb = cp.empty((n,), np.bool_)

    def same(got, want):
        got = got.to_torch()
        if got.dtype != want.dtype or got.shape != want.shape:
            return 'MISMATCH (dtype / shape)'
        if want.dtype == torch.float32:
            ok = bool(((got == want) | (got.isnan() & want.isnan())).all())
            return 'exact, NaNs in the same places' if ok else 'MISMATCH'
        return 'bit-exact' if bool(torch.equal(got, want)) else 'MISMATCH'
    for name, f, want, nbytes in (
            ('where(mask, x, y) f32 2^28 (bool mask + two operands)', lambda: cp.where(fm, fx, fy), lambda: torch.where(tm, tx, ty), 13 * n),
            ('greater(x, y) -> bool f32 2^28', lambda: cp.greater(fx, fy, out=fb), lambda: tx > ty, 9 * n),
            ('isnan(x) -> bool f32 2^28', lambda: cp.isnan(fx, out=fb), lambda: torch.isnan(tx), 5 * n),
            ('floor_divide(x, y) f32 2^28', lambda: cp.floor_divide(fx, fy, out=fo), lambda: torch.floor(tx / ty), 12 * n),
            ('clip(x, -1, 1) f32 2^28', lambda: cp.clip(fx, -1, 1, out=fo), lambda: torch.clamp(tx, -1, 1), 8 * n),
            ('floor(x) f32 2^28', lambda: cp.floor(fx, out=fo), lambda: torch.floor(tx), 8 * n)):
        ms, _ = _median_ms(f, iters=it)
        entries.append(_entry(name, nbytes, ms, peak, same(f(), want())))
    ms, _ = _median_ms(lambda: cp.nanmean(fx), iters=it)
    ref = float(torch.nanmean(tx.double()))
    err = abs(float(cp.nanmean(fx).get()) - ref)
    entries.append(_entry('nanmean(x) f32 2^28 (full; sum and count of the non-NaN elements in one pass)', 4 * n, ms, peak,
                          ('ok' if err < 1e-5 else 'MISMATCH') + ' abs_err=%.1e' % err))
    del ty, tm, fy, fm, fo, fb
    entries.extend(compaction_rows(cp, torch, fx, tx, peak, it))
    del tx, fx
    torch.cuda.empty_cache()
    return entries, ref_note


def compaction_rows(cp, torch, fx, tx, peak, it):
    """The scan's callers (cupy_b200/_core/_compaction.py) on 2^28 float32: flags -> int32 ranks (prebuilt
    bool -> int32 scan) -> one scatter; each call reads the hit count back once, as the reference does
    (cupy/_core/_routines_indexing.pyx:123).  Bytes = the input read once + the compacted output written once."""
    n = fx.size
    rows = []
    hits = int((tx > 0.75).sum())
    ms, _ = _median_ms(lambda: cp.flatnonzero(fx > 0.75), iters=it)
    ok = bool(torch.equal(cp.flatnonzero(fx > 0.75).to_torch(), torch.nonzero(tx > 0.75).flatten()))
    rows.append(_entry('flatnonzero(x > 0.75) f32 2^28 (compare + rank scan + scatter, one count read-back; %d hits)' % hits,
                       4 * n + 8 * hits, ms, peak, 'bit-exact' if ok else 'MISMATCH', launches=3))
    ms, _ = _median_ms(lambda: fx[fx > 0.75], iters=it)
    ok = bool(torch.equal(fx[fx > 0.75].to_torch(), tx[tx > 0.75]))
    rows.append(_entry('x[x > 0.75] f32 2^28 (boolean-mask select: compare + rank scan + gather)', 4 * n + 4 * hits, ms, peak,
                       'bit-exact' if ok else 'MISMATCH', launches=3))
    return rows


def oracle_cindexer(shape):
    from oracle.ref_gpu import cindexer
    return cindexer(shape)


def _c3_ref(e, refjit, refcub, t, op, ax, sfx, m, isz, nb, alloc, it):
    """The reference's call chain for one C3 case."""
    import ctypes
    import torch
    from oracle.ref_gpu import carray, cindexer, CUB_SUM, CUB_MAX
    dt = 'float32' if sfx == 'f32' else 'float16'
    tdt = t.dtype
    ptr = t.data_ptr()
    flat_in = carray(ptr, (m * m,), (isz,))
    if op in ('sum', 'max') and ax == 1:
        y = torch.empty(m, device='cuda', dtype=tdt)
        f = refcub.segmented_reduce(ptr, y.data_ptr(), m, m, CUB_SUM if op == 'sum' else CUB_MAX, dt, alloc)
        _ref(e, nb, f, 'reference cub_device_segmented_reduce (cub.pyx:210-273)' +
             ('; accumulates in __half' if sfx == 'f16' and op == 'sum' else ''), it)
        return
    if op == 'argmax' and ax == 1:
        y = torch.empty(m, device='cuda', dtype=torch.int64)
        f = refjit.cub_block('ref_cub_argmax_' + sfx, ptr, y.data_ptr(), m, m)
        _ref(e, nb, f, 'reference CUB-block JIT template (_cub_reduction.pyx:33-242): one 512-thread block per row, '
                       '4 items per thread, BlockReduce per 2048-element tile', it)
        f()
        e['ref_gpu']['same_indices'] = bool(torch.equal(y, torch.argmax(t, dim=1))) or 'ties differ from torch.argmax'
        return
    if op in ('sum', 'max', 'argmax'):       # axis=0: generic reduction, 1-D collapsed input
        odt = torch.int64 if op == 'argmax' else tdt
        y = torch.empty(m, device='cuda', dtype=odt)
        f = refjit.reduction('ref_%s_%s' % (op, sfx), [flat_in], [], carray(y.data_ptr(), (m,), (y.element_size(),)),
                             (m * m,), (m,), m)
        _ref(e, nb, f, 'reference generic reduction template (_reduction.pyx:59-112), geometry %s' % f.geometry, it)
        return
    # var: mean (keepdims) then the second pass over x and the broadcast mean
    mean = torch.empty(m, device='cuda', dtype=tdt)
    y = torch.empty(m, device='cuda', dtype=tdt)
    if ax == 0:
        f1 = refjit.reduction('ref_mean_' + sfx, [flat_in], [], carray(mean.data_ptr(), (m,), (isz,)), (m * m,), (m,), m)
        x2 = carray(ptr, (m, m), (m * isz, isz))
        mb = carray(mean.data_ptr(), (m, m), (0, isz))
        what = 'reference _var: cupy_mean generic reduction + cupy_var_core second pass (2 reads of x)'
    else:
        fs = refcub.segmented_reduce(ptr, mean.data_ptr(), m, m, CUB_SUM, dt, alloc)

        def f1():
            fs()
            mean.div_(m)           # stands in for the reference's true_divide launch on 32768 items
        x2 = carray(ptr, (m, m), (isz, m * isz))          # transposed to (reduce, out)
        mb = carray(mean.data_ptr(), (m, m), (0, isz))
        what = ('reference _var: CUB segmented sum + true_divide + cupy_var_core second pass with args transposed '
                'to (reduce, out) (2 reads of x)')
    f2 = refjit.reduction('ref_var_core_%s_ax%d' % (sfx, ax), [x2, mb], [ctypes.c_float(1.0 / m)],
                          carray(y.data_ptr(), (m,), (isz,)), (m, m), (m,), m)

    def both():
        f1()
        f2()
    _ref(e, nb, both, what + ', geometry %s' % f2.geometry, it)


# ------------------------------------------------------------------------------------------------
# C5: 1-D float32 2^32 elements sharded over the ranks (strong scaling), sum and var
# ------------------------------------------------------------------------------------------------
def run_c5(world, rank, comm, peak, log2_total=32, iters=20):
    import torch
    import torch.distributed as dist
    import cupy_b200 as cp
    from cupy_b200 import distributed as cdist
    n_total = 1 << log2_total
    n = n_total // world
    g = torch.Generator(device='cuda')
    g.manual_seed(77 + rank)
    tx = torch.empty(n, device='cuda', dtype=torch.float32)
    step = 1 << 28
    for lo in range(0, n, step):
        tx[lo:lo + step] = torch.rand(min(step, n - lo), device='cuda', generator=g) * 2 - 1
    x = cp.from_torch(tx)
    ref = torch.zeros(3, device='cuda', dtype=torch.float64)          # n, sum, sum of squares
    for lo in range(0, n, step):
        c = tx[lo:lo + step].double()
        ref[0] += c.numel(); ref[1] += c.sum(); ref[2] += (c * c).sum()
    if world > 1:
        dist.all_reduce(ref)
    want_sum = float(ref[1])
    mean = float(ref[1] / ref[0])
    want_var = float(ref[2] / ref[0] - mean * mean)

    class Solo:
        _n_devices, rank = 1, 0

        def all_reduce(self, a, b, op='sum', stream=None):
            pass
    c = comm if comm is not None else Solo()
    fused = bool(comm is not None and comm.peer_exchange() is not None)
    out = {}
    for name, f, want in (('sum', lambda: cdist.sharded_sum(x, c), want_sum),
                          ('var', lambda: cdist.sharded_var(x, c), want_var)):
        for _ in range(3):
            r = f()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            r = f()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda', dtype=torch.float64)
        got = torch.tensor([float(r.get()) if hasattr(r, 'get') else float(r)], device='cuda', dtype=torch.float64)
        lo_, hi_ = got.clone(), got.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        rel = abs(float(got) - want) / max(abs(want), 1e-30) if name == 'var' else abs(float(got) - want) / max(1.0, abs(want))
        ms = float(ms)
        out[name] = {'ms': round(ms, 5), 'gbs_aggregate': round(4 * n_total / ms / 1e6, 1),
                     'frac_per_gpu': round(4 * n_total / ms / 1e6 / world / peak, 4),
                     'check': ('ok' if rel <= 1e-5 else 'MISMATCH') + ' rel_err=%.1e vs float64 (tol 1e-5)' % rel,
                     'identical_on_all_ranks': bool(float(lo_) == float(hi_))}
    out['combine'] = ('fused: per-GPU partials exchanged through NVLink peer memory inside the reduction kernel '
                      '(b200_reduce_run_sharded), one launch per GPU' if fused else
                      'single GPU' if world == 1 else 'NCCL all-reduce / all-gather of the partials (fallback route)')
    out['elements_total'] = n_total
    out['elements_per_gpu'] = n
    out['scaling'] = 'strong (2^%d float32 total, contiguous shards); efficiency = ms(N=1) / (N * ms(N))' % log2_total
    del x, tx
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# DistributedArray over the ranks: sharded rows, lazy SUM-mode reduction, resharding
# ------------------------------------------------------------------------------------------------
def run_darray_check(world, rank, comm):
    import cupy_b200 as cp
    from cupy_b200.distributed import array as da
    rows = 64 * world
    base = np.arange(rows * 96, dtype=np.float32).reshape(rows, 96)
    imap = {r: slice(64 * r, 64 * (r + 1)) for r in range(world)}
    d = da.distributed_array(base, imap, da.REPLICA, comm=comm)
    s = d.sum(axis=0)                                   # one engine kernel per rank, no exchange
    lazy = s.mode is da.SUM
    ok = bool(np.allclose(s.get(), base.sum(axis=0), rtol=1e-6))
    cols = {r: (slice(None), slice(96 * r // world, 96 * (r + 1) // world)) for r in range(world)}
    t = (d * d).reshard(cols)                           # row shards -> column shards over send / recv
    ok = ok and bool(np.array_equal(t.get(), base * base))
    ok = ok and bool(np.array_equal(t.max(axis=1).change_mode(da.REPLICA).get(), (base * base).max(axis=1)))
    return {'check': 'ok' if ok else 'MISMATCH', 'lazy_sum_mode': lazy, 'ranks': world,
            'what': 'distributed_array(row shards).sum(axis=0) stays in SUM mode until get(); (d*d).reshard(column '
                    'shards); max(axis=1).change_mode(REPLICA) -- all compared with NumPy'}
