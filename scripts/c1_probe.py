"""C1 (4096^2 float32, L2-resident) kernel time apart from the per-call host cost: per-call events (the bench
protocol), batched launches between two events, and replays of a CUDA graph holding the same launches."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402


def batched(f, n=200):
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        f()
    b.record()
    host = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, host


def graphed(f, n=200, inner=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            f()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            keep = f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n // inner):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (n // inner * inner) * 1e3, keep


x = cp.from_torch(torch.rand(4096, 4096, device='cuda') * 2 - 1)
fz = cp.fuse(kernel_name='c1_x2p1')(lambda a: a * 2 + 1)
out = cp.empty((4096, 4096), np.float32)
rows = cp.empty((4096,), np.float32)
small = cp.arange(1000)
cases = [('fuse x*2+1', lambda: fz(x), 2 * 4 * 4096 * 4096),
         ('multiply(x,2,out) ', lambda: cp.multiply(x, 2, out=out), 2 * 4 * 4096 * 4096),
         ('x.sum(axis=1)', lambda: x.sum(axis=1), 4 * 4096 * 4096),
         ('x.sum(axis=1,out)', lambda: x.sum(axis=1, out=rows), 4 * 4096 * 4096),
         ('x.sum(axis=0)', lambda: x.sum(axis=0), 4 * 4096 * 4096),
         ('x.sum()', lambda: x.sum(), 4 * 4096 * 4096),
         ('x.max(axis=1)', lambda: x.max(axis=1), 4 * 4096 * 4096),
         ('cumsum(x, axis=1)', lambda: cp.cumsum(x, axis=1), 2 * 4 * 4096 * 4096),
         ('cumsum(x)', lambda: cp.cumsum(x), 2 * 4 * 4096 * 4096),
         ('arange(1000).sum()', lambda: small.sum(), 8000)]
for name, f, nbytes in cases:
    us_b, host = batched(f)
    try:
        us_g, keep = graphed(f)
        ok = ''
        if name.startswith('x.sum(axis=1)'):
            ok = ' graph result ok' if bool(torch.allclose(keep.to_torch(), x.to_torch().sum(1), atol=1e-3)) else ' GRAPH MISMATCH'
        gtxt = '%7.2f us/launch in a graph (%6.0f GB/s)%s' % (us_g, nbytes / us_g / 1e3, ok)
    except Exception as ex:        # noqa: BLE001
        gtxt = 'graph capture failed: %s: %s' % (type(ex).__name__, str(ex)[:120])
        torch.cuda.synchronize()
    print('%-20s batched %7.2f us/call (host %6.2f us) | %s' % (name, us_b, host, gtxt), flush=True)
