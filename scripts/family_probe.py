"""Throughput of a few members of the widened elementwise family at 2^28 float32 elements (same FLAT tiler as
axpy; the algorithmic bytes are the operands read + the result written)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402

PEAK = 6546.9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass


def timed(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


n = 1 << 28
x = cp.from_torch(torch.rand(n, device='cuda') * 8 - 4)
y = cp.from_torch(torch.rand(n, device='cuda') * 3 + 1)
m = cp.from_torch(torch.rand(n, device='cuda') > 0.5)
out = cp.empty((n,), np.float32)
outb = cp.empty((n,), np.bool_)
cases = [
    ('floor(x)', lambda: cp.floor(x, out=out), 8 * n),
    ('clip(x, -1, 1)', lambda: cp.clip(x, -1, 1, out=out), 8 * n),
    ('floor_divide(x, y)', lambda: cp.floor_divide(x, y, out=out), 12 * n),
    ('remainder(x, y)', lambda: cp.remainder(x, y, out=out), 12 * n),
    ('where(m, x, y)', lambda: cp.where(m, x, y), 13 * n),
    ('isnan(x)', lambda: cp.isnan(x, out=outb), 5 * n),
    ('arctan(x)', lambda: cp.arctan(x, out=out), 8 * n),
    ('logaddexp(x, y)', lambda: cp.logaddexp(x, y, out=out), 12 * n),
    ('nanmean(x)', lambda: cp.nanmean(x), 4 * n),
]
for name, f, nbytes in cases:
    ms = timed(f)
    gbs = nbytes / ms / 1e6
    print(json.dumps({'case': name + ' f32 2^28', 'ms': round(ms, 4), 'GBps': round(gbs, 1), 'pct_of_measured_peak': round(100 * gbs / PEAK, 1)}))

# the compaction rows of the bench `configs` table, through the very function bench.py calls
import bench_configs  # noqa: E402
tx = torch.rand(n, device='cuda') * 8 - 4
tx[::97] = float('nan')
for e in bench_configs.compaction_rows(cp, torch, cp.from_torch(tx), tx, PEAK, 10):
    print(json.dumps(e))
