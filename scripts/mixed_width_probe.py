"""FLAT loops over operands of different item sizes (bool masks, casts): vector width x elements per thread sweep
(cupy_b200._core._kernel.tunables['flat_mixed_vec'] / ['flat_mixed_items']) at 2^28 elements."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from cupy_b200._core._kernel import tunables  # noqa: E402

PEAK = 6546.9


def timed(f, n=15):
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
i8 = cp.from_torch((torch.rand(n, device='cuda') * 100).to(torch.int8))
out = cp.empty((n,), np.float32)
outb = cp.empty((n,), np.bool_)
outh = cp.empty((n,), np.float16)
cast = cp.ElementwiseKernel('float32 a', 'float16 b', 'b = a', 'probe_cast_f32_f16')
cases = [
    ('greater(x, y) -> bool', lambda: cp.greater(x, y, out=outb), 9 * n),
    ('isnan(x) -> bool', lambda: cp.isnan(x, out=outb), 5 * n),
    ('where(m, x, y)', lambda: cp.where(m, x, y), 13 * n),
    ('add(int8, f32)', lambda: cp.add(i8, x, out=out), 9 * n),
    ('cast f32 -> f16 (user kernel)', lambda: cast(x, outh), 6 * n),
    ('logical_not(m)', lambda: cp.logical_not(m, out=outb), 2 * n),
]
want = {}
for vec, items in ((0, 16), (8, 16), (8, 32), (16, 16), (16, 32)):
    tunables['flat_mixed_vec'] = vec
    tunables['flat_mixed_items'] = items
    for name, f, nbytes in cases:
        ms = timed(f)
        got = f()
        got = got.get()[:1 << 20] if hasattr(got, 'get') else None
        if name not in want:
            want[name] = got
        same = bool(got is None or np.array_equal(got, want[name], equal_nan=True))
        gbs = nbytes / ms / 1e6
        print(json.dumps({'vec': vec, 'items': items, 'case': name, 'ms': round(ms, 4), 'GBps': round(gbs, 1),
                          'pct': round(100 * gbs / PEAK, 1), 'same_as_default': same}))
