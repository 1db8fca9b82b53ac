"""Axis-1 (contiguous axis) reductions and scans of matrices with very short rows, replayed in a CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from bench_configs import _graph_us  # noqa: E402

for shape in ((1 << 24, 3), (1 << 24, 4), (1 << 23, 7), (1 << 22, 16), (1 << 21, 32), (1 << 20, 64), (1 << 20, 100), (1 << 19, 128)):
    t = torch.rand(*shape, device='cuda') * 2 - 1
    x = cp.from_torch(t)
    nbytes = 4 * shape[0] * shape[1]
    row = []
    for name, f, mult in (('sum1', lambda: x.sum(axis=1), 1), ('max1', lambda: x.max(axis=1), 1), ('argmax1', lambda: x.argmax(axis=1), 1),
                          ('var1', lambda: x.var(axis=1), 1), ('cumsum1', lambda: cp.cumsum(x, axis=1), 2)):
        us = _graph_us(f, inner=5, replays=4)
        row.append('%s %7.1f us %5.0f GB/s' % (name, us, mult * nbytes / us / 1e3))
    ok = bool(torch.allclose(x.sum(axis=1).to_torch(), t.sum(1), rtol=1e-4, atol=1e-4)) and bool(torch.equal(x.argmax(axis=1).to_torch(), t.argmax(1))) \
        and bool(torch.allclose(cp.cumsum(x, axis=1).to_torch(), torch.cumsum(t, 1), rtol=1e-4, atol=1e-4))
    print('%-16s %s %s' % (shape, ' | '.join(row), 'ok' if ok else 'MISMATCH'), flush=True)
