"""Axis-0 reductions and scans of tall, narrow matrices (point clouds, feature tables), replayed in a CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from bench_configs import _graph_us  # noqa: E402

for shape in ((1 << 24, 3), (1 << 24, 4), (1 << 22, 16), (1 << 21, 32), (1 << 20, 64), (1 << 20, 100), (1 << 19, 128), (1 << 18, 256)):
    t = torch.rand(*shape, device='cuda') * 2 - 1
    x = cp.from_torch(t)
    nbytes = 4 * shape[0] * shape[1]
    row = []
    for name, f, mult in (('sum0', lambda: x.sum(axis=0), 1), ('max0', lambda: x.max(axis=0), 1), ('argmax0', lambda: x.argmax(axis=0), 1),
                          ('var0', lambda: x.var(axis=0), 1), ('cumsum0', lambda: cp.cumsum(x, axis=0), 2)):
        us = _graph_us(f, inner=5, replays=4)
        row.append('%s %7.1f us %5.0f GB/s' % (name, us, mult * nbytes / us / 1e3))
    ok = bool(torch.allclose(x.sum(axis=0).to_torch(), t.sum(0), rtol=1e-4, atol=1)) and bool(torch.equal(x.argmax(axis=0).to_torch(), t.argmax(0)))
    print('%-16s %s %s' % (shape, ' | '.join(row), 'ok' if ok else 'MISMATCH'), flush=True)
