"""Column reductions of mid-size (L2-resident to a few hundred MB) matrices, replayed in a CUDA graph."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from bench_configs import _graph_us  # noqa: E402

for shape in ((4096, 4096), (1024, 16384), (16384, 1024), (8192, 8192), (2048, 2048), (512, 512), (65536, 256),
              (262144, 128), (1 << 20, 64), (1 << 22, 64), (1 << 20, 256), (30000, 1000)):
    t = torch.rand(*shape, device='cuda') * 2 - 1
    x = cp.from_torch(t)
    nbytes = 4 * shape[0] * shape[1]
    row = []
    for name, f in (('sum0', lambda: x.sum(axis=0)), ('max0', lambda: x.max(axis=0)), ('argmax0', lambda: x.argmax(axis=0)),
                    ('var0', lambda: x.var(axis=0)), ('sum1', lambda: x.sum(axis=1))):
        us = _graph_us(f)
        row.append('%s %6.1f us %5.0f GB/s' % (name, us, nbytes / us / 1e3))
    ok = bool(torch.allclose(x.sum(axis=0).to_torch(), t.sum(0), atol=1e-2)) and bool(torch.equal(x.argmax(axis=0).to_torch(), t.argmax(0)))
    print('%-14s %s %s' % (shape, ' | '.join(row), 'ok' if ok else 'MISMATCH'), flush=True)
