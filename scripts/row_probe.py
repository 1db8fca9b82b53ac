"""ROWWISE tiler sweep: dense (2^20, 256) float32 + broadcast row v (config 4a's second launch)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import cupy_b200 as cp
from cupy_b200._core import _kernel
from scripts.perf_probe import report, randn

nel = 1 << 28
tmp = cp.from_torch(randn((1024, 1024, 256), torch.float32))
out = cp.empty((1024, 1024, 256), np.float32)
v = cp.from_torch(randn((256,), torch.float32))
col = cp.from_torch(randn((1024, 1024, 1), torch.float32))
k = cp.ElementwiseKernel('T x, T v', 'T z', 'z = x + v', 'row_addv')
report('prebuilt tmp + v', 8 * nel, lambda: cp.add(tmp, v, out=out), iters=10)
report('prebuilt tmp + col (stride 0 inner)', 8 * nel, lambda: cp.add(tmp, col, out=out), iters=10)
for un in (1, 2, 4, 8):
    for bps in (0, 8, 16, 64):
        _kernel.tunables['row_unroll'] = un
        _kernel.tunables['blocks_per_sm'] = bps
        report('JIT x + v unroll=%d blocks/SM=%d' % (un, bps), 8 * nel, lambda: k(tmp, v, out), iters=10)
