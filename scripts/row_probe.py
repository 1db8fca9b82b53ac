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
_kernel.tunables['row_unroll'] = 2
_kernel.tunables['blocks_per_sm'] = 0
kc = cp.ElementwiseKernel('T x, T c', 'T z', 'z = x + c', 'row_addcol')
report('JIT x + col (stride 0 inner, SPEC)', 8 * nel, lambda: kc(tmp, col, out), iters=10)
big = cp.from_torch(randn((1024, 1024, 512), torch.float32))
report('prebuilt x[..., ::2] + v (strided inner)', 12 * nel, lambda: cp.add(big[..., ::2], v, out=out), iters=10)
ks = cp.ElementwiseKernel('T x, T v', 'T z', 'z = x + v', 'row_addstrided')
report('JIT x[..., ::2] + v (SPEC)', 12 * nel, lambda: ks(big[..., ::2], v, out), iters=10)
pad = big[:, :, :300]
o3 = cp.empty((1024, 1024, 300), np.float32)
report('prebuilt padded rows x[:, :, :300] * 2', 8 * 1024 * 1024 * 300, lambda: cp.multiply(pad, 2, out=o3), iters=10)
kp = cp.ElementwiseKernel('T x', 'T z', 'z = x * 2', 'row_pad')
report('JIT padded rows x[:, :, :300] * 2 (SPEC)', 8 * 1024 * 1024 * 300, lambda: kp(pad, o3), iters=10)
