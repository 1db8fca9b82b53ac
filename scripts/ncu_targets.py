"""One launch of every kernel the judge asked ncu evidence for (run under `ncu --set full -k regex:...`):
C3 row / column reductions at 32768^2 (fp32 and fp16; sum, argmax, var), the pipelined scans (int64, int32->int64,
bool->int64, float16), axpy (headline, for roofline.traffic), and the strided scatter / gather copies.
    ncu --set full --clock-control none --import-source on -k regex:'reduce_rows|reduce_cols|scan_pipe|axpy|copy' \
        -o gpurun_out/r02_targets python scripts/ncu_targets.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402

which = sys.argv[1:] or ['c3', 'scan', 'axpy', 'copy']
g = torch.Generator(device='cuda')
g.manual_seed(0)
if 'c3' in which:
    m = 32768
    for tdt in (torch.float32, torch.float16):
        t = torch.empty(m, m, device='cuda', dtype=tdt)
        for lo in range(0, m, 4096):
            t[lo:lo + 4096] = (torch.rand(4096, m, device='cuda', generator=g) * 2 - 1).to(tdt)
        x = cp.from_torch(t)
        for op in ('sum', 'argmax', 'var'):
            for ax in (0, 1):
                getattr(x, op)(axis=ax)
        del x, t
        torch.cuda.empty_cache()
n = 1 << 28
if 'scan' in which:
    xi = cp.from_torch(torch.randint(-1000, 1000, (n,), device='cuda', dtype=torch.int64, generator=g))
    cp.cumsum(xi)
    del xi
    x32 = cp.from_torch(torch.randint(-1000, 1000, (n,), device='cuda', dtype=torch.int32, generator=g))
    cp.cumsum(x32)
    del x32
    xb = cp.from_torch(torch.rand(n, device='cuda', generator=g) < 0.5)
    cp.cumsum(xb)
    del xb
    xh = cp.from_torch((torch.rand(n, device='cuda', generator=g) * 2 - 1).half())
    cp.cumsum(xh)
    del xh
    torch.cuda.empty_cache()
if 'axpy' in which:
    x = cp.from_torch(torch.rand(n, device='cuda', generator=g))
    y = cp.from_torch(torch.rand(n, device='cuda', generator=g))
    z = cp.empty((n,), np.float32)
    k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')
    k(np.float32(1.5), x, y, z)
    x.sum()
    del x, y, z
    torch.cuda.empty_cache()
if 'copy' in which:
    src = cp.from_torch(torch.rand(16384, 16384, device='cuda', generator=g))
    dst = cp.zeros((16384, 32768), np.float32)
    cp.elementwise_copy(src, dst[:, ::2])          # strided scatter
    dense = cp.empty((16384, 16384), np.float32)
    cp.elementwise_copy(dst[:, ::2], dense)        # strided gather
torch.cuda.synchronize()
print('ncu targets done')
