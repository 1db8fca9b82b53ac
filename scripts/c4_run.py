"""Config 4a launches only (for ncu captures): fused exp(x^T)+v and the transposed copy."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import cupy_b200 as cp

base = cp.from_torch((torch.rand((256, 1024, 1024), device='cuda') * 2 - 1))
xt = base.transpose(2, 1, 0)
v = cp.from_torch(torch.rand((256,), device='cuda'))
out = cp.empty((1024, 1024, 256), np.float32)
fused = cp.ElementwiseKernel('T x, T v', 'T z', 'z = exp(x) + v', 'expadd')
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    fused(xt, v, out)
    cp.elementwise_copy(xt, out)
torch.cuda.synchronize()
