"""int64 cumsum 2^28 launches only (for ncu captures)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import torch
import cupy_b200 as cp
n = 1 << 28
xi = cp.from_torch(torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', dtype=torch.int64))
yo = cp.empty((n,), np.int64)
for _ in range(4):
    cp.cumsum(xi, out=yo)
torch.cuda.synchronize()
