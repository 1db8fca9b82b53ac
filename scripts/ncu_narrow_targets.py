"""One launch of each narrow / short-row kernel for `ncu --set full` (profiles/r02_narrow_ncu_full.csv):
    ncu --set full --clock-control none -k regex:'narrow|short_rows' -o /tmp/r02_narrow python scripts/ncu_narrow_targets.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402

for shape in ((1 << 26, 3), (1 << 23, 32)):
    x = cp.from_torch(torch.rand(*shape, device='cuda') * 2 - 1)
    x.sum(axis=0)
    x.argmax(axis=0)
    x.var(axis=0)
    cp.cumsum(x, axis=0)
    cp.cumsum(x, axis=1)
    x.sum(axis=1) if shape[1] == 32 else None
    del x
    torch.cuda.empty_cache()
x4 = cp.from_torch(torch.rand(1 << 25, 4, device='cuda'))
x4.argmax(axis=1)
x4.sum(axis=1)
torch.cuda.synchronize()
print('narrow ncu targets done')
