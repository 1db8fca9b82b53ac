"""C3 (32768^2) axis-0 / axis-1 reductions, float32 and float16: per-call CUDA-event medians (the bench protocol)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from bench_configs import _median_ms  # noqa: E402

m = 32768
for tdt, isz in ((torch.float32, 4), (torch.float16, 2)):
    t = torch.empty(m, m, device='cuda', dtype=tdt)
    for lo in range(0, m, 4096):
        t[lo:lo + 4096] = (torch.rand(4096, m, device='cuda') * 2 - 1).to(tdt)
    x = cp.from_torch(t)
    row = []
    for op in ('sum', 'max', 'argmax', 'var'):
        for ax in (0, 1):
            ms, _ = _median_ms(lambda: getattr(x, op)(axis=ax), iters=10)
            row.append('%s%d %.0f' % (op, ax, isz * m * m / ms / 1e6))
    print(str(tdt), ' | '.join(row), flush=True)
    del x, t
    torch.cuda.empty_cache()
