import sys, torch
sys.path.insert(0, ".")
import cupy_b200 as cp
from bench_configs import _median_ms
for lg in (29, 30, 32):
    n = 1 << lg
    t = torch.empty(n, device="cuda")
    for lo in range(0, n, 1 << 28):
        t[lo:lo + (1 << 28)] = torch.rand(1 << 28, device="cuda") * 2 - 1
    x = cp.from_torch(t)
    for op in ("sum", "var", "max"):
        ms, mn = _median_ms(lambda: getattr(x, op)(), iters=20)
        print("2^%d %s: median %.4f ms min %.4f  %.1f GB/s" % (lg, op, ms, mn, 4 * n / ms / 1e6), flush=True)
    del x, t
    torch.cuda.empty_cache()
