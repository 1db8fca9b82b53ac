"""Scan kernel sweep on a GPU: correctness vs torch.cumsum (int64, bit-exact) and device
time per configuration of the TMA scan (B200_SCAN_CFG is read once per process, so each
configuration runs in its own subprocess).

    python scripts/scan_probe.py            # sweep
    python scripts/scan_probe.py --one      # run in this process with the current env
"""
import os
import subprocess
import sys

sys.path.insert(0, '.')


def one():
    import numpy as np
    import torch
    import cupy_b200 as cp
    cfg = os.environ.get('B200_SCAN_CFG', 'default')
    ok = True
    for dt in (torch.int64, torch.int32, torch.float32, torch.float64):
        for n in ((1 << 20), (1 << 20) + 1, (1 << 22) + 12345, (1 << 24) - 7, 1 << 26):
            if dt in (torch.int64, torch.int32):
                t = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', dtype=dt)
            else:
                t = torch.rand(n, device='cuda', dtype=dt) - 0.5
            x = cp.from_torch(t)
            kw = {'dtype': np.int32} if dt == torch.int32 else {}
            got = x.cumsum(**kw).to_torch()
            if dt in (torch.int64, torch.int32):
                want = torch.cumsum(t, 0, dtype=dt)
                good = bool(torch.equal(got, want))
            else:
                want = torch.cumsum(t.double(), 0)
                err = float((got.double() - want).abs().max())
                good = err < (1e-2 if dt == torch.float32 else 1e-8) * max(1.0, n ** 0.5)
            ok &= good
            if not good:
                bad = (got != want).nonzero()[:4].flatten().tolist() if dt in (torch.int64, torch.int32) else err
                print('cfg', cfg, 'MISMATCH', dt, n, bad, flush=True)
    n = 1 << 28
    t = torch.randint(-(1 << 20), 1 << 20, (n,), device='cuda', dtype=torch.int64)
    x = cp.from_torch(t)
    out = cp.empty((n,), np.int64)
    for _ in range(3):
        x.cumsum(out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record()
    for i in range(20):
        x.cumsum(out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    med = ts[10]
    exact = bool(torch.equal(out.to_torch(), torch.cumsum(t, 0)))
    print('cfg %-8s ok=%s exact2^28=%s  int64 2^28: %.3f ms  %.1f GB/s (%.1f%% of 6650)' % (
        cfg, ok, exact, med, 16 * n / med / 1e6, 100 * 16 * n / med / 1e6 / 6650), flush=True)
    # fp32 2^28
    tf = torch.rand(n, device='cuda', dtype=torch.float32) - 0.5
    xf = cp.from_torch(tf)
    of = cp.empty((n,), np.float32)
    for _ in range(3):
        xf.cumsum(out=of)
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(20):
        xf.cumsum(out=of)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    print('cfg %-8s f32 2^28: %.3f ms  %.1f GB/s' % (cfg, ts[10], 8 * n / ts[10] / 1e6), flush=True)


def main():
    if '--one' in sys.argv:
        return one()
    for cfg in sys.argv[1:] or ['-1', '12', '13', '14', '16', '22', '23', '24']:
        env = dict(os.environ, B200_SCAN_CFG=cfg)
        try:
            r = subprocess.run([sys.executable, __file__, '--one'], env=env, timeout=240, capture_output=True, text=True)
            print(r.stdout.strip() or r.stderr.strip()[-800:], flush=True)
        except subprocess.TimeoutExpired:
            print('cfg', cfg, 'TIMEOUT', flush=True)


if __name__ == '__main__':
    main()
