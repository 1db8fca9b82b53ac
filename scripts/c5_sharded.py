"""BASELINE.json config 5: 1-D float32 2^32-element sum / var sharded over N GPUs
(one process per GPU, torchrun), per-GPU single-pass partials + one NCCL exchange.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 scripts/c5_sharded.py [--log2 32]

Checks the results against float64 references computed shard by shard (sum: rel 1e-5;
var: rel 1e-5), times with CUDA events (max over ranks) and prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--log2', type=int, default=32)
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    import cupy_b200 as cp
    from cupy_b200 import distributed as cdist
    comm = cdist.init_process_group(world, rank, backend='nccl') if world > 1 else None
    n_total = 1 << args.log2
    n = n_total // world
    g = torch.Generator(device='cuda')
    g.manual_seed(77 + rank)
    tx = torch.empty(n, device='cuda', dtype=torch.float32)
    step = 1 << 28
    for lo in range(0, n, step):        # generate in slices: no 2x temporary
        tx[lo:lo + step] = torch.rand(min(step, n - lo), device='cuda', dtype=torch.float32, generator=g) * 2 - 1
    x = cp.from_torch(tx)

    # float64 references, shard by shard
    ref = torch.zeros(3, device='cuda', dtype=torch.float64)   # n, sum, sumsq about 0
    for lo in range(0, n, step):
        c = tx[lo:lo + step].double()
        ref[0] += c.numel(); ref[1] += c.sum(); ref[2] += (c * c).sum()
    if world > 1:
        dist.all_reduce(ref)
    want_sum = float(ref[1])
    mean = float(ref[1] / ref[0])
    want_var = float(ref[2] / ref[0] - mean * mean)

    class Solo:
        def all_reduce(self, a, b, op='sum'):
            pass
    c = comm if comm is not None else Solo()

    def f_sum():
        return cdist.sharded_sum(x, c)

    def f_var():
        return cdist.sharded_var(x, c)

    got_sum = float(f_sum().get())
    got_var = float(f_var().item())
    ok_sum = abs(got_sum - want_sum) <= 1e-5 * np.sqrt(n_total) * 1.0 + 1e-5 * abs(want_sum)
    ok_var = abs(got_var - want_var) <= 1e-5 * abs(want_var)

    def timeit(f):
        for _ in range(3):
            f()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            f()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_sum, ms_var = timeit(f_sum), timeit(f_var)
    if rank == 0:
        print(json.dumps({
            'config': 'c5: 1-D float32 2^%d elements sharded over %d GPU(s)' % (args.log2, world),
            'n_gpus': world, 'elements_per_gpu': n,
            'sum': {'ms': round(ms_sum, 4), 'GBps_aggregate': round(4.0 * n_total / ms_sum / 1e6, 1), 'got': got_sum,
                    'want_f64': want_sum, 'ok': bool(ok_sum)},
            'var': {'ms': round(ms_var, 4), 'GBps_aggregate': round(4.0 * n_total / ms_var / 1e6, 1), 'got': got_var,
                    'want_f64': want_var, 'ok': bool(ok_var)},
        }), flush=True)
    assert ok_sum and ok_var, (got_sum, want_sum, got_var, want_var)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
