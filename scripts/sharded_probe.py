"""Per-GPU 2^29 float32: local sum / var (no exchange) vs the fused sharded calls vs the NCCL route, same process group.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/sharded_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
import cupy_b200 as cp  # noqa: E402
from cupy_b200 import distributed as cdist  # noqa: E402

comm = cdist.init_process_group(world, rank, backend='nccl')
n = 1 << 29
t = torch.empty(n, device='cuda')
for lo in range(0, n, 1 << 28):
    t[lo:lo + (1 << 28)] = torch.rand(1 << 28, device='cuda') * 2 - 1
x = cp.from_torch(t)


def timeit(f, iters=30):
    for _ in range(5):
        f()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device='cuda', dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


res = {}
res['local sum'] = timeit(lambda: x.sum())
res['local var'] = timeit(lambda: x.var())
assert comm.peer_exchange() is not None
res['fused sum'] = timeit(lambda: cdist.sharded_sum(x, comm))
res['fused var'] = timeit(lambda: cdist.sharded_var(x, comm))
comm._exchange = False          # NCCL route
res['nccl sum'] = timeit(lambda: cdist.sharded_sum(x, comm))
res['nccl var'] = timeit(lambda: cdist.sharded_var(x, comm))
if rank == 0:
    for k, v in res.items():
        print('N=%d  %-10s %.4f ms' % (world, k, v), flush=True)
dist.barrier()
dist.destroy_process_group()
