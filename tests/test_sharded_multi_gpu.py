"""GPU tier, needs >= 2 GPUs (skipped on a single-GPU box; `bench.py --gpus N` carries the same checks in its `c5` and
`darray` records on the driver's scaling run): the fused cross-GPU combine of sharded sum / var
(b200_reduce_run_sharded over NVLink peer memory) against float64 references and against the NCCL route, and a
DistributedArray round trip over NCCL send / recv."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
world, rank = int(os.environ['WORLD_SIZE']), int(os.environ['RANK'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
import cupy_b200 as cp
from cupy_b200 import distributed as cdist
from cupy_b200.distributed import array as da
comm = cdist.init_process_group(world, rank, backend='nccl')
assert comm.peer_exchange() is not None, 'symmetric memory unavailable'
g = torch.Generator(device='cuda'); g.manual_seed(100 + rank)
n = (1 << 22) + 12345
t = torch.rand(n, device='cuda', generator=g) * 2 - 1
x = cp.from_torch(t)
ref = torch.stack([t.double().sum(), (t.double() ** 2).sum()]); dist.all_reduce(ref)
tot = float(ref[0]); mean = tot / (n * world); var = float(ref[1]) / (n * world) - mean * mean
for rep in range(5):                       # tags alternate parity: several calls in a row
    s = float(cdist.sharded_sum(x, comm).get())
    v = float(cdist.sharded_var(x, comm).get())
    assert abs(s - tot) <= 1e-5 * n * world, (s, tot)
    assert abs(v - var) <= 1e-5 * var, (v, var)
both = torch.tensor([s, v], device='cuda', dtype=torch.float64)
lo, hi = both.clone(), both.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert bool((lo == hi).all()), 'results differ between ranks'
xi = cp.asarray(np.arange(1000, dtype=np.int64) * (rank + 1))
si = int(cdist.sharded_sum(xi, comm).get())
assert si == 499500 * world * (world + 1) // 2, si
comm._exchange = False                     # the NCCL route gives the same numbers within tolerance
s2 = float(cdist.sharded_sum(x, comm).get()); v2 = float(cdist.sharded_var(x, comm).get())
assert abs(s2 - s) <= 1e-5 * n * world and abs(v2 - v) <= 1e-5 * var
base = np.arange(64 * world * 48, dtype=np.float32).reshape(64 * world, 48)
d = da.distributed_array(base, {r: slice(64 * r, 64 * (r + 1)) for r in range(world)}, comm=comm)
sm = d.sum(axis=0)
assert sm.mode is da.SUM and np.allclose(sm.get(), base.sum(axis=0))
cols = {r: (slice(None), slice(48 * r // world, 48 * (r + 1) // world)) for r in range(world)}
assert np.array_equal((d * d).reshard(cols).get(), base * base)
dist.barrier()
if rank == 0:
    print('sharded multi-gpu ok')
dist.destroy_process_group()
''' % ROOT


def test_fused_sharded_reductions_and_distributed_array_over_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29517', str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'sharded multi-gpu ok' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
