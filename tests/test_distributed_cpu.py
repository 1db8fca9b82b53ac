"""CPU tier: the N > 1 host logic with world_size-2 `gloo` process groups (no GPU): the
NCCLBackend wrapper's collectives and the rank-ordered moment merge behind sharded var."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    from cupy_b200 import distributed as cdist
    comm = cdist.init_process_group(world, rank, backend='gloo')
    # all_reduce of arange(24) == world * arange (tests/cupyx_tests/distributed_tests/comm_runner.py:88-107)
    t = torch.arange(24, dtype=torch.float32)
    out = torch.zeros(24)
    comm.all_reduce(t, out, 'sum')
    ok = bool((out == world * torch.arange(24)).all())
    mx = torch.tensor([float(rank)])
    comm.all_reduce(mx, mx, 'max')
    ok = ok and float(mx) == world - 1
    b = torch.full((4,), float(rank))
    comm.broadcast(b, root=1)
    ok = ok and bool((b == 1).all())
    # the other collectives of the reference's NCCLBackend (cupyx/distributed/_nccl_comm.py:187-295;
    # expectations follow tests/cupyx_tests/distributed_tests/comm_runner.py)
    g_in = torch.full((3,), float(rank + 1))
    g_out = torch.zeros(world, 3)
    comm.gather(g_in, g_out, root=0)
    if rank == 0:
        ok = ok and bool((g_out == torch.arange(1, world + 1, dtype=torch.float32)[:, None]).all())
    s_in = torch.arange(world * 2, dtype=torch.float32).reshape(world, 2) + 100 * rank
    s_out = torch.zeros(2)
    comm.scatter(s_in, s_out, root=1)
    ok = ok and bool((s_out == torch.arange(2) + 2 * rank + 100).all())
    ag_out = torch.zeros(world * 3)
    comm.all_gather(g_in, ag_out, 3)
    ok = ok and bool((ag_out.reshape(world, 3) == torch.arange(1, world + 1, dtype=torch.float32)[:, None]).all())
    rs_in = torch.arange(world * 2, dtype=torch.float32)
    rs_out = torch.zeros(2)
    try:
        comm.reduce_scatter(rs_in, rs_out, 2)
        ok = ok and bool((rs_out == world * (torch.arange(2) + 2 * rank)).all())
    except RuntimeError:
        pass                      # gloo builds without reduce_scatter: NCCL has it
    a2a_in = torch.arange(world, dtype=torch.float32).reshape(world, 1) + 10 * rank
    a2a_out = torch.zeros(world, 1)
    try:
        comm.all_to_all(a2a_in, a2a_out)
        ok = ok and bool((a2a_out.reshape(-1) == 10 * torch.arange(world) + rank).all())
    except RuntimeError:
        pass
    peer = (rank + 1) % world
    sr_out = torch.zeros(2)
    comm.send_recv(torch.full((2,), float(rank)), sr_out, peer)
    ok = ok and bool((sr_out == float(peer)).all()) if world == 2 else ok
    for bad in (lambda: comm.scatter(torch.zeros(world + 1, 2), s_out), lambda: comm.gather(g_in, torch.zeros(world + 1, 3))):
        try:
            bad()
            ok = False
        except RuntimeError:
            pass
    # sharded var merge: every rank folds the gathered (n, mean, M2) in rank order
    rs = np.random.RandomState(0)
    full = rs.rand(1000 * world + 37)
    lo = rank * 1000
    hi = (rank + 1) * 1000 if rank < world - 1 else full.size
    shard = full[lo:hi]
    loc = torch.tensor([float(shard.size), shard.mean(), shard.var() * shard.size], dtype=torch.float64)
    allv = torch.empty(world * 3, dtype=torch.float64)
    import torch.distributed as dist
    dist.all_gather_into_tensor(allv, loc)
    allv = allv.reshape(world, 3)
    n, m, m2 = cdist.combine_moments([allv[k, 0] for k in range(world)], [allv[k, 1] for k in range(world)],
                                     [allv[k, 2] for k in range(world)])
    q.put((rank, ok, float(n), float(m), float(m2 / n), float(full.mean()), float(full.var())))
    comm.barrier()
    comm.stop()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    for rank, ok, n, m, v, fm, fv in res:
        assert ok
        assert n == 2037
        assert abs(m - fm) < 1e-12 and abs(v - fv) < 1e-12
    assert res[0][2:5] == res[1][2:5]          # bit-identical on every rank


def test_combine_moments_matches_numpy_and_handles_empty_shards():
    from cupy_b200.distributed import combine_moments
    rs = np.random.RandomState(1)
    parts = [rs.rand(k) * 10 for k in (5, 0, 1, 1000)]
    c = [torch.tensor(float(p.size), dtype=torch.float64) for p in parts]
    m = [torch.tensor(float(p.mean()) if p.size else 0.0, dtype=torch.float64) for p in parts]
    s = [torch.tensor(float(p.var() * p.size) if p.size else 0.0, dtype=torch.float64) for p in parts]
    n, mean, m2 = combine_moments(c, m, s)
    full = np.concatenate(parts)
    assert float(n) == full.size
    assert abs(float(mean) - full.mean()) < 1e-12 and abs(float(m2 / n) - full.var()) < 1e-11


def test_init_process_group_argument_checks():
    from cupy_b200 import distributed as cdist
    with pytest.raises(ValueError, match='Invalid number of devices'):
        cdist.init_process_group(0, 0)
    with pytest.raises(ValueError, match='Invalid number of rank'):
        cdist.init_process_group(2, 2)
    with pytest.raises(ValueError, match='not supported'):
        cdist.init_process_group(2, 0, backend='mpi')
