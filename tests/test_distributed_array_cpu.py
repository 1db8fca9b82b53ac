"""CPU tier for cupy_b200.distributed.array: the index arithmetic (the reference's own randomised tests,
tests/cupyx_tests/distributed_tests/test_index_arith.py) and the DistributedArray semantics on a NumPy backend, in a
single-rank world and in a world of two ranks over gloo (real send / recv / broadcast between processes)."""
import math
import os
import random
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def test_extgcd_and_slice_intersection_randomised():
    from cupy_b200.distributed import array as da
    rnd = random.Random(0)
    for _ in range(300):
        a, b = rnd.randint(1, 100), rnd.randint(1, 100)
        g, x = da._extgcd(a, b)
        assert g == math.gcd(a, b) and (g - a * x) % b == 0
    max_value = 100

    def all_indices(s0, s1=slice(None)):
        return set(list(range(max_value))[s0][s1])

    for _ in range(300):
        a_start, b_start = rnd.randint(0, max_value - 1), rnd.randint(0, max_value - 1)
        a = slice(a_start, rnd.randint(a_start + 1, max_value), rnd.randint(1, max_value // 3))
        b = slice(b_start, rnd.randint(b_start + 1, max_value), rnd.randint(1, max_value // 3))
        c = da._slice_intersection(a, b, max_value)
        if c is None:
            assert not (all_indices(a) & all_indices(b))
        else:
            assert all_indices(c) == all_indices(a) & all_indices(b)
            assert all_indices(c) == all_indices(a, da._index_for_subslice(a, c, max_value))


def test_index_map_normalisation_and_2d_map():
    from cupy_b200.distributed import array as da
    m = da._normalize_index_map((9, 9), {1: (slice(1, None), 2), 0: [slice(None, None, 2), (slice(2), slice(2))]})
    assert list(m) == [1, 0]
    assert m[1] == [(slice(1, 9, 1), slice(2, 3, 1))]
    assert m[0] == [(slice(0, 2, 1), slice(0, 2, 1)), (slice(0, 9, 2), slice(0, 9, 1))]     # sorted
    with pytest.raises(IndexError):
        da._normalize_index((4,), (0, 0))
    with pytest.raises(ValueError):
        da._normalize_index((4,), slice(2, 2))
    with pytest.raises(ValueError):
        da._normalize_index((4,), slice(None, None, -1))
    im = da.make_2d_index_map([0, 2, 4], [0, 3, 5], [[{0}, {1}], [{2}, {0, 1}]])
    assert im == {0: [(slice(0, 2), slice(0, 3)), (slice(2, 4), slice(3, 5))],
                  1: [(slice(0, 2), slice(3, 5)), (slice(2, 4), slice(3, 5))],
                  2: [(slice(2, 4), slice(0, 3))]}
    assert da.MAX.identity_of(np.dtype('int8')) == -128 and da.MIN.identity_of(np.dtype('f')) == np.inf
    assert da.SUM.identity_of(np.dtype('q')) == 0 and da.PROD.identity_of(np.dtype('d')) == 1.0


def _user_kernels():
    import _numpy_backend
    _numpy_backend.USER_KERNELS['custom'] = lambda x, y: ((x - y) * (x - y)).astype(np.float32)


def test_semantics_single_rank_numpy_backend():
    import cupy_b200 as cp
    from cupy_b200.distributed import array as da
    import _numpy_backend
    import darray_scenarios
    _user_kernels()
    da._set_backend(_numpy_backend.NumpyBackend())
    try:
        assert darray_scenarios.run_all(cp, da, None, 1) > 10
    finally:
        da._set_backend(None)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    try:
        import cupy_b200 as cp
        from cupy_b200 import distributed as cdist
        from cupy_b200.distributed import array as da
        import _numpy_backend
        import darray_scenarios
        _numpy_backend.USER_KERNELS['custom'] = lambda x, y: ((x - y) * (x - y)).astype(np.float32)
        da._set_backend(_numpy_backend.NumpyBackend())
        comm = cdist.init_process_group(world, rank, backend='gloo')
        n = darray_scenarios.run_all(cp, da, comm, world)
        # the lazy SUM mode: a reduction over the sharded axis moves nothing until REPLICA is asked for
        base = np.arange(64, dtype='q').reshape(8, 8)
        d = da.distributed_array(base, {0: slice(4), 1: slice(4, None)}, comm=comm)
        s = d.sum(axis=0)
        local = s._xp.to_host(s._chunks[0].array)
        ok = bool((local == base[4 * rank:4 * rank + 4].sum(axis=0)).all()) and s.mode is da.SUM
        r = s.change_mode(da.REPLICA)
        ok = ok and bool((r._xp.to_host(r._chunks[0].array) == base.sum(axis=0)).all())
        q.put((rank, n, ok, ''))
        comm.barrier()
        comm.stop()
    except Exception as e:                # surface the failure instead of a queue timeout
        import traceback
        q.put((rank, 0, False, traceback.format_exc()[-1500:] + repr(e)))


def test_semantics_world_of_two_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, n, ok, err in res:
        assert ok and n > 10, err
