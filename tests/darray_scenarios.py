"""Scenarios for DistributedArray shared by the CPU tier (NumPy backend, gloo world of 2) and the GPU tier (engine
backend).  They restate the reference's tests/cupyx_tests/distributed_tests/test_array_nccl.py:26-270 -- same
shapes, index maps and expectations -- for a world of `world` ranks.  `cp` is cupy_b200, `da` its
distributed.array module, `comm` the communicator (None in a single-rank world)."""
import numpy as np

size = 256
_2d = [
    {0: [(slice(8), slice(None, None))], 1: [(slice(8, None), slice(None, None))]},
    {0: [(slice(8), slice(4)), (slice(8), slice(4, None))], 1: [(slice(8, None), slice(None, None))]},
    # overlapping chunks and a strided chunk (cf. the reference docs' example, _array.py:838-844), covering the array
    {0: [(slice(12), slice(None)), slice(None, None, 2)], 1: [(slice(6, None), slice(5, None)), (slice(10, None), slice(8))]},
]
_3d = [
    {0: [(slice(4), slice(None, None), slice(None, None))], 1: [(slice(4, None), slice(None, None), slice(None, None, None))]},
    {0: [(slice(4), slice(4), slice(None, None)), (slice(4), slice(4, None), slice(None, None))],
     1: [(slice(4, None), slice(None, None), slice(None, None, None))]},
]


def fold(index_map, world):
    """The reference's maps name devices 0 and 1; a world of one rank keeps every chunk on rank 0."""
    out = {}
    for dev, idxs in index_map.items():
        idxs = idxs if isinstance(idxs, list) else [idxs]
        out.setdefault(dev % world, []).extend(idxs)
    return out


def run_all(cp, da, comm, world):
    REPLICA, SUM, MAX, MIN, PROD = da.REPLICA, da.SUM, da.MAX, da.MIN, da.PROD
    cases = [((16, 16), fold(m, world)) for m in _2d] + [((8, 8, 4), fold(m, world)) for m in _3d]
    custom = cp.ElementwiseKernel('float32 x, float32 y', 'float32 z', 'z = (x - y) * (x - y)', 'custom')
    n = 0
    for shape, imap in cases:
        base = np.arange(size, dtype='q').reshape(shape)
        for mode in (REPLICA, SUM, MAX):
            # creation + get + chunk contents (test_array_nccl.py:26-66)
            d = da.distributed_array(base, imap, mode, comm=comm)
            assert d.shape == shape and d.mode is mode
            np.testing.assert_array_equal(d.get(), base)
            if mode is REPLICA:
                norm = da._normalize_index_map(shape, imap)
                for c, idx in zip(d._chunks, norm.get(d.rank, [])):
                    np.testing.assert_array_equal(d._xp.to_host(c.array), base[idx])
            # change_mode round trips (:68-100)
            for target in (REPLICA, SUM, MAX, MIN):
                e = d.change_mode(target)
                assert e.mode is target
                np.testing.assert_array_equal(e.get(), base)
                np.testing.assert_array_equal(d.get(), base)
            # reductions stay in their op mode; values only meet on get() / change_mode (:214-250)
            for axis in range(len(shape)):
                s = d.sum(axis=axis)
                assert s.mode is SUM
                np.testing.assert_array_equal(s.get(), base.sum(axis=axis))
                np.testing.assert_array_equal(s.change_mode(REPLICA).get(), base.sum(axis=axis))
                np.testing.assert_array_equal(d.max(axis=axis).get(), base.max(axis=axis))
                np.testing.assert_array_equal(d.min(axis=axis).get(), base.min(axis=axis))
            n += 1
        # ufuncs and user kernels on arrays in different modes (:102-126)
        a = np.arange(size).reshape(shape)
        b = a * 2
        for ma, mb in ((REPLICA, REPLICA), (SUM, REPLICA), (MAX, SUM)):
            d_a = da.distributed_array(a, imap, ma, comm=comm)
            d_b = da.distributed_array(b, imap, mb, comm=comm)
            r = cp.multiply(d_a, d_b)
            np.testing.assert_array_equal(r.get(), a * b)
            assert r.mode is REPLICA
            np.testing.assert_array_equal((d_a + d_b).get(), a + b)
            fa, fb = a.astype(np.float32), b.astype(np.float32)
            r = custom(da.distributed_array(fa, imap, ma, comm=comm), da.distributed_array(fb, imap, mb, comm=comm))
            np.testing.assert_allclose(r.get(), (fa - fb) * (fa - fb))
        # prod reduction (:252-259)
        rs = np.random.RandomState(7)
        p = rs.rand(*shape) + 0.5
        d_p = da.distributed_array(p, imap, REPLICA, comm=comm)
        for axis in range(len(shape)):
            np.testing.assert_allclose(d_p.prod(axis=axis).get(), p.prod(axis=axis), rtol=1e-6)
    # resharding between every pair of 2-D maps, all modes (:183-199), and operands with different maps (:201-212)
    shape = (16, 16)
    base = np.arange(size, dtype='q').reshape(shape)
    maps = [fold(m, world) for m in _2d]
    for ia in maps:
        for ib in maps:
            for mode in (REPLICA, SUM, MAX):
                d_a = da.distributed_array(base, ia, mode, comm=comm)
                d_b = d_a.reshard(ib)
                assert d_b.mode is mode
                np.testing.assert_array_equal(d_b.get(), base)
                np.testing.assert_array_equal(d_a.get(), base)
                if mode is REPLICA:
                    for c, idx in zip(d_b._chunks, da._normalize_index_map(shape, ib).get(d_b.rank, [])):
                        np.testing.assert_array_equal(d_b._xp.to_host(c.array), base[idx])
            d_a = da.distributed_array(base, ia, REPLICA, comm=comm)
            d_b = da.distributed_array(base * 2, ib, SUM, comm=comm)
            np.testing.assert_array_equal((d_a + d_b).get(), base * 3)          # maps differ: b is resharded
            d_c = d_a + d_b.reshard(ia)
            np.testing.assert_array_equal(d_c.reshard(ib).max(axis=0).get(), (base * 3).max(axis=0))
            n += 1
    # (a * b).max(axis=0) * c with c laid out like the reduced map (:261-277)
    rs = np.random.RandomState(11)
    A, B, C = rs.randint(0, 1 << 10, shape), rs.randint(0, 1 << 10, shape), rs.randint(0, 1 << 10, shape[1:])
    d_a, d_b = da.distributed_array(A, maps[0], comm=comm), da.distributed_array(B, maps[1], comm=comm)
    map_c = {dev: [idx[1:] for idx in idxs] for dev, idxs in d_a.index_map.items()}
    d_c2 = (d_a.reshard(maps[1]) * d_b).max(axis=0)
    np.testing.assert_array_equal(d_c2.get(), (A * B).max(axis=0))
    d_d = d_c2.reshard(map_c) * da.distributed_array(C, map_c, comm=comm)
    np.testing.assert_array_equal(d_d.get(), (A * B).max(axis=0) * C)
    # errors (:176-181, :261)
    d = da.distributed_array(base, maps[0], comm=comm)
    for bad in (lambda: cp.argmax(d, axis=0), lambda: d[0], lambda: d.var()):
        try:
            bad()
            raise AssertionError('expected an error')
        except (RuntimeError, NotImplementedError):
            pass
    try:
        d.sum()
        raise AssertionError('axis=None must be rejected')
    except RuntimeError:
        pass
    return n
