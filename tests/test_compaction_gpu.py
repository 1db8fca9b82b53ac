"""GPU tier: the scan's callers -- nonzero / argwhere / flatnonzero / where(condition), boolean-mask and
integer-array indexing, take / compress / extract (cupy_b200/_core/_compaction.py) -- bit-exact against NumPy,
the reference's own oracle for them (tests/cupy_tests/sorting_tests/test_search.py `TestNonzero`, `TestFlatNonzero`,
`TestArgwhere`; core_tests/test_ndarray_adv_indexing.py; indexing_tests/test_indexing.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RS = np.random.RandomState(23)


@pytest.fixture(scope='module')
def cp():
    import cupy_b200
    return cupy_b200


def sparse(shape, dt, density=0.3):
    a = (RS.rand(*shape) * 9 + 1).astype(dt)
    a[RS.rand(*shape) >= density] = 0
    return a


@pytest.mark.parametrize('dt', ['?', 'int8', 'int32', 'float16', 'float32', 'float64'])
@pytest.mark.parametrize('shape', [(1,), (7,), (4097,), (100003,), (33, 65), (5, 6, 7), (3, 1, 4, 5), (1 << 20,)])
def test_nonzero_family(cp, shape, dt):
    a = sparse(shape, dt)
    d = cp.asarray(a)
    got = cp.nonzero(d)
    want = np.nonzero(a)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.dtype == np.int64
        np.testing.assert_array_equal(g.get(), w)
    np.testing.assert_array_equal(cp.argwhere(d).get(), np.argwhere(a))
    np.testing.assert_array_equal(cp.flatnonzero(d).get(), np.flatnonzero(a))
    for g, w in zip(cp.where(d), np.where(a)):
        np.testing.assert_array_equal(g.get(), w)
    for g, w in zip(d.nonzero(), a.nonzero()):
        np.testing.assert_array_equal(g.get(), w)


def test_nonzero_edge_cases(cp):
    for a in (np.zeros((50,), 'f'), np.ones((50,), 'f'), np.zeros((0,), 'f'), np.zeros((4, 0, 3), 'i'),
              np.array([0.0, -0.0, np.nan, np.inf], 'f')):
        for g, w in zip(cp.nonzero(cp.asarray(a)), np.nonzero(a)):
            assert g.shape == w.shape
            np.testing.assert_array_equal(g.get(), w)
        assert cp.argwhere(cp.asarray(a)).shape == np.argwhere(a).shape
    t = sparse((40, 50), 'float32')
    np.testing.assert_array_equal(cp.flatnonzero(cp.asarray(t).T).get(), np.flatnonzero(t.T))          # a view: C order of the VIEW
    np.testing.assert_array_equal(cp.argwhere(cp.asarray(t)[::2, 1::3]).get(), np.argwhere(t[::2, 1::3]))
    with pytest.raises(ValueError):
        cp.nonzero(cp.asarray(np.float32(3)))


@pytest.mark.parametrize('dt', ['int8', 'int64', 'float16', 'float32', 'float64'])
def test_boolean_mask_indexing(cp, dt):
    a = (RS.rand(37, 21, 5) * 100).astype(dt)
    d = cp.asarray(a)
    m3 = RS.rand(37, 21, 5) > 0.6
    m2 = RS.rand(37, 21) > 0.5
    m1 = RS.rand(37) > 0.5
    for m in (m3, m2, m1):
        got = d[cp.asarray(m)]
        assert got.dtype == a.dtype
        np.testing.assert_array_equal(got.get(), a[m])
        np.testing.assert_array_equal(d[m].get(), a[m])                         # a host mask is accepted as an index
    np.testing.assert_array_equal(d.T[cp.asarray(m3.T)].get(), a.T[m3.T])      # strided source
    # comparisons feed it directly
    np.testing.assert_array_equal(d[d > 50].get(), a[a > 50])
    assert d[cp.asarray(np.zeros((37, 21, 5), bool))].shape == (0,)
    assert d[cp.asarray(np.zeros((37,), bool))].shape == (0, 21, 5)
    with pytest.raises(IndexError):
        d[cp.asarray(np.zeros((36,), bool))]
    # assignment: scalar, broadcast row, one value per hit
    for m, v in ((m3, 7), (m2, np.arange(5).astype(dt)), (m1, 3), (m3, np.arange(int(m3.sum())).astype(dt)),
                 (m1, (RS.rand(int(m1.sum()), 21, 5) * 50).astype(dt))):
        want = a.copy()
        want[m] = v
        dd = cp.asarray(a)
        dd[cp.asarray(m)] = cp.asarray(v) if isinstance(v, np.ndarray) else v
        np.testing.assert_array_equal(dd.get(), want)
    dd = cp.asarray(a)
    dd[dd > 50] = 50
    np.testing.assert_array_equal(dd.get(), np.minimum(a, 50))
    with pytest.raises(ValueError):
        dd[cp.asarray(m3)] = cp.asarray(np.arange(3).astype(dt))


@pytest.mark.parametrize('dt', ['int32', 'float16', 'float32', 'float64'])
def test_take_and_integer_indexing(cp, dt):
    a = (RS.rand(23, 17, 9) * 100).astype(dt)
    d = cp.asarray(a)
    for axis in (None, 0, 1, 2, -1):
        n = a.size if axis is None else a.shape[axis]
        for ishape in ((), (5,), (3, 4)):
            idx = RS.randint(-n, n, size=ishape)
            got = cp.take(d, cp.asarray(idx), axis=axis)
            want = np.take(a, idx, axis=axis)
            assert got.dtype == want.dtype and got.shape == want.shape
            np.testing.assert_array_equal(got.get(), want)
    np.testing.assert_array_equal(d.take(cp.asarray(np.array([1, 2, 30])), axis=0).get(), np.take(a, [1, 2, 30], axis=0, mode='wrap'))
    np.testing.assert_array_equal(cp.take(d.transpose(2, 0, 1), cp.asarray(np.array([0, 3])), axis=1).get(),
                                  np.take(a.transpose(2, 0, 1), [0, 3], axis=1))
    out = cp.empty((23, 4, 9), dt)
    assert cp.take(d, cp.asarray(np.array([0, 5, 5, 16])), axis=1, out=out) is out
    np.testing.assert_array_equal(out.get(), a[:, [0, 5, 5, 16]])
    i0, i1 = RS.randint(-23, 23, size=(11,)), RS.randint(-17, 17, size=(11,))
    np.testing.assert_array_equal(d[cp.asarray(i0)].get(), a[i0])
    np.testing.assert_array_equal(d[cp.asarray(i0), cp.asarray(i1)].get(), a[i0, i1])
    np.testing.assert_array_equal(d[[1, 4, -2]].get(), a[[1, 4, -2]])
    np.testing.assert_array_equal(d[cp.asarray(i0.reshape(11, 1)), cp.asarray(i1[:4])].get(), a[i0.reshape(11, 1), i1[:4]])
    # nonzero's output indexes the array it came from
    s = a.copy()
    s[s < 60] = 0
    ds = cp.asarray(s)
    np.testing.assert_array_equal(ds[cp.nonzero(ds)].get(), s[np.nonzero(s)])
    # assignment through unique indices
    u0 = RS.permutation(23)[:9]
    want = a.copy()
    want[u0] = 1
    dd = cp.asarray(a)
    dd[cp.asarray(u0)] = 1
    np.testing.assert_array_equal(dd.get(), want)
    u1 = RS.permutation(17)[:9]
    v = (RS.rand(9, 9) * 10).astype(dt)
    want = a.copy()
    want[u0, u1] = v
    dd = cp.asarray(a)
    dd[cp.asarray(u0), cp.asarray(u1)] = cp.asarray(v)
    np.testing.assert_array_equal(dd.get(), want)
    with pytest.raises(IndexError):
        d[cp.asarray(np.array([0.5]))]
    with pytest.raises(NotImplementedError):
        d[:, cp.asarray(i1)]


def test_compress_extract(cp):
    a = (RS.rand(19, 31) * 100).astype('float32')
    d = cp.asarray(a)
    c0, c1 = RS.rand(19) > 0.5, RS.rand(31) > 0.5
    np.testing.assert_array_equal(cp.compress(cp.asarray(c0), d, axis=0).get(), np.compress(c0, a, axis=0))
    np.testing.assert_array_equal(cp.compress(cp.asarray(c1), d, axis=1).get(), np.compress(c1, a, axis=1))
    np.testing.assert_array_equal(d.compress(cp.asarray(c1[:20]), axis=1).get(), a.compress(c1[:20], axis=1))
    cf = RS.rand(19 * 31) > 0.7
    np.testing.assert_array_equal(cp.compress(cp.asarray(cf), d).get(), np.compress(cf, a))
    np.testing.assert_array_equal(cp.compress(cp.asarray(np.array([0, 2, 0, 1])), d, axis=0).get(), np.compress([0, 2, 0, 1], a, axis=0))
    m = RS.rand(19, 31) > 0.5
    np.testing.assert_array_equal(cp.extract(cp.asarray(m), d).get(), np.extract(m, a))
    np.testing.assert_array_equal(cp.extract(d > 50, d).get(), np.extract(a > 50, a))


def test_mask_rank_is_the_int32_scan(cp):
    """cumsum(bool, dtype=int32): the prebuilt pair the compaction ranks with (int64 past 2^31 - 1 elements)."""
    for n in (1, 4097, (1 << 22) + 3):
        m = RS.rand(n) > 0.5
        got = cp.cumsum(cp.asarray(m), dtype=np.int32)
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got.get(), np.cumsum(m, dtype=np.int32))
    m2 = RS.rand(300, 257) > 0.5
    for ax in (0, 1):
        np.testing.assert_array_equal(cp.cumsum(cp.asarray(m2), axis=ax, dtype=np.int32).get(), np.cumsum(m2, axis=ax, dtype=np.int32))


def test_compaction_full_size(cp):
    """2^28 float32: every hit lands in order (sortedness + checksum: size-independent properties)."""
    import torch
    n = 1 << 28
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    t = torch.rand(n, device='cuda', generator=g)
    x = cp.from_torch(t)
    idx = cp.flatnonzero(x > 0.75)
    ti = idx.to_torch()
    want = int((t > 0.75).sum())
    assert ti.shape == (want,)
    assert bool((ti[1:] > ti[:-1]).all())                                   # strictly increasing: C order, no duplicates
    assert bool((t[ti] > 0.75).all())                                       # every index is a hit
    sel = x[x > 0.75]
    assert sel.shape == (want,) and bool(torch.equal(sel.to_torch(), t[t > 0.75]))
    del t, x, idx, ti, sel
    torch.cuda.empty_cache()
