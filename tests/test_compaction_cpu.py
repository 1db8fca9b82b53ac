"""CPU tier for cupy_b200/_core/_compaction.py (dry-run: host logic + NVRTC compile, nothing launched; the one
device read of a compaction -- the last rank -- is stubbed)."""
import numpy as np
import pytest

import cupy_b200 as cp
from cupy_b200._core import _ndarray


def names(log):
    return [d['name'].split('__')[0] if 'name' in d else '%s:%s->%s' % (d['kind'], d['in_dtype'], d['out_dtype']) for d in log]


RANK = 'prebuilt_scan:bool->int32'          # the flags are ranked by the prebuilt bool -> int32 scan


@pytest.fixture
def hits(monkeypatch):
    """Pretend every ranking scan ended at 100."""
    monkeypatch.setattr(_ndarray.ndarray, 'item', lambda self: 100)
    return 100


def test_nonzero_routes(dry, hits):
    f = cp.empty((30, 20), 'f')
    del dry[:]
    r = cp.nonzero(f)
    assert len(r) == 2 and all(x.shape == (hits,) and x.dtype == np.int64 for x in r)
    # flags, their int32 rank (prebuilt bool -> int32 scan), one scatter of the coordinates
    assert names(dry) == ['cupy_not_equal', RANK, 'cupy_nonzero_kernel']
    assert 'CIndexer<2>' in dry[-1]['source']                 # the un-collapsed shape: coordinates, not a flat index
    del dry[:]
    assert cp.flatnonzero(cp.empty((600,), '?')).shape == (hits,)
    assert names(dry) == [RANK, 'cupy_nonzero_1d']     # boolean input: no flag pass
    assert cp.argwhere(f).shape == (hits, 2)
    assert cp.where(f)[1].shape == (hits,) and f.nonzero()[0].shape == (hits,)
    assert cp.nonzero(cp.empty((0, 3), 'f'))[0].shape == (0,)
    with pytest.raises(ValueError):
        cp.nonzero(cp.empty((), 'f'))


def test_mask_indexing_routes(dry, hits):
    f = cp.empty((30, 20), 'f')
    m, m1 = cp.empty((30, 20), '?'), cp.empty((30,), '?')
    del dry[:]
    assert f[m].shape == (hits,) and f[m].dtype == np.float32
    assert names(dry)[-2:] == [RANK, 'cupy_getitem_mask']
    assert f[m1].shape == (hits // 20, 20)
    assert f[np.zeros((30,), bool)].shape == (hits // 20, 20)
    del dry[:]
    f[m] = 0
    assert names(dry) == ['cupy_fill_mask']                     # a scalar needs no ranking
    del dry[:]
    f[m] = cp.empty((hits,), 'd')
    assert names(dry)[-2:] == [RANK, 'cupy_setitem_mask']
    f[m1] = cp.empty((20,), 'f')
    f[m1] = cp.empty((hits // 20, 20), 'f')
    with pytest.raises(ValueError):
        f[m] = cp.empty((3,), 'f')
    with pytest.raises(IndexError):
        f[cp.empty((29,), '?')]
    with pytest.raises(IndexError):
        f[cp.empty((30, 20, 2), '?')]


def test_take_routes(dry, hits):
    f = cp.empty((30, 20), 'f')
    i = cp.empty((7,), 'l')
    assert cp.take(f, i).shape == (7,) and cp.take(f, i, axis=1).shape == (30, 7)
    assert f.take(cp.empty((2, 3), 'i'), axis=0).shape == (2, 3, 20)
    assert f[i].shape == (7, 20) and f[i, i].shape == (7,) and f[[1, 2, 3]].shape == (3, 20)
    assert cp.take(f, 3, axis=0).shape == (20,)
    del dry[:]
    f[i] = 1.5
    f[i, i] = cp.empty((7,), 'f')
    assert names(dry) == ['cupy_scatter_update', 'cupy_scatter_fold_index', 'cupy_scatter_fold_index', 'cupy_scatter_update']
    out = cp.empty((30, 7), 'f')
    assert cp.take(f, i, axis=1, out=out) is out
    with pytest.raises(ValueError):
        cp.take(f, i, axis=1, out=cp.empty((30, 8), 'f'))
    with pytest.raises(TypeError):
        cp.take(f, i, axis=1, out=cp.empty((30, 7), 'd'))
    with pytest.raises(IndexError):
        f[cp.empty((3,), 'f')]
    with pytest.raises(NotImplementedError):
        f[:, i]
    with pytest.raises(cp.AxisError):
        cp.take(f, i, axis=2)
    assert cp.compress(cp.empty((30,), '?'), f, axis=0).shape == (hits, 20)
    assert cp.extract(cp.empty((30, 20), '?'), f).shape == (hits,)
    with pytest.raises(ValueError):
        cp.compress(cp.empty((30, 20), '?'), f)
