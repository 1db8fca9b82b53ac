"""GPU tier for cupy_b200.distributed.array on the ENGINE backend (cupy_b200 arrays, ufuncs, ElementwiseKernel and
reduction kernels through the __cupy_override_* hooks): the reference's DistributedArray scenarios
(tests/cupyx_tests/distributed_tests/test_array_nccl.py) in a single-rank world -- several, overlapping and strided
chunks on one GPU.  The multi-rank transfers are covered over gloo on the CPU tier and over NCCL by
`bench.py --gpus N` (record `darray`)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_reference_scenarios_on_the_engine_backend():
    import cupy_b200 as cp
    from cupy_b200.distributed import array as da
    import darray_scenarios
    assert da._get_backend().name == 'cupy_b200'
    assert darray_scenarios.run_all(cp, da, None, 1) > 10


def test_chunks_are_device_arrays_and_reductions_run_engine_kernels():
    import cupy_b200 as cp
    from cupy_b200.distributed import array as da
    base = np.arange(4096, dtype=np.float32).reshape(64, 64)
    d = da.distributed_array(cp.asarray(base), {0: [slice(0, 40), slice(24, None)]}, da.REPLICA)
    assert all(isinstance(c, cp.ndarray) for c in d.all_chunks()[0])
    s = d.sum(axis=1)                       # overlapping rows are counted once (REPLICA -> SUM sets identities)
    assert s.mode is da.SUM and isinstance(s.all_chunks()[0][0], cp.ndarray)
    np.testing.assert_allclose(s.get(), base.sum(axis=1), rtol=1e-6)
    m = (d * d).max(axis=0)
    np.testing.assert_array_equal(m.get(), (base * base).max(axis=0))
    with pytest.raises(RuntimeError, match='Mixing'):
        cp.add(d, cp.asarray(base))
