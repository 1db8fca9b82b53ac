"""TEST INFRASTRUCTURE: a NumPy look-alike of the array backend `cupy_b200.distributed.array` runs on, so that
the host logic of DistributedArray (index arithmetic, modes, resharding, the chunk-pair transfer order over a
real gloo process group) is checked on the CPU tier.  The product backend is the engine (cupy_b200 arrays on a
GPU); this one is injected with `array._set_backend` by tests only."""
import numpy
import torch

UFUNCS = {'cupy_add': numpy.add, 'cupy_subtract': numpy.subtract, 'cupy_multiply': numpy.multiply,
          'cupy_cos': numpy.cos, 'cupy_negative': numpy.negative, 'cupy_maximum': numpy.maximum,
          'cupy_minimum': numpy.minimum, 'cupy_true_divide': numpy.true_divide}
REDUCTIONS = {'cupy_sum': numpy.sum, 'cupy_prod': numpy.prod, 'cupy_max': numpy.max, 'cupy_min': numpy.min}
USER_KERNELS = {}     # ElementwiseKernel name -> NumPy statement of it, registered by the test


class NumpyBackend:
    name = 'numpy (tests)'
    ndarray = numpy.ndarray

    def empty(self, shape, dtype):
        return numpy.empty(shape, dtype)

    def full(self, shape, value, dtype):
        return numpy.full(shape, value, dtype)

    def from_host(self, a):
        return numpy.array(a, copy=True)

    def to_host(self, a):
        return numpy.array(a, copy=True)

    def contiguous(self, a):
        return numpy.ascontiguousarray(a)

    def copy(self, a):
        return a.copy()

    def assign(self, a, idx, value):
        a[idx] = value

    def combine(self, func_name, a, idx, value):
        a[idx] = getattr(numpy, func_name)(a[idx], value)

    def wire(self, a):
        assert a.flags.c_contiguous
        return torch.from_numpy(a)

    def run_elementwise(self, kernel, arrays, kwargs):
        f = UFUNCS.get(kernel.name) or USER_KERNELS[kernel.name]
        return numpy.asarray(f(*arrays, **kwargs))

    def run_reduction(self, kernel, array, axis, dtype):
        return numpy.asarray(REDUCTIONS[kernel.name](array, axis=axis, dtype=dtype) if kernel.name in ('cupy_sum', 'cupy_prod')
                             else REDUCTIONS[kernel.name](array, axis=axis))
