"""cupy_b200 -- a B200-native (sm_100a) implementation of CuPy's data-parallel
engine: ufunc / ElementwiseKernel launcher, ReductionKernel / axis reductions and
the cumsum scan, behind CuPy's own Python surface.  See DESIGN.md.

Importing this package loads libcupy_b200.so and fails loudly if it is missing:
there is no CPU or NumPy fallback.
"""
from cupy_b200 import _lib  # noqa: F401  (loads the C-ABI library; ImportError if absent)
from cupy_b200._lib import B200Error, CompileException  # noqa: F401
from cupy_b200._core._ndarray import (  # noqa: F401
    ndarray, AxisError, empty, empty_like, zeros, zeros_like, ones, ones_like, full, arange,
    asarray, array, asnumpy, from_torch, from_cuda_array_interface, empty_pinned)
from cupy_b200._core._kernel import (  # noqa: F401
    ElementwiseKernel, ufunc, create_ufunc, elementwise_copy, may_share_bounds)
from cupy_b200._core._reduction import ReductionKernel, create_reduction_func  # noqa: F401
from cupy_b200._core._routines_math import (  # noqa: F401
    add, subtract, multiply, true_divide, divide, negative, absolute, square, sqrt, exp, log,
    expm1, exp2, log2, log10, log1p, sin, cos, tan, tanh, sinh, cosh, arctan2, hypot,
    maximum, minimum, power, fma, greater, greater_equal, less, less_equal, equal, not_equal,
    sum, prod, cumsum, cumprod)
from cupy_b200._core._routines_binary import (  # noqa: F401
    bitwise_and, bitwise_or, bitwise_xor, bitwise_not, invert, left_shift, right_shift)
from cupy_b200._core._routines_statistics import (  # noqa: F401
    amax, amin, argmax, argmin, mean, var, std)

from cupy_b200._core._routines_more import (  # noqa: F401,E402
    all, any, count_nonzero, nansum, nanprod, nanmin, nanmax, nanargmin, nanargmax, ptp,
    nanmean, nanvar, nanstd, nancumsum, nancumprod, average)
from cupy_b200._core._routines_elementwise import (  # noqa: F401,E402
    arcsin, arccos, arctan, arcsinh, arccosh, arctanh, deg2rad, rad2deg, radians, degrees,
    logaddexp, logaddexp2, rint, floor, ceil, trunc, fix, around, round, round_,
    reciprocal, positive, floor_divide, remainder, mod, divmod, fmod, modf, float_power,
    signbit, copysign, nextafter, ldexp, frexp, cbrt, fabs, sign, heaviside, fmax, fmin,
    clip, nan_to_num, gcd, lcm, logical_and, logical_or, logical_not, logical_xor,
    isfinite, isinf, isnan, isneginf, isposinf, isclose, allclose, array_equal, where)
from cupy_b200._core._compaction import (  # noqa: F401,E402
    nonzero, argwhere, flatnonzero, compress, extract, take)
from cupy_b200 import cuda  # noqa: F401,E402
from cupy_b200._core.fusion import fuse  # noqa: F401,E402

from cupy_b200._core._accelerator import (  # noqa: F401,E402
    set_routine_accelerators, set_reduction_accelerators, get_routine_accelerators, get_reduction_accelerators)

abs = absolute
max = amax
min = amin

__version__ = '0.1.0'
