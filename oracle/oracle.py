"""TEST INFRASTRUCTURE, not product code: the CPU oracle.

A NumPy restatement of what the reference computes on the hot path, function by
function, each citing the reference code it follows.  NumPy is the reference's
own stated oracle (every hot-path test of the reference is NumPy-vs-CuPy,
cupy/testing/_loops.py:270-376), so most functions ARE the
NumPy call plus the places where the reference deliberately differs from NumPy
(accumulator types, -ftz / FMA contraction, result dtypes).

PINNING: the oracle is pinned against the known-answer tests of the reference's
own test-suite for this path (tests/test_oracle_pinning.py lists each with its
file:line) and, on a GPU, against the reference's native CUB kernels compiled
from its own sources (oracle/_ref, tests/test_reference_cub.py).  The reference
package itself cannot be imported here (no cupy wheel, no GPU in the build
container), so there are no reference-generated fixtures.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def clib():
    """liboracle.so (oracle.c), built by oracle/Makefile."""
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'], stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(path)
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.oracle_axpy_f32.argtypes = [ctypes.c_float, f32p, f32p, f32p, ctypes.c_size_t]
        lib.oracle_axpy_f32_noftz.argtypes = [ctypes.c_float, f32p, f32p, f32p, ctypes.c_size_t]
        lib.oracle_add_f32.argtypes = [f32p, f32p, f32p, ctypes.c_size_t]
        lib.oracle_mul_f32.argtypes = [f32p, f32p, f32p, ctypes.c_size_t]
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_cumsum_i64.argtypes = [i64p, i64p, ctypes.c_size_t]
        lib.oracle_sum_f32.argtypes = [f32p, ctypes.c_size_t]
        lib.oracle_sum_f32.restype = ctypes.c_double
        _clib = lib
    return _clib


def _f32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# ---------------------------------------------------------------------------
# elementwise
# ---------------------------------------------------------------------------
def ftz32(a):
    """Flush float32 denormals to signed zero (-ftz=true, cupy/cuda/compiler.py:667)."""
    a = np.ascontiguousarray(a, np.float32).copy()
    u = a.view(np.uint32)
    den = (u & np.uint32(0x7f800000)) == 0
    u[den] &= np.uint32(0x80000000)
    return a


def axpy(a, x, y):
    """`z = a * x + y` of an ElementwiseKernel on float32 (BASELINE.json config 2).

    The reference's JIT contracts mul+add into one FMA and flushes denormals
    (SURVEY.md compile probe 2: FFMA.FTZ); this is that, exactly (oracle.c)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    z = np.empty_like(x)
    clib().oracle_axpy_f32(np.float32(a), _f32p(x), _f32p(y), _f32p(z), x.size)
    return z


def binary_f32(name, x, y):
    """add / multiply on float32: IEEE-exact plus -ftz (cupy/_core/_routines_math.pyx:878-895)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    z = np.empty_like(x)
    fn = {'add': clib().oracle_add_f32, 'multiply': clib().oracle_mul_f32}[name]
    fn(_f32p(x), _f32p(y), _f32p(z), x.size)
    return z


def ufunc(name, *args, **kw):
    """Every other ufunc: NumPy's ufunc of the same name IS the reference's oracle
    (NEP-50 promotion included: cupy/_core/_kernel.pyx:1103-1144, 1656-1753)."""
    return getattr(np, name)(*args, **kw)


def exp_exact(x):
    """exp rounded from float64: the reference's `expf` (cupy/_math/explog.py:8-14) is
    within 2 ulp of this (CUDA math API accuracy table); tests state the tolerance."""
    x = np.asarray(x)
    return np.exp(x.astype(np.float64)).astype(x.dtype)


def ulp_diff(got, want):
    """Distance in units in the last place between two float arrays of one dtype."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.dtype == want.dtype
    it = {2: np.int16, 4: np.int32, 8: np.int64}[got.dtype.itemsize]
    g = got.view(it).astype(np.int64)
    w = want.view(it).astype(np.int64)
    sign = np.int64(1) << (8 * got.dtype.itemsize - 1)
    g = np.where(g < 0, -(g & (sign - 1)), g)        # sign-magnitude -> monotone integers
    w = np.where(w < 0, -(w & (sign - 1)), w)
    d = np.abs(g - w)
    both_nan = np.isnan(got) & np.isnan(want)
    return np.where(both_nan, 0, d)


# ---------------------------------------------------------------------------
# the division / rounding members of the elementwise family: the places where the reference's
# routine text is NOT NumPy's algorithm (cupy_b200/_core/_routines_elementwise.py follows the reference)
# ---------------------------------------------------------------------------
def floor_divide(x, y):
    """cupy/_core/include/cupy/carray.cuh:671-700 `_floor_divide`: integers round toward minus infinity and a zero
    divisor yields 0; floats are floor(x / y) in the operand precision (NumPy derives the float quotient from
    fmod, which differs when x / y rounds up to an integer)."""
    x, y = np.asarray(x), np.asarray(y)
    dt = np.result_type(x, y)
    x, y = x.astype(dt), y.astype(dt)
    with np.errstate(all='ignore'):
        if dt.kind == 'f':
            ct = np.float32 if dt == np.float16 else dt        # float16 operands compute in float
            return np.floor(x.astype(ct) / y.astype(ct)).astype(dt)
        safe = np.where(y == 0, 1, y)
        q = np.floor_divide(x, safe)
        return np.where(y == 0, 0, q).astype(dt)


def remainder(x, y):
    """cupy/_core/_routines_math.pyx:1099-1110: `in0 - _floor_divide(in0, in1) * in1`, times `(in1 != 0)` for
    integers.  (The device compiler may contract the float form into one FMA -- a single rounding -- so bit equality
    with this two-rounding restatement holds where q * y is exact, which is the domain the parity tests use.)"""
    x, y = np.asarray(x), np.asarray(y)
    dt = np.result_type(x, y)
    x, y = x.astype(dt), y.astype(dt)
    with np.errstate(all='ignore'):
        if dt.kind == 'f':
            ct = np.float32 if dt == np.float16 else dt
            q = np.floor(x.astype(ct) / y.astype(ct))
            return (x.astype(ct) - q * y.astype(ct)).astype(dt)
        return (x - floor_divide(x, y) * y) * (y != 0)


def around_int(x, decimals):
    """Integers rounded to a negative number of decimals: cupy/_core/core.pyx:2796-2814 `_round_ufunc_neg_uint`
    (split at the last two digits of the scaled value, round those half to even in float, unscale)."""
    x = np.asarray(x)
    assert x.dtype.kind in 'iu' and decimals < 0
    d = -decimals
    scale = np.int64(10) ** (d - 1)
    xi = x.astype(np.int64)
    q = (np.abs(xi) // scale // 100) * np.sign(xi)             # C division truncates toward zero
    r = (xi - q * scale * 100).astype(np.int32)
    t = np.rint(r.astype(np.float32) / np.float32(scale * 10.0)).astype(np.int64)
    return ((q * 100 + t * 10) * scale).astype(x.dtype)


def sign(x):
    """cupy/_math/misc.py:224-262: integers (x > 0) - (x < 0); floats copysign(1, x) off zero, x - x at zero and NaN
    (so sign(-0.) is +0., NaN stays NaN)."""
    x = np.asarray(x)
    if x.dtype.kind == 'f':
        with np.errstate(invalid='ignore'):
            return np.where((x < 0) | (x > 0), np.copysign(x.dtype.type(1), x), x - x).astype(x.dtype)
    return ((x > 0).astype(np.int64) - (x < 0).astype(np.int64)).astype(x.dtype)


def clip(x, lo, hi):
    """cupy/_core/_routines_math.pyx:1146-1151: `lo > hi ? hi : (x < lo ? lo : (x > hi ? hi : x))`."""
    x = np.asarray(x)
    lo, hi = np.asarray(lo, x.dtype), np.asarray(hi, x.dtype)
    return np.where(lo > hi, hi, np.where(x < lo, lo, np.where(x > hi, hi, x))).astype(x.dtype)


# ---------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------
def sum_dtype(dtype):
    """Result dtype of sum/prod without dtype= (cupy/_core/_routines_math.pyx:762-775:
    '?->l','b->l','B->L', ... 'e->e','f->f','d->d')."""
    dtype = np.dtype(dtype)
    if dtype.kind in 'bi':
        return np.dtype('int64')
    if dtype.kind == 'u':
        return np.dtype('uint64')
    return dtype


def sum(x, axis=None, dtype=None, keepdims=False):
    """ndarray.sum (cupy/_core/_routines_math.pyx:109-128, 777-797).  Integers are
    exact (wrap-around in the 64-bit accumulator); float16 accumulates in float
    (`('e->e', (None, None, None, 'float'))`); floats are summed in float64 here and
    the caller applies the reduction tolerance (summation order is unspecified)."""
    x = np.asarray(x)
    out_dt = sum_dtype(x.dtype) if dtype is None else np.dtype(dtype)
    if out_dt.kind == 'f':
        r = x.sum(axis=axis, dtype=np.float64, keepdims=keepdims)
        return np.asarray(r).astype(out_dt)
    with np.errstate(over='ignore'):
        return np.asarray(x.sum(axis=axis, dtype=out_dt, keepdims=keepdims))


def prod(x, axis=None, dtype=None, keepdims=False):
    x = np.asarray(x)
    out_dt = sum_dtype(x.dtype) if dtype is None else np.dtype(dtype)
    acc = np.float64 if out_dt.kind == 'f' else out_dt
    with np.errstate(over='ignore'):
        return np.asarray(x.prod(axis=axis, dtype=acc, keepdims=keepdims)).astype(out_dt)


def amax(x, axis=None, keepdims=False):
    """NaN propagates (cupy/_core/_routines_statistics.pyx:213-244 my_max_float) == numpy.max."""
    return np.asarray(np.max(x, axis=axis, keepdims=keepdims))


def amin(x, axis=None, keepdims=False):
    return np.asarray(np.min(x, axis=axis, keepdims=keepdims))


def argmax(x, axis=None, keepdims=False):
    """Ties -> lowest index, NaN wins (cupy/_core/_routines_statistics.pyx:255-276,
    342-353); among several NaNs NumPy returns the first, which the reference's
    unordered tree may or may not -- the build follows NumPy.  Result int64."""
    return np.asarray(np.argmax(x, axis=axis, keepdims=keepdims)).astype(np.int64)


def argmin(x, axis=None, keepdims=False):
    return np.asarray(np.argmin(x, axis=axis, keepdims=keepdims)).astype(np.int64)


def mean_dtype(dtype):
    """cupy/_core/_routines_statistics.pyx:132-146: ints/bool -> float64, float16 -> float16
    (summed in float32), else same."""
    dtype = np.dtype(dtype)
    return np.dtype('float64') if dtype.kind in 'iub' else dtype


def mean(x, axis=None, keepdims=False):
    x = np.asarray(x)
    return np.asarray(x.mean(axis=axis, dtype=np.float64, keepdims=keepdims)).astype(mean_dtype(x.dtype))


def var(x, axis=None, ddof=0, keepdims=False):
    """cupy/_core/_routines_statistics.pyx:556-600: mean (keepdims), then
    sum(|x-mean|^2) * 1/max(n-ddof,0) (NaN when that is 0).  Evaluated in float64;
    result dtype float64 for ints/bool, else the input dtype."""
    x = np.asarray(x)
    xf = x.astype(np.float64)
    m = xf.mean(axis=axis, keepdims=True)
    ss = ((xf - m) ** 2).sum(axis=axis, keepdims=keepdims)
    n = x.size // max(np.asarray(ss).size, 1) if x.size else 0
    div = max(n - ddof, 0)
    alpha = 1.0 / div if div != 0 else np.nan
    return np.asarray(ss * alpha).astype(mean_dtype(x.dtype))


# ---------------------------------------------------------------------------
# scan
# ---------------------------------------------------------------------------
def scan_dtype(dtype):
    """cupy/_core/_routines_math.pyx:704-714."""
    dtype = np.dtype(dtype)
    if dtype.kind in 'bi':
        return np.dtype('int64')
    if dtype.kind == 'u':
        return np.dtype('uint64')
    return dtype


def cumsum(x, axis=None, dtype=None):
    """Integers: exact (NumPy).  Floats: float64 scan rounded to the result dtype --
    the device scan order differs from a serial loop, callers apply a tolerance."""
    x = np.asarray(x)
    out_dt = scan_dtype(x.dtype) if dtype is None else np.dtype(dtype)
    if out_dt.kind == 'f':
        return np.cumsum(x, axis=axis, dtype=np.float64).astype(out_dt)
    with np.errstate(over='ignore'):
        return np.cumsum(x, axis=axis, dtype=out_dt)


def cumprod(x, axis=None, dtype=None):
    x = np.asarray(x)
    out_dt = scan_dtype(x.dtype) if dtype is None else np.dtype(dtype)
    if out_dt.kind == 'f':
        return np.cumprod(x, axis=axis, dtype=np.float64).astype(out_dt)
    with np.errstate(over='ignore'):
        return np.cumprod(x, axis=axis, dtype=out_dt)


# ---------------------------------------------------------------------------
# CPU baseline kernels (bench.py cpu_baseline / --impl reference): the hot path
# as the reference's CPU counterpart runs it -- NumPy on the host cores.
# ---------------------------------------------------------------------------
def numpy_axpy(a, x, y, out):
    np.multiply(x, a, out=out)
    np.add(out, y, out=out)
    return out


def numpy_sum(x):
    return x.sum()
