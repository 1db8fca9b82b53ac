"""Known-answer cases of the REFERENCE's own test-suite for the hot path
(SURVEY.md section 8c).  The reference holds no binary golden files: its
expectations are literals or NumPy results on `shaped_arange` / `shaped_random`
inputs.  Each case below cites the reference test it restates and carries the
literal expectation where the reference gives one.  Two users:

  * tests/test_oracle_pinning.py  -- pins oracle/oracle.py on the CPU;
  * tests/test_parity_gpu.py      -- runs the same cases through cupy_b200 on the GPU.

The input helpers restate cupy/testing/_helper.py:136-160 (shaped_arange: 1..N,
exact in every dtype; bool = even numbers) and :189-221 (shaped_random:
numpy.random.RandomState(seed=0).rand(*shape) * 10).
"""
import numpy as np

ALL_DTYPES = ['?', 'int8', 'uint8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'uint64',
              'float16', 'float32', 'float64']
FLOAT_DTYPES = ['float16', 'float32', 'float64']


def shaped_arange(shape, dtype='float32', order='C'):
    dtype = np.dtype(dtype)
    a = np.arange(1, int(np.prod(shape)) + 1, 1)
    if dtype == '?':
        a = a % 2 == 0
    return np.array(a.astype(dtype).reshape(shape), order=order)


def shaped_random(shape, dtype='float32', scale=10, seed=0, order='C'):
    rng = np.random.RandomState(seed)
    dtype = np.dtype(dtype)
    if dtype == '?':
        a = rng.randint(2, size=shape)
    else:
        a = rng.rand(*shape) * scale
    return np.asarray(a, dtype=dtype, order=order)


# (id, reference test file:line, input builder, op name, kwargs, literal expectation or None)
KNOWN_ANSWERS = []


def _case(cid, ref, build, op, kwargs=None, expect=None):
    KNOWN_ANSWERS.append((cid, ref, build, op, kwargs or {}, expect))


for _dt in ALL_DTYPES:
    # sorting_tests/test_search.py:62-66  test_argmax_tie: [0,5,2,3,4,5] -> 1 (lowest index of the max)
    _case('argmax_tie_' + _dt, 'tests/cupy_tests/sorting_tests/test_search.py:62-66',
          lambda dt=_dt: np.array([0, 5, 2, 3, 4, 5], dt), 'argmax', {}, 1 if _dt != '?' else 1)
    _case('argmin_tie_' + _dt, 'tests/cupy_tests/sorting_tests/test_search.py:152-156',
          lambda dt=_dt: np.array([0, 1, 2, 3, 0, 5], dt), 'argmin', {}, 0)
    # core_tests/test_scan.py:12-23  ones(n) -> arange(1..n), bit exact, every dtype
    _n = 100 if _dt in ('int8', 'uint8', 'float16') else 10000
    _case('scan_ones_' + _dt, 'tests/cupy_tests/core_tests/test_scan.py:12-23',
          lambda dt=_dt, n=_n: np.ones((n,), dt), 'cumsum_same_dtype', {},
          np.arange(1, _n + 1).astype(_dt) if _dt != '?' else None)

for _dt in FLOAT_DTYPES:
    # sorting_tests/test_search.py:26-30  test_argmax_nan: [nan,-1,1] -> 0
    _case('argmax_nan_' + _dt, 'tests/cupy_tests/sorting_tests/test_search.py:26-30',
          lambda dt=_dt: np.array([np.nan, -1, 1], dt), 'argmax', {}, 0)
    _case('argmin_nan_' + _dt, 'tests/cupy_tests/sorting_tests/test_search.py:116-120',
          lambda dt=_dt: np.array([np.nan, -1, 1], dt), 'argmin', {}, 0)
    # core_tests/test_ndarray_reduction.py:93-97  test_max_nan: [nan,1,-1] -> nan
    _case('max_nan_' + _dt, 'tests/cupy_tests/core_tests/test_ndarray_reduction.py:93-97',
          lambda dt=_dt: np.array([np.nan, 1, -1], dt), 'max', {}, np.nan)
    _case('min_nan_' + _dt, 'tests/cupy_tests/core_tests/test_ndarray_reduction.py:168-172',
          lambda dt=_dt: np.array([np.nan, 1, -1], dt), 'min', {}, np.nan)
    # core_tests/test_ndarray_reduction.py:111-116  test_max_inf (cupy/cupy#8180): [-inf,-inf] -> -inf
    _case('max_inf_' + _dt, 'tests/cupy_tests/core_tests/test_ndarray_reduction.py:111-116',
          lambda dt=_dt: np.array([-np.inf, -np.inf], dt), 'max', {}, -np.inf)
    _case('min_inf_' + _dt, 'tests/cupy_tests/core_tests/test_ndarray_reduction.py:186-191',
          lambda dt=_dt: np.array([np.inf, np.inf], dt), 'min', {}, np.inf)

# math_tests/test_sumprod.py:26-174: shaped_arange sums are integer-exact in every dtype
for _dt in ALL_DTYPES:
    _case('sum_all_' + _dt, 'tests/cupy_tests/math_tests/test_sumprod.py:26-37',
          lambda dt=_dt: shaped_arange((2, 3, 4), dt), 'sum', {},
          300 if _dt != '?' else 12)
    _case('sum_axis1_' + _dt, 'tests/cupy_tests/math_tests/test_sumprod.py:56-61',
          lambda dt=_dt: shaped_arange((2, 3, 4), dt), 'sum', {'axis': 1}, None)
    _case('sum_axes_' + _dt, 'tests/cupy_tests/math_tests/test_sumprod.py:95-106',
          lambda dt=_dt: shaped_arange((2, 3, 4, 5), dt), 'sum', {'axis': (1, 3)}, None)
    _case('sum_transposed_' + _dt, 'tests/cupy_tests/math_tests/test_sumprod.py:77-93',
          lambda dt=_dt: shaped_arange((20, 30, 40), dt).transpose(2, 0, 1), 'sum', {'axis': 1}, None)
    _case('sum_keepdims_' + _dt, 'tests/cupy_tests/math_tests/test_sumprod.py:165-174',
          lambda dt=_dt: shaped_arange((2, 3, 4), dt), 'sum', {'axis': 1, 'keepdims': True}, None)

# statistics_tests/test_meanvar.py:193-330: mean/var on shaped_arange (2,3)/(2,3,4), ddof
for _dt in ['int8', 'int32', 'int64', 'float16', 'float32', 'float64']:
    _case('mean_all_' + _dt, 'tests/cupy_tests/statistics_tests/test_meanvar.py:193-204',
          lambda dt=_dt: shaped_arange((2, 3), dt), 'mean', {}, 3.5)
    _case('mean_axis_' + _dt, 'tests/cupy_tests/statistics_tests/test_meanvar.py:206-217',
          lambda dt=_dt: shaped_arange((2, 3, 4), dt), 'mean', {'axis': 1}, None)
    _case('var_all_' + _dt, 'tests/cupy_tests/statistics_tests/test_meanvar.py:232-243',
          lambda dt=_dt: shaped_arange((2, 3), dt), 'var', {}, 35.0 / 12.0)
    _case('var_ddof_' + _dt, 'tests/cupy_tests/statistics_tests/test_meanvar.py:245-256',
          lambda dt=_dt: shaped_arange((2, 3), dt), 'var', {'ddof': 1}, 3.5)
    _case('var_axis_ddof_' + _dt, 'tests/cupy_tests/statistics_tests/test_meanvar.py:271-282',
          lambda dt=_dt: shaped_arange((2, 3, 4), dt), 'var', {'axis': 1, 'ddof': 1}, None)

# core_tests/test_reduction.py:58-93: int8 sums (wrap into the int64 accumulator: exact) over 2^i, 2^i +- 1
for _i in (7, 10, 13, 16):
    for _d in (-1, 0, 1):
        _n = (1 << _i) + _d
        _case('sum_int8_%d' % _n, 'tests/cupy_tests/core_tests/test_reduction.py:58-66',
              lambda n=_n: np.ones((n,), 'int8'), 'sum', {}, _n)
_case('sum_int8_axis0_large', 'tests/cupy_tests/core_tests/test_reduction.py:68-75',
      lambda: np.ones((1025, 1000), 'int8'), 'sum', {'axis': 0}, None)
_case('sum_int8_axis1_large', 'tests/cupy_tests/core_tests/test_reduction.py:77-84',
      lambda: np.ones((1025, 1000), 'int8'), 'sum', {'axis': 1}, None)
