"""CPU tier for cupy_b200/_core/_routines_elementwise.py: the loop tables pick NumPy's result dtypes, every
ufunc's routine text compiles for sm_100a (NVRTC, dry-run: nothing is launched), and the Python-level wrappers
(clip / around / where / isclose / nan_to_num / NaN-ignoring moments) route as the reference's do."""
import numpy as np
import pytest

import cupy_b200 as cp
from cupy_b200._core import _routines_elementwise as E
from cupy_b200._core._kernel import ufunc

ALL = '?bBhHiIlLefd'


def names(log):
    """Kernel names without the dtype suffix the generator appends."""
    return [d['name'].split('__')[0] for d in log]
UFUNCS = sorted(n for n in dir(E) if isinstance(getattr(E, n), ufunc) and hasattr(np, n))


def _numpy_result(name, chars):
    args = [np.ones(3, c) for c in chars]
    try:
        with np.errstate(all='ignore'):
            r = getattr(np, name)(*args)
    except TypeError:
        return None
    return tuple(x.dtype for x in r) if isinstance(r, tuple) else (r.dtype,)


@pytest.mark.parametrize('name', UFUNCS)
def test_loop_selection_matches_numpy(name):
    """Same output dtype as NumPy for every input dtype (both operands of one dtype), or both refuse."""
    uf = getattr(E, name)
    if name in ('ldexp',):
        pytest.skip('mixed float / integer operands: covered below')
    for c in ALL:
        in_types = (np.dtype(c),) * uf.nin
        op = uf._ops._guess_routine_from_in_types(in_types)
        want = _numpy_result(name, c * uf.nin)
        if name in ('positive', 'gcd', 'lcm') and c == '?':
            # NumPy computes these on booleans; the reference raises (cupy/_core/_routines_math.pyx:951-954,
            # cupy/_math/rational.py:6-11)
            with pytest.raises(TypeError):
                op.check_valid()
            continue
        if name in ('floor', 'ceil', 'trunc', 'fix', 'rint') and c == '?' and want is not None:
            # the reference keeps bool (cupy/_math/rounding.py:55); NumPy < 2.1 promoted to float16
            assert op is not None
            continue
        if want is None:
            # NumPy refuses booleans where the reference's table casts them to its first integer loop
            # (sign, reciprocal, floor_divide ...: no '?' loop, `can_cast(bool, int8)` holds)
            assert op is None or c == '?', (name, c, op)
            continue
        assert op is not None, (name, c)
        if name == 'float_power' and want == (np.dtype('float64'),):
            assert op.out_types == (np.dtype('float64'),)
            continue
        if op.out_types != want:
            # the documented differences: integer / bool inputs of float-only functions take the first loop
            # they can be cast to, which is NumPy's choice as well -- so any mismatch is a bug
            raise AssertionError((name, c, op.out_types, want))


def test_ldexp_loops():
    for f, i, want in (('e', 'i', 'e'), ('f', 'i', 'f'), ('f', 'l', 'f'), ('d', 'i', 'd'), ('d', 'l', 'd'), ('e', 'l', 'e')):
        op = E.ldexp._ops._guess_routine_from_in_types((np.dtype(f), np.dtype(i)))
        assert op is not None and op.out_types == (np.dtype(want),)
        assert np.ldexp(np.ones(2, f), np.ones(2, i)).dtype == np.dtype(want)


@pytest.mark.parametrize('name', sorted(n for n in dir(E) if isinstance(getattr(E, n), ufunc) and not n.startswith('__')))
def test_every_routine_compiles_for_sm100a(dry, name):
    """One integer and one float loop of every table go through codegen + NVRTC."""
    uf = getattr(E, name)
    seen = set()
    for op in uf._ops.ops:
        kind = op.in_types[0].kind + str(op.in_types[0].itemsize)
        if op.error_func is not None or kind in seen or kind not in ('i4', 'u8', 'f2', 'f8', 'b1'):
            continue
        seen.add(kind)
        args = [cp.empty((257, 3), t) for t in op.in_types]
        res = uf(*args)
        res = res if isinstance(res, tuple) else (res,)
        assert tuple(r.dtype for r in res) == op.out_types
        assert dry[-1]['kind'] in ('jit_elementwise', 'prebuilt_ufunc')
    assert seen


def test_wrappers_route_like_the_reference(dry):
    f = cp.empty((50, 20), 'f')
    i = cp.empty((50, 20), 'i')
    b = cp.empty((50, 20), '?')
    # clip: a missing bound becomes the dtype's own limit, one cupy_clip launch either way
    for a in (f, i):
        del dry[:]
        assert cp.clip(a, None, 5).dtype == a.dtype and a.clip(2).dtype == a.dtype
        assert names(dry) == ['cupy_clip', 'cupy_clip']
    # around: integers with negative decimals take the digit-splitting kernel
    del dry[:]
    cp.around(i, -1), cp.around(i, 2), cp.around(f, -1), f.round(1)
    assert names(dry) == ['cupy_round_neg_uint', 'cupy_round', 'cupy_round', 'cupy_round']
    # where: a non-boolean condition is compared with zero first
    del dry[:]
    assert cp.where(b, f, i).dtype == np.float64
    assert cp.where(f, f, 0).dtype == np.float32
    assert names(dry) == ['cupy_where', 'cupy_not_equal', 'cupy_where']
    with pytest.raises(ValueError):
        cp.where(b, f)
    # isclose on integers compares in float64
    del dry[:]
    assert cp.isclose(i, i).dtype == np.bool_
    assert names(dry)[-1] == 'cupy_is_close' and 'double' in dry[-1]['source']
    assert cp.allclose(f, f).shape == () and cp.array_equal(f, f).shape == ()
    # nan_to_num leaves integers alone
    del dry[:]
    assert cp.nan_to_num(i, copy=False) is i and not dry
    assert cp.nan_to_num(f).dtype == np.float32 and names(dry)[-1].startswith('cupy_nan_to_num')
    # operators
    q, r = divmod(f, 3)
    assert q.dtype == r.dtype == np.float32
    assert (i // 2).dtype == np.int32 and (7 % i).dtype == np.int32 and (i % 2.5).dtype == np.float64
    # NumPy's protocol: numpy.floor(device array) is this package's ufunc
    got = np.floor(f)
    assert isinstance(got, cp.ndarray) and names(dry)[-1] == 'cupy_floor'
    got = np.logical_and(f, i)
    assert isinstance(got, cp.ndarray) and got.dtype == np.bool_


def test_nan_moments_route(dry):
    f = cp.empty((300, 200), 'f')
    del dry[:]
    assert cp.nanmean(f, axis=1).shape == (300,)
    assert len(dry) == 1 and dry[0]['name'].startswith('cupy_nanmean')
    del dry[:]
    assert cp.nanvar(f, axis=0, ddof=1).dtype == np.float32
    assert [d['name'].split('_rows')[0].split('_cols')[0] for d in dry] == ['cupy_count_non_nan', 'cupy_nansum', 'cupy_nanvar_core']
    assert cp.nanstd(f, keepdims=True).shape == (1, 1)
    h = cp.empty((64, 100), 'e')
    del dry[:]
    assert cp.nanvar(h, axis=1).dtype == np.float16
    assert dry[-1]['name'].startswith('cupy_nanvar_core_float16')
    i = cp.empty((64, 100), 'i')
    assert cp.nanmean(i).dtype == np.float64 and cp.nanvar(i, axis=0).dtype == np.float64
