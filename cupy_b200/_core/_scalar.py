"""dtype <-> C type names / dtype ids, and by-value scalars with NEP-50 weakness.

Mirrors cupy/_core/_scalar.pyx (`_typenames` :25-58 via get_typename :86-125,
`CScalar` :320-408, `numpy_dtype_from_pyscalar` :302-318) and
cupy/_core/_dtype.pyx (`get_dtype`, `_raise_if_invalid_cast` :149-174).
"""
from __future__ import annotations

import numpy

from cupy_b200 import _lib

_DTYPE_IDS = {
    numpy.dtype('int8'): _lib.TYPE_INT8, numpy.dtype('uint8'): _lib.TYPE_UINT8,
    numpy.dtype('int16'): _lib.TYPE_INT16, numpy.dtype('uint16'): _lib.TYPE_UINT16,
    numpy.dtype('int32'): _lib.TYPE_INT32, numpy.dtype('uint32'): _lib.TYPE_UINT32,
    numpy.dtype('int64'): _lib.TYPE_INT64, numpy.dtype('uint64'): _lib.TYPE_UINT64,
    numpy.dtype('float16'): _lib.TYPE_FLOAT16, numpy.dtype('float32'): _lib.TYPE_FLOAT32,
    numpy.dtype('float64'): _lib.TYPE_FLOAT64, numpy.dtype('bool'): _lib.TYPE_BOOL,
    numpy.dtype('complex64'): _lib.TYPE_COMPLEX64, numpy.dtype('complex128'): _lib.TYPE_COMPLEX128,
}

# C type names user code strings are compiled against (same spelling as the
# reference so that preambles / operations port unchanged; `float16` is
# b200::float16, pulled into the global namespace by the generated prologue).
_TYPENAMES = {
    numpy.dtype('bool'): 'bool',
    numpy.dtype('int8'): 'signed char', numpy.dtype('uint8'): 'unsigned char',
    numpy.dtype('int16'): 'short', numpy.dtype('uint16'): 'unsigned short',
    numpy.dtype('int32'): 'int', numpy.dtype('uint32'): 'unsigned int',
    numpy.dtype('int64'): 'long long', numpy.dtype('uint64'): 'unsigned long long',
    numpy.dtype('float16'): 'float16', numpy.dtype('float32'): 'float',
    numpy.dtype('float64'): 'double',
}


def get_dtype(t):
    return t if isinstance(t, numpy.dtype) else numpy.dtype(t)


def dtype_id(dtype):
    try:
        return _DTYPE_IDS[get_dtype(dtype)]
    except KeyError:
        raise TypeError('Unsupported dtype %s' % (dtype,))


def get_typename(dtype):
    if dtype is None:
        raise TypeError('dtype is None')
    try:
        return _TYPENAMES[get_dtype(dtype)]
    except KeyError:
        raise ValueError('Unsupported dtype %s (complex / structured dtypes are outside the hot path)' % (dtype,))


def raise_if_invalid_cast(from_dt, to_dt, casting, argname='array data'):
    """cupy/_core/_dtype.pyx:149-174."""
    from_dt, to_dt = get_dtype(from_dt), get_dtype(to_dt)
    if from_dt == to_dt:
        return
    if casting == 'same_kind' and from_dt.kind == to_dt.kind:
        return
    if casting == 'unsafe':
        return
    if numpy.can_cast(from_dt, to_dt, casting=casting):
        return
    raise TypeError('Cannot cast %s from %r to %r according to the rule %r'
                    % (argname, from_dt, to_dt, casting))


class CScalar:
    """A Python / NumPy scalar passed by value to a kernel.

    `weak_t` is the Python type (bool / int / float / complex) for Python
    scalars -- they take part in loop selection "weakly" (NEP 50) -- and False
    for NumPy scalars (cupy/_core/_scalar.pyx:320-349).
    """

    __slots__ = ('value', 'descr', 'weak_t')

    def __init__(self, value):
        if isinstance(value, numpy.generic):
            self.value = value
            self.descr = value.dtype
            self.weak_t = False
        elif isinstance(value, bool):
            self.value, self.descr, self.weak_t = value, numpy.dtype('bool'), bool
        elif isinstance(value, int):
            self.value, self.weak_t = value, int
            # numpy_dtype_from_pyscalar (cupy/_core/_scalar.pyx:302-318)
            if -(1 << 63) <= value < (1 << 63):
                self.descr = numpy.dtype('int64')
            elif 0 <= value < (1 << 64):
                self.descr = numpy.dtype('uint64')
            else:
                raise OverflowError('Python int too large to convert to C long')
        elif isinstance(value, float):
            self.value, self.descr, self.weak_t = value, numpy.dtype('float64'), float
        elif isinstance(value, complex):
            self.value, self.descr, self.weak_t = value, numpy.dtype('complex128'), complex
        elif isinstance(value, numpy.ndarray) and value.ndim == 0:
            self.value = value[()]
            self.descr = value.dtype
            self.weak_t = False
        else:
            raise TypeError('Unsupported type %s' % type(value))

    @property
    def dtype(self):
        return self.descr

    def apply_dtype(self, dtype):
        """Cast to the kernel's parameter type (cupy/_core/_scalar.pyx:382-398).

        Python ints that do not fit raise OverflowError (NEP 50; reference tests
        tests/cupy_tests/core_tests/test_elementwise.py:89-145)."""
        dtype = get_dtype(dtype)
        if self.weak_t is int and dtype.kind in 'iu':
            info = numpy.iinfo(dtype)
            if not (info.min <= self.value <= info.max):
                raise OverflowError('Python integer %d out of bounds for %s' % (self.value, dtype))
        if self.weak_t is not False and self.weak_t is not bool and dtype.kind == 'b':
            self.value = bool(self.value)
        with numpy.errstate(over='ignore', invalid='ignore'):
            self.value = numpy.asarray(self.value).astype(dtype, casting='unsafe')[()]
        self.descr = dtype
        self.weak_t = False

    def raw_bytes(self):
        b = numpy.asarray(self.value, dtype=self.descr).tobytes()
        return b + b'\0' * (16 - len(b))
