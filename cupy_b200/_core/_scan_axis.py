"""cumsum / cumprod along one axis ("next").

Reference: `_proc_as_batch` + `_batch_scan_op` (cupy/_core/_routines_math.pyx:
499-699): roll the axis to the end, reshape to (lines, n), scan every line.
Round-1 implementation: correct and simple -- the lines are made contiguous and
each line is scanned serially by one thread through a `raw` ElementwiseKernel.
The flat look-back scan (the measured config) does not come through here.
"""
from __future__ import annotations

from cupy_b200 import _lib
from cupy_b200._core import _kernel
from cupy_b200._core._ndarray import ndarray

_memo = {}


def _line_kernel(op):
    k = _memo.get(op)
    if k is None:
        sym = '+' if op == _lib.OP_CUMSUM else '*'
        k = _kernel.ElementwiseKernel(
            'int64 n', 'raw T y',
            'T acc = y[i * n]; for (long long j = 1; j < n; ++j) { acc = acc %s y[i * n + j]; y[i * n + j] = acc; }' % sym,
            'cupy_scan_lines_' + ('sum' if op == _lib.OP_CUMSUM else 'prod'))
        _memo[op] = k
    return k


def scan_axis(a, axis, op, dtype, out):
    nd = a.ndim
    if a.shape[axis] == 0 or a.size == 0:
        res = ndarray(a.shape, dtype)
    else:
        perm = [i for i in range(nd) if i != axis] + [axis]
        t = a.transpose(perm)
        lines = ndarray(t.shape, dtype)            # C-contiguous, scanned axis innermost
        _kernel.elementwise_copy(t, lines)
        n = a.shape[axis]
        _line_kernel(op)(n, lines, size=lines.size // n)
        inv = [perm.index(i) for i in range(nd)]
        res = lines.transpose(inv)
    if out is not None:
        if out.shape != a.shape:
            raise ValueError('Provided out is the wrong size for the reduction')
        _kernel.elementwise_copy(res, out)
        return out
    return res.copy() if not res._c_contiguous else res
