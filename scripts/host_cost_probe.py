"""Where the ~14 us of a small call goes on the GPU box's host: the pieces of the memoised launch path, timed alone."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cupy_b200 as cp  # noqa: E402
from cupy_b200 import _lib  # noqa: E402
from cupy_b200._core import _workspace, _kernel, _scalar  # noqa: E402
from cupy_b200._core._ndarray import ndarray, current_stream_ptr  # noqa: E402


def per_call(f, n=5000):
    for _ in range(200):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return dt


x = cp.arange(1000)
xf = cp.arange(1000, dtype=np.float32)
f32 = np.dtype('float32')
st = current_stream_ptr()
out = cp.empty((), np.int64)
layout_desc = _lib.ReduceDesc(_lib.OP_SUM, _lib.RED_FULL, _scalar.dtype_id(x.dtype), _scalar.dtype_id(out.dtype), 1, 1000, 1, 0.0)
need = ctypes.c_size_t()
_lib.check(_lib.lib.b200_reduce_workspace_bytes(ctypes.byref(layout_desc), ctypes.byref(need)))
ws_ptr, ws_bytes = _workspace.get(need.value, st)
rows = [
    ('x.sum()  (whole call)', lambda: x.sum()),
    ('cp.add(xf, xf)  (whole call)', lambda: cp.add(xf, xf)),
    ('cp.add(xf, xf, out=)', (lambda o: (lambda: cp.add(xf, xf, out=o)))(cp.empty((1000,), np.float32))),
    ('torch.empty(4000 B, cuda) + free', lambda: torch.empty(4000, dtype=torch.uint8, device='cuda')),
    ('ndarray._fresh((1000,), f32)', lambda: ndarray._fresh((1000,), f32, (4,), 1000)),
    ('current_stream_ptr()', current_stream_ptr),
    ('torch.cuda.current_device()', torch.cuda.current_device),
    ('_workspace.get(need, st)', lambda: _workspace.get(need.value, st)),
    ('b200_reduce_run ctypes call (launch included)',
     lambda: _lib.lib.b200_reduce_run(ctypes.byref(layout_desc), x.ptr, out.ptr, ws_ptr, ws_bytes, st)),
    ('torch: xt.sum()', (lambda t: (lambda: t.sum()))(x.to_torch())),
    ('torch: torch.add(t, t)', (lambda t: (lambda: torch.add(t, t)))(xf.to_torch())),
]
for name, f in rows:
    print('%-50s %6.2f us' % (name, per_call(f)), flush=True)
