"""Stream-keyed scratch memory for the single-pass kernels.

The C ABI never allocates (SURVEY.md section 8b "ownership"): callers hand in a
workspace whose 16 KiB ticket header is zero at first use and is left zeroed by
every kernel.  One growing, zero-initialised buffer per (device, stream) gives
that without a per-call memset; the reference allocates its scratch from the
memory pool on every call instead (cupy/cuda/cub.pyx:170-191,
cupy/_core/_cub_reduction.pyx:449-452).
"""
from __future__ import annotations

import torch

_TICKET_BYTES = 16384
_pool = {}


def get(nbytes, stream_ptr):
    """Returns (device pointer, size) of a workspace of at least `nbytes`."""
    nbytes = max(int(nbytes), _TICKET_BYTES)
    from cupy_b200._core import _dryrun
    if _dryrun.enabled:
        return _dryrun.fake_alloc(nbytes), nbytes
    key = (torch.cuda.current_device(), int(stream_ptr))
    buf = _pool.get(key)
    if buf is None or buf.numel() < nbytes:
        size = max(nbytes, 1 << 20)
        size = (size + 255) // 256 * 256
        # a fresh buffer replaces the old one; kernels in flight on this stream keep
        # using the old allocation, which the caching allocator only reuses in stream order
        buf = torch.zeros(size, dtype=torch.uint8, device='cuda')
        _pool[key] = buf
    return buf.data_ptr(), buf.numel()


def clear():
    _pool.clear()
