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
        # The buffer belongs to the stream the kernels run on, not to torch's current stream: it is
        # allocated under that stream (so the caching allocator reuses a replaced buffer only in THAT
        # stream's order -- kernels still in flight on it keep their old allocation intact) and zeroed
        # by a memset enqueued on that stream, ahead of the first kernel that reads the tickets.
        with torch.cuda.stream(_as_torch_stream(stream_ptr)):
            buf = torch.empty(size, dtype=torch.uint8, device='cuda')
        from cupy_b200 import _lib
        _lib.check(_lib.lib.b200_workspace_init(buf.data_ptr(), size, stream_ptr))
        _pool[key] = buf
    return buf.data_ptr(), buf.numel()


def _as_torch_stream(stream_ptr):
    cur = torch.cuda.current_stream()
    if cur.cuda_stream == stream_ptr:
        return cur
    if stream_ptr == 0:
        return torch.cuda.default_stream()
    return torch.cuda.ExternalStream(stream_ptr)


def clear():
    _pool.clear()
