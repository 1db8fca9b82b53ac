"""cupy_b200.cuda -- the stream / event / pinned-memory surface that sits on either side
of the hot path (host <-> device copies overlapped with the kernels).

Mirrors the names and behaviour of cupy.cuda (cupy/cuda/stream.pyx:101-190 `Event`,
:194-520 `Stream`, `get_current_stream`; cupy/cuda/pinned_memory.pyx `alloc_pinned_memory`)
over torch's stream plumbing: the current stream is thread-local per device
(cupy_backends/cuda/stream.pyx:34-53), `with stream:` / `stream.use()` make it current, and
every cupy_b200 kernel launch and `ndarray.get/set` is enqueued on the current stream.
"""
from __future__ import annotations

import torch

from cupy_b200._core._ndarray import empty_pinned  # noqa: F401


class Event:
    """cupy/cuda/stream.pyx:101-166."""

    def __init__(self, block=False, disable_timing=False, interprocess=False):
        if interprocess and not disable_timing:
            raise ValueError('Timing must be disabled for interprocess events')
        self._ev = torch.cuda.Event(enable_timing=not disable_timing, blocking=block, interprocess=interprocess)

    @property
    def done(self):
        return self._ev.query()

    def record(self, stream=None):
        self._ev.record(_torch_stream(stream))

    def synchronize(self):
        self._ev.synchronize()


def get_elapsed_time(start_event, end_event):
    """Milliseconds between two recorded events (cupy/cuda/stream.pyx:168-190)."""
    return start_event._ev.elapsed_time(end_event._ev)


class Stream:
    """cupy/cuda/stream.pyx:459-530.  `Stream.null` is the legacy default stream."""

    null = None

    def __init__(self, null=False, non_blocking=False, ptds=False, priority=None):
        if ptds:
            raise NotImplementedError('per-thread default streams are not supported')
        if null:
            self._st = torch.cuda.default_stream()
        else:
            # torch streams are created with cudaStreamNonBlocking; `non_blocking` is accepted for parity
            self._st = torch.cuda.Stream() if priority is None else torch.cuda.Stream(priority=priority)
        self._ctx = []
        self.non_blocking = bool(non_blocking)

    @classmethod
    def from_external(cls, obj):
        s = cls.__new__(cls)
        if isinstance(obj, torch.cuda.Stream):
            s._st = obj
        else:
            s._st = torch.cuda.ExternalStream(int(getattr(obj, 'ptr', obj)))
        s._ctx, s.non_blocking = [], True
        return s

    @property
    def ptr(self):
        return self._st.cuda_stream

    @property
    def device_id(self):
        return self._st.device_index

    def __eq__(self, other):
        return isinstance(other, Stream) and self.ptr == other.ptr and self.device_id == other.device_id

    def __hash__(self):
        return hash((self.ptr, self.device_id))

    def __repr__(self):
        return '<Stream %d (device %d)>' % (self.ptr, self.device_id)

    def __cuda_stream__(self):
        return (0, self.ptr)

    def __enter__(self):
        c = torch.cuda.stream(self._st)
        c.__enter__()
        self._ctx.append(c)
        return self

    def __exit__(self, *args):
        self._ctx.pop().__exit__(*args)

    def use(self):
        """Make this the current stream of the calling thread (cupy/cuda/stream.pyx:254-264)."""
        torch.cuda.set_stream(self._st)
        return self

    @property
    def done(self):
        return self._st.query()

    def synchronize(self):
        self._st.synchronize()

    def record(self, event=None):
        if event is None:
            event = Event(disable_timing=True)
        event._ev.record(self._st)
        return event

    def wait_event(self, event):
        self._st.wait_event(event._ev)


Stream.null = Stream(null=True) if torch.cuda.is_available() else None


def get_current_stream(device_id=None):
    return Stream.from_external(torch.cuda.current_stream(device_id))


def _torch_stream(stream):
    if stream is None:
        return torch.cuda.current_stream()
    if isinstance(stream, Stream):
        return stream._st
    if isinstance(stream, torch.cuda.Stream):
        return stream
    return torch.cuda.ExternalStream(int(getattr(stream, 'ptr', stream)))


def alloc_pinned_memory(nbytes):
    """Page-locked host bytes (numpy uint8 view); cupy/cuda/pinned_memory.pyx `alloc_pinned_memory`."""
    import numpy
    return empty_pinned((int(nbytes),), numpy.uint8)
