"""Dry-run mode for machines without a GPU (the CPU test tier).

When enabled, arrays get fake device pointers (never dereferenced), every kernel
the engine would launch is still planned, generated and COMPILED for sm_100a by
NVRTC, but nothing is loaded or launched; each would-be launch is appended to
`log` so tests can assert on the launcher's decisions (variant, vector width,
prebuilt vs NVRTC route, reduction layout).  This is test infrastructure for the
host logic only -- it computes nothing and is never a fallback.
"""
from __future__ import annotations

import contextlib

enabled = False
log = []
_next_ptr = [1 << 40]


def fake_alloc(nbytes):
    p = _next_ptr[0]
    _next_ptr[0] += (max(int(nbytes), 1) + 511) // 512 * 512
    return p


def record(kind, **info):
    info['kind'] = kind
    log.append(info)


@contextlib.contextmanager
def dry_run():
    global enabled
    prev = enabled
    enabled = True
    del log[:]
    try:
        yield log
    finally:
        enabled = prev
