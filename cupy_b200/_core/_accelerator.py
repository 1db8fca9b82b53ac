"""Accelerator switch (cupy/_core/_accelerator.pyx:6-59, env `CUPY_ACCELERATORS`).

The reference lets a user choose which backends its routines and reductions try (`cub`,
`cutensor`, ...), in order, before the generic JIT kernel.  Here the list selects between the code
paths of this engine, with the same setters / getters / environment variable:

  'b200'      the prebuilt single-pass kernels and the structured NVRTC skeletons (default)
  'generic'   only the generic strided NVRTC reduction (any layout; the A/B arm for the fast paths)
  'reference' an externally registered checker backend (`register_reference_backend`): TESTS hang the
              reference's own kernels (oracle/_ref) here to run whole call chains through them.  The
              package itself registers nothing and never imports oracle/; selecting 'reference' without
              a registered backend raises.
Names the reference accepts ('cub', 'cutensor', 'cutensornet') are accepted and mean 'b200'.
"""
from __future__ import annotations

import os

ACCELERATOR_B200 = 'b200'
ACCELERATOR_GENERIC = 'generic'
ACCELERATOR_REFERENCE = 'reference'
_ALIASES = {'cub': ACCELERATOR_B200, 'cutensor': ACCELERATOR_B200, 'cutensornet': ACCELERATOR_B200}
_KNOWN = (ACCELERATOR_B200, ACCELERATOR_GENERIC, ACCELERATOR_REFERENCE)

_routine_accelerators = [ACCELERATOR_B200]
_reduction_accelerators = [ACCELERATOR_B200]
_reference_backend = None


def _normalize(accelerators):
    if isinstance(accelerators, str):
        accelerators = [a for a in accelerators.split(',') if a]
    out = []
    for a in accelerators:
        a = _ALIASES.get(a, a)
        if a not in _KNOWN:
            raise ValueError('Unknown accelerator: %s' % a)
        if a not in out:
            out.append(a)
    return out


def set_routine_accelerators(accelerators):
    global _routine_accelerators
    _routine_accelerators = _normalize(accelerators)


def set_reduction_accelerators(accelerators):
    global _reduction_accelerators
    _reduction_accelerators = _normalize(accelerators)


def get_routine_accelerators():
    return list(_routine_accelerators)


def get_reduction_accelerators():
    return list(_reduction_accelerators)


def register_reference_backend(backend):
    """backend(kind, name, array, **kw) -> ndarray or None (None = not handled).  kind is 'reduction'
    or 'scan'.  Pass None to unregister."""
    global _reference_backend
    _reference_backend = backend


def try_reference(kind, name, array, **kw):
    """Called by the routers when 'reference' is the first accelerator of that family."""
    if _reference_backend is None:
        raise RuntimeError("accelerator 'reference' is selected but no reference backend is registered "
                           '(tests register oracle/_ref; the package has none)')
    return _reference_backend(kind, name, array, **kw)


def fast_paths_enabled():
    return ACCELERATOR_B200 in _reduction_accelerators


def reference_first(routine=False):
    lst = _routine_accelerators if routine else _reduction_accelerators
    return bool(lst) and lst[0] == ACCELERATOR_REFERENCE


def _set_default_accelerators():
    env = os.getenv('CUPY_ACCELERATORS', ACCELERATOR_B200)
    set_routine_accelerators(env)
    set_reduction_accelerators(env)


_set_default_accelerators()
