import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# A kernel cache left in the tree (`.kcache/`, filled by `scripts/dry_gpu_tests.py`: the NVRTC cubins of the kernels
# the GPU tier launches, content-addressed by source + options + header digest) travels with the checkout like the
# built .so does; when present the tests start warm instead of recompiling ~2000 kernels on the GPU box.  An
# explicit CUPY_B200_CACHE_DIR wins; without the directory nothing changes (the default per-user cache).
_KCACHE = os.path.join(ROOT, '.kcache')
if os.path.isdir(_KCACHE) and 'CUPY_B200_CACHE_DIR' not in os.environ:
    os.environ['CUPY_B200_CACHE_DIR'] = _KCACHE


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def dry():
    """Host logic + NVRTC compilation without a GPU (cupy_b200/_core/_dryrun.py)."""
    from cupy_b200._core import _dryrun
    with _dryrun.dry_run() as log:
        yield log
