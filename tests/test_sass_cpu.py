"""CPU tier: what the generated kernels look like in SASS (cuobjdump of the NVRTC cubin, no GPU needed) -- the
vector widths DESIGN.md section 3.1 claims are properties of the compiled code, so they are pinned here."""
import collections
import os
import re
import shutil
import subprocess

import pytest

import cupy_b200 as cp
from cupy_b200._core import _jit

CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason='cuobjdump not found')


def sass_ops(source, name, options=()):
    cubin = _jit.compile_to_cubin(source, options, name + '.cu')
    path = '/tmp/_b200_sass_%d.cubin' % os.getpid()
    with open(path, 'wb') as f:
        f.write(cubin)
    try:
        text = subprocess.check_output([CUOBJDUMP, '-sass', path]).decode()
    finally:
        os.remove(path)
    return collections.Counter(re.findall(r'\b(LDG[.\w]*|STG[.\w]*|UTMALDG[.\w]*)', text))


def test_flat_axpy_moves_128_bits_per_access(dry):
    n = 1 << 28
    k = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'sass_axpy')
    k(1.5, cp.empty((n,), 'f'), cp.empty((n,), 'f'), cp.empty((n,), 'f'))
    ops = sass_ops(dry[-1]['source'], 'sass_axpy')
    assert dry[-1]['vec'] == 4
    assert ops['LDG.E.128'] >= 8 and ops['STG.E.128'] >= 4          # 2 operands x 4 unrolled steps; 4 stores
    assert ops['LDG.E.64'] == 0


def test_mixed_item_sizes_keep_the_narrow_operand_wide(dry):
    """where(mask, x, y): 8-element vectors -- the 1-byte mask moves 8 bytes per access, the float operands two
    16-byte accesses (cupy_b200._core._kernel.tunables['flat_mixed_vec'])."""
    n = 1 << 28
    cp.where(cp.empty((n,), '?'), cp.empty((n,), 'f'), cp.empty((n,), 'f'))
    assert dry[-1]['vec'] == 8
    ops = sass_ops(dry[-1]['source'], dry[-1]['name'])
    assert ops['LDG.E.64'] >= 2                                      # the mask, one per unrolled step
    assert ops['LDG.E.128'] >= 8 and ops['STG.E.128'] >= 4
    # a misaligned mask falls back to the planner's vector, never to a misaligned wide access
    m = cp.empty((n + 4,), '?')[4:]
    cp.where(m, cp.empty((n,), 'f'), cp.empty((n,), 'f'))
    assert dry[-1]['vec'] == 4
