"""Generates tests/golden/hotpath_v1.npz: seeded inputs and the expected outputs of the hot
path, computed by the CPU oracle (oracle/oracle.py, oracle/oracle.c).

The reference package (cupy) cannot be imported or run in the build container (no wheel,
no GPU), so these are ORACLE-generated vectors, not reference-generated ones; the oracle is
pinned separately (tests/test_oracle_pinning.py).  Their job is to freeze today's expected
values so that a later change of the oracle, of NumPy, or of a kernel cannot drift silently.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def main():
    rs = np.random.RandomState(20261017)
    out = {}
    # config 2 (small): axpy, bit exact
    x = (rs.rand(4099) * 2 - 1).astype(np.float32)
    y = (rs.rand(4099) * 2 - 1).astype(np.float32)
    out['axpy_x'], out['axpy_y'] = x, y
    out['axpy_a'] = np.float32(1.5)
    out['axpy_z'] = oracle.axpy(1.5, x, y)
    # config 1: x*2+1 (IEEE exact, two ufuncs)
    out['affine_z'] = (x * np.float32(2) + np.float32(1)).astype(np.float32)
    # config 3 (small): axis reductions of float32 / float16
    a = (rs.rand(67, 131) * 2 - 1).astype(np.float32)
    a[5, 7] = a[5, 90] = 0.99999           # a planted tie per row 5
    out['red_a'] = a
    h = a.astype(np.float16)
    out['red_h'] = h
    for name, arr in (('f32', a), ('f16', h)):
        for ax in (0, 1):
            out['sum_%s_ax%d' % (name, ax)] = oracle.sum(arr, axis=ax)
            out['max_%s_ax%d' % (name, ax)] = oracle.amax(arr, axis=ax)
            out['argmax_%s_ax%d' % (name, ax)] = oracle.argmax(arr, axis=ax)
            out['var_%s_ax%d' % (name, ax)] = oracle.var(arr, axis=ax)
    # config 4a (small): exp of a transposed view + broadcast row vector
    t = (rs.rand(12, 40, 36) * 2 - 1).astype(np.float32)
    v = (rs.rand(12) * 2 - 1).astype(np.float32)
    out['exp_t'], out['exp_v'] = t, v
    out['exp_z'] = (oracle.exp_exact(t.transpose(2, 1, 0)).astype(np.float64) + v).astype(np.float32)
    # config 4b (small): int64 cumsum, bit exact
    xi = rs.randint(-(1 << 20), 1 << 20, size=10007).astype(np.int64)
    out['scan_x'] = xi
    out['scan_y'] = oracle.cumsum(xi)
    # config 5 (small): sharded sum / var of a 1-D array (what every rank must end up with)
    s = (rs.rand(8 * 1031) * 2 - 1).astype(np.float32)
    out['shard_x'] = s
    out['shard_sum'] = oracle.sum(s)
    out['shard_var'] = np.asarray(np.var(s.astype(np.float64)))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hotpath_v1.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
