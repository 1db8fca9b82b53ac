"""CPU tier: the committed golden vectors still agree with the oracle (guards drift of the
oracle / NumPy between the build container and the GPU box)."""
import os

import numpy as np

from oracle import oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hotpath_v1.npz'))


def test_golden_axpy_and_affine():
    np.testing.assert_array_equal(oracle.axpy(G['axpy_a'], G['axpy_x'], G['axpy_y']), G['axpy_z'])
    np.testing.assert_array_equal(G['axpy_x'] * np.float32(2) + np.float32(1), G['affine_z'])


def test_golden_reductions():
    for name, arr in (('f32', G['red_a']), ('f16', G['red_h'])):
        for ax in (0, 1):
            np.testing.assert_array_equal(oracle.sum(arr, axis=ax), G['sum_%s_ax%d' % (name, ax)])
            np.testing.assert_array_equal(oracle.amax(arr, axis=ax), G['max_%s_ax%d' % (name, ax)])
            np.testing.assert_array_equal(oracle.argmax(arr, axis=ax), G['argmax_%s_ax%d' % (name, ax)])
            np.testing.assert_array_equal(oracle.var(arr, axis=ax), G['var_%s_ax%d' % (name, ax)])
    assert G['argmax_f32_ax1'][5] == 7            # the planted tie resolves to the lowest index


def test_golden_scan_and_exp():
    np.testing.assert_array_equal(oracle.cumsum(G['scan_x']), G['scan_y'])
    want = (oracle.exp_exact(G['exp_t'].transpose(2, 1, 0)).astype(np.float64) + G['exp_v']).astype(np.float32)
    np.testing.assert_array_equal(want, G['exp_z'])
