"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the NumPy statement and for the
kernel code (host build), and the reset sampler / action stream of the kernels against NumPy."""
import ctypes as C

import numpy as np

import hostsim
import philox_ref as pr
from parity_utils import env_params

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    L = hostsim.lib()
    for ctr, key, exp in KAT:
        assert tuple(int(x) for x in pr.philox4x32_10(np.array(ctr, np.uint32), key)) == exp
        out = (C.c_uint32 * 4)()
        L.hs_philox(*ctr, *key, out)
        assert tuple(out) == exp


def test_reset_sampler_matches_numpy_statement():
    ep = env_params()
    n, off, seed = 257, 10_000_000_000, 0x1234_5678_9ABC
    hb = hostsim.HostSimBatch(ep, n)
    hb.envs["episode"] = np.arange(n) % 5
    from rl_rocket_6dof_b200._lib import R6Buffers
    b = R6Buffers()
    hostsim.lib().hs_reset(C.byref(hb.p), C.byref(b), hb.envs.ctypes.data, n, off, seed)
    ic = pr.normalize_ic_quaternion(pr.sample_ic(ep.ic_low, ep.ic_high, seed, off + np.arange(n), np.arange(n) % 5))
    assert np.array_equal(hb.envs["y"], ic.astype(np.float64))
    assert np.array_equal(hb.envs["m0"], ic[:, 13])
    assert np.array_equal(hb.envs["episode"], np.arange(n) % 5 + 1)
    # samples live in the init box and the constant components are exact
    cols = [0, 1, 2, 3, 4, 5, 10, 11, 12, 13]      # the quaternion is re-normalised afterwards
    assert np.all(ic[:, cols] >= ep.ic_low[cols]) and np.all(ic[:, cols] <= ep.ic_high[cols])
    assert np.all(ic[:, [2, 5]] == 0)
    assert np.abs(np.linalg.norm(ic[:, 6:10].astype(np.float64), axis=1) - 1).max() < 2e-7


def test_action_stream_matches_numpy_statement():
    L = hostsim.lib()
    a = (C.c_float * 3)()
    for genv, step in [(0, 0), (5, 17), (2 ** 33 + 7, 2 ** 32 + 3), (123456789, 999)]:
        L.hs_philox_action(42, genv, step, a)
        ref = pr.actions(42, np.array([genv]), np.array([step]))[0]
        assert np.array_equal(np.array(a[:], np.float32), ref)
        assert np.all(np.abs(ref) <= 1)
