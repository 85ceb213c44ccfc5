"""Philox4x32-10: Random123 known answers for the oracle, and oracle == kernel source bit for bit."""
import numpy as np

from oracle import philox

from .util import HostHarness

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        np.testing.assert_array_equal(philox.philox4x32_10(ctr, key), np.array(want, dtype=np.uint32))
        np.testing.assert_array_equal(HostHarness.philox(ctr, key), np.array(want, dtype=np.uint32))


def test_kernel_source_equals_oracle_on_random_counters():
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr = rng.integers(0, 2 ** 32, 4, dtype=np.uint64); key = rng.integers(0, 2 ** 32, 2, dtype=np.uint64)
        np.testing.assert_array_equal(HostHarness.philox(ctr, key), philox.philox4x32_10(ctr.astype(np.uint32), key.astype(np.uint32)))


def test_uniform_range_and_exactness():
    x = np.array([0, 255, 256, 0xffffffff, 0x80000000], dtype=np.uint32)
    u = philox.u01(x)
    assert u[0] == 0 and u[1] == 0 and u[2] == np.float32(2.0 ** -24) and u[3] < 1.0 and u[4] == 0.5
    v = philox.uniform(x, -1.5, 1.5)
    assert v.dtype == np.float32 and v.min() >= -1.5 and v.max() < 1.5
