"""ORACLE (test infrastructure, not product code): Philox4x32-10 and the engine's draw layout.

The reference draws its reset noise from JAX threefry keys (train_brax_ppo.py:259-261)
or NumPy PCG64 (envs/hover_env.py:220,227; utils/state.py:98).  The north star replaces
both by a counter-based Philox stream so that a reset is a pure function of
(seed, global env id, episode index); this file is the NumPy statement of that
stream, bit-for-bit what csrc/philox.cuh computes.

Philox4x32-10 is Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11);
the known-answer vectors in tests/test_philox.py are the Random123 ones.

Draw layout (all uint32):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (global_env_id, episode_or_step, block, stream)
    stream  : 0 = reset noise, 1 = policy sampling noise, 2 = synthetic random actions
Uniform float32 in [0, 1): (x >> 8) * 2**-24.  Range map: fl32(fl32(u * (hi - lo)) + lo).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

STREAM_RESET, STREAM_POLICY, STREAM_ACTION = 0, 1, 2


def philox4x32_10(counter, key):
    """counter: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4]."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(key, dtype=np.uint32, copy=True)
    k = np.broadcast_to(k, c.shape[:-1] + (2,)).copy()
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + W0
            k1 = k1 + W1
    return np.stack([c0, c1, c2, c3], axis=-1)


def u01(x):
    """uint32 -> float32 uniform in [0, 1) with 24 random bits."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def uniform(x, lo, hi):
    """fl32(fl32(u*(hi-lo)) + lo), the exact float32 op sequence of the kernel."""
    lo = np.float32(lo); hi = np.float32(hi)
    rng = np.float32(hi - lo)
    return (u01(x) * rng).astype(np.float32) + lo


def draw_blocks(seed, env_ids, second, nblocks, stream):
    """uint32[len(env_ids), 4*nblocks] raw draws for the given counter tuple."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    second = np.broadcast_to(np.asarray(second, dtype=np.uint32), env_ids.shape)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    out = []
    for b in range(nblocks):
        ctr = np.stack([env_ids, second, np.full_like(env_ids, b), np.full_like(env_ids, stream)], axis=-1)
        out.append(philox4x32_10(ctr, key))
    return np.concatenate(out, axis=-1)


def normal_pair(x0, x1):
    """Box-Muller from two uint32 draws -> two float32 standard normals.

    u0 = (x0 >> 8 + 1) * 2^-24 in (0, 1], u1 = (x1 >> 8) * 2^-24 in [0, 1).
    The kernel evaluates the same formula in float32 with its own logf/sincosf,
    so agreement is to float32 rounding of those functions, not bit-exact.
    """
    u0 = ((np.asarray(x0, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    u1 = (np.asarray(x1, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    r = np.sqrt(-2.0 * np.log(u0))
    return r * np.cos(2.0 * np.pi * u1), r * np.sin(2.0 * np.pi * u1)
