"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

NumPy restatement of Philox4x32-10 (Salmon et al., SC'11; Random123 constants)
and of this repo's draw-slot contract (DESIGN.md "RNG contract").  The
reference uses one global MT19937 stream seeded once (PKG/trainer.py:45); a
batched device run cannot share a sequential stream, so the contract below
replaces it and the parity harness feeds the SAME draws to the reference
modules by patching ``np.random.uniform/randint/normal``.

Counter  = (env index in population, global step index, purpose, population id)
Key      = (seed & 0xffffffff, seed >> 32)
purpose 0 (every agent step)  : out[0] explore draw, out[1] random action,
                                out[2] table-pick draw (consumed, ignored: PKG/double_q_learning.py:102)
purpose 1 (every episode reset): out[0] x_init radius / uniform, out[1] Box-Muller angle,
                                out[2] platform phase (u32 turns)
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

PURPOSE_STEP = 0
PURPOSE_RESET = 1
PURPOSE_RESET_NOISE = 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over arrays of uint32 counters; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def draws(seed: int, population: int, env, step, purpose: int):
    """The contract: 4 uint32 words for (env, step, purpose) of one population."""
    env = np.asarray(env, dtype=np.uint64)
    step = np.broadcast_to(np.asarray(step, dtype=np.uint64), env.shape)
    pur = np.full(env.shape, purpose, dtype=np.uint64)
    pop = np.full(env.shape, population, dtype=np.uint64)
    return philox4x32_10(env, step, pur, pop, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def u24(x):
    """Top 24 bits of a draw; the uniform is u24 * 2**-24, exact in fp32 and fp64."""
    return np.asarray(x, dtype=np.uint32) >> np.uint32(8)


def uniform01(x) -> np.ndarray:
    return u24(x).astype(np.float64) * (2.0 ** -24)


def random_action(x) -> np.ndarray:
    """randint(3) replacement: mulhi(x, 3)."""
    return ((np.asarray(x, dtype=np.uint64) * np.uint64(3)) >> np.uint64(32)).astype(np.int64)
