"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the engine's counter-based RNG.

The reference draws from the global, never-seeded ``numpy.random`` state (SURVEY.md 8b
"Threading"), so no stream-level parity with it is possible.  The engine instead uses
Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) keyed by
(seed, purpose tag) and counted by (draw index, a, b, chain); this module is the bit-exact CPU
statement of that generator, so the oracle can reproduce every device-side draw.
Known-answer vectors from the Random123 distribution are checked in tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)

# purpose tags (xor-ed into the high key word); must match smjp_b200/csrc/philox.cuh
TAG_CANDIDATES = 0x1001
TAG_BACKWARD = 0x1002
TAG_PRIOR = 0x1003
TAG_PF_PROPAGATE = 0x1004
TAG_PF_RESAMPLE = 0x1005
TAG_DATA = 0x1006


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcastable uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3, k0, k1 = [np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3, k0, k1)]
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    c0, c1, c2, c3, k0, k1 = [x.copy() for x in (c0, c1, c2, c3, k0, k1)]
    for r in range(10):
        if r > 0:
            k0 = (k0 + W0) & MASK
            k1 = (k1 + W1) & MASK
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK
        hi1, lo1 = p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
    return c0, c1, c2, c3


def uniform53(x_hi, x_lo):
    """Two 32-bit words -> double in [0, 1) with 53 random bits."""
    x_hi = np.asarray(x_hi, dtype=np.uint64)
    x_lo = np.asarray(x_lo, dtype=np.uint64)
    return ((x_hi >> np.uint64(5)) * np.float64(67108864.0) + (x_lo >> np.uint64(6))) * np.float64(2.0 ** -53)


def u01(seed, tag, r, a, b, chain):
    """Uniform double number ``r`` of stream (seed, tag, a, b, chain).

    counter = (r >> 1, a, b, chain); key = (seed & 0xffffffff, (seed >> 32) ^ tag);
    even r takes output words (0, 1), odd r words (2, 3).
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64(((seed >> 32) ^ tag) & 0xFFFFFFFF)
    r = np.asarray(r, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10((r >> np.uint64(1)) & MASK, a, b, chain, k0, k1)
    odd = (r & np.uint64(1)).astype(bool)
    hi = np.where(odd, x2, x0)
    lo = np.where(odd, x3, x1)
    return uniform53(hi, lo)
