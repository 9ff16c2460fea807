"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) in numpy — the oracle's copy of the counter
RNG that the CUDA kernels use (csrc/philox.cuh).  Written from the published algorithm, pinned to
the Random123 known-answer vectors in tests/test_philox.py.  Test infrastructure only."""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_SWEEP = 0x53574550
STREAM_INIT = 0x494E4954
STREAM_SWEEP3 = 0x53575033
STREAM_NOISE = 0x4E4F4953


def philox4x32_10(c0, c1, c2, c3, key):
    """Vectorised over broadcastable uint32 counter arrays; key = 64-bit int seed.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)])
    k0 = int(key) & 0xFFFFFFFF
    k1 = (int(key) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))
