"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) pin the oracle's RNG; the CUDA copy
is compared with the oracle's in test_ising_gpu.py (hot start and sweeps are bit-exact)."""
import numpy as np

from oracle.philox import philox4x32_10

KAT = [
    ((0, 0, 0, 0), 0, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, 0xFFFFFFFFFFFFFFFF, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0x299F31D0 << 32) | 0xA4093822,
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(*ctr, key)
        assert tuple(int(x) for x in got) == want


def test_vectorised_matches_scalar():
    c0 = np.arange(7, dtype=np.uint64)
    out = philox4x32_10(c0, 5, 9, 3, 1234)
    for n in range(7):
        one = philox4x32_10(n, 5, 9, 3, 1234)
        assert [int(o[n]) for o in out] == [int(x) for x in one]
