"""Pins the oracle's XXH64 (oracle/xxh64.h) against the public algorithm:
canonical vectors, python-xxhash 3.7.0 and (when present) the system libxxhash.
Rapid's ring order and configuration id hang off this hash
(MembershipView.java:47, :548-553, :568, :580-581; zero-allocation-hashing 0.8 is not vendored)."""
import ctypes
import random
import struct

import pytest


def test_canonical_vectors(orc):
    assert orc.xxh64(b"", 0) == 0xEF46DB3751D8E999
    assert orc.xxh64(b"a", 0) == 0xD24EC4F1A98C6E5B
    assert orc.xxh64(b"abc", 0) == 0x44BC2CF5AD770999
    # "Nobody inspects the spammish repetition" — the vector in the xxHash README family
    assert orc.xxh64(b"Nobody inspects the spammish repetition", 0) == 0xFBCEA83C8A378BF1


def test_against_python_xxhash(orc):
    xxhash = pytest.importorskip("xxhash")
    rng = random.Random(7)
    for n in list(range(0, 80)) + [127, 128, 129, 1000]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        for seed in (0, 1, 2, 9, 0xDEADBEEF, 2**63 + 5):
            assert orc.xxh64(data, seed) == xxhash.xxh64(data, seed=seed).intdigest(), (n, seed)


def test_hash_int_long_are_le_bytes(orc):
    xxhash = pytest.importorskip("xxhash")
    for seed in range(10):
        for v in (0, 1, -1, 1234, 65535, 2**31 - 1, -(2**31)):
            assert orc.xx_hash_int(v, seed) == xxhash.xxh64(struct.pack("<i", v), seed=seed).intdigest()
        for v in (0, 1, -1, 2**63 - 1, -(2**63), 0x0123456789ABCDEF):
            assert orc.xx_hash_long(v, seed) == xxhash.xxh64(struct.pack("<q", v), seed=seed).intdigest()


def test_against_system_libxxhash(orc):
    try:
        lx = ctypes.CDLL("libxxhash.so.0")
    except OSError:
        pytest.skip("no system libxxhash")
    lx.XXH64.restype = ctypes.c_uint64
    lx.XXH64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
    rng = random.Random(11)
    for n in (0, 1, 3, 4, 7, 8, 9, 15, 31, 32, 33, 63, 64, 100):
        data = bytes(rng.getrandbits(8) for _ in range(n))
        for seed in (0, 3, 9):
            assert orc.xxh64(data, seed) == lx.XXH64(data, n, seed)
