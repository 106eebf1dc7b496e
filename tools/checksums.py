"""Vectorised, order-sensitive 64-bit checksum of a uint32/int32 array (numpy only).

Used for the per-frame fingerprints in tests/golden/fingerprints.json (`*_mix64` fields) so that
the GPU tests and bench.py can verify all 154 reference frames without executing oracle/ code.
"""
import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)


def mix64(arr) -> str:
    a = np.ascontiguousarray(arr).view(np.uint32).astype(np.uint64).ravel()
    with np.errstate(over="ignore"):
        h = (a + _M1) * (np.arange(a.size, dtype=np.uint64) * np.uint64(2) + np.uint64(1))
        h ^= h >> np.uint64(29)
        h *= _M2
        h ^= h >> np.uint64(32)
        total = np.add.reduce(h, dtype=np.uint64) if a.size else np.uint64(0)
    return f"{int(total) ^ a.size:016x}"
