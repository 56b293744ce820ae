"""Shared helpers for the parity tests: byte <-> big-int conversion in the C-ABI layouts."""
import numpy as np

from oracle.pyref import bls12_381 as bls
from oracle.pyref.algos import SplitMix64

R, P = bls.R, bls.P


def fr_mont_array(vals) -> np.ndarray:
    return np.frombuffer(b"".join(bls.fr_to_mont_bytes(v) for v in vals), dtype=np.uint8).copy()


def fr_from_mont_array(buf) -> list:
    b = bytes(buf)
    return [bls.fr_from_mont_bytes(b[i:i + 32]) for i in range(0, len(b), 32)]


def scalars_array(vals) -> np.ndarray:
    return np.frombuffer(b"".join(bls.int_to_le(v % R, 32) for v in vals), dtype=np.uint8).copy()


def g1_array(pts) -> np.ndarray:
    return np.frombuffer(b"".join(bls.g1_to_ffi(p) for p in pts), dtype=np.uint8).copy()


def g2_array(pts) -> np.ndarray:
    return np.frombuffer(b"".join(bls.g2_to_ffi(p) for p in pts), dtype=np.uint8).copy()


def g1_list(buf) -> list:
    b = bytes(buf)
    return [bls.g1_from_ffi(b[i:i + 96]) for i in range(0, len(b), 96)]


def g2_list(buf) -> list:
    b = bytes(buf)
    return [bls.g2_from_ffi(b[i:i + 192]) for i in range(0, len(b), 192)]


def rand_fr(seed: int, n: int) -> list:
    g = SplitMix64(seed)
    return [g.fr() for _ in range(n)]


def rand_fr_bytes_fast(seed: int, n: int) -> np.ndarray:
    """n random 32-byte values < 2^254 (< r): valid both as canonical scalars and as Montgomery words."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    a[:, 31] &= 0x3F
    return a.reshape(-1)


def le_ints(buf, width: int = 32) -> list:
    b = bytes(buf)
    return [int.from_bytes(b[i:i + width], "little") for i in range(0, len(b), width)]
