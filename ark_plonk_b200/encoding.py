"""Wire encodings of the C ABI: field elements as little-endian u64 limbs.

Host-side helpers only (Python ints <-> numpy limb arrays); they do no field arithmetic
beyond the Montgomery change of representation needed to talk to the library.
"""
from __future__ import annotations

import numpy as np

# moduli; index = curve id
FR_MODULUS = (
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
)
FQ_MODULUS = (
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
)


def ints_to_limbs(values, nlimbs: int) -> np.ndarray:
    """canonical ints -> (n, nlimbs) uint64 little-endian limbs."""
    nb = 8 * nlimbs
    buf = b"".join(int(v).to_bytes(nb, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, nlimbs).copy()


def limbs_to_ints(arr: np.ndarray) -> list:
    arr = np.ascontiguousarray(arr, dtype="<u8")
    nb = 8 * arr.shape[-1]
    raw = arr.tobytes()
    return [int.from_bytes(raw[i:i + nb], "little") for i in range(0, len(raw), nb)]


def fr_to_mont(curve: int, values) -> np.ndarray:
    p = FR_MODULUS[curve]
    return ints_to_limbs([(v % p) * (1 << 256) % p for v in values], 4)


def fr_from_mont(curve: int, arr: np.ndarray) -> list:
    p = FR_MODULUS[curve]
    rinv = pow(1 << 256, -1, p)
    return [v * rinv % p for v in limbs_to_ints(arr)]


def fq_to_mont(curve: int, values) -> np.ndarray:
    p = FQ_MODULUS[curve]
    return ints_to_limbs([(v % p) * (1 << 384) % p for v in values], 6)


def fq_from_mont(curve: int, arr: np.ndarray) -> list:
    p = FQ_MODULUS[curve]
    rinv = pow(1 << 384, -1, p)
    return [v * rinv % p for v in limbs_to_ints(arr)]


def g1_affine_to_mont(curve: int, points) -> np.ndarray:
    """[(x, y) | None] -> (n, 12) uint64 packed records; None (infinity) -> zeros."""
    flat = []
    for P in points:
        if P is None:
            flat += [0, 0]
        else:
            flat += [P[0], P[1]]
    return fq_to_mont(curve, flat).reshape(-1, 12)


def g1_from_xyz(curve: int, xyz: np.ndarray):
    """normalised Jacobian (X, Y, Z in {0, 1}) as returned by apb_msm -> (x, y) or None."""
    xyz = np.asarray(xyz, dtype="<u8").reshape(3, 6)
    X, Y, Z = fq_from_mont(curve, xyz)
    if Z == 0:
        return None
    assert Z == 1, "apb_msm returns normalised points"
    return (X, Y)
