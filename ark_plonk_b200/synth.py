"""Synthetic inputs for benches and large parity tests (host-side Python ints).

`progression_bases(curve, a, b, n)` returns P_i = [a + i*b]G as affine points using one affine
addition per point; an MSM over them has the closed form [sum_i s_i (a + i b)]G, which lets a
test check an MSM of any size with one scalar multiplication.  Seeded scalars come from
SHA-256 in counter mode.
"""
from __future__ import annotations

import hashlib

import numpy as np

from .encoding import FQ_MODULUS, FR_MODULUS

G1_GENERATOR = (
    (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
     0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1),
    (81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
     241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030),
)


def _affine_add(p, P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return (x3, (lam * (x1 - x3) - y1) % p)


def scalar_mul(curve: int, P, k: int):
    p = FQ_MODULUS[curve]
    k %= FR_MODULUS[curve]
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = _affine_add(p, acc, acc)
        if bit == "1":
            acc = _affine_add(p, acc, P)
    return acc


def progression_bases(curve: int, a: int, b: int, n: int):
    """[(a + i*b) G for i < n] as affine (x, y) ints (None for the identity)."""
    p = FQ_MODULUS[curve]
    G = G1_GENERATOR[curve]
    cur = scalar_mul(curve, G, a)
    step = scalar_mul(curve, G, b)
    out = []
    for _ in range(n):
        out.append(cur)
        cur = _affine_add(p, cur, step)
    return out


def progression_expected(curve: int, a: int, b: int, scalars):
    r = FR_MODULUS[curve]
    e = sum(int(s) * (a + i * b) for i, s in enumerate(scalars)) % r
    return scalar_mul(curve, G1_GENERATOR[curve], e)


def seeded_scalars(curve: int, n: int, seed: bytes = b"apb") -> np.ndarray:
    """n canonical scalars < r as an (n, 4) uint64 array (SHA-256 counter mode, top bits masked)."""
    r = FR_MODULUS[curve]
    nbytes = n * 32
    blocks = []
    ctr = 0
    # hash 32 bytes at a time is slow in Python for 2^26; expand each digest with numpy's PCG instead
    h = hashlib.sha256(seed).digest()
    rng = np.random.Generator(np.random.PCG64(int.from_bytes(h[:16], "little")))
    arr = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    shift = 256 - r.bit_length() + 1            # keep strictly below r: clear top bits so value < 2^(bits-1) < r
    arr[:, 3] >>= np.uint64(shift)
    return arr


def limbs_to_int_list(arr: np.ndarray):
    raw = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    w = 8 * arr.shape[1]
    return [int.from_bytes(raw[i:i + w], "little") for i in range(0, len(raw), w)]
