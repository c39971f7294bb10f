"""ark-serialize 0.3 compressed encodings used by the transcript and the Proof.

Oracle = test infrastructure (see oracle/__init__.py).

Call sites: plonk-core/src/transcript.rs:27-33 (`append` = CanonicalSerialize::serialize
into a Vec), proof.rs:41-103 (Proof derive).  ark-serialize is not vendored; format restated
from SURVEY.md Appendix B:
  * Fr / Fq: canonical integer, little-endian, 32 / 48 bytes.
  * G1Affine compressed: x little-endian, 48 bytes; flags in the top bits of the LAST byte:
    bit 7 = y is the lexicographically larger of (y, -y) ; bit 6 = infinity (x = 0).
  * Option<T> = bool byte (+ T); Vec/String = u64-LE length + items; usize = u64-LE.
"""
from __future__ import annotations

from .curves import Curve
from .fields import Field


def ser_field(field: Field, v: int) -> bytes:
    return (v % field.p).to_bytes(field.nbytes, "little")


def ser_g1(curve: Curve, P) -> bytes:
    nb = curve.fq.nbytes
    if P is None:
        out = bytearray(nb)
        out[-1] |= 0x40
        return bytes(out)
    x, y = P
    p = curve.fq.p
    out = bytearray(x.to_bytes(nb, "little"))
    if y > (p - y) % p:
        out[-1] |= 0x80
    return bytes(out)


def deser_g1(curve: Curve, b: bytes):
    nb = curve.fq.nbytes
    assert len(b) == nb
    flags = b[-1] & 0xC0
    if flags & 0x40:
        return None
    x = int.from_bytes(b[:-1] + bytes([b[-1] & 0x3F]), "little")
    p = curve.fq.p
    rhs = (x * x * x + curve.b) % p
    y = sqrt_mod(rhs, p)
    if y is None:
        raise ValueError("x not on curve")
    neg = (p - y) % p
    big, small = (y, neg) if y > neg else (neg, y)
    return (x, big if flags & 0x80 else small)


def sqrt_mod(a: int, p: int):
    """Tonelli-Shanks (BLS12-381 q = 3 mod 4; BLS12-377 q has high 2-adicity)."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


def ser_u64(v: int) -> bytes:
    return v.to_bytes(8, "little")


def ser_string(s: str) -> bytes:
    b = s.encode()
    return ser_u64(len(b)) + b


def ser_kzg_proof(curve: Curve, w) -> bytes:
    """kzg10::Proof { w: G1Affine, random_v: Option<Fr> } with random_v = None."""
    return ser_g1(curve, w) + b"\x00"
