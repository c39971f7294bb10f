"""Prime fields of BLS12-381 / BLS12-377 and their Montgomery limb encodings.

Oracle = test infrastructure (see oracle/__init__.py).

Restates ark-ff 0.3 `Fp256<FrParameters>` / `Fp384<FqParameters>` as used by
plonk-core (dependency pinned `ark-ff = "0.3"`, plonk-core/Cargo.toml:50-59):
elements are stored fully reduced in Montgomery form a*R mod p with
R = 2^256 (Fr) or 2^384 (Fq), as little-endian u64 limbs.  Constants are the
ones listed in SURVEY.md Appendix D and are re-derived in
tests/test_oracle_fields.py (primality, generator order, two-adicity).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Field:
    name: str
    p: int            # modulus
    limbs64: int      # number of u64 limbs (4 for Fr, 6 for Fq)
    generator: int = 0        # multiplicative generator (Fr only)
    two_adicity: int = 0      # Fr only

    @property
    def bits(self) -> int:
        return self.p.bit_length()

    @property
    def R(self) -> int:
        return 1 << (64 * self.limbs64)

    @property
    def nbytes(self) -> int:
        return 8 * self.limbs64

    # -- Montgomery form -------------------------------------------------
    def to_mont(self, a: int) -> int:
        return (a % self.p) * self.R % self.p

    def from_mont(self, am: int) -> int:
        return am * pow(self.R, -1, self.p) % self.p

    def mont_bytes(self, a: int) -> bytes:
        """a (canonical) -> little-endian bytes of its Montgomery limbs."""
        return self.to_mont(a).to_bytes(self.nbytes, "little")

    def from_mont_bytes(self, b: bytes) -> int:
        return self.from_mont(int.from_bytes(b, "little"))

    # -- arithmetic ------------------------------------------------------
    def inv(self, a: int) -> int:
        return pow(a, -1, self.p)

    def two_adic_root(self) -> int:
        """generator^((p-1)/2^two_adicity): ark-ff FftParameters::TWO_ADIC_ROOT_OF_UNITY."""
        return pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)

    # Montgomery constants the CUDA side needs (checked in tests)
    def n0inv32(self) -> int:
        """-p^{-1} mod 2^32."""
        return (-pow(self.p, -1, 1 << 32)) % (1 << 32)

    def r2(self) -> int:
        return self.R * self.R % self.p


# ---- BLS12-381 (ark-bls12-381 0.3) ------------------------------------------
FR381 = Field(
    "Fr381",
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    4, generator=7, two_adicity=32)
FQ381 = Field(
    "Fq381",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    6)

# ---- BLS12-377 (ark-bls12-377 0.3) ------------------------------------------
FR377 = Field(
    "Fr377",
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    4, generator=22, two_adicity=47)
FQ377 = Field(
    "Fq377",
    0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    6)

CURVE_BLS12_381 = 0
CURVE_BLS12_377 = 1

FR = {CURVE_BLS12_381: FR381, CURVE_BLS12_377: FR377}
FQ = {CURVE_BLS12_381: FQ381, CURVE_BLS12_377: FQ377}


def pack_mont(field: Field, values) -> bytes:
    """list of canonical ints -> concatenated Montgomery LE limb bytes (the C-ABI wire format)."""
    R, p, nb = field.R, field.p, field.nbytes
    return b"".join(((v % p) * R % p).to_bytes(nb, "little") for v in values)


def unpack_mont(field: Field, buf: bytes) -> list:
    nb = field.nbytes
    rinv = pow(field.R, -1, field.p)
    return [int.from_bytes(buf[i:i + nb], "little") * rinv % field.p
            for i in range(0, len(buf), nb)]


def pack_canonical(field: Field, values) -> bytes:
    nb = field.nbytes
    return b"".join((v % field.p).to_bytes(nb, "little") for v in values)


def unpack_canonical(field: Field, buf: bytes) -> list:
    nb = field.nbytes
    return [int.from_bytes(buf[i:i + nb], "little") for i in range(0, len(buf), nb)]
