"""ark-serialize 0.3 wire formats of the artefacts around the hot path:
readers and writers so that keys / parameters / proofs produced by arkworks can be loaded into the
resident structures of this library and vice versa.

    reference type                                   (derive at)                         here
    ------------------------------------------------ ----------------------------------- -------------------------
    Proof<F, PC>                                     proof_system/proof.rs:41-103        parse_proof / write_proof
    VerifierKey<F, PC>                               proof_system/widget/mod.rs:137-176  VerifierKey
    VerifierData<F, PC>                              circuit.rs:25-41                    VerifierData
    PublicInputs<F>  (BTreeMap<usize, F>)            proof_system/pi.rs:28-36            write/read_public_inputs
    ProverKey<F>  (serialize_unchecked)              proof_system/widget/mod.rs:285-328  prover_key_to_bytes / Prover.load_prover_key
    MultiSet<F>                                      lookup/multiset.rs:22-34            (inside ProverKey)
    kzg10::UniversalParams / sonic_pc::CommitterKey  (ark-poly-commit, not vendored)     read_powers_of_g / write_powers_of_g

ALL conventions that are RECALLED from the un-vendored arkworks 0.3 crates (they could not be checked
against a Rust build here) live in this one module:
  * field element: canonical integer, little-endian, 32 (Fr) / 48 (Fq) bytes;
  * G1Affine compressed: x, flags in the two top bits of the LAST byte - bit 7: y is the larger of (y, -y),
    bit 6: infinity; uncompressed: x || y with the infinity flag in the last byte of y;
  * usize / lengths: u64 little-endian; Vec<T> / String / BTreeMap: length + items; Option<T>: bool byte + T;
  * DensePolynomial<F> = Vec<F> of coefficients (no trailing zeros); Evaluations<F> = Vec<F> followed by its
    GeneralEvaluationDomain = tag byte (0 = Radix2) + Radix2EvaluationDomain { size: u64, log_size_of_group: u32,
    size_as_field_element, size_inv, group_gen, group_gen_inv, generator_inv };
  * kzg10::UniversalParams and sonic_pc::CommitterKey both START with `powers_of_g: Vec<G1Affine>`.
Nothing here touches the oracle; heavy conversions (canonical <-> Montgomery) run on the device.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from . import encoding as enc
from ._lib import ApbError, Lib, get_lib

CURVE_B = (4, 1)                      # y^2 = x^3 + b
FR_GENERATOR = (7, 22)
FR_TWO_ADICITY = (32, 47)
FLAG_Y_LARGER, FLAG_INFINITY = 0x80, 0x40

# derive order of the reference structs (names = the keys this package uses for commitments / polynomials)
ARITH = ("q_m", "q_l", "q_r", "q_o", "q_4", "q_c", "q_arith")
SIGMAS = ("left_sigma", "right_sigma", "out_sigma", "fourth_sigma")
TABLES = ("table_1", "table_2", "table_3", "table_4")
VK_ORDER = ARITH + ("q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add") + SIGMAS + ("q_lookup",) + TABLES


class SerializationError(ValueError):
    """ark_serialize::SerializationError (InvalidData / UnexpectedFlags / NotEnoughSpace)"""


class Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(bytes(data))
        self.pos = 0

    def take(self, n: int) -> memoryview:
        if self.pos + n > len(self.b):
            raise SerializationError("unexpected end of input at byte %d (+%d)" % (self.pos, n))
        v = self.b[self.pos:self.pos + n]
        self.pos += n
        return v

    def u8(self) -> int:
        return self.take(1)[0]

    def u32(self) -> int:
        return struct.unpack("<I", self.take(4))[0]

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def done(self) -> bool:
        return self.pos == len(self.b)


# ---- scalars and points ---------------------------------------------------------------------------------
def write_fr(curve: int, v: int) -> bytes:
    return (int(v) % enc.FR_MODULUS[curve]).to_bytes(32, "little")


def read_fr(curve: int, r: Reader) -> int:
    v = int.from_bytes(r.take(32), "little")
    if v >= enc.FR_MODULUS[curve]:
        raise SerializationError("scalar not reduced")
    return v


def sqrt_mod(a: int, p: int):
    """square root in F_p (Tonelli-Shanks; BLS12-381 q = 3 mod 4, BLS12-377 q - 1 has 46 factors of two)"""
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
    m, c, t, x = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, x = t * c % p, x * b % p
    return x


def write_g1(curve: int, P, compressed: bool = True) -> bytes:
    """P = (x, y) ints or None (infinity)"""
    p = enc.FQ_MODULUS[curve]
    if P is None:
        out = bytearray(48 if compressed else 96)
        out[-1] |= FLAG_INFINITY
        return bytes(out)
    x, y = P
    if compressed:
        out = bytearray(x.to_bytes(48, "little"))
        if y > p - y:
            out[-1] |= FLAG_Y_LARGER
        return bytes(out)
    return x.to_bytes(48, "little") + y.to_bytes(48, "little")


def read_g1(curve: int, r: Reader, compressed: bool = True, check: bool = True):
    p = enc.FQ_MODULUS[curve]
    if compressed:
        raw = bytearray(r.take(48))
        flags = raw[-1] & 0xC0
        raw[-1] &= 0x3F
        x = int.from_bytes(raw, "little")
        if flags & FLAG_INFINITY:
            if x != 0 or flags & FLAG_Y_LARGER:
                raise SerializationError("unexpected flags on the point at infinity")
            return None
        if x >= p:
            raise SerializationError("x not reduced")
        y = sqrt_mod(x * x * x + CURVE_B[curve], p)
        if y is None:
            raise SerializationError("x is not the abscissa of a curve point")
        if (y > p - y) != bool(flags & FLAG_Y_LARGER):
            y = p - y
        return (x, y)
    raw = bytearray(r.take(96))
    flags = raw[-1] & 0xC0
    raw[-1] &= 0x3F
    x, y = int.from_bytes(raw[:48], "little"), int.from_bytes(raw[48:], "little")
    if flags & FLAG_INFINITY:
        return None
    if check and (x >= p or y >= p or (y * y - x * x * x - CURVE_B[curve]) % p):
        raise SerializationError("point not on the curve")
    return (x, y)


def g1_from_record(curve: int, xyz: np.ndarray):
    """normalised (X, Y, Z) record of apb_msm -> (x, y) | None"""
    return enc.g1_from_xyz(curve, xyz)


# ---- Proof (proof_system/proof.rs:51-103) -------------------------------------------------------------------
PROOF_COMMITMENTS = ("a_comm", "b_comm", "c_comm", "d_comm", "z_comm", "f_comm", "h_1_comm", "h_2_comm", "z_2_comm",
                     "t_1_comm", "t_2_comm", "t_3_comm", "t_4_comm")
PROOF_EVALS = ("a_eval", "b_eval", "c_eval", "d_eval",                                           # WireEvaluations
               "left_sigma_eval", "right_sigma_eval", "out_sigma_eval", "permutation_eval",      # PermutationEvaluations
               "q_lookup_eval", "z2_next_eval", "h1_eval", "h1_next_eval", "h2_eval", "f_eval",  # LookupEvaluations
               "table_eval", "table_next_eval")


def parse_proof(curve: int, blob: bytes) -> dict:
    """serialized Proof -> {name: point | int, "custom_evals": [(label, value)]}; the inverse of write_proof"""
    r = Reader(blob)
    out = {}
    for name in PROOF_COMMITMENTS:
        out[name] = read_g1(curve, r)
    for name in ("aw_opening", "saw_opening"):             # kzg10::Proof { w: G1Affine, random_v: Option<Fr> }
        out[name] = read_g1(curve, r)
        has_v = r.u8()
        if has_v > 1:
            raise SerializationError("bad Option tag")
        out[name + "_random_v"] = read_fr(curve, r) if has_v else None
    for name in PROOF_EVALS:
        out[name] = read_fr(curve, r)
    custom = []
    for _ in range(r.u64()):                               # CustomEvaluations: Vec<(String, F)>
        label = bytes(r.take(r.u64())).decode()
        custom.append((label, read_fr(curve, r)))
    out["custom_evals"] = custom
    if not r.done():
        raise SerializationError("trailing bytes after the proof")
    return out


def write_proof(curve: int, proof: dict) -> bytes:
    out = b"".join(write_g1(curve, proof[name]) for name in PROOF_COMMITMENTS)
    for name in ("aw_opening", "saw_opening"):
        out += write_g1(curve, proof[name])
        v = proof.get(name + "_random_v")
        out += b"\x00" if v is None else b"\x01" + write_fr(curve, v)
    out += b"".join(write_fr(curve, proof[name]) for name in PROOF_EVALS)
    out += struct.pack("<Q", len(proof["custom_evals"]))
    for label, v in proof["custom_evals"]:
        lb = label.encode()
        out += struct.pack("<Q", len(lb)) + lb + write_fr(curve, v)
    return out


# ---- PublicInputs, VerifierKey, VerifierData ----------------------------------------------------------------
def write_public_inputs(curve: int, pi: dict) -> bytes:
    """BTreeMap<usize, F>: length, then (position, value) in ascending position order; zero values are never stored"""
    items = sorted((int(k), int(v) % enc.FR_MODULUS[curve]) for k, v in pi.items())
    items = [(k, v) for k, v in items if v]
    return struct.pack("<Q", len(items)) + b"".join(struct.pack("<Q", k) + write_fr(curve, v) for k, v in items)


def read_public_inputs(curve: int, r: Reader) -> dict:
    out = {}
    for _ in range(r.u64()):
        k = r.u64()
        out[k] = read_fr(curve, r)
    return out


@dataclass
class VerifierKey:
    """widget/mod.rs:148-176: n, then 20 commitments in derive order (VK_ORDER)"""
    curve: int
    n: int
    commitments: dict = field(default_factory=dict)       # name -> (x, y) | None

    def to_bytes(self, compressed: bool = True) -> bytes:
        return struct.pack("<Q", self.n) + b"".join(write_g1(self.curve, self.commitments[k], compressed) for k in VK_ORDER)

    @classmethod
    def from_bytes(cls, curve: int, blob, compressed: bool = True, check: bool = True) -> "VerifierKey":
        r = blob if isinstance(blob, Reader) else Reader(blob)
        n = r.u64()
        vk = cls(curve, n, {k: read_g1(curve, r, compressed, check) for k in VK_ORDER})
        if not isinstance(blob, Reader) and not r.done():
            raise SerializationError("trailing bytes after the verifier key")
        return vk

    @classmethod
    def from_compressed_commitments(cls, curve: int, n: int, comms: dict) -> "VerifierKey":
        """from `plonk.ProverKey.commitments` (name -> 48 compressed bytes, as the transcript absorbs them)"""
        return cls(curve, n, {k: read_g1(curve, Reader(comms[k])) for k in VK_ORDER})


@dataclass
class VerifierData:
    """circuit.rs:32-41: key, then the circuit's PublicInputs"""
    key: VerifierKey
    pi: dict

    def to_bytes(self) -> bytes:
        return self.key.to_bytes() + write_public_inputs(self.key.curve, self.pi)

    @classmethod
    def from_bytes(cls, curve: int, blob: bytes) -> "VerifierData":
        r = Reader(blob)
        key = VerifierKey.from_bytes(curve, r)
        pi = read_public_inputs(curve, r)
        if not r.done():
            raise SerializationError("trailing bytes after the verifier data")
        return cls(key, pi)


# ---- SRS: powers_of_g -----------------------------------------------------------------------------------------
def _fq_limbs_to_mont(lib: Lib, curve: int, canon: np.ndarray) -> np.ndarray:
    """(m, 6) uint64 canonical -> Montgomery, on the device (apb_field_op to_mont)"""
    return lib.field_op(1 if curve == 0 else 3, 3, np.ascontiguousarray(canon), None)


def _fr_limbs(lib: Lib, curve: int, arr: np.ndarray, to_mont: bool) -> np.ndarray:
    return lib.field_op(0 if curve == 0 else 2, 3 if to_mont else 4, np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4), None)


def read_powers_of_g(curve: int, blob: bytes, compressed: bool = True, max_points: int | None = None,
                     lib: Lib | None = None) -> np.ndarray:
    """The leading `powers_of_g: Vec<G1Affine>` of a serialized kzg10::UniversalParams or sonic_pc::CommitterKey
    -> (n, 12) uint64 packed Montgomery records, ready for `apb_ck_upload` / `kzg.CommitterKey` (what `PC::trim`
    hands to the prover, circuit.rs:236,276).  The uncompressed form is vectorised; the compressed form costs one
    modular square root per point on the host."""
    lib = lib or get_lib()
    r = Reader(blob)
    n = r.u64()
    if max_points is not None:
        n = min(n, max_points)
    p = enc.FQ_MODULUS[curve]
    if compressed:
        pts = [read_g1(curve, r, True) for _ in range(n)]
        canon = enc.ints_to_limbs([c for P in pts for c in (P if P is not None else (0, 0))], 6)
    else:
        raw = np.frombuffer(r.take(96 * n), dtype=np.uint8).reshape(n, 96).copy()
        inf = (raw[:, 95] & FLAG_INFINITY) != 0
        raw[:, 95] &= 0x3F
        raw[inf] = 0
        canon = raw.view("<u8").reshape(2 * n, 6)
        # on-curve check of the first and last point (the rest is `deserialize_unchecked` territory)
        for i in {0, n - 1} if n else ():
            x, y = enc.limbs_to_ints(canon[2 * i:2 * i + 2])
            if not inf[i] and (y * y - x * x * x - CURVE_B[curve]) % p:
                raise SerializationError("powers_of_g[%d] is not on the curve" % i)
    return _fq_limbs_to_mont(lib, curve, canon).reshape(n, 12)


def write_powers_of_g(curve: int, records_mont: np.ndarray, compressed: bool = True, lib: Lib | None = None) -> bytes:
    """(n, 12) Montgomery records (e.g. `CommitterKey.download`) -> the `Vec<G1Affine>` bytes"""
    lib = lib or get_lib()
    rec = np.ascontiguousarray(records_mont, dtype=np.uint64).reshape(-1, 12)
    n = rec.shape[0]
    canon = lib.field_op(1 if curve == 0 else 3, 4, rec.reshape(2 * n, 6), None)
    vals = enc.limbs_to_ints(canon)
    out = [struct.pack("<Q", n)]
    for i in range(n):
        x, y = vals[2 * i], vals[2 * i + 1]
        out.append(write_g1(curve, None if x == 0 and y == 0 else (x, y), compressed))
    return b"".join(out)


# ---- ProverKey (serialize_unchecked; widget/mod.rs:292-328) ------------------------------------------------------
def _root_of_unity(curve: int, size: int) -> int:
    p = enc.FR_MODULUS[curve]
    root = pow(FR_GENERATOR[curve], (p - 1) >> FR_TWO_ADICITY[curve], p)
    return pow(root, 1 << (FR_TWO_ADICITY[curve] - (size.bit_length() - 1)), p)


def write_domain(curve: int, size: int) -> bytes:
    """GeneralEvaluationDomain::Radix2 of `size` points"""
    p = enc.FR_MODULUS[curve]
    w = _root_of_unity(curve, size)
    fields = [size % p, pow(size, -1, p), w, pow(w, -1, p), pow(FR_GENERATOR[curve], -1, p)]
    return b"\x00" + struct.pack("<QI", size, size.bit_length() - 1) + b"".join(write_fr(curve, v) for v in fields)


def read_domain(curve: int, r: Reader) -> int:
    if r.u8() != 0:
        raise SerializationError("only Radix2 evaluation domains are supported")
    size, log = r.u64(), r.u32()
    if size != 1 << log:
        raise SerializationError("inconsistent domain size")
    got = r.take(5 * 32)
    if bytes(got) != write_domain(curve, size)[13:]:
        raise SerializationError("domain constants do not match the field's radix-2 domain of size %d" % size)
    return size


def _write_vec_fr(canon: np.ndarray) -> bytes:
    canon = np.ascontiguousarray(canon, dtype="<u8").reshape(-1, 4)
    return struct.pack("<Q", canon.shape[0]) + canon.tobytes()


def _read_vec_fr(r: Reader) -> np.ndarray:
    n = r.u64()
    return np.frombuffer(r.take(32 * n), dtype="<u8").reshape(n, 4).copy()


def prover_key_to_bytes(pk) -> bytes:
    """`ProverKey::serialize_unchecked` of a device-resident `plonk.ProverKey`: downloads the coefficient and
    4n-evaluation vectors, converts Montgomery -> canonical on the device, writes arkworks' layout."""
    lib, curve, n, A = pk.arena.lib, pk.curve, pk.n, pk.arena
    p = enc.FR_MODULUS[curve]
    dom4 = write_domain(curve, 4 * n)

    def canon(off, count):
        return _fr_limbs(lib, curve, A.download(off, count), to_mont=False)

    def poly_evals(name):
        if name not in pk.poly:                               # identically-zero custom selector: no resident vectors
            return struct.pack("<Q", 0) + _write_vec_fr(np.zeros((4 * n, 4), dtype=np.uint64)) + dom4
        c = canon(pk.poly[name], n)
        nz = np.flatnonzero(c.any(axis=1))
        c = c[: int(nz[-1]) + 1] if nz.size else c[:0]        # DensePolynomial keeps no trailing zeros
        return _write_vec_fr(c) + _write_vec_fr(canon(pk.ev4[name], 4 * n)) + dom4

    out = [struct.pack("<Q", n)]
    out += [poly_evals(s) for s in ARITH]
    out += [poly_evals("q_range"), poly_evals("q_logic")]
    out += [poly_evals("q_lookup")] + [_write_vec_fr(canon(off, n)) for off in pk.tables]       # lookup::ProverKey
    out += [poly_evals("q_fixed_group_add"), poly_evals("q_variable_group_add")]
    out += [poly_evals(s) for s in SIGMAS]                                                        # permutation::ProverKey
    out += [_write_vec_fr(canon(pk.ev4["linear"], 4 * n)) + dom4]
    vh_inv = enc.fr_from_mont(curve, pk.vh_inv)
    vh = enc.ints_to_limbs([pow(v, -1, p) for v in vh_inv], 4)                                    # X^n - 1 on the coset: period 4
    out += [_write_vec_fr(np.tile(vh, (n, 1))) + dom4]
    return b"".join(out)


def read_prover_key_arrays(curve: int, blob: bytes) -> dict:
    """parses `ProverKey::serialize_unchecked` bytes into canonical limb arrays:
    {"n", "poly": {name: (len, 4)}, "ev4": {name: (4n, 4)}, "tables": [4 x (n, 4)], "linear": (4n, 4), "v_h": (4n, 4)}"""
    r = Reader(blob)
    n = r.u64()
    if n == 0 or n & (n - 1):
        raise SerializationError("circuit size %d is not a power of two" % n)
    out = {"n": n, "poly": {}, "ev4": {}, "tables": []}

    def evals():
        e = _read_vec_fr(r)
        if read_domain(curve, r) != 4 * n or e.shape[0] != 4 * n:
            raise SerializationError("expected evaluations over the 4n domain")
        return e

    def poly_evals(name):
        c = _read_vec_fr(r)
        if c.shape[0] > n:
            raise SerializationError("polynomial %s has more than n coefficients" % name)
        out["poly"][name] = c
        out["ev4"][name] = evals()

    for s in ARITH:
        poly_evals(s)
    poly_evals("q_range")
    poly_evals("q_logic")
    poly_evals("q_lookup")
    for _ in range(4):
        t = _read_vec_fr(r)
        if t.shape[0] != n:
            raise SerializationError("lookup table column is not padded to n")
        out["tables"].append(t)
    poly_evals("q_fixed_group_add")
    poly_evals("q_variable_group_add")
    for s in SIGMAS:
        poly_evals(s)
    out["linear"] = evals()
    out["v_h"] = evals()
    if not r.done():
        raise SerializationError("trailing bytes after the prover key")
    return out


__all__ = ["SerializationError", "Reader", "write_fr", "read_fr", "write_g1", "read_g1", "parse_proof", "write_proof",
           "write_public_inputs", "read_public_inputs", "VerifierKey", "VerifierData", "read_powers_of_g",
           "write_powers_of_g", "prover_key_to_bytes", "read_prover_key_arrays", "ApbError"]
