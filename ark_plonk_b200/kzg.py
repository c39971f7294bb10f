"""Host-side mirror of the reference's commitment interface over the apb C ABI.

Mirrors `HomomorphicCommitment` / `KZG10<E>` (= SonicKZG10) as plonk-core uses them
(plonk-core/src/commitment.rs:8-49): `trim` -> CommitterKey, `commit`, `open`,
`multi_scalar_mul`.  Polynomials are (n, 4) uint64 arrays of Fr Montgomery limbs
(DensePolynomial coefficients as arkworks stores them); commitments are affine points
(x, y) of Python ints or None for the identity, plus their ark-serialize compressed bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import encoding as enc
from ._lib import ApbError, Lib, get_lib


class CommitterKey:
    """Resident `powers_of_g` (SonicKZG10 CommitterKey after `PC::trim`, circuit.rs:236,276)."""

    def __init__(self, curve: int, powers_of_g_mont: np.ndarray, lib: Lib | None = None):
        self.lib = lib or get_lib()
        self.curve = curve
        pts = np.ascontiguousarray(powers_of_g_mont, dtype=np.uint64).reshape(-1, 12)
        self.n = pts.shape[0]
        h = C.c_void_p()
        self.lib.check(self.lib.c.apb_ck_upload(curve, pts.ctypes.data_as(C.c_void_p), self.n, C.byref(h)))
        self._h = h

    @classmethod
    def from_tau(cls, curve: int, tau: int, n: int, lib: Lib | None = None, generator=None) -> "CommitterKey":
        """`PC::setup` + `trim` with a known tau: [tau^i]G, i < n, generated on the device.
        `generator`: affine (x, y) ints to use instead of the standard G1 generator (KZG10::setup draws a random
        one; a rank of a point-split key passes tau^(first index) G to get its slice of the powers)."""
        from .synth import G1_GENERATOR
        self = cls.__new__(cls)
        self.lib = lib or get_lib()
        self.curve = curve
        self.n = n
        gen = enc.g1_affine_to_mont(curve, [generator if generator is not None else G1_GENERATOR[curve]])
        t = enc.fr_to_mont(curve, [tau])
        h = C.c_void_p()
        self.lib.check(self.lib.c.apb_ck_from_tau(curve, gen.ctypes.data, t.ctypes.data, n, C.byref(h)))
        self._h = h
        return self

    def download(self, first: int, count: int) -> np.ndarray:
        out = np.zeros((count, 12), dtype=np.uint64)
        self.lib.check(self.lib.c.apb_ck_download(self._h, first, count, out.ctypes.data))
        return out

    @property
    def max_degree(self) -> int:
        return self.n - 1

    def close(self):
        if getattr(self, "_h", None):
            self.lib.c.apb_ck_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _skip_leading_zeros(poly: np.ndarray):
    """kzg10 `skip_leading_zeros_and_convert_to_bigints`: drop low-order zero coefficients."""
    nz = np.flatnonzero(poly.any(axis=1))
    if nz.size == 0:
        return 0, poly[:0]
    k = int(nz[0])
    return k, poly[k:int(nz[-1]) + 1]          # also strips trailing zeros (DensePolynomial invariant)


def multi_scalar_mul(ck: CommitterKey, scalars: np.ndarray, base_offset: int = 0, montgomery: bool = False):
    """VariableBaseMSM::multi_scalar_mul over the resident bases -> affine point or None."""
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(18, dtype=np.uint64)
    ck.lib.check(ck.lib.c.apb_msm(ck._h, base_offset, s.ctypes.data_as(C.c_void_p) if s.shape[0] else None,
                                  s.shape[0], 1 if montgomery else 0, out.ctypes.data_as(C.c_void_p)))
    return out


def commit(ck: CommitterKey, polys):
    """`PC::commit(ck, polys, None)`: one MSM per polynomial, batched in a single pass.

    Returns a list of normalised Jacobian records (18 uint64); use `compress` / `to_affine`.
    Raises ApbError(TooManyCoefficients) like kzg10's degree check.
    """
    polys = [np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4) for p in polys]
    k = len(polys)
    if k == 0:
        return []
    trimmed = [_skip_leading_zeros(p) for p in polys]
    for off, t in trimmed:
        if off + t.shape[0] > ck.n:
            raise ApbError(2, "polynomial of degree %d exceeds the %d supported powers" % (off + t.shape[0] - 1, ck.n))
    ptrs = (C.c_void_p * k)(*[t.ctypes.data_as(C.c_void_p) if t.shape[0] else None for _, t in trimmed])
    offs = (C.c_size_t * k)(*[o for o, _ in trimmed])
    lens = (C.c_size_t * k)(*[t.shape[0] for _, t in trimmed])
    out = np.zeros((k, 18), dtype=np.uint64)
    ck.lib.check(ck.lib.c.apb_msm_batch(ck._h, k, ptrs, offs, lens, 1, out.ctypes.data_as(C.c_void_p)))
    return [out[i] for i in range(k)]


def open(ck: CommitterKey, polys, point: int, opening_challenge: int):
    """`PC::open(ck, polys, _, &point, opening_challenge, _, None)` (sonic_pc): p = sum challenge^i p_i,
    witness = p / (X - point) (remainder dropped), proof.w = commit(witness); random_v = None.
    The combination, the division and the MSM run on the device.  Returns the normalised point record."""
    from .plonk import Arena
    lib, curve = ck.lib, ck.curve
    p = enc.FR_MODULUS[curve]
    polys = [np.ascontiguousarray(q, dtype=np.uint64).reshape(-1, 4) for q in polys]
    m = max((q.shape[0] for q in polys), default=0)
    out = np.zeros(18, dtype=np.uint64)
    if m <= 1:
        return out                                         # constant or zero polynomial: witness is zero
    if m - 1 > ck.n:
        raise ApbError(2, "witness polynomial of degree %d exceeds the %d supported powers" % (m - 2, ck.n))
    arena = Arena(lib, (len(polys) + 2) * m + 8)
    try:
        offs = []
        for q in polys:
            o = arena.alloc(max(q.shape[0], 1))
            if q.shape[0]:
                arena.upload(o, q)
            offs.append(o)
        comb, wit = arena.alloc(m), arena.alloc(m)
        k = len(polys)
        ptrs = (C.c_void_p * k)(*[arena.ptr(o) for o in offs])
        lens = (C.c_size_t * k)(*[q.shape[0] for q in polys])
        sc = np.ascontiguousarray(enc.fr_to_mont(curve, [pow(opening_challenge, i, p) for i in range(k)]))
        lib.check(lib.c.apb_fr_lincomb(curve, k, ptrs, lens, sc.ctypes.data, arena.ptr(comb), m))
        z = np.ascontiguousarray(enc.fr_to_mont(curve, [point]))
        lib.check(lib.c.apb_poly_divide_linear(curve, arena.ptr(comb), m, z.ctypes.data, arena.ptr(wit)))
        lib.check(lib.c.apb_msm_dev(ck._h, 0, arena.ptr(wit), m - 1, 1, out.ctypes.data))
    finally:
        arena.close()
    return out


def compress(ck_or_curve, xyz: np.ndarray, lib: Lib | None = None) -> bytes:
    """ark-serialize compressed bytes of a commitment (what the transcript absorbs)."""
    if isinstance(ck_or_curve, CommitterKey):
        return ck_or_curve.lib.g1_compress(ck_or_curve.curve, xyz)
    return (lib or get_lib()).g1_compress(ck_or_curve, xyz)


def to_affine(curve: int, xyz: np.ndarray):
    return enc.g1_from_xyz(curve, xyz)
