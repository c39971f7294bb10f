"""G1 of BLS12-381 / BLS12-377: short Weierstrass y^2 = x^3 + b, a = 0.

Oracle = test infrastructure (see oracle/__init__.py).

Restates ark-ec 0.3 `short_weierstrass_jacobian::{GroupAffine, GroupProjective}`
and `msm::VariableBaseMSM::multi_scalar_mul` (called at
plonk-core/src/commitment.rs:45 and inside every KZG commit/open,
plonk-core/src/proof_system/prover.rs:213,290,313,316,362,388,459,579,582,606,609).
Points here are affine tuples (x, y) of canonical ints, `None` = infinity.
"""
from __future__ import annotations

from dataclasses import dataclass

from .fields import FQ377, FQ381, FR377, FR381, Field, CURVE_BLS12_377, CURVE_BLS12_381


@dataclass(frozen=True)
class Curve:
    cid: int
    name: str
    fq: Field
    fr: Field
    b: int
    gx: int
    gy: int

    @property
    def G(self):
        return (self.gx, self.gy)

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        x, y = P
        p = self.fq.p
        return (y * y - x * x * x - self.b) % p == 0

    # ---- Jacobian arithmetic (X, Y, Z), Z == 0 -> infinity ----------------
    def to_jac(self, P):
        return (0, 1, 0) if P is None else (P[0], P[1], 1)

    def to_affine(self, J):
        X, Y, Z = J
        if Z == 0:
            return None
        p = self.fq.p
        zi = pow(Z, -1, p)
        zi2 = zi * zi % p
        return (X * zi2 % p, Y * zi2 % p * zi % p)

    def jdouble(self, J):
        X, Y, Z = J
        p = self.fq.p
        if Z == 0 or Y == 0:
            return (0, 1, 0)
        A = X * X % p
        B = Y * Y % p
        C = B * B % p
        D = 2 * ((X + B) * (X + B) - A - C) % p
        E = 3 * A % p
        F = E * E % p
        X3 = (F - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def jadd(self, J1, J2):
        p = self.fq.p
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        if Z1 == 0:
            return J2
        if Z2 == 0:
            return J1
        Z1Z1 = Z1 * Z1 % p
        Z2Z2 = Z2 * Z2 % p
        U1 = X1 * Z2Z2 % p
        U2 = X2 * Z1Z1 % p
        S1 = Y1 * Z2 % p * Z2Z2 % p
        S2 = Y2 * Z1 % p * Z1Z1 % p
        if U1 == U2:
            if S1 == S2:
                return self.jdouble(J1)
            return (0, 1, 0)
        H = (U2 - U1) % p
        Rr = (S2 - S1) % p
        HH = H * H % p
        HHH = H * HH % p
        V = U1 * HH % p
        X3 = (Rr * Rr - HHH - 2 * V) % p
        Y3 = (Rr * (V - X3) - S1 * HHH) % p
        Z3 = Z1 * Z2 % p * H % p
        return (X3, Y3, Z3)

    def jadd_mixed(self, J1, P2):
        if P2 is None:
            return J1
        return self.jadd(J1, (P2[0], P2[1], 1))

    # ---- affine convenience ------------------------------------------------
    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.fq.p)

    def add(self, P, Q):
        return self.to_affine(self.jadd(self.to_jac(P), self.to_jac(Q)))

    def mul(self, P, k: int):
        """[k]P by left-to-right double-and-add (k reduced mod r)."""
        k %= self.fr.p
        acc = (0, 1, 0)
        if P is None or k == 0:
            return None
        J = self.to_jac(P)
        for bit in bin(k)[2:]:
            acc = self.jdouble(acc)
            if bit == "1":
                acc = self.jadd(acc, J)
        return self.to_affine(acc)

    # ---- MSM ---------------------------------------------------------------
    def msm_naive(self, bases, scalars):
        acc = (0, 1, 0)
        for P, s in zip(bases, scalars):
            Q = self.mul(P, s)
            if Q is not None:
                acc = self.jadd(acc, self.to_jac(Q))
        return self.to_affine(acc)

    def msm_pippenger(self, bases, scalars):
        """ark-ec 0.3 msm/variable_base.rs algorithm (unsigned windows, c from ln-ish rule).

        size = min(len); c = 3 if size < 32 else ceil_log2(size)*69/100 + 2; windows over
        MODULUS_BITS step c; per window 2^c - 1 Jacobian buckets with mixed adds; running-sum;
        fold high -> low with c doublings.
        """
        size = min(len(bases), len(scalars))
        if size == 0:
            return None
        if size < 32:
            c = 3
        else:
            lg = (size - 1).bit_length()           # ceil(log2(size)) == ark_std::log2
            c = lg * 69 // 100 + 2
        nbits = self.fr.bits
        window_sums = []
        for w_start in range(0, nbits, c):
            res = (0, 1, 0)
            buckets = [(0, 1, 0)] * ((1 << c) - 1)
            for P, s in zip(bases[:size], scalars[:size]):
                if s == 0:
                    continue
                if s == 1:
                    if w_start == 0:
                        res = self.jadd_mixed(res, P)
                    continue
                d = (s >> w_start) % (1 << c)
                if d != 0:
                    buckets[d - 1] = self.jadd_mixed(buckets[d - 1], P)
            running = (0, 1, 0)
            for bkt in reversed(buckets):
                running = self.jadd(running, bkt)
                res = self.jadd(res, running)
            window_sums.append(res)
        lowest = window_sums[0]
        total = (0, 1, 0)
        for ws in reversed(window_sums[1:]):
            total = self.jadd(total, ws)
            for _ in range(c):
                total = self.jdouble(total)
        return self.to_affine(self.jadd(lowest, total))


BLS12_381 = Curve(
    CURVE_BLS12_381, "BLS12-381", FQ381, FR381, 4,
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)

BLS12_377 = Curve(
    CURVE_BLS12_377, "BLS12-377", FQ377, FR377, 1,
    81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
    241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030)

CURVES = {CURVE_BLS12_381: BLS12_381, CURVE_BLS12_377: BLS12_377}


def powers_of_tau_g1(curve: Curve, tau: int, n: int):
    """[tau^i]G for i < n as affine points (what KZG10 `powers_of_g` holds).

    Uses one scalar multiplication per point in Jacobian form and one shared batch inversion.
    """
    r, p = curve.fr.p, curve.fq.p
    jac = []
    t = 1
    # fixed-base windowed table for speed: 4-bit windows over 256 bits
    win = 8
    nwin = (curve.fr.bits + win - 1) // win
    table = []
    base = curve.to_jac(curve.G)
    for _ in range(nwin):
        row = [(0, 1, 0)]
        for _j in range((1 << win) - 1):
            row.append(curve.jadd(row[-1], base))
        table.append(row)
        for _j in range(win):
            base = curve.jdouble(base)
    for _ in range(n):
        acc = (0, 1, 0)
        k = t
        for w in range(nwin):
            d = (k >> (w * win)) & ((1 << win) - 1)
            if d:
                acc = curve.jadd(acc, table[w][d])
        jac.append(acc)
        t = t * tau % r
    # batch to-affine
    out = [None] * n
    prefix = []
    acc = 1
    for (_, _, Z) in jac:
        prefix.append(acc)
        if Z:
            acc = acc * Z % p
    inv = pow(acc, -1, p)
    for i in range(n - 1, -1, -1):
        X, Y, Z = jac[i]
        if Z == 0:
            continue
        zi = inv * prefix[i] % p
        inv = inv * Z % p
        zi2 = zi * zi % p
        out[i] = (X * zi2 % p, Y * zi2 % p * zi % p)
    return out
