"""TEST INFRASTRUCTURE (see oracle/__init__.py).  Custom gate constraint polynomials of the reference's
widgets, over python ints, and the embedded twisted Edwards curves they are parameterised by.

Each `*_constraints` function is GateConstraint::constraints of one widget; the quotient term is
selector(X) * constraints(separation challenge, wire values, next-row wire values) evaluated point-wise
(plonk-core/src/proof_system/widget/mod.rs:105-133) and the linearisation term is
selector_poly * constraints(evaluations) (:135-150).

Embedded curves (third-party crates, absent from /root/reference; constants restated from the crates'
published parameters and pinned by tests/test_oracle_gates.py through the group order check):
  ark-ed-on-bls12-381 (JubJub):  a = -1, d = -(10240/10241)
  ark-ed-on-bls12-377:           a = -1, d = 3021
"""
from __future__ import annotations

from .curves import BLS12_377, BLS12_381, Curve

_JUBJUB_D = 19257038036680949359750312669786877991949435402254120286184196891950884077233
# (a, d, prime subgroup order, cofactor)
EMBEDDED = {
    BLS12_381.name: (BLS12_381.fr.p - 1, _JUBJUB_D,
                     6554484396890773809930967563523245729705921265872317281365359162392183254199, 8),
    BLS12_377.name: (BLS12_377.fr.p - 1, 3021,
                     2111115437357092606062206234695386632838870926408408195193685246394721360383, 4),
}


def embedded_params(curve: Curve):
    a, d, _, _ = EMBEDDED[curve.name]
    return a, d


# ---- twisted Edwards arithmetic (affine, complete for a = -1 with non-square d) --------------------
def te_add(P, Q, a, d, p):
    x1, y1 = P
    x2, y2 = Q
    t = d * x1 % p * x2 % p * y1 % p * y2 % p
    x3 = (x1 * y2 + y1 * x2) % p * pow(1 + t, -1, p) % p
    y3 = (y1 * y2 - a * x1 % p * x2) % p * pow(1 - t, -1, p) % p
    return (x3, y3)


def te_neg(P, p):
    return ((-P[0]) % p, P[1])


def te_mul(P, k, a, d, p):
    R = (0, 1)
    for bit in bin(k)[2:] if k else "":
        R = te_add(R, R, a, d, p)
        if bit == "1":
            R = te_add(R, P, a, d, p)
    return R


def te_on_curve(P, a, d, p):
    x, y = P
    return (a * x * x + y * y - 1 - d * x * x % p * y * y) % p == 0


def fr_sqrt(v, p):
    """Tonelli-Shanks (the scalar fields have 2-adicity 32 / 47); None for a non-residue"""
    v %= p
    if v == 0:
        return 0
    if pow(v, (p - 1) // 2, p) != 1:
        return None
    s, q = 0, p - 1
    while q % 2 == 0:
        s += 1
        q //= 2
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(v, q, p), pow(v, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return r


def te_point_from_x(curve: Curve, x0: int):
    """first point with x >= x0 on the embedded curve, multiplied into the prime-order subgroup"""
    a, d, _, cof = EMBEDDED[curve.name]
    p = curve.fr.p
    x = x0
    while True:
        num = (1 - a * x * x) % p
        den = (1 - d * x * x) % p
        y = fr_sqrt(num * pow(den, -1, p) % p, p) if den else None
        if y is not None:
            P = te_mul((x % p, y), cof, a, d, p)
            if P != (0, 1):
                return P
        x += 1


def find_wnaf2(k: int):
    """ark-ff BigInteger::find_wnaf(2): digits in {-1, 0, 1}, least significant first"""
    out = []
    while k:
        if k & 1:
            z = k % 4
            if z >= 2:
                z -= 4
            k -= z
        else:
            z = 0
        out.append(z)
        k >>= 1
    return out


# ---- gate constraints ------------------------------------------------------------------------------
def delta(f, p):
    """f (f-1) (f-2) (f-3)   (widget/range.rs:64-73, widget/logic.rs:100-108)"""
    return f * (f - 1) % p * (f - 2) % p * (f - 3) % p


def range_constraints(sep, a, b, c, d, d_next, p):
    """Range::constraints (widget/range.rs:46-62)"""
    kappa = sep * sep % p
    kappa_sq = kappa * kappa % p
    kappa_cu = kappa_sq * kappa % p
    b1 = delta((c - 4 * d) % p, p)
    b2 = delta((b - 4 * c) % p, p) * kappa
    b3 = delta((a - 4 * b) % p, p) * kappa_sq
    b4 = delta((d_next - 4 * a) % p, p) * kappa_cu
    return (b1 + b2 + b3 + b4) % p * sep % p


def delta_xor_and(a, b, w, c, q_c, p):
    """widget/logic.rs:119-141"""
    F = w * (w * (4 * w - 18 * (a + b) + 81) % p + 18 * (a * a + b * b) - 81 * (a + b) + 83) % p
    E = (3 * (a + b + c) - 2 * F) % p
    B = q_c * (9 * c - 3 * (a + b)) % p
    return (B + E) % p


def logic_constraints(sep, a_val, b_val, c_val, d_val, a_next, b_next, d_next, q_c, p):
    """Logic::constraints (widget/logic.rs:66-98)"""
    kappa = sep * sep % p
    kappa_sq = kappa * kappa % p
    kappa_cu = kappa_sq * kappa % p
    kappa_qu = kappa_cu * kappa % p
    a = (a_next - 4 * a_val) % p
    b = (b_next - 4 * b_val) % p
    d = (d_next - 4 * d_val) % p
    w = c_val
    c0 = delta(a, p)
    c1 = delta(b, p) * kappa
    c2 = delta(d, p) * kappa_sq
    c3 = (w - a * b) % p * kappa_cu
    c4 = delta_xor_and(a, b, w, d, q_c, p) * kappa_qu
    return (c0 + c1 + c2 + c3 + c4) % p * sep % p


def fixed_base_constraints(sep, a_val, b_val, c_val, d_val, a_next, b_next, d_next, q_l, q_r, q_c, A, D, p):
    """FixedBaseScalarMul::constraints (widget/ecc/fixed_base_scalar_mul.rs:88-156)"""
    kappa = sep * sep % p
    kappa_sq = kappa * kappa % p
    kappa_cu = kappa_sq * kappa % p
    acc_x, acc_y, xy_alpha, acc_bit = a_val, b_val, c_val, d_val
    bit = (d_next - 2 * acc_bit) % p
    bit_consistency = bit * (bit - 1) % p * (bit + 1) % p
    y_alpha = (bit * bit % p * (q_r - 1) + 1) % p
    x_alpha = q_l * bit % p
    xy_consistency = (bit * q_c - xy_alpha) % p * kappa % p
    t = xy_alpha * acc_x % p * acc_y % p * D % p
    x_acc = ((a_next + a_next * t) - (x_alpha * acc_y + y_alpha * acc_x)) % p * kappa_sq % p
    y_acc = ((b_next - b_next * t) - (y_alpha * acc_y - A * x_alpha % p * acc_x)) % p * kappa_cu % p
    return (bit_consistency + x_acc + y_acc + xy_consistency) % p * sep % p


def curve_add_constraints(sep, a_val, b_val, c_val, d_val, a_next, b_next, d_next, A, D, p):
    """CurveAddition::constraints (widget/ecc/curve_addition.rs:62-96)"""
    x1, y1, x2, y2, x3, y3, x1_y2 = a_val, b_val, c_val, d_val, a_next, b_next, d_next
    kappa = sep * sep % p
    xy_consistency = (x1 * y2 - x1_y2) % p
    y1_x2, y1_y2, x1_x2 = y1 * x2 % p, y1 * y2 % p, x1 * x2 % p
    t = D * x1_y2 % p * y1_x2 % p
    x3_consistency = ((x1_y2 + y1_x2) - (x3 + x3 * t)) % p * kappa % p
    y3_consistency = ((y1_y2 - A * x1_x2) - (y3 - y3 * t)) % p * kappa % p * kappa % p
    return (xy_consistency + x3_consistency + y3_consistency) % p * sep % p


def custom_gate_sum(sel, seps, wires, nexts, q_l, q_r, q_c, A, D, p):
    """sum over the four custom gates of selector * constraints (quotient_poly.rs:231-264 point-wise;
    linearisation_poly.rs:382-410 with `sel` the selector polynomials' scalars).  `sel` = (range, logic,
    fixed, variable) selector values; `seps` the four separation challenges; returns the 4 products."""
    a, b, c, d = wires
    an, bn, dn = nexts
    return (
        sel[0] * range_constraints(seps[0], a, b, c, d, dn, p) % p if sel[0] is not None else 0,
        sel[1] * logic_constraints(seps[1], a, b, c, d, an, bn, dn, q_c, p) % p if sel[1] is not None else 0,
        sel[2] * fixed_base_constraints(seps[2], a, b, c, d, an, bn, dn, q_l, q_r, q_c, A, D, p) % p
        if sel[2] is not None else 0,
        sel[3] * curve_add_constraints(seps[3], a, b, c, d, an, bn, dn, A, D, p) % p if sel[3] is not None else 0,
    )
