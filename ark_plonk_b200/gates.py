"""Host-side scalar evaluation of the custom gate constraints at the opening point.

The 4n-point evaluation of these polynomials is the device quotient kernel's work
(csrc/poly.cu: gate_range / gate_logic / gate_fixed_base / gate_curve_add).  The linearisation
polynomial needs each of them once, at the evaluation challenge, as the scalar that multiplies the
gate's selector polynomial (plonk-core/src/proof_system/linearisation_poly.rs:382-410,
widget/mod.rs:135-150) -- a few dozen field multiplications on python ints.
"""
from __future__ import annotations

from .encoding import FR_MODULUS

# embedded twisted Edwards curves a x^2 + y^2 = 1 + d x^2 y^2 (TEModelParameters COEFF_A, COEFF_D of
# ark-ed-on-bls12-381 and ark-ed-on-bls12-377), indexed by curve id
EMBEDDED_A = (FR_MODULUS[0] - 1, FR_MODULUS[1] - 1)
EMBEDDED_D = (19257038036680949359750312669786877991949435402254120286184196891950884077233, 3021)


def _quad(f, p):
    return f * (f - 1) % p * (f - 2) % p * (f - 3) % p


def range_scalar(sep, w, nxt, p):
    """widget/range.rs:46-62"""
    a, b, c, d = w
    k = sep * sep % p
    acc = _quad((nxt[2] - 4 * a) % p, p)
    acc = (acc * k + _quad((a - 4 * b) % p, p)) % p
    acc = (acc * k + _quad((b - 4 * c) % p, p)) % p
    acc = (acc * k + _quad((c - 4 * d) % p, p)) % p
    return acc * sep % p


def logic_scalar(sep, w, nxt, q_c, p):
    """widget/logic.rs:66-141"""
    k = sep * sep % p
    a = (nxt[0] - 4 * w[0]) % p
    b = (nxt[1] - 4 * w[1]) % p
    d = (nxt[2] - 4 * w[3]) % p
    wv = w[2]
    s = (a + b) % p
    f = wv * ((wv * ((4 * wv - 18 * s + 81) % p) + 18 * (a * a + b * b) - 81 * s + 83) % p) % p
    xor_and = (q_c * (9 * d - 3 * s) + 3 * (s + d) - 2 * f) % p
    acc = xor_and
    acc = (acc * k + (wv - a * b)) % p
    acc = (acc * k + _quad(d, p)) % p
    acc = (acc * k + _quad(b, p)) % p
    acc = (acc * k + _quad(a, p)) % p
    return acc * sep % p


def fixed_base_scalar(sep, w, nxt, q_l, q_r, q_c, curve, p):
    """widget/ecc/fixed_base_scalar_mul.rs:88-156"""
    A, D = EMBEDDED_A[curve], EMBEDDED_D[curve]
    k = sep * sep % p
    acc_x, acc_y, xy_alpha, acc_bit = w
    bit = (nxt[2] - 2 * acc_bit) % p
    y_alpha = (bit * bit % p * (q_r - 1) + 1) % p
    x_alpha = q_l * bit % p
    t = xy_alpha * acc_x % p * acc_y % p * D % p
    bit_ok = bit * (bit - 1) % p * (bit + 1) % p
    xy_ok = (bit * q_c - xy_alpha) % p
    x_ok = (nxt[0] * (1 + t) - x_alpha * acc_y - y_alpha * acc_x) % p
    y_ok = (nxt[1] * (1 - t) - y_alpha * acc_y + A * x_alpha % p * acc_x) % p
    return (((y_ok * k + x_ok) % p * k + xy_ok) % p * k + bit_ok) % p * sep % p


def curve_add_scalar(sep, w, nxt, curve, p):
    """widget/ecc/curve_addition.rs:62-96"""
    A, D = EMBEDDED_A[curve], EMBEDDED_D[curve]
    k = sep * sep % p
    x1, y1, x2, y2 = w
    x3, y3, x1y2 = nxt
    y1x2 = y1 * x2 % p
    t = D * x1y2 % p * y1x2 % p
    xy_ok = (x1 * y2 - x1y2) % p
    x_ok = (x1y2 + y1x2 - x3 * (1 + t)) % p
    y_ok = (y1 * y2 - A * x1 % p * x2 - y3 * (1 - t)) % p
    return ((y_ok * k + x_ok) % p * k + xy_ok) % p * sep % p
