"""Oracle-side checks of the custom gates: embedded curve constants, gadget semantics, and
prove -> verify of gadget circuits (accept valid, reject invalid), as the reference's gadget tests do
(constraint_system/range.rs:196-260, logic.rs:351-420, ecc/scalar_mul/fixed_base.rs:186-250)."""
import random

import pytest

import gadget_cases
from ark_plonk_b200 import gates as product_gates
from oracle import gates
from oracle import plonk as op
from oracle import plonk_verify as pv
from oracle.curves import CURVES


@pytest.mark.parametrize("curve_id", [0, 1])
def test_embedded_curve_constants(curve_id):
    """a = -1, d pinned by the group order: cofactor * r * P = O for a point found on the curve"""
    curve = CURVES[curve_id]
    a, d, r, cof = gates.EMBEDDED[curve.name]
    p = curve.fr.p
    assert (a, d) == (product_gates.EMBEDDED_A[curve_id], product_gates.EMBEDDED_D[curve_id])
    if curve_id == 0:
        assert d == (-10240 * pow(10241, -1, p)) % p            # JubJub
    P = gates.te_point_from_x(curve, 2)
    assert gates.te_on_curve(P, a, d, p)
    assert gates.te_mul(P, r, a, d, p) == (0, 1) and gates.te_mul(P, 5, a, d, p) != (0, 1)


def test_wnaf2():
    rng = random.Random(3)
    for _ in range(50):
        k = rng.getrandbits(200)
        digits = gates.find_wnaf2(k)
        assert set(digits) <= {-1, 0, 1}
        assert sum(d << i for i, d in enumerate(digits)) == k
        assert all(not (digits[i] and digits[i + 1]) for i in range(len(digits) - 1))


@pytest.mark.parametrize("curve_id", [0, 1])
def test_product_gate_scalars_match_oracle(curve_id):
    curve = CURVES[curve_id]
    p = curve.fr.p
    A, D = gates.embedded_params(curve)
    rng = random.Random(1)
    for _ in range(20):
        sep, a, b, c, d, an, bn, dn, ql, qr, qc = [rng.randrange(p) for _ in range(11)]
        w, nx = (a, b, c, d), (an, bn, dn)
        assert product_gates.range_scalar(sep, w, nx, p) == gates.range_constraints(sep, a, b, c, d, dn, p)
        assert product_gates.logic_scalar(sep, w, nx, qc, p) == gates.logic_constraints(sep, a, b, c, d, an, bn, dn, qc, p)
        assert product_gates.fixed_base_scalar(sep, w, nx, ql, qr, qc, curve_id, p) == \
            gates.fixed_base_constraints(sep, a, b, c, d, an, bn, dn, ql, qr, qc, A, D, p)
        assert product_gates.curve_add_scalar(sep, w, nx, curve_id, p) == \
            gates.curve_add_constraints(sep, a, b, c, d, an, bn, dn, A, D, p)


def _prove_verify(cs):
    curve = cs.curve
    n = cs.circuit_bound()
    tau = random.Random(9).randrange(curve.fr.p)
    kzg = op.Kzg(curve, tau, n + 8)
    pk = op.preprocess(cs, kzg)
    _, blob = op.prove(cs, pk, kzg, b"t")
    return pv.verify(curve, pk.commitments, n, blob, tau, b"t", public_inputs=cs.public_inputs)


@pytest.mark.parametrize("curve_id", [0, 1])
@pytest.mark.parametrize("kind", gadget_cases.KINDS)
def test_gadget_circuits_verify(curve_id, kind):
    assert _prove_verify(gadget_cases.build_composer(curve_id, kind))


@pytest.mark.parametrize("curve_id", [0, 1])
def test_bad_gadget_circuits_are_rejected(curve_id):
    curve = CURVES[curve_id]
    bl = list(range(1, 9))
    cs = op.Composer(curve, bl)                                   # range.rs:205: not a 32-bit number
    cs.add_dummy_lookup_table()
    cs.range_gate(cs.add_input(2 ** 34 + 5), 32)
    assert not _prove_verify(cs)
    cs = op.Composer(curve, bl)                                   # logic.rs:385: xor result constrained to the and
    cs.add_dummy_lookup_table()
    r = cs.xor_gate(cs.add_input(139), cs.add_input(33), 10)
    cs.constrain_to_constant(r, 139 & 33)
    assert not _prove_verify(cs)
    cs = op.Composer(curve, bl)                                   # wrong public point
    cs.add_dummy_lookup_table()
    G = gates.te_point_from_x(curve, 2)
    x, y = cs.fixed_base_scalar_mul(cs.add_input(12345), G)
    cs.constrain_to_constant(x, 0, pi=-G[0])
    assert not _prove_verify(cs)
    cs = op.Composer(curve, bl)                                   # wrong public input value
    cs.add_dummy_lookup_table()
    cs.constrain_to_constant(cs.add_input(5), 0, pi=-5)
    assert _prove_verify(cs)
    cs.public_inputs = {k: v + 1 for k, v in cs.public_inputs.items()}
    assert not _prove_verify(cs)
