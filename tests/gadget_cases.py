"""Circuits built from the reference's gadgets (range, logic, curve addition, fixed-base scalar
multiplication, public inputs), proved by the oracle and by the CUDA prover -- shared by the CPU-emulation
and GPU tests.  The oracle Composer restates constraint_system/{range,logic,arithmetic}.rs and ecc/*;
the circuits follow the reference's own gadget tests (range.rs:196-260, logic.rs:351-420,
ecc/curve_addition/variable_base_gate.rs:140-250, ecc/scalar_mul/fixed_base.rs:186-250)."""
import random

import numpy as np

from ark_plonk_b200 import bench_circuit as bc
from ark_plonk_b200 import kzg
from ark_plonk_b200 import plonk as gp
from oracle import gates
from oracle import plonk as op
from oracle import plonk_verify as pv
from oracle.curves import CURVES

KINDS = ("range", "logic", "curve_add", "fixed_base", "mixed")


def build_composer(curve_id: int, kind: str, seed: int = 11):
    curve = CURVES[curve_id]
    p = curve.fr.p
    rng = random.Random(seed * 31 + curve_id)
    cs = op.Composer(curve, [rng.randrange(p) for _ in range(8)])
    cs.add_dummy_lookup_table()
    A, D = gates.embedded_params(curve)
    if kind in ("range", "mixed"):
        cs.range_gate(cs.add_input(0xDEADBEEF), 32)
        cs.range_gate(cs.add_input(513), 10)
    if kind in ("logic", "mixed"):
        r = cs.xor_gate(cs.add_input(500), cs.add_input(357), 10)
        cs.constrain_to_constant(r, 500 ^ 357)
        r = cs.and_gate(cs.add_input(469), cs.add_input(321), 10)
        cs.constrain_to_constant(r, 0, pi=-(469 & 321))
    if kind in ("curve_add", "mixed"):
        P, Q = gates.te_point_from_x(curve, 2), gates.te_point_from_x(curve, 9)
        x3, y3 = cs.point_addition_gate((cs.add_input(P[0]), cs.add_input(P[1])), (cs.add_input(Q[0]), cs.add_input(Q[1])))
        S = gates.te_add(P, Q, A, D, p)
        cs.constrain_to_constant(x3, 0, pi=-S[0])           # assert_equal_public_point (ecc/mod.rs)
        cs.constrain_to_constant(y3, 0, pi=-S[1])
    if kind in ("fixed_base", "mixed"):
        G = gates.te_point_from_x(curve, 2)
        k = rng.randrange(p)
        x, y = cs.fixed_base_scalar_mul(cs.add_input(k), G)
        S = gates.te_mul(G, k, A, D, p)
        cs.constrain_to_constant(x, 0, pi=-S[0])
        cs.constrain_to_constant(y, 0, pi=-S[1])
    if kind == "mixed":
        cs.add_dummy_constraints()
    return cs


def arrays_from_composer(cs, curve_id: int) -> bc.CircuitArrays:
    """plain-array view of a composer (what the CUDA prover's front end consumes)"""
    n = cs.circuit_bound()
    rows = cs.n
    values, index = [], {}

    def vid(v):
        if v not in index:
            index[v] = len(values)
            values.append(v)
        return index[v]

    zero = vid(0)
    sel = {s: np.full(n, zero, dtype=np.int64) for s in cs.SELECTORS}
    for s in cs.SELECTORS:
        col = getattr(cs, s)
        sel[s][:len(col)] = [vid(v) for v in col]
    wires = np.full((4, n), zero, dtype=np.int64)
    for c in range(4):
        assert len(cs.w[c]) == rows
        wires[c, :rows] = [vid(cs.variables[v]) for v in cs.w[c]]
    sigma = np.empty((4, n, 2), dtype=np.int64)
    sigma[:, :, 0] = np.arange(4)[:, None]
    sigma[:, :, 1] = np.arange(n)[None, :]
    for wl in cs.variable_map:
        for k, (col, row) in enumerate(wl):
            sigma[col, row] = wl[(k + 1) % len(wl)]
    return bc.CircuitArrays(curve=curve_id, n=n, rows=rows, selectors=sel, wires=wires, values=values, sigma=sigma,
                            table=[list(r) for r in cs.lookup_table], public_inputs=dict(cs.public_inputs))


def prove_gadget_case(lib, curve_id: int, kind: str, tamper: bool = False):
    """oracle proof == CUDA proof byte for byte, and the restated verifier accepts it"""
    curve = CURVES[curve_id]
    cs = build_composer(curve_id, kind)
    circ = arrays_from_composer(cs, curve_id)            # before the oracle pads the composer
    n = cs.circuit_bound()
    tau = random.Random(77).randrange(curve.fr.p)
    okzg = op.Kzg(curve, tau, n + 8)
    opk = op.preprocess(cs, okzg)
    _, want = op.prove(cs, opk, okzg, b"gadgets")
    assert pv.verify(curve, opk.commitments, n, want, tau, b"gadgets", public_inputs=cs.public_inputs)

    ck = kzg.CommitterKey.from_tau(curve_id, tau, n + 1, lib=lib)
    pr = gp.Prover(curve_id, ck, lib=lib)
    pk = pr.preprocess(circ, commit_verifier_key=True)
    for name, comp in pk.commitments.items():             # verifier key: same commitments as the oracle's
        assert comp == op.ser_g1(curve, opk.commitments[name]), name
    wires = gp.wires_to_mont(circ)
    if tamper:                                            # break one witness: the proof must be rejected
        wires = wires.copy()
        row = next(i for i, v in enumerate(cs.q_range) if v)
        assert not np.array_equal(wires[0, row], wires[0, 1])
        wires[0, row] = wires[0, 1]                       # a range accumulator replaced by a blinding value
    got = pr.prove(pk, wires, b"gadgets")
    pk.arena.close()
    ck.close()
    if tamper:
        assert not pv.verify(curve, opk.commitments, n, got, tau, b"gadgets", public_inputs=cs.public_inputs)
        return got
    assert got == want, "CUDA proof differs from the oracle's"
    return got


def _pi_composer(curve_id: int, a: int, b: int, seed: int = 5):
    """one structure, witness-dependent public inputs: and_gate(a, b) with PI = -(a & b), plus a range gate"""
    curve = CURVES[curve_id]
    rng = random.Random(seed)
    cs = op.Composer(curve, [rng.randrange(curve.fr.p) for _ in range(8)])
    cs.add_dummy_lookup_table()
    r = cs.and_gate(cs.add_input(a), cs.add_input(b), 10)
    cs.constrain_to_constant(r, 0, pi=-(a & b))
    cs.range_gate(cs.add_input(a), 10)
    return cs


def prove_two_public_input_assignments(lib, curve_id: int = 0):
    """ONE compiled key, TWO witnesses with different public inputs (the reference keeps PI out of the
    ProverKey and reads them per proof: prover.rs:182,392; circuit.rs gen_proof): both proofs equal the
    oracle's byte for byte and verify against their own public inputs, not against the other's"""
    curve = CURVES[curve_id]
    tau = random.Random(78).randrange(curve.fr.p)
    cases = [(469, 321), (1000, 731)]
    comps = [_pi_composer(curve_id, a, b) for a, b in cases]
    circs = [arrays_from_composer(cs, curve_id) for cs in comps]
    n = comps[0].circuit_bound()
    assert comps[1].circuit_bound() == n and comps[0].public_inputs != comps[1].public_inputs
    ck = kzg.CommitterKey.from_tau(curve_id, tau, n + 1, lib=lib)
    pr = gp.Prover(curve_id, ck, lib=lib)
    pk = pr.preprocess(circs[0], commit_verifier_key=True)                 # compiled once, from the first witness
    try:
        for cs, circ in zip(comps, circs):
            pis = dict(cs.public_inputs)
            okzg = op.Kzg(curve, tau, n + 8)
            opk = op.preprocess(cs, okzg)
            for name, comp in pk.commitments.items():                      # structure only: same verifier key
                assert comp == op.ser_g1(curve, opk.commitments[name]), name
            _, want = op.prove(cs, opk, okzg, b"pi")
            got = pr.prove(pk, gp.wires_to_mont(circ), b"pi", public_inputs=pis)
            assert got == want
            assert pv.verify(curve, opk.commitments, n, got, tau, b"pi", public_inputs=pis)
            other = comps[1].public_inputs if cs is comps[0] else comps[0].public_inputs
            assert not pv.verify(curve, opk.commitments, n, got, tau, b"pi", public_inputs=other)
    finally:
        pk.arena.close()
        ck.close()
