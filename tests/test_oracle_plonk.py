"""Pins the oracle prover (oracle/plonk.py): the proof it produces satisfies the identity the
reference verifier checks, r(z) + r0 == 0 with r0 computed by the verifier's own formula
(proof_system/proof.rs:428-486), the quotient is a polynomial of degree < 4n, the KZG openings
satisfy w(tau)(tau - z) = p(tau) - p(z), and the committed golden vectors are reproduced."""
import hashlib
import json
import os
import random

import pytest

from oracle import plonk as op
from oracle.curves import CURVES
from oracle.ntt import Domain, poly_eval

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_proofs.json")))


def verifier_r0(p, n, proof, T):
    """compute_r0 (proof.rs:428-486) with an empty public-input vector"""
    alpha, beta, gamma, delta, epsilon = T["alpha"], T["beta"], T["gamma"], T["delta"], T["epsilon"]
    zc, ls = T["z_challenge"], T["lookup_sep"]
    a, b, c, d = proof["wire_evals"]
    s1, s2, s3, zhat = proof["perm_evals"]
    _ql, z2n, _h1e, h1n, h2e, _fe, _te, _tn = proof["lookup_evals"]
    zh = (pow(zc, n, p) - 1) % p
    l1 = zh * pow(n * (zc - 1) % p, -1, p) % p
    bb = (a + beta * s1 + gamma) * (b + beta * s2 + gamma) * (c + beta * s3 + gamma) * ((d + gamma) * zhat * alpha) % p
    cc = l1 * alpha * alpha % p
    eopd = epsilon * (1 + delta) % p
    dd = ls * ls * z2n * (eopd + delta * h2e) * (eopd + h2e + delta * h1n) % p
    ee = ls * ls * ls * l1 % p
    return (0 - bb - cc - dd - ee) % p


@pytest.mark.parametrize("curve,degree", [(0, 5), (1, 5), (0, 7)])
def test_oracle_proof_satisfies_verifier_identity(curve, degree):
    cv = CURVES[curve]
    rnd = random.Random(degree * 3 + curve)
    bl = [rnd.randrange(cv.fr.p) for _ in range(8)]
    tau = rnd.randrange(1, cv.fr.p)
    kz = op.Kzg(cv, tau, 1 << degree)
    pk = op.preprocess(op.bench_circuit(cv, degree, bl), kz)
    cs = op.bench_circuit(cv, degree, [rnd.randrange(cv.fr.p) for _ in range(8)])    # gen_proof: fresh blinders
    T = {}
    proof, blob = op.prove(cs, pk, kz, b"ark", T)
    p, n = cv.fr.p, pk.n
    assert cs.n == (1 << (degree - 1)) + 2 and n == 1 << degree           # 2^(k-1)+2 real rows
    assert len(T["t_poly"]) <= 4 * n - 4                                  # the numerator is divisible by Z_H
    assert (poly_eval(cv.fr, T["lin_poly"], T["z_challenge"]) + verifier_r0(p, n, proof, T)) % p == 0
    # openings: w(tau) (tau - z) == p(tau) - p(z)
    zc = T["z_challenge"]
    zw = zc * Domain.for_size(cv.fr, n).group_gen % p
    P = pk.polys
    aw = [T["lin_poly"], P["left_sigma"], P["right_sigma"], P["out_sigma"], T["f_poly"], T["h2_poly"], T["table_poly"]] + T["w_polys"]
    saw = [T["z_poly"], T["w_polys"][0], T["w_polys"][1], T["w_polys"][3], T["h1_poly"], T["z2_poly"], T["table_poly"]]
    for polys, chal, point, wit, comm in ((aw, T["aw_challenge"], zc, T["aw_witness"], proof["aw_opening"]),
                                          (saw, T["saw_challenge"], zw, T["saw_witness"], proof["saw_opening"])):
        ct = cz = 0
        cur = 1
        for q in polys:
            ct = (ct + cur * poly_eval(cv.fr, q, tau)) % p
            cz = (cz + cur * poly_eval(cv.fr, q, point)) % p
            cur = cur * chal % p
        assert poly_eval(cv.fr, wit, tau) * (tau - point) % p == (ct - cz) % p
        assert comm == cv.mul(cv.G, poly_eval(cv.fr, wit, tau))
    assert len(blob) == 13 * 48 + 2 * 49 + 16 * 32 + 8 + sum(8 + len(k) + 32 for k, _ in proof["custom_evals"])


def test_combine_split_reference_vector():
    """lookup/multiset.rs:335-391 `test_combine_split`: the Plonkup paper example in the doc comment
    (multiset.rs:119-123): t = {2,4,1,3}, f = {2,3,3,2} -> h1 = {2,2,1,3}, h2 = {2,4,3,3}"""
    h1, h2 = op.combine_split([2, 4, 1, 3], [2, 3, 3, 2])
    assert h1 == [2, 2, 1, 3] and h2 == [2, 4, 3, 3]
    with pytest.raises(ValueError):
        op.combine_split([1, 2], [3])


@pytest.mark.parametrize("case", [g for g in GOLDEN if g["degree"] <= 8], ids=lambda g: "c%d-2^%d" % (g["curve"], g["degree"]))
def test_oracle_reproduces_golden(case):
    cv = CURVES[case["curve"]]
    tau = int(case["tau"], 16)
    bl = [int(b, 16) for b in case["blinders"]]
    kz = op.Kzg(cv, tau, 1 << case["degree"])
    pk = op.preprocess(op.bench_circuit(cv, case["degree"], bl), kz)
    _, blob = op.prove(op.bench_circuit(cv, case["degree"], bl), pk, kz, b"ark")
    assert hashlib.sha256(blob).hexdigest() == case["proof_sha256"] and blob.hex() == case["proof"]


def _vk_points(cv, case):
    from oracle.serialize import deser_g1
    return {k: deser_g1(cv, bytes.fromhex(v)) for k, v in case["vk"].items()}


@pytest.mark.parametrize("case", GOLDEN, ids=lambda g: "c%d-2^%d" % (g["curve"], g["degree"]))
def test_restated_verifier_accepts_golden_and_rejects_tampering(case):
    """proof.rs:111-426 restated (oracle/plonk_verify.py): every golden proof verifies; flipping one bit of an
    evaluation, a commitment or an opening makes it fail"""
    from oracle import plonk_verify as pv
    cv = CURVES[case["curve"]]
    tau = int(case["tau"], 16)
    n = 1 << case["degree"]
    vk = _vk_points(cv, case)
    blob = bytes.fromhex(case["proof"])
    assert pv.verify(cv, vk, n, blob, tau)
    if case["degree"] <= 8:
        for pos in (13 * 48 + 2 * 49 + 5, 13 * 48 + 2 * 49 + 16 * 32 + 8 + 8 + 12 + 3):
            bad = bytearray(blob)
            bad[pos] ^= 1
            assert not pv.verify(cv, vk, n, bytes(bad), tau)
        # a different (valid) commitment in place of a: take b's bytes
        bad = bytearray(blob)
        bad[0:48] = blob[48:96]
        assert not pv.verify(cv, vk, n, bytes(bad), tau)
