"""Wire-format round trips of the product's ark-serialize module (ark_plonk_b200/serialize.py), mirroring the
reference's own serde tests: proof_system/widget/mod.rs:438-572 (ProverKey `serialize_unchecked` round trip,
VerifierKey round trip), proof.rs:686-707 (Proof), circuit.rs:392-463 (VerifierData inside test_full).
Shared by the CPU-emulation and the GPU tests; the independent check of the byte layout is the ORACLE's
serializer (oracle/serialize.py, oracle/plonk.py:serialize_proof)."""
import random

import numpy as np

from ark_plonk_b200 import bench_circuit as bc
from ark_plonk_b200 import encoding as enc
from ark_plonk_b200 import kzg
from ark_plonk_b200 import plonk as gp
from ark_plonk_b200 import serialize as ser
from oracle.curves import CURVES, powers_of_tau_g1
from oracle.serialize import deser_g1, ser_g1

import gadget_cases
import prover_cases


def check_points_and_scalars(curve_id):
    cv = CURVES[curve_id]
    rnd = random.Random(41 + curve_id)
    pts = [None, cv.G] + [cv.mul(cv.G, rnd.randrange(1, cv.fr.p)) for _ in range(6)]
    for P in pts:
        b = ser.write_g1(curve_id, P)
        assert b == ser_g1(cv, P)                                   # same bytes as the oracle's encoder
        assert ser.read_g1(curve_id, ser.Reader(b)) == P == deser_g1(cv, b)
        u = ser.write_g1(curve_id, P, compressed=False)
        assert len(u) == 96 and ser.read_g1(curve_id, ser.Reader(u), compressed=False) == P
    for bad in (b"\xff" * 48, bytes(47) + b"\xc0"):
        try:
            ser.read_g1(curve_id, ser.Reader(bad))
        except ser.SerializationError:
            pass
        else:
            raise AssertionError("malformed point accepted")
    v = rnd.randrange(cv.fr.p)
    assert ser.read_fr(curve_id, ser.Reader(ser.write_fr(curve_id, v))) == v
    try:
        ser.read_fr(curve_id, ser.Reader(b"\xff" * 32))
    except ser.SerializationError:
        pass
    else:
        raise AssertionError("unreduced scalar accepted")


def check_proof_roundtrip(case):
    """golden proof bytes (made by the oracle's serializer) -> parse -> write == the same bytes"""
    blob = bytes.fromhex(case["proof"])
    proof = ser.parse_proof(case["curve"], blob)
    assert ser.write_proof(case["curve"], proof) == blob
    assert [l for l, _ in proof["custom_evals"]] == ["q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "a_next_eval",
                                                      "b_next_eval", "d_next_eval"]
    assert proof["aw_opening_random_v"] is None and len(blob) == 13 * 48 + 2 * 49 + 16 * 32 + 8 + sum(
        8 + len(l) + 32 for l, _ in proof["custom_evals"])
    try:
        ser.parse_proof(case["curve"], blob + b"\x00")
    except ser.SerializationError:
        pass
    else:
        raise AssertionError("trailing byte accepted")


def check_srs_roundtrip(lib, curve_id, n=40):
    """powers_of_g of a UniversalParams / CommitterKey: bytes -> resident key -> same commitments"""
    cv = CURVES[curve_id]
    tau = 0x1234567 + curve_id
    pts = powers_of_tau_g1(cv, tau, n)
    rec = enc.g1_affine_to_mont(curve_id, pts)
    for compressed in (True, False):
        blob = ser.write_powers_of_g(curve_id, rec, compressed, lib=lib)
        assert blob[:8] == n.to_bytes(8, "little") and len(blob) == 8 + n * (48 if compressed else 96)
        if compressed:
            assert blob[8:] == b"".join(ser_g1(cv, P) for P in pts)
        tail = b"\x01\x02\x03"                              # the G2 part etc. of the real structs follows: ignored
        got = ser.read_powers_of_g(curve_id, blob + tail, compressed, lib=lib)
        assert np.array_equal(got, rec)
        assert np.array_equal(ser.read_powers_of_g(curve_id, blob, compressed, max_points=7, lib=lib), rec[:7])
    ck = kzg.CommitterKey(curve_id, got, lib=lib)
    ck2 = kzg.CommitterKey.from_tau(curve_id, tau, n, lib=lib)
    assert np.array_equal(ck.download(0, n), ck2.download(0, n))
    ck.close()
    ck2.close()


def check_keys_roundtrip(lib, curve_id=0, degree=5, kind=None):
    """compile -> VerifierKey / VerifierData / ProverKey bytes -> load -> the loaded key proves to the SAME proof
    (widget/mod.rs:438-504: assert_eq!(prover_key, obtained_pk); circuit.rs gen_proof takes the deserialized key)"""
    cv = CURVES[curve_id]
    if kind is None:
        case = prover_cases.golden_case(curve_id, degree)
        tau = int(case["tau"], 16)
        circ = bc.build(curve_id, degree, [int(b, 16) for b in case["blinders"]])
        label, pis = b"ark", {}
    else:
        cs = gadget_cases.build_composer(curve_id, kind)
        circ = gadget_cases.arrays_from_composer(cs, curve_id)
        tau, label, pis = 0xABCDEF123, b"serde", dict(cs.public_inputs)
    n = circ.n
    ck = kzg.CommitterKey.from_tau(curve_id, tau, n + 1, lib=lib)
    pr = gp.Prover(curve_id, ck, lib=lib)
    pk = pr.preprocess(circ, commit_verifier_key=True)
    want = pr.prove(pk, gp.wires_to_mont(circ), label)
    if kind is None and case.get("proof"):
        assert want.hex() == case["proof"]
    # VerifierKey: 8 + 20 * 48 bytes, identity commitments for the all-zero custom selectors
    comms = dict(pk.commitments)
    for s in ser.VK_ORDER:
        comms.setdefault(s, ser.write_g1(curve_id, None))
    vk = ser.VerifierKey.from_compressed_commitments(curve_id, n, comms)
    vb = vk.to_bytes()
    assert len(vb) == 8 + 20 * 48 and vb[:8] == n.to_bytes(8, "little")
    assert vb[8:] == b"".join(comms[s] for s in ser.VK_ORDER)
    assert ser.VerifierKey.from_bytes(curve_id, vb) == vk
    assert ser.VerifierKey.from_bytes(curve_id, vk.to_bytes(compressed=False), compressed=False) == vk
    vd = ser.VerifierData(vk, pis)
    assert ser.VerifierData.from_bytes(curve_id, vd.to_bytes()) == ser.VerifierData(vk, {k: v % cv.fr.p for k, v in pis.items() if v % cv.fr.p})
    # ProverKey: serialize_unchecked -> deserialize -> byte-identical re-serialization and the same proof
    blob = ser.prover_key_to_bytes(pk)
    assert blob[:8] == n.to_bytes(8, "little")
    pk2 = pr.load_prover_key(blob)
    assert pk2.custom == pk.custom and pk2.n == n
    assert ser.prover_key_to_bytes(pk2) == blob
    got = pr.prove(pk2, gp.wires_to_mont(circ), label, public_inputs=pis)
    assert got == want, "the deserialized prover key proves to different bytes"
    bad = bytearray(blob)
    bad[8 + 8 + 32 + 5] ^= 0xFF                            # q_m's SECOND coefficient (the constant one drops out of the opening witness)
    try:
        pk3 = pr.load_prover_key(bytes(bad))
        diff = pr.prove(pk3, gp.wires_to_mont(circ), label, public_inputs=pis)
        pk3.arena.close()
        assert diff != want
    except ser.SerializationError:
        pass
    try:
        pr.load_prover_key(blob[:-1])
    except ser.SerializationError:
        pass
    else:
        raise AssertionError("truncated key accepted")
    pk.arena.close()
    pk2.arena.close()
    ck.close()
