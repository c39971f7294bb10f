"""GPU/emulated prover against the golden proofs (byte-identical) - shared by CPU-emu and GPU tests."""
import hashlib
import json
import os

from ark_plonk_b200 import bench_circuit as bc
from ark_plonk_b200 import kzg
from ark_plonk_b200 import plonk as gp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_proofs.json")))


def golden_case(curve, degree):
    for g in GOLDEN:
        if g["curve"] == curve and g["degree"] == degree:
            return g
    raise KeyError((curve, degree))


def prove_case(lib, case, faithful=True, repeat=1):
    curve, degree = case["curve"], case["degree"]
    tau = int(case["tau"], 16)
    bl = [int(b, 16) for b in case["blinders"]]
    circ = bc.build(curve, degree, bl)
    ck = kzg.CommitterKey.from_tau(curve, tau, circ.n + 1, lib=lib)
    pr = gp.Prover(curve, ck, lib=lib)
    pk = pr.preprocess(circ, commit_verifier_key=False)
    wires = gp.wires_to_mont(circ)
    blob = None
    for _ in range(repeat):
        m0, n0 = pr.msm_calls, pr.ntt_calls
        blob = pr.prove(pk, wires, b"ark", faithful=faithful)
        assert pr.msm_calls - m0 == (29 if faithful else 15)            # 29 MSMs per proof
        assert hashlib.sha256(blob).hexdigest() == case["proof_sha256"], "proof differs from the oracle's"
    assert blob.hex() == case["proof"]
    pk.arena.close()
    ck.close()
    return blob
