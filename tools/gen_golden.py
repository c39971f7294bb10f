#!/usr/bin/env python3
"""Generates tests/golden/plonk_proofs.json with the ORACLE prover (oracle/plonk.py).

The reference holds no byte-level vectors and cannot be run here (no Rust),
so these are oracle-made known-answer vectors: seeded tau + 8 blinders -> serialized Proof.
The oracle prover is itself pinned by the verifier identity r(z) + r0 == 0
(proof.rs:428-486, tests/test_oracle_plonk.py).  Usage: python tools/gen_golden.py
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import plonk as op  # noqa: E402
from oracle.curves import CURVES  # noqa: E402

CASES = [(0, 5), (0, 8), (0, 10), (1, 6), (1, 10), (0, 12), (0, 14), (0, 16)]


def seeded(curve, degree):
    cv = CURVES[curve]
    rnd = random.Random("apb-golden-%d-%d" % (curve, degree))
    return rnd.randrange(1, cv.fr.p), [rnd.randrange(cv.fr.p) for _ in range(8)]


def main():
    out = []
    for curve, degree in CASES:
        cv = CURVES[curve]
        tau, bl = seeded(curve, degree)
        cs = op.bench_circuit(cv, degree, bl)
        kz = op.Kzg(cv, tau, 1 << degree)
        pk = op.preprocess(cs, kz)
        cs2 = op.bench_circuit(cv, degree, bl)
        _, blob = op.prove(cs2, pk, kz, b"ark")
        from oracle.serialize import ser_g1
        vk = {k: ser_g1(cv, v).hex() for k, v in pk.commitments.items()}
        out.append({"curve": curve, "degree": degree, "tau": hex(tau), "blinders": [hex(b) for b in bl],
                    "proof_sha256": hashlib.sha256(blob).hexdigest(), "proof": blob.hex(), "vk": vk})
        print(curve, degree, out[-1]["proof_sha256"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "plonk_proofs.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
