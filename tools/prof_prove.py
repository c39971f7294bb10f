#!/usr/bin/env python3
"""Short prove workload for ncu: key + prover key at 2^LOG, then N proofs.  usage: prof_prove.py [log=18] [proofs=2]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import bench_circuit as bc, kzg, plonk as gp
from ark_plonk_b200._lib import get_lib
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
proofs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = get_lib(); lib.init(0)
circ = bc.build(0, log_n, [1000 + i for i in range(8)])
ck = kzg.CommitterKey.from_tau(0, 0x1234567890ABCDEF1234567890ABCDEF, circ.n + 1)
pr = gp.Prover(0, ck)
pk = pr.preprocess(circ, commit_verifier_key=False)
w = gp.wires_to_mont(circ)
off = pr.upload_wires(pk, w)
for _ in range(proofs):
    blob = pr.prove(pk, None, b"ark", wires_resident=off)
print("ok", len(blob))
