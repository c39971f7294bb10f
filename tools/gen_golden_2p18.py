#!/usr/bin/env python3
"""Generates tests/golden/plonk_bench_2p18.json: the ORACLE proof (oracle/plonk.py, Python big
integers, known-tau commitments) of the exact instance bench.py measures by default -
BenchCircuit at 2^18 gates, BLS12-381, tau and the 8 blinders of bench.py's `prove` workload
(benches/plonk.rs:95-98 runs degrees 5..18; 18 is the headline).  Takes tens of minutes of
pure-Python NTTs once, offline; the -m gpu parity test and bench.py compare the CUDA proof's
bytes with it.  Usage: python tools/gen_golden_2p18.py [log_n]
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import plonk as op  # noqa: E402
from oracle.curves import CURVES  # noqa: E402

BENCH_TAU = 0x1234567890ABCDEF1234567890ABCDEF          # bench.py: run_b200, workload "prove", rank 0
BENCH_BLINDERS = [1000 + i for i in range(8)]


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
    cv = CURVES[0]
    t0 = time.time()
    cs = op.bench_circuit(cv, log_n, BENCH_BLINDERS)
    kz = op.Kzg(cv, BENCH_TAU, 1 << log_n)
    pk = op.preprocess(cs, kz)
    print("preprocess %.0f s" % (time.time() - t0), flush=True)
    cs2 = op.bench_circuit(cv, log_n, BENCH_BLINDERS)
    _, blob = op.prove(cs2, pk, kz, b"ark")
    print("prove %.0f s" % (time.time() - t0), flush=True)
    out = {"curve": 0, "degree": log_n, "tau": hex(BENCH_TAU), "blinders": [hex(b) for b in BENCH_BLINDERS],
           "transcript_label": "ark", "proof_sha256": hashlib.sha256(blob).hexdigest(), "proof": blob.hex(),
           "generator": "tools/gen_golden_2p18.py (oracle/plonk.py)"}
    name = "plonk_bench_2p%d.json" % log_n
    with open(os.path.join(ROOT, "tests", "golden", name), "w") as fh:
        json.dump(out, fh, indent=1)
    print(name, out["proof_sha256"], flush=True)


if __name__ == "__main__":
    main()
