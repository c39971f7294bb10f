#!/usr/bin/env python3
"""Times the device-resident prover (BenchCircuit, BLS12-381 / KZG10) at 2^degree gates."""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import bench_circuit as bc, kzg, plonk as gp
from ark_plonk_b200._lib import get_lib

lib = get_lib(); lib.init(0)
res = []
for degree in [int(a) for a in sys.argv[1:]] or [16, 18]:
    tau = 0x1234567890ABCDEF1234567890ABCDEF
    bl = [1000 + i for i in range(8)]
    t0 = time.perf_counter(); circ = bc.build(0, degree, bl); t_build = time.perf_counter() - t0
    t0 = time.perf_counter(); ck = kzg.CommitterKey.from_tau(0, tau, circ.n + 1); t_srs = time.perf_counter() - t0
    pr = gp.Prover(0, ck)
    t0 = time.perf_counter(); pk = pr.preprocess(circ); t_pre = time.perf_counter() - t0
    wires = gp.wires_to_mont(circ)
    times = []
    for faithful in (True, False):
        ts = []
        for i in range(6):
            t0 = time.perf_counter()
            blob = pr.prove(pk, wires, b"ark", faithful=faithful)
            ts.append((time.perf_counter() - t0) * 1e3)
        times.append(ts)
    lib.set_profiling(True)
    pr.phase_log = []
    t0 = time.perf_counter(); pr.prove(pk, wires, b"ark", faithful=True); t_prof = (time.perf_counter() - t0) * 1e3
    for k, ms, ph in pr.phase_log:
        print("   commit k=%d total %.3f ms  %s" % (k, ms, {a: round(b, 3) for a, b in ph.items()}), flush=True)
    print("   sum of commit calls %.2f ms of %.2f ms" % (sum(m for _, m, _ in pr.phase_log), t_prof), flush=True)
    pr.phase_log = None
    lib.set_profiling(False)
    rec = dict(degree=degree, n=circ.n, rows=circ.rows, build_s=t_build, srs_s=t_srs, preprocess_s=t_pre,
               prove_ms_faithful=float(np.median(times[0][2:])), prove_ms_no_dead_commits=float(np.median(times[1][2:])),
               all=times, sha=hashlib.sha256(blob).hexdigest())
    print(json.dumps(rec), flush=True)
    res.append(rec)
    pk.arena.close(); ck.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "prove_bench.json"), "w"), indent=1)
