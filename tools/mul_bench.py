#!/usr/bin/env python3
"""Montgomery-product throughput vs occupancy / ILP (dependent chains), as a fraction of the
measured IMAD.WIDE peak.  Writes gpurun_out/mul_bench.json."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200._lib import get_lib
lib = get_lib(); lib.init(0)
wide, n32 = lib.imad_peak()
res = {"imad_wide_per_s": wide, "rows": []}
for field, name, imads in ((4, "Fq381_28", 300), (1, "Fq381", 300)):
    for threads, bps in ((128, 1), (128, 2), (128, 3), (128, 4), (128, 8), (256, 8)):
        for ilp in (1, 2):
            m = lib.mul_bench(field, threads, bps, ilp, 1500)
            row = dict(field=name, threads_per_sm=threads * bps, ilp=ilp, muls_per_s=m, imad_frac=m * imads / wide)
            res["rows"].append(row)
            print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mul_bench.json"), "w"), indent=1)
