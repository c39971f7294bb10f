#!/usr/bin/env python3
"""First-light measurements on a B200: INT32 multiply-add peak, NTT and MSM sweeps.

Writes gpurun_out/sweep.json.  Timing: CUDA events on the library's stream (device-resident
inputs), >= 3 warm-ups, median of the timed repetitions.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import encoding as enc, kzg, synth  # noqa: E402
from ark_plonk_b200._lib import get_lib  # noqa: E402
from ark_plonk_b200.domain import Radix2EvaluationDomain  # noqa: E402

out = {}
lib = get_lib()
lib.init(0)
stream = torch.cuda.ExternalStream(lib.c.apb_stream())


def timed(fn, warm=3, reps=7):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            fn()
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


wide, n32 = lib.imad_peak()
out["imad_wide_per_s"] = wide
out["imad32_per_s"] = n32
print("IMAD.WIDE/s %.3e   IMAD32/s %.3e" % (wide, n32), flush=True)

what = sys.argv[1:] or ["ntt", "msm"]
if "ntt" in what:
    out["ntt"] = []
    for log_n in (16, 18, 20, 22, 24, 26):
        n = 1 << log_n
        d = Radix2EvaluationDomain(0, n)
        x = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda")
        y = torch.empty_like(x)
        torch.cuda.synchronize()
        for kind, name in ((0, "fft"), (2, "coset_fft"), (3, "coset_ifft")):
            med, best = timed(lambda: d.ntt_dev(kind, x.data_ptr(), n, y.data_ptr()))
            gbs = 2 * n * 32 / (med * 1e-3) / 1e9
            imad = (n / 2 * log_n * 136) / (med * 1e-3)
            rec = dict(log_n=log_n, kind=name, ms=med, best_ms=best, gbs=gbs, imad_per_s=imad)
            out["ntt"].append(rec)
            print(rec, flush=True)
        if log_n == 20:     # quarter-filled input (prover's 4n coset fft of n coefficients)
            med, best = timed(lambda: d.ntt_dev(2, x.data_ptr(), n // 4, y.data_ptr()))
            print(dict(log_n=log_n, kind="coset_fft_quarter", ms=med), flush=True)
            out["ntt"].append(dict(log_n=log_n, kind="coset_fft_quarter", ms=med, best_ms=best))
        d.close()
        del x, y
if "msm" in what:
    out["msm"] = []
    lib.set_profiling(True)
    for log_n in (10, 14, 16, 18, 20):
        n = 1 << log_n
        t0 = time.time()
        pts = synth.progression_bases(0, 12345, 67891, n)
        tb = time.time() - t0
        t0 = time.time()
        ck = kzg.CommitterKey(0, enc.g1_affine_to_mont(0, pts))
        tu = time.time() - t0
        S = synth.seeded_scalars(0, n)
        dS = torch.from_numpy(S.view(np.int64)).cuda()
        outxyz = np.zeros(18, dtype=np.uint64)
        ts, phases = [], []
        for i in range(8):
            lib.check(lib.c.apb_msm_dev(ck._h, 0, dS.data_ptr(), n, 0, outxyz.ctypes.data))
            if i >= 3:
                ts.append(lib.last_device_ms())
                phases.append(lib.msm_phase_ms())
        med = float(np.median(ts))
        ph = {k: float(np.median([p[k] for p in phases])) for k in phases[0]}
        t0 = time.time()
        kzg.multi_scalar_mul(ck, S)
        e2e = (time.time() - t0) * 1e3
        rec = dict(log_n=log_n, ms=med, mpts=n / med / 1e3, phases=ph, e2e_ms=e2e, upload_s=tu, synth_s=tb,
                   imad_frac=(n * 48000 / (med * 1e-3)) / wide if wide else None)
        out["msm"].append(rec)
        print(rec, flush=True)
        ck.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
