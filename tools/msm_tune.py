#!/usr/bin/env python3
"""Sweeps the MSM's stage variants on one GPU: number of batched-affine levels x batch size (optionally
APB_MSM_STEP / other environment knobs set by the caller), for 2^18-point commits (the prover's shape) and
single MSMs of 2^18 / 2^20 / 2^22.  Prints per-phase CUDA-event times; every result is checked against the
first variant of its shape.  Output: gpurun_out/msm_tune.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from ark_plonk_b200 import kzg, synth
from ark_plonk_b200._lib import get_lib
import ctypes as C

lib = get_lib(); lib.init(0)
lib.set_profiling(True)
out = []
shapes = [(18, 1), (18, 2), (18, 4), (18, 8), (18, 16), (20, 1), (22, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
variants = [(2, l) for l in (0, 1, 2, 3, 4, 5, 6)]
if os.environ.get("APB_TUNE_VARIANTS"):          # e.g. "2:3,2:4,1:3" = pairs kernel : levels
    variants = [tuple(int(x) for x in v.split(":")) for v in os.environ["APB_TUNE_VARIANTS"].split(",")]
keys = {}
for log_n, k in shapes:
    n = 1 << log_n
    if log_n not in keys:
        for ck in keys.values():
            ck.close()
        keys = {log_n: kzg.CommitterKey.from_tau(0, 0x1234567, n)}
    ck = keys[log_n]
    S = torch.from_numpy(np.concatenate([synth.seeded_scalars(0, n, seed=b"tune%d" % j) for j in range(k)]).view(np.int64)).cuda()
    so = (C.c_size_t * k)(*[j * n for j in range(k)])
    bo = (C.c_size_t * k)(*([0] * k))
    ln = (C.c_size_t * k)(*([n] * k))
    ref = None
    for pairs, levels in variants:
        env = {"APB_MSM_AFFINE_LEVELS": str(levels), "APB_MSM_AFFINE_MIN": "0"}
        os.environ.update(env)
        res = np.zeros((k, 18), dtype=np.uint64)
        ms, ph = [], []
        for it in range(5):
            lib.check(lib.c.apb_msm_batch_dev(ck._h, k, S.data_ptr(), so, bo, ln, 0, res.ctypes.data))
            if it >= 2:
                ms.append(lib.last_device_ms()); ph.append(lib.msm_phase_ms())
        if ref is None:
            ref = res.copy()
        ok = bool(np.array_equal(ref, res))
        rec = dict(log_n=log_n, k=k, pairs=pairs, levels=levels, plan=lib.msm_last_plan(), ms=float(np.mean(ms)),
                   mpts=k * n / np.mean(ms) / 1e3, same_result=ok, **{a: float(np.mean([p[a] for p in ph])) for a in ph[0]})
        print(json.dumps(rec), flush=True)
        out.append(rec)
        assert ok
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "msm_tune.json"), "w"), indent=1)
