#!/usr/bin/env python3
"""Short workload for ncu: batched commits of K polynomials of 2^LOG coefficients (the prover's MSM shape) through the
batched-affine pair levels.  usage: prof_pairs.py [log=18] [k=8] [reps=2]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import kzg, synth  # noqa: E402
from ark_plonk_b200._lib import get_lib  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib = get_lib()
lib.init(0)
n = 1 << log_n
ck = kzg.CommitterKey.from_tau(0, 0x1234567, n)
S = torch.from_numpy(np.concatenate([synth.seeded_scalars(0, n, seed=b"prof%d" % j) for j in range(k)]).view(np.int64)).cuda()
torch.cuda.synchronize()
so = (C.c_size_t * k)(*[j * n for j in range(k)])
bo = (C.c_size_t * k)(*([0] * k))
ln = (C.c_size_t * k)(*([n] * k))
res = np.zeros((k, 18), dtype=np.uint64)
for _ in range(reps):
    lib.check(lib.c.apb_msm_batch_dev(ck._h, k, S.data_ptr(), so, bo, ln, 0, res.ctypes.data))
print("ms", lib.last_device_ms(), lib.msm_last_plan())
