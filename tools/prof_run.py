#!/usr/bin/env python3
"""Short workload for ncu: one resident key of 2^LOG points, a few MSMs and NTTs.
usage: prof_run.py [log_msm=18] [log_ntt=20]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import encoding as enc, kzg, synth  # noqa: E402
from ark_plonk_b200._lib import get_lib  # noqa: E402
from ark_plonk_b200.domain import Radix2EvaluationDomain  # noqa: E402

log_msm = int(sys.argv[1]) if len(sys.argv) > 1 else 18
log_ntt = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = get_lib()
lib.init(0)
n = 1 << log_msm
pts = synth.progression_bases(0, 12345, 67891, n)
ck = kzg.CommitterKey(0, enc.g1_affine_to_mont(0, pts))
S = synth.seeded_scalars(0, n)
dS = torch.from_numpy(S.view(np.int64)).cuda()
out = np.zeros(18, dtype=np.uint64)
for _ in range(3):
    lib.check(lib.c.apb_msm_dev(ck._h, 0, dS.data_ptr(), n, 0, out.ctypes.data))
print("msm ms", lib.last_device_ms())
N = 1 << log_ntt
d = Radix2EvaluationDomain(0, N)
x = torch.randint(0, 2**62, (N, 4), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()        # filled on torch's stream; the library uses its own
y = torch.empty_like(x)
for _ in range(3):
    d.ntt_dev(2, x.data_ptr(), N, y.data_ptr(), sync=True)
print("done")
