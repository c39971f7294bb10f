#!/usr/bin/env python3
"""NTT tile-shape sweep: APB_NTT_MAX_LOG_TILE x APB_NTT_LOG_COLS for a few sizes (each in a subprocess)."""
import json, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from ark_plonk_b200._lib import get_lib
    from ark_plonk_b200.domain import Radix2EvaluationDomain
    lib = get_lib(); lib.init(0)
    stream = torch.cuda.ExternalStream(lib.c.apb_stream())
    res = {}
    for log_n in (18, 20, 24):
        n = 1 << log_n
        d = Radix2EvaluationDomain(0, n)
        x = torch.randint(0, 2**60, (n, 4), dtype=torch.int64, device="cuda"); y = torch.empty_like(x); torch.cuda.synchronize()
        for _ in range(3): d.ntt_dev(0, x.data_ptr(), n, y.data_ptr(), sync=True)
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(); d.ntt_dev(0, x.data_ptr(), n, y.data_ptr()); e1.record()
            e1.synchronize(); ts.append(e0.elapsed_time(e1))
        res[log_n] = float(np.median(ts))
        d.close()
    print(json.dumps(res))
else:
    # APB_NTT_LOG_RADIX: 2 / 3 = register-resident pass with 4 / 8 elements per thread, 0 = radix-2 stages in shared memory
    for lr in (2, 3, 0):
        for mt in ((10, 9, 8) if lr else (10, 8, 7)):
            for lc in (0, 1, 2):
                env = dict(os.environ, APB_NTT_LOG_RADIX=str(lr), APB_NTT_MAX_LOG_TILE=str(mt), APB_NTT_LOG_COLS=str(lc))
                out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
                print("log_radix", lr, "max_tile", mt, "log_cols", lc, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
