#!/usr/bin/env python3
"""BASELINE configs #4 / #5: standalone G1 MSM sweep (2^16..2^26, both curves) and Fr NTT sweep
(2^16..2^28, both fields) on one B200.  Bases = [tau^i]G generated on the device.  EVERY result is checked:
MSM against [e]G with e = sum s_i tau^i (one host scalar multiplication; e from Python big integers up to 2^20 and
from the device's Fr Horner kernel - an independent code path - at every size), NTT by round trip on slices
(Horner spot checks of the large transforms live in tests/test_gpu_pins.py)."""
import ctypes as C
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from ark_plonk_b200 import encoding as enc, kzg, synth
from ark_plonk_b200._lib import get_lib
from ark_plonk_b200.domain import Radix2EvaluationDomain
lib = get_lib(); lib.init(0)
wide, _ = lib.imad_peak()
out = {"imad_wide_per_s": wide, "msm": [], "ntt": []}
stream = torch.cuda.ExternalStream(lib.c.apb_stream())
max_msm = int(sys.argv[1]) if len(sys.argv) > 1 else 26
max_ntt = int(sys.argv[2]) if len(sys.argv) > 2 else 28
min_msm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lib.set_profiling(True)
for curve in (0, 1):
    for log_n in [16, 18, 20, 22, 24, 26]:
        if log_n > max_msm or log_n < min_msm: continue
        n = 1 << log_n
        tau = 0xABCDEF0123456789ABCDEF + curve
        t0 = time.time(); ck = kzg.CommitterKey.from_tau(curve, tau, n); t_setup = time.time() - t0
        S = synth.seeded_scalars(curve, n, seed=b"sweep%d" % log_n)
        dS = torch.from_numpy(S.view(np.int64)).cuda()
        o = np.zeros(18, dtype=np.uint64)
        ts, ph = [], []
        for i in range(5):
            lib.check(lib.c.apb_msm_dev(ck._h, 0, dS.data_ptr(), n, 0, o.ctypes.data))
            if i >= 2: ts.append(lib.last_device_ms()); ph.append(lib.msm_phase_ms())
        ms = float(np.median(ts))
        # e = sum s_i tau^i: Montgomery Horner over the raw scalar limbs at the point tau*R returns exactly that integer
        r = enc.FR_MODULUS[curve]
        ptr = (C.c_void_p * 1)(dS.data_ptr()); ln = (C.c_size_t * 1)(n)
        pt = enc.fr_to_mont(curve, [tau % r]); ev = np.zeros((1, 4), dtype=np.uint64)
        lib.check(lib.c.apb_poly_eval(curve, 1, ptr, ln, pt.ctypes.data, ev.ctypes.data))
        e_dev = enc.limbs_to_ints(ev)[0] % r
        ok = enc.g1_from_xyz(curve, o) == synth.scalar_mul(curve, synth.G1_GENERATOR[curve], e_dev)
        by = "device Fr Horner e = sum s_i tau^i, then [e]G on the host"
        if log_n <= 20:       # also the closed form in Python big integers (cheap enough up to here)
            e, tp = 0, 1
            for s in synth.limbs_to_int_list(S):
                e = (e + s * tp) % r; tp = tp * tau % r
            ok = ok and e == e_dev
            by += " + Python big-integer e"
        rec = dict(curve=curve, log_n=log_n, ms=ms, mpts=n / ms / 1e3, accumulate_ms=float(np.median([p["accumulate"] for p in ph])),
                   sort_ms=float(np.median([p["sort"] for p in ph])), reduce_ms=float(np.median([p["reduce"] for p in ph])),
                   imad_frac=n * 48000 / (ms * 1e-3) / wide, setup_s=t_setup, verified=bool(ok), verified_by=by, plan=lib.msm_last_plan())
        print(rec, flush=True); out["msm"].append(rec)
        ck.close(); del dS
lib.set_profiling(False)        # (profiling mode synchronises inside every NTT call)
for curve in (0, 1):
    for log_n in [16, 18, 20, 22, 24, 26, 28]:
        if log_n > max_ntt: continue
        n = 1 << log_n
        d = Radix2EvaluationDomain(curve, n)
        x = torch.randint(0, 2**60, (n, 4), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()        # filled on torch's stream; the library uses its own
        y = torch.empty_like(x)
        rec = dict(curve=curve, log_n=log_n)
        for kind, name in ((0, "fft"), (2, "coset_fft")):
            for _ in range(2): d.ntt_dev(kind, x.data_ptr(), n, y.data_ptr(), sync=True)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(); d.ntt_dev(kind, x.data_ptr(), n, y.data_ptr()); e1.record()
                e1.synchronize(); ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            rec[name + "_ms"] = ms; rec[name + "_gbs"] = 2 * n * 32 / (ms * 1e-3) / 1e9
            rec[name + "_imad_frac"] = (n / 2 * log_n + n) * 136 / (ms * 1e-3) / wide
        # round trip on the device: coset_ifft(coset_fft(x)) == x
        d.ntt_dev(2, x.data_ptr(), n, y.data_ptr(), sync=True); d.ntt_dev(3, y.data_ptr(), n, y.data_ptr(), sync=True)
        rec["roundtrip_ok"] = bool(torch.equal(x[: 1 << 16], y[: 1 << 16]) and torch.equal(x[-(1 << 16):], y[-(1 << 16):]))
        print(rec, flush=True); out["ntt"].append(rec)
        d.close(); del x, y
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "big_sweep.json"), "w"), indent=1)
