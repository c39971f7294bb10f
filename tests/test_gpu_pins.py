"""Pins of every size the benches measure or the docs claim ("pin every measured config").

  * the exact instances bench.py proves (tau, blinders) at 2^10 / 2^16 / 2^18: byte-identical to the
    ORACLE proof (tests/golden/plonk_bench_2p*.json, made by tools/gen_golden_2p18.py),
  * MSM at 2^24 and 2^26 on both curves by the KZG identity MSM(tau^i G, s) == [p_s(tau)]G, with p_s(tau)
    evaluated by the C oracle (Horner) - includes the bucket-range-slice path and its unbalanced fallback,
  * single transforms at 2^26 / 2^28: fft(p)[i] == p(w^i) and coset_fft(p)[i] == p(g w^i) at spot indices
    (C oracle Horner over the full vector), plus five round trips at 2^28 on a caller stream with no
    host synchronisation between producer and transform.
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


@pytest.mark.parametrize("degree", [10, 16, 18])
def test_bench_instance_matches_oracle_golden(gpu_lib, degree):
    """BASELINE configs #1-#3 with bench.py's own tau / blinders: proof bytes == the oracle's"""
    import prover_cases
    case = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_bench_2p%d.json" % degree)))
    assert case["degree"] == degree and case["curve"] == 0
    blob = prover_cases.prove_case(gpu_lib, case)
    assert hashlib.sha256(blob).hexdigest() == case["proof_sha256"]


def _msm_tau_check(lib, curve, n, S, montgomery, tau, ck=None):
    """MSM over resident powers of tau against [p_S(tau)]G, p_S(tau) by the C oracle's Horner"""
    from ark_plonk_b200 import encoding as enc, kzg, synth
    from oracle import cbuild
    own = ck is None
    if own:
        ck = kzg.CommitterKey.from_tau(curve, tau, n, lib=lib)
    try:
        out = kzg.multi_scalar_mul(ck, S, montgomery=montgomery)
        tau_m = enc.fr_to_mont(curve, [tau])[0]
        h = cbuild.fr_horner(curve, S, tau_m, threads=host_threads())
        # canonical scalars read as Montgomery residues are s_i / R: Horner then returns (e / R) * R = e
        e = enc.fr_from_mont(curve, h.reshape(1, 4))[0] if montgomery else enc.limbs_to_ints(h.reshape(1, 4))[0]
        assert enc.g1_from_xyz(curve, out) == synth.scalar_mul(curve, synth.G1_GENERATOR[curve], e), (curve, n)
    finally:
        if own:
            ck.close()
    return lib.msm_last_plan()


@pytest.mark.parametrize("curve,log_n", [(0, 24), (1, 24), (0, 26), (1, 26)])
def test_msm_large_tau_identity(gpu_lib, curve, log_n):
    """config #4's large sizes, both curves; 2^26 takes the bucket-range-slice path"""
    from ark_plonk_b200 import synth
    n = 1 << log_n
    S = synth.seeded_scalars(curve, n, seed=b"pin%d-%d" % (curve, log_n))
    plan = _msm_tau_check(gpu_lib, curve, n, S, False, 0xC0FFEE1234567 + curve)
    assert plan["pair_levels"] >= 2 and plan["unbalanced"] == 0
    if log_n >= 26:
        assert plan["slices"] > 1


@pytest.mark.parametrize("curve", [0, 1])
def test_msm_slices_and_unbalanced_fallback(gpu_lib, curve):
    """the sliced stage (forced by a small level-array budget) with uniform scalars, and with skewed scalars
    that overflow one slice -> the plain-accumulate fallback (msm.cu, `unbalanced`)"""
    from ark_plonk_b200 import kzg, synth
    n, tau = 1 << 18, 0xBADC0DE77 + curve
    ck = kzg.CommitterKey.from_tau(curve, tau, n, lib=gpu_lib)
    try:
        with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_AFFINE_MAX_BYTES=150 << 20):
            S = synth.seeded_scalars(curve, n, seed=b"slices")
            plan = _msm_tau_check(gpu_lib, curve, n, S, False, tau, ck=ck)
            assert plan["slices"] > 1 and plan["pair_levels"] >= 1 and plan["unbalanced"] == 0, plan
            skew = S.copy()
            skew[: n - n // 8] = np.array([0x0123012301230123] * 4, dtype=np.uint64) >> np.uint64(4)   # 7/8 of the digits land in 4 buckets
            plan = _msm_tau_check(gpu_lib, curve, n, skew, False, tau, ck=ck)
            assert plan["slices"] > 1 and plan["unbalanced"] == 1, plan
            plan = _msm_tau_check(gpu_lib, curve, n, skew, True, tau, ck=ck)      # same bytes read as Montgomery residues
    finally:
        ck.close()


def test_msm_pass_split_at_the_entry_bound(gpu_lib):
    pc.check_msm_pass_split(gpu_lib, 0, n=5000, k=4, limit_split=250000, limit_fail=60000)


def _ntt_spot_check(lib, curve, log_n, spots=3, seed=31):
    """one large transform resident in HBM: Horner spot checks of fft / coset_fft against the C oracle and
    round trips, all on a caller-provided stream with no host synchronisation in between"""
    import torch
    from ark_plonk_b200 import encoding as enc
    from ark_plonk_b200.domain import Radix2EvaluationDomain
    from oracle import cbuild
    from oracle.ntt import Domain
    n = 1 << log_n
    f = pc.FR[curve]
    od = Domain(f, log_n)
    d = Radix2EvaluationDomain(curve, n, lib=lib)
    side = torch.cuda.Stream()
    lib.set_stream(side.cuda_stream)
    try:
        with torch.cuda.stream(side):
            x = torch.randint(0, 2 ** 60, (n, 4), dtype=torch.int64, device="cuda")      # < 2^252 < r: valid residues
            y = torch.empty_like(x)
            rnd = random.Random(seed + log_n)
            idx = [0, 1, n - 1, n // 2 + 1] + [rnd.randrange(n) for _ in range(spots)]
            got = {}
            for kind in (0, 2):
                d.ntt_dev(kind, x.data_ptr(), n, y.data_ptr())           # enqueued behind randint on `side`
                got[kind] = y[idx].cpu().numpy().view(np.uint64)
            hx = x.cpu().numpy().view(np.uint64)
        for kind, shift in ((0, 1), (2, od.coset_gen)):
            for i, row in zip(idx, got[kind]):
                pt = shift * pow(od.group_gen, i, f.p) % f.p
                want = cbuild.fr_horner(curve, hx, enc.fr_to_mont(curve, [pt])[0], threads=host_threads())
                assert np.array_equal(row, want), ("fft" if kind == 0 else "coset_fft", log_n, i)
        del hx
        with torch.cuda.stream(side):
            for fwd, inv in ((2, 3), (0, 1)):
                d.ntt_dev(fwd, x.data_ptr(), n, y.data_ptr())
                d.ntt_dev(inv, y.data_ptr(), n, y.data_ptr())
                assert torch.equal(x, y)
    finally:
        lib.set_stream(None)
        d.close()
    del x, y
    torch.cuda.empty_cache()


@pytest.mark.parametrize("curve,log_n", [(0, 26), (1, 26), (0, 28)])
def test_ntt_large_horner_spot_checks(gpu_lib, curve, log_n):
    _ntt_spot_check(gpu_lib, curve, log_n)


def test_ntt_2p28_roundtrip_five_times_on_caller_stream(gpu_lib):
    """round 1 saw this fail once in four runs when the input was produced on another stream; with
    apb_set_stream the transform is ordered behind its producer without a host synchronisation"""
    import torch
    from ark_plonk_b200.domain import Radix2EvaluationDomain
    n = 1 << 28
    d = Radix2EvaluationDomain(0, n, lib=gpu_lib)
    side = torch.cuda.Stream()
    gpu_lib.set_stream(side.cuda_stream)
    try:
        with torch.cuda.stream(side):
            y = torch.empty((n, 4), dtype=torch.int64, device="cuda")
            for rep in range(5):
                x = torch.randint(0, 2 ** 60, (n, 4), dtype=torch.int64, device="cuda")
                d.ntt_dev(2, x.data_ptr(), n, y.data_ptr())
                assert not torch.equal(x[:1024], y[:1024])
                d.ntt_dev(3, y.data_ptr(), n, y.data_ptr())
                assert torch.equal(x, y), rep
                del x
    finally:
        gpu_lib.set_stream(None)
        d.close()
    del y
    torch.cuda.empty_cache()


def test_foreign_stream_input_without_host_sync(gpu_lib):
    """a tensor produced by a long chain of kernels on a foreign stream goes straight into apb_ntt_dev /
    apb_msm_dev on that stream: results equal the synchronised computation"""
    import torch
    from ark_plonk_b200 import kzg
    from ark_plonk_b200.domain import Radix2EvaluationDomain
    n = 1 << 20
    d = Radix2EvaluationDomain(0, n, lib=gpu_lib)
    ck = kzg.CommitterKey.from_tau(0, 0x77777, n, lib=gpu_lib)
    side = torch.cuda.Stream()
    try:
        base = torch.randint(0, 2 ** 59, (n, 4), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        y = torch.empty_like(base)
        gpu_lib.set_stream(side.cuda_stream)
        with torch.cuda.stream(side):
            x = base.clone()
            for _ in range(200):                      # keeps `side` busy for a while: x is final only at the end
                x = (x + 1) & ((1 << 59) - 1)
            d.ntt_dev(0, x.data_ptr(), n, y.data_ptr())
            out_async = np.zeros(18, dtype=np.uint64)
            gpu_lib.check(gpu_lib.c.apb_msm_dev(ck._h, 0, x.data_ptr(), n, 1, out_async.ctypes.data))
        torch.cuda.synchronize()
        gpu_lib.set_stream(None)
        y2 = torch.empty_like(base)
        d.ntt_dev(0, x.data_ptr(), n, y2.data_ptr(), sync=True)
        out_sync = np.zeros(18, dtype=np.uint64)
        gpu_lib.check(gpu_lib.c.apb_msm_dev(ck._h, 0, x.data_ptr(), n, 1, out_sync.ctypes.data))
        assert torch.equal(y, y2) and np.array_equal(out_async, out_sync)
    finally:
        gpu_lib.set_stream(None)
        d.close()
        ck.close()
