#!/usr/bin/env python3
"""Benchmark of the B200 proving hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload msm|ntt] [--log-n L]
    python bench.py --impl reference ...      # the CPU arm (arkworks-algorithm restatement)

One "step" = one pass of the hot path over one batch of synthetic input:
  prove (default): one PLONK proof of the reference's BenchCircuit at 2^L (default 2^18) gates,
        BLS12-381 / KZG10: 29 MSMs + 17 size-n + 11 size-4n transforms (10 coset FFTs + 1 coset IFFT;
        the reference's other three 4n transforms are of key polynomials and are precomputed with the
        key) + the pointwise kernels; the proof bytes are compared with the ORACLE's golden proof of the
        same instance (tests/golden/plonk_bench_2p*.json -> "golden_match"); metric = ms per proof.
        With --gpus N the default is ONE proof over N GPUs, SPMD (strong scaling): every rank runs the
        prover, each commit batch's MSM work is split evenly (144-byte partial sums all-reduced, folded by
        one C call), the 10 coset FFTs of round 4 are spread by polynomial and the quotient by index range
        (in-place NVLink all-gathers of the evaluation vectors); `phase_split` then is rank 0's share and
        carries the time spent waiting at the partial-sum exchanges; --prove-mode replicas = N independent
        proofs (weak scaling).
  msm : one KZG10 commitment MSM over 2^L (default 2^18) BLS12-381 G1 points per GPU
        (resident powers + precomputed table in HBM, seeded uniform scalars); --total-log-n T fixes the
        TOTAL size at 2^T points split over the ranks instead (strong scaling)
  ntt : one coset FFT of 2^L (default 2^20) Fr elements per GPU
`value` is device-resident throughput (inputs in HBM when the timed region starts); `e2e` is
the same metric through the blocking C-ABI call with HOST buffers (pinned scalars in, result
out).  With --gpus N (torchrun, one rank per GPU) every rank owns a shard of N*2^L points; the
partial sums (144 B each) are exchanged with one NCCL all-gather per step: weak scaling.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its version
# banner to fd 1 from C, past sys.stdout) and the result line goes to a private duplicate of the original stdout
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict) -> None:
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ncu --set full of the accumulation stage (tools/ncu_summary.py); the newest committed capture
NCU_STAGE_CAPTURE = next((c for c in ("r02_ncu_msm_stage.json", "r01_ncu_msm_stage.json")
                          if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", c))), "r01_ncu_msm_stage.json")
MSM_IMAD_PER_POINT = 48_000          # 16 windows x 10 Fq-mul x 300 wide multiply-adds
FR_MUL_IMAD = 136


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="prove", choices=["prove", "msm", "ntt"])
    ap.add_argument("--log-n", type=int, default=0)
    ap.add_argument("--total-log-n", type=int, default=0, help="msm: total points 2^T split over the ranks (strong scaling)")
    ap.add_argument("--curve", type=int, default=0, choices=[0, 1], help="msm / ntt: 0 = BLS12-381, 1 = BLS12-377")
    ap.add_argument("--prove-mode", default="split", choices=["split", "replicas"],
                    help="--gpus N > 1: 'split' = ONE proof, the polynomials of every commit batch spread over the ranks "
                         "(strong scaling); 'replicas' = N independent proofs (weak scaling)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons during the timed region (NVML)"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def ncu_traffic(kernel_prefix: str, per: str | None = None, capture: str = "r01_ncu_prove_kernels.json"):
    """average DRAM bytes per launch of the kernels matching `kernel_prefix` in a committed ncu capture
    (profiles/), or None.  `per`: count one unit per launch of THAT kernel instead (a stage made of
    several kernels: bytes of all matching kernels per launch of the stage's last kernel)."""
    try:
        rows = json.load(open(os.path.join(ROOT, "profiles", capture)))
    except Exception:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, cnt = 0.0, 0
    if per:
        units = sum(1 for r in rows if per in r.get("Kernel Name", ""))
        for r in rows:
            if kernel_prefix not in r.get("Kernel Name", ""):
                continue
            for k, v in r.items():
                if k.startswith("dram__bytes_read.sum") or k.startswith("dram__bytes_write.sum"):
                    unit = k[k.index("[") + 1:-1] if "[" in k else "byte"
                    tot += float(v) * scale.get(unit, 1.0)
        return tot / units if units else None
    for r in rows:
        if kernel_prefix not in r.get("Kernel Name", ""):
            continue
        b = 0.0
        for k, v in r.items():
            if k.startswith("dram__bytes_read.sum") or k.startswith("dram__bytes_write.sum"):
                unit = k[k.index("[") + 1:-1] if "[" in k else "byte"
                b += float(v) * scale.get(unit, 1.0)
        tot += b
        cnt += 1
    return tot / cnt if cnt else None


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------
def cpu_msm_baseline(log_sample: int, threads: int, seed: bytes):
    """arkworks-algorithm restatement (oracle/c) on host cores over a bounded sample."""
    from ark_plonk_b200 import encoding as enc
    from ark_plonk_b200 import synth
    from oracle import cbuild
    n = 1 << log_sample
    pts = synth.progression_bases(0, 12345, 67891, n)
    B = enc.g1_affine_to_mont(0, pts)
    S = synth.seeded_scalars(0, n, seed=seed)
    t0 = time.perf_counter()
    out = cbuild.msm(0, B, S, threads=threads)
    dt = time.perf_counter() - t0
    exp = synth.progression_expected(0, 12345, 67891, synth.limbs_to_int_list(S))
    assert out is not None and tuple(enc.fq_from_mont(0, out.reshape(2, 6))) == exp
    return n / dt / 1e6, dt


def cpu_ntt_baseline(log_n: int, threads: int):
    from oracle import cbuild
    rng = np.random.default_rng(7)
    X = rng.integers(0, 1 << 62, size=(1 << log_n, 4), dtype=np.uint64)
    t0 = time.perf_counter()
    cbuild.ntt(0, 2, X, log_n, threads=threads)
    dt = time.perf_counter() - t0
    return 2 * (1 << log_n) * 32 / dt / 1e9, dt


def cpu_prove_schedule(log_n: int, threads: int):
    """ONE execution of the hot-path schedule of Prover::prove_with_preprocessed (prover.rs:163-638, SURVEY.md
    3.2) on the arkworks-algorithm C port (oracle/c): 29 x VariableBaseMSM with the proof's lengths
    (27 of n coefficients, the two opening witnesses n - 1), 13 ifft(n) + 4 fft(n), 13 coset_fft(4n) of
    n coefficients + 1 coset_ifft(4n) - every call really executed, with all `threads` host threads.
    The reference's serial pointwise loops (quotient_poly.rs:167-173,208-266; permutation/mod.rs:694-744)
    are NOT included, so this is a lower bound of the reference's CPU time.  Returns (ms, breakdown)."""
    from ark_plonk_b200 import encoding as enc
    from ark_plonk_b200 import synth
    from oracle import cbuild
    n = 1 << log_n
    log_b = min(log_n, 14)                       # bases: a 2^14 progression tiled (any points time the same)
    pts = synth.progression_bases(0, 12345, 67891, 1 << log_b)
    B = np.tile(enc.g1_affine_to_mont(0, pts), (n >> log_b, 1))
    rng = np.random.default_rng(7)
    X = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    t_msm = t_n = t_4n = 0.0
    t_all = time.perf_counter()
    for j, ln in enumerate([n] * 27 + [n - 1] * 2):
        S = synth.seeded_scalars(0, ln, seed=b"cpu-prove-%d" % j)
        t0 = time.perf_counter()
        cbuild.msm(0, B[:ln], S, threads=threads)
        t_msm += time.perf_counter() - t0
    for kind in [1] * 13 + [0] * 4:
        t0 = time.perf_counter()
        cbuild.ntt(0, kind, X, log_n, threads=threads)
        t_n += time.perf_counter() - t0
    for kind in [2] * 13 + [3]:
        t0 = time.perf_counter()
        cbuild.ntt(0, kind, X, log_n + 2, threads=threads)
        t_4n += time.perf_counter() - t0
    total = time.perf_counter() - t_all
    return total * 1e3, dict(msm_s=t_msm, ntt_n_s=t_n, ntt_4n_s=t_4n, total_s=total)


def cpu_prove_sample_text(log_n, parts):
    return ("one full execution of the proof's hot-path schedule on the arkworks-algorithm C port: 29 MSM (27 x 2^%d + 2 x (2^%d - 1) "
            "points, %.2f s), 17 NTT(2^%d) (%.2f s), 14 NTT(2^%d) (%.2f s), every call executed; the reference's serial pointwise "
            "loops are excluded (lower bound)" % (log_n, log_n, parts["msm_s"], log_n, parts["ntt_n_s"], log_n + 2, parts["ntt_4n_s"]))


def reference_prove(args, log_n, cores):
    """every step really executes the whole schedule; the number of steps is bounded by a wall-clock budget
    (the line reports the steps actually executed, not the ones requested)"""
    from oracle import cbuild
    cbuild.build()
    budget_s = float(os.environ.get("APB_REF_BUDGET_S", "150"))
    t_start = time.perf_counter()
    vals, parts, warm = [], None, 0
    if args.warmup > 0:                          # at most one untimed warm-up schedule (page-in, thread start-up)
        warm_log = log_n if log_n <= 14 else 14
        cpu_prove_schedule(warm_log, cores)
        warm = 1
    while len(vals) < max(args.steps, 1):
        v, parts = cpu_prove_schedule(log_n, cores)
        vals.append(v)
        if time.perf_counter() - t_start + v * 1e-3 > budget_s:
            break
    value = float(np.mean(vals))
    sample = cpu_prove_sample_text(log_n, parts)
    return {
        "impl": "reference", "metric": "plonk_prove_ms", "value": value, "unit": "ms", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": value / 20184.0 if log_n == 18 else None,
        "dtype": "u64-limb integers", "data": "synthetic",
        "config": {"workload": "BLS12-381 KZG10 prove, BenchCircuit 2^%d gates (with lookups), %d real rows" % (log_n, (1 << (log_n - 1)) + 2),
                   "curve": "BLS12-381", "msm_per_proof": 29, "ntt_n_per_proof": 17, "ntt_4n_per_proof": 14,
                   "ntt_4n_note": "the reference's schedule: 13 coset FFTs + 1 coset IFFT; the B200 arm executes 11 of them per proof "
                                  "(the 3 coset FFTs of key polynomials are resident with the prover key)"},
        "cpu_baseline": {"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "arkworks-0.3-algorithm restatement in C (oracle/c); the Rust reference cannot be built here. "
                "Steps are bounded by a %.0f s wall-clock budget: `steps` is what was executed (the one warm-up runs the same schedule "
                "at 2^%d). Published reference: 20.184 s on a Ryzen 7 3700X (README.md:107)" % (budget_s, min(log_n, 14)),
    }


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (C restatement; the Rust reference cannot
    be built in this image) with all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbuild
    cbuild.build()
    cores = host_cores()
    vals, times = [], []
    if args.workload == "prove":
        log_n = args.log_n or 18
        line = reference_prove(args, log_n, cores)
        emit(line)
        return
    if args.workload == "msm":
        log_n = args.log_n or 18
        log_sample = min(log_n, 15)
        for i in range(args.warmup + args.steps):
            v, dt = cpu_msm_baseline(log_sample, cores, b"ref%d" % i)
            if i >= args.warmup:
                vals.append(v)
                times.append(dt)
        metric, unit = "msm_mpts_per_s", "Mpts/s"
        workload = "KZG10 commitment MSM, BLS12-381 G1, 2^%d points per GPU" % log_n
        sample = "VariableBaseMSM over 2^%d of the 2^%d points per step" % (log_sample, log_n)
    else:
        log_n = args.log_n or 20
        log_sample = min(log_n, 18)
        for i in range(args.warmup + args.steps):
            v, dt = cpu_ntt_baseline(log_sample, cores)
            if i >= args.warmup:
                vals.append(v)
                times.append(dt)
        metric, unit = "ntt_gb_per_s", "GB/s"
        workload = "coset FFT, BLS12-381 Fr, 2^%d elements per GPU" % log_n
        sample = "coset_fft of 2^%d elements per step" % log_sample
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32-limb integers", "data": "synthetic",
        "config": {"workload": workload, "curve": "BLS12-381"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "arkworks-0.3-algorithm restatement in C (oracle/c); the Rust reference cannot be built here",
    }
    emit(line)


# --------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from ark_plonk_b200 import encoding as enc
    from ark_plonk_b200 import kzg, synth
    from ark_plonk_b200._lib import get_lib
    from ark_plonk_b200.domain import Radix2EvaluationDomain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = get_lib()
    lib.init(local)
    stream = torch.cuda.ExternalStream(lib.c.apb_stream())
    K, W = args.steps, args.warmup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wide_peak, imad32_peak = lib.imad_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    line = {}
    if args.workload == "prove":
        import hashlib

        from ark_plonk_b200 import bench_circuit as bc
        from ark_plonk_b200 import plonk as gp
        log_n = args.log_n or 18
        split = args.prove_mode == "split"           # one proof over all ranks (SPMD); at N = 1 the same single proof
        salt = 0 if split else rank
        tau = 0x1234567890ABCDEF1234567890ABCDEF + salt
        circ = bc.build(0, log_n, [1000 + 8 * salt + i for i in range(8)])
        n = circ.n
        ck = kzg.CommitterKey.from_tau(0, tau, n + 1)
        committer = None
        if split and world > 1:
            from ark_plonk_b200 import parallel
            committer = parallel.DistributedCommitter(0, ck, device="cuda")
        # (a torch-backed arena at N > 1: the round-4 coset FFTs / quotient slices are all-gathered on arena views)
        pr = gp.Prover(0, ck, committer=committer, arena_device="cuda" if committer is not None else None)
        pk = pr.preprocess(circ, commit_verifier_key=False)
        wires = gp.wires_to_mont(circ)
        wires_pinned = torch.from_numpy(wires.view(np.int64)).pin_memory()
        w_res = pr.upload_wires(pk, wires)
        lib.set_profiling(True)
        proof = None
        for _ in range(W):
            proof = pr.prove(pk, None, b"ark", wires_resident=w_res)
        sampler = ClockSampler(local)
        sampler.start()
        lib.msm_totals(reset=True)
        lib.msm_work(reset=True)
        lib.msm_call_ms(reset=True)
        launches0 = lib.kernel_launches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(K):
            proof = pr.prove(pk, None, b"ark", wires_resident=w_res)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        launches = lib.kernel_launches() - launches0
        acc_total_ms, pts_total = lib.msm_totals()
        madds_model, madds_issued = lib.msm_work()
        msm_calls_ms = lib.msm_call_ms()
        clocks = sampler.result()
        ms_per_step = total_ms / K
        # end to end: pinned host witness in, proof bytes out
        for _ in range(W):
            proof_e2e = pr.prove(pk, wires_pinned.data_ptr(), b"ark")
        assert proof_e2e == proof
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            proof_e2e = pr.prove(pk, wires_pinned.data_ptr(), b"ark")
        barrier()
        e2e_local = (time.perf_counter() - t0) * 1e3
        # per-phase split of one profiled proof (outside the timed regions): MSM / NTT / other
        lib.ntt_totals(reset=True)
        pr.phase_log = []
        if committer is not None:
            committer.exchange_ms = committer.fold_ms = 0.0
        barrier()
        t0 = time.perf_counter()
        pr.prove(pk, None, b"ark", wires_resident=w_res)
        prof_ms = (time.perf_counter() - t0) * 1e3
        log, pr.phase_log = pr.phase_log, None
        comm_split = None
        if committer is not None:
            comm_split = {"partial_sum_exchange_ms": committer.exchange_ms, "host_fold_ms": committer.fold_ms}
        msm_ms = sum(m for _, m, _ in log)
        ntt_ms, ntt_cnt = lib.ntt_totals()
        phase_split = {"msm_ms": msm_ms, "ntt_ms": ntt_ms, "other_ms": max(prof_ms - msm_ms - ntt_ms, 0.0), "profiled_proof_ms": prof_ms,
                       "msm_calls": len(log), "msm_count": sum(k for k, _, _ in log), "ntt_transforms": int(ntt_cnt),
                       "msm_accumulate_ms": sum(p["accumulate"] for _, _, p in log), "msm_sort_ms": sum(p["sort"] for _, _, p in log),
                       "msm_reduce_ms": sum(p["reduce"] for _, _, p in log)}
        if comm_split:
            phase_split.update(comm_split)
        # extra (not the headline): the same proof without the 14 commitments whose results the reference discards
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(K // 2, 1)):
            proof_nd = pr.prove(pk, None, b"ark", wires_resident=w_res, faithful=False)
        barrier()
        no_dead_ms = (time.perf_counter() - t0) * 1e3 / max(K // 2, 1)
        assert proof_nd == proof
        e2e_ms = max_over_ranks(e2e_local) / K
        sha = hashlib.sha256(proof).hexdigest()
        golden_match = None                      # the oracle's proof of this exact instance (tools/gen_golden_2p18.py)
        gpath = os.path.join(ROOT, "tests", "golden", "plonk_bench_2p%d.json" % log_n)
        if salt == 0 and os.path.exists(gpath):
            golden_match = json.load(open(gpath))["proof_sha256"] == sha
            assert golden_match, "proof differs from the oracle's golden proof of the bench instance"
        stage_s = acc_total_ms * 1e-3
        achieved = pts_total * MSM_IMAD_PER_POINT / stage_s / 1e12
        ncalls = max(len(log), 1)
        acc_per_launch = acc_total_ms / (K * ncalls)
        line = {
            "metric": "plonk_prove_ms", "value": ms_per_step, "unit": "ms", "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong" if split else "weak",
            "vs_baseline": ms_per_step / 20184.0 if log_n == 18 else None,
            "config": {"workload": "BLS12-381 KZG10 prove, BenchCircuit 2^%d gates (with lookups), %d real rows" % (log_n, circ.rows),
                       "curve": "BLS12-381", "msm_per_proof": 29, "ntt_n_per_proof": 17, "ntt_4n_per_proof": 11,
                       "multi_gpu": ("one proof, SPMD over %d ranks: every rank runs the prover, the points of each of the %d commit "
                                     "batches are split evenly (one all-reduce of 144-byte partial sums per batch), the 10 coset FFTs "
                                     "of round 4 are spread by polynomial and the quotient by index range (NVLink all-gathers); "
                                     "phase_split = rank 0's share" % (world, ncalls))
                                    if split and world > 1 else ("%d independent proofs" % world if world > 1 else "single GPU"),
                       "proof_sha256": sha, "golden_match": golden_match,
                       "extra_ms_without_discarded_commitments": no_dead_ms,
                       "phase_split_of_one_profiled_proof": phase_split,
                       "l2": "inputs exceed L2 (resident key table %d MB, prover key %d MB)" % ((n + 1) * 16 * 96 >> 20, pk.arena.elems * 32 >> 20),
                       "vs_baseline_note": "this value / published 20184 ms (Ryzen 7 3700X CPU, reference README.md:107); < 1 is faster"},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": 4 * n * 32, "d2h_bytes_per_step": len(proof) + 29 * 3024},
            "roofline": {"kernel": "bucket accumulation stage of the MSM: k_msm_pairs_coop (batched-affine levels) + k_msm_accumulate (XYZZ)",
                         "bound": "int32", "achieved": achieved, "peak": wide_peak / 1e12,
                         "unit": "T wide-IMAD/s", "frac": achieved / (wide_peak / 1e12),
                         "frac_whole_msm": pts_total * MSM_IMAD_PER_POINT / (msm_calls_ms * 1e-3) / wide_peak if msm_calls_ms else None,
                         "frac_whole_step": pts_total * MSM_IMAD_PER_POINT / (total_ms * 1e-3) / wide_peak,
                         "traffic": ncu_traffic("k_msm_", per="k_msm_accumulate", capture=NCU_STAGE_CAPTURE),
                         "issued": {"achieved": madds_issued / stage_s / 1e12,
                                    "frac": madds_issued / stage_s / wide_peak,
                                    "issued_over_model": madds_issued / madds_model if madds_model else None,
                                    "note": "multiply-adds actually issued: a batched-affine pair addition costs 6 Fq products, the XYZZ "
                                            "mixed addition of the cost model 10; `achieved`/`frac` keep SURVEY 8(d)'s algorithmic figure"},
                         "traffic_note": "DRAM bytes of the stage (pair levels + accumulate) per commit call, ncu --set full (profiles/%s); "
                                         "algorithmic bytes per call = entries x (4 B id + 96 B point) = %.2e (mean of the %d calls)" % (
                                             NCU_STAGE_CAPTURE, pts_total / K * 16 * 100 / ncalls, ncalls),
                         "kernel_ms_per_launch": acc_per_launch, "launches_per_step": ncalls,
                         "kernel_share_of_step": acc_total_ms / K / ms_per_step,
                         "whole_msm_ms_per_step": msm_calls_ms / K,
                         "frac_note": "peak = independent (carry-free) IMAD.WIDE chains; the carry-chained IMAD.WIDE.X of a Montgomery product "
                                      "issues at half that rate, so 0.5 is the ceiling of THIS carry-chain formulation (not of the chip); "
                                      "`frac` times the accumulation stage only, `frac_whole_msm` all MSM calls (sort, reduction, copy-out, "
                                      "host epilogue included), `frac_whole_step` the whole proof",
                         "peak_source": "measured live: independent mad.wide.u32 chains (apb_imad_peak)",
                         "note": "algorithmic 48000 wide multiply-adds per point x %d points per proof (SURVEY 8d)" % (pts_total // K)},
        }
        if rank == 0:
            v, parts = cpu_prove_schedule(log_n, host_cores())
            line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": host_cores(), "kind": "port",
                                    "sample": cpu_prove_sample_text(log_n, parts)}
    elif args.workload == "msm":
        curve = args.curve
        cname = "BLS12-381" if curve == 0 else "BLS12-377"
        strong = args.total_log_n > 0
        if strong:                               # fixed total size split over the ranks
            total_n = 1 << args.total_log_n
            n = total_n // world
            log_n = args.total_log_n
        else:
            log_n = args.log_n or 18
            n = 1 << log_n
            total_n = world * n
        # rank r owns points [r*n, (r+1)*n) of the (world*n)-point MSM over powers of tau (what KZG really uses);
        # the slice of the key is generated on the device: [tau^(r n + i)] G = [tau^i] (tau^(r n) G)
        tau = 0x9E3779B97F4A7C15F39CC0605CEDC835 + curve
        r_mod = enc.FR_MODULUS[curve]
        shift = pow(tau, rank * n, r_mod)
        gen = synth.scalar_mul(curve, synth.G1_GENERATOR[curve], shift)
        ck = kzg.CommitterKey.from_tau(curve, tau, n, generator=gen)
        S = synth.seeded_scalars(curve, n, seed=b"bench%d" % rank)
        dS = torch.from_numpy(S.view(np.int64)).cuda()
        S_pinned = torch.from_numpy(S.view(np.int64)).pin_memory()
        out = np.zeros(18, dtype=np.uint64)
        gathered = torch.zeros(world * 18, dtype=torch.int64, device="cuda") if world > 1 else None

        def fold():             # exchange the 144-byte partial sums; every rank folds them on the host
            mine = torch.from_numpy(out.view(np.int64)).cuda()
            dist.all_gather_into_tensor(gathered, mine)
            parts = gathered.cpu().numpy().view(np.uint64).reshape(world, 18)
            acc = parts[0]
            for r in range(1, world):
                acc = lib.g1_add(curve, acc, parts[r])
            return acc

        def step_dev():
            lib.check(lib.c.apb_msm_dev(ck._h, 0, dS.data_ptr(), n, 0, out.ctypes.data))
            return fold() if world > 1 else out

        def step_e2e():
            lib.check(lib.c.apb_msm(ck._h, 0, S_pinned.data_ptr(), n, 0, out.ctypes.data))
            return fold() if world > 1 else out

        lib.set_profiling(True)
        res = None
        for _ in range(W):
            res = step_dev()
        # correctness of the WHOLE result, outside the timed region: MSM(tau^i G, s) == [sum_i s_i tau^i] G with the
        # exponent evaluated by the oracle's Horner (C, host) over every rank's scalars
        from oracle import cbuild
        cbuild.build()
        tau_m = enc.fr_to_mont(curve, [tau])[0]
        h = cbuild.fr_horner(curve, S, tau_m, threads=host_cores())
        e_local = enc.limbs_to_ints(h.reshape(1, 4))[0] * shift % r_mod
        if world > 1:
            lst = [None] * world
            dist.all_gather_object(lst, e_local)
            e_total = sum(lst) % r_mod
        else:
            e_total = e_local
        verified = enc.g1_from_xyz(curve, res) == synth.scalar_mul(curve, synth.G1_GENERATOR[curve], e_total)
        assert verified, "MSM result mismatch"
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = lib.kernel_launches()
        lib.msm_call_ms(reset=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc_ms, phases = [], []
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(K):
            step_dev()
            ph = lib.msm_phase_ms()
            acc_ms.append(ph["accumulate"])
            phases.append(ph)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        launches = lib.kernel_launches() - launches0
        msm_calls_ms = lib.msm_call_ms()
        clocks = sampler.result()
        ms_per_step = total_ms / K
        value = total_n / (ms_per_step * 1e-3) / 1e6
        # end to end through the blocking C-ABI call with host scalars
        for _ in range(W):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
        acc = float(np.mean(acc_ms))
        achieved = n * MSM_IMAD_PER_POINT / (acc * 1e-3) / 1e12
        plan = lib.msm_last_plan()
        line = {
            "metric": "msm_mpts_per_s", "value": value, "unit": "Mpts/s", "ms_per_step": ms_per_step,
            "scaling": "strong" if strong else "weak",
            "config": {"workload": "KZG10 commitment MSM, %s G1, 2^%d points %s" % (cname, log_n, "in total" if strong else "per GPU"),
                       "curve": cname, "points_per_gpu": n, "points_total": total_n, "verified": bool(verified),
                       "digit_bits": plan["digit_bits"], "pair_levels": plan["pair_levels"], "bucket_range_slices": plan["slices"],
                       "phase_ms": {k: float(np.mean([p[k] for p in phases])) for k in phases[0]},
                       "l2": "inputs exceed L2 (resident base table %d MB)" % (n * 16 * 96 >> 20)},
            "e2e": {"value": total_n / (e2e_ms * 1e-3) / 1e6, "unit": "Mpts/s", "h2d_bytes_per_step": n * 32,
                    "d2h_bytes_per_step": 144 + 15 * 192},
            "roofline": {"kernel": "bucket accumulation stage of the MSM: k_msm_pairs_coop (batched-affine levels) + k_msm_accumulate (XYZZ)",
                         "bound": "int32", "achieved": achieved, "peak": wide_peak / 1e12,
                         "unit": "T wide-IMAD/s", "frac": achieved / (wide_peak / 1e12),
                         "frac_whole_msm": n * K * MSM_IMAD_PER_POINT / (msm_calls_ms * 1e-3) / wide_peak if msm_calls_ms else None,
                         "traffic": None,
                         "kernel_ms": acc, "kernel_share_of_step": acc / ms_per_step,
                         "peak_source": "measured live: independent mad.wide.u32 chains (apb_imad_peak)",
                         "note": "algorithmic 48000 wide multiply-adds per point (SURVEY 8d: XYZZ cost model); carry-chained IMAD.WIDE.X "
                                 "issues at half the plain IMAD.WIDE rate, so 0.5 is the ceiling of this carry-chain formulation for a pure "
                                 "XYZZ accumulation - the pair levels issue 6 products per addition instead of 10 and can exceed it"},
        }
        if rank == 0:
            v, dt = cpu_msm_baseline(min(log_n, 15), host_cores(), b"cpu")
            line["cpu_baseline"] = {"value": v, "unit": "Mpts/s", "cores": host_cores(), "kind": "port",
                                    "sample": "arkworks-algorithm VariableBaseMSM (oracle/c) over 2^%d points, %.1f s" % (min(log_n, 15), dt)}
    else:
        log_n = args.log_n or 20
        n = 1 << log_n
        dom = Radix2EvaluationDomain(args.curve, n)
        x = torch.randint(0, 2 ** 62, (n, 4), dtype=torch.int64, device="cuda")
        y = torch.empty_like(x)
        hx = x.cpu().pin_memory()
        hy = np.empty((n, 4), dtype=np.uint64)

        def step_dev():
            dom.ntt_dev(2, x.data_ptr(), n, y.data_ptr())

        def step_e2e():
            lib.check(lib.c.apb_ntt(dom._h, 2, hx.data_ptr(), n, hy.ctypes.data))

        for _ in range(W):
            step_dev()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = lib.kernel_launches()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(K):
                step_dev()
            e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        launches = lib.kernel_launches() - launches0
        clocks = sampler.result()
        ms_per_step = total_ms / K
        gbs = 2 * n * 32 / (ms_per_step * 1e-3) / 1e9
        for _ in range(W):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step_e2e()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
        imad = (n / 2 * log_n + n) * FR_MUL_IMAD / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": "ntt_gb_per_s", "value": world * gbs, "unit": "GB/s", "ms_per_step": ms_per_step,
            "config": {"workload": "coset FFT, %s Fr, 2^%d elements per GPU" % ("BLS12-381" if args.curve == 0 else "BLS12-377", log_n),
                       "curve": "BLS12-381" if args.curve == 0 else "BLS12-377",
                       "l2": "vector %d MB %s L2" % (n * 32 >> 20, "exceeds" if n * 32 > 126 << 20 else "fits in")},
            "e2e": {"value": world * 2 * n * 32 / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n * 32,
                    "d2h_bytes_per_step": n * 32},
            "roofline": {"kernel": "k_ntt_pass_reg (4 elements per thread, two stages per shared-memory exchange)", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "traffic": None, "peak_source": hbm_src,
                         "int32": {"achieved": imad, "peak": wide_peak / 1e12, "unit": "T wide-IMAD/s", "frac": imad / (wide_peak / 1e12)},
                         "note": "2*N*32 algorithmic bytes; the transform is INT32-issue bound (SURVEY 8d), both fractions reported"},
        }
        if rank == 0:
            v, dt = cpu_ntt_baseline(min(log_n, 18), host_cores())
            line["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": host_cores(), "kind": "port",
                                    "sample": "arkworks-algorithm coset_fft (oracle/c) of 2^%d elements, %.2f s" % (min(log_n, 18), dt)}

    line.setdefault("higher_is_better", True)
    line.setdefault("vs_baseline", None)
    line.setdefault("scaling", "weak")
    line.update({"n_gpus": world, "steps": K, "warmup": W,
                 "dtype": "u32-limb integers (381/255-bit Montgomery)", "data": "synthetic", "gpu_launches": int(launches),
                 "clocks": clocks})
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
