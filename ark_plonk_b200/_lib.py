"""ctypes binding of the apb C ABI (include/apb.h).

The product library is ark_plonk_b200/lib/libapb.so (built by ark_plonk_b200/build.py with
nvcc for sm_100a).  There is no CPU fallback: if the library is missing or no CUDA device is
present, calls raise.  (`Lib(path)` also lets the CPU test-suite bind the emulation build of
the same kernel sources; the package itself never does.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "lib", "libapb.so")

CURVE_BLS12_381 = 0
CURVE_BLS12_377 = 1
NTT_FFT, NTT_IFFT, NTT_COSET_FFT, NTT_COSET_IFFT = 0, 1, 2, 3

STATUS_NAMES = {
    1: "InvalidArgument", 2: "TooManyCoefficients", 3: "InvalidEvalDomainSize",
    4: "BadHandle", 5: "CudaError", 6: "OutOfMemory",
}


class ApbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (STATUS_NAMES.get(code, "status %d" % code), msg))
        self.code = code


class Lib:
    def __init__(self, path: str = DEFAULT_LIB):
        if not os.path.exists(path):
            raise ApbError(5, "native library %s not found - run `python -m ark_plonk_b200.build` "
                              "(there is no CPU fallback)" % path)
        self.path = path
        self.c = C.CDLL(path)
        c = self.c
        vp, sz, u64p, ci = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_int
        c.apb_init.argtypes = [ci]
        c.apb_last_error.restype = C.c_char_p
        c.apb_version.restype = C.c_char_p
        c.apb_ck_upload.argtypes = [ci, vp, sz, C.POINTER(vp)]
        c.apb_ck_from_tau.argtypes = [ci, vp, vp, sz, C.POINTER(vp)]
        c.apb_ck_download.argtypes = [vp, sz, sz, vp]
        c.apb_ck_size.argtypes = [vp, C.POINTER(sz)]
        c.apb_ck_free.argtypes = [vp]
        c.apb_ck_free.restype = None
        c.apb_msm.argtypes = [vp, sz, vp, sz, ci, vp]
        c.apb_msm_batch.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(sz), C.POINTER(sz), ci, vp]
        c.apb_msm_dev.argtypes = [vp, sz, vp, sz, ci, vp]
        c.apb_msm_batch_dev.argtypes = [vp, sz, vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), ci, vp]
        c.apb_g1_compress.argtypes = [ci, vp, vp]
        c.apb_g1_add.argtypes = [ci, vp, vp, vp]
        c.apb_g1_fold.argtypes = [ci, sz, vp, vp, sz, vp]
        c.apb_msm_totals.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), ci]
        c.apb_msm_totals.restype = None
        c.apb_ntt_totals.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), ci]
        c.apb_ntt_totals.restype = None
        c.apb_domain_new.argtypes = [ci, C.c_uint32, C.POINTER(vp)]
        c.apb_domain_size.argtypes = [vp, C.POINTER(sz)]
        c.apb_domain_free.argtypes = [vp]
        c.apb_domain_free.restype = None
        c.apb_ntt.argtypes = [vp, ci, vp, sz, vp]
        c.apb_ntt_dev.argtypes = [vp, ci, vp, sz, vp, ci]
        c.apb_ntt_batch_dev.argtypes = [vp, ci, vp, sz, sz, vp, sz, sz, ci]
        c.apb_dev_alloc.argtypes = [sz, C.POINTER(vp)]
        c.apb_dev_free.argtypes = [vp]
        c.apb_dev_upload.argtypes = [vp, vp, sz]
        c.apb_dev_download.argtypes = [vp, vp, sz]
        c.apb_stream.restype = vp
        c.apb_set_stream.argtypes = [vp]
        c.apb_field_op.argtypes = [ci, ci, vp, vp, vp, sz]
        c.apb_kernel_launches.restype = C.c_uint64
        c.apb_imad_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        c.apb_last_device_ms.restype = C.c_double
        vpp, szp = C.POINTER(vp), C.POINTER(sz)
        c.apb_fr_lincomb.argtypes = [ci, sz, vpp, szp, vp, vp, sz]
        c.apb_plonk_lookup_f.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, vp, sz]
        c.apb_plonk_combine_split.argtypes = [ci, vp, vp, sz, vp, vp]
        c.apb_plonk_perm_z.argtypes = [vp, vpp, vpp, vp, vp, vp]
        c.apb_plonk_lookup_z2.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        c.apb_plonk_quotient.argtypes = [ci, vpp, vp, vp, vp, sz]
        c.apb_plonk_quotient_full.argtypes = [ci, vpp, vp, vp, vp, sz]
        c.apb_plonk_quotient_range.argtypes = [ci, vpp, vp, vp, vp, sz, sz, sz]
        c.apb_poly_eval.argtypes = [ci, sz, vpp, szp, vp, vp]
        c.apb_poly_divide_linear.argtypes = [ci, vp, sz, vp, vp]
        c.apb_transcript_new.argtypes = [C.c_char_p, sz, vpp]
        c.apb_transcript_append.argtypes = [vp, C.c_char_p, sz, C.c_char_p, sz]
        c.apb_transcript_challenge.argtypes = [vp, C.c_char_p, sz, vp, sz]
        c.apb_transcript_free.argtypes = [vp]
        c.apb_transcript_free.restype = None
        c.apb_dev_sync.argtypes = []
        c.apb_mul_bench.argtypes = [ci, ci, ci, ci, C.c_uint32, C.POINTER(C.c_double)]
        c.apb_set_profiling.argtypes = [ci]
        c.apb_set_profiling.restype = None
        c.apb_msm_phase_ms.argtypes = [C.POINTER(C.c_double)]
        c.apb_msm_phase_ms.restype = None
        c.apb_msm_work.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), ci]
        c.apb_msm_work.restype = None
        c.apb_msm_call_ms.argtypes = [C.POINTER(C.c_double), ci]
        c.apb_msm_call_ms.restype = None
        c.apb_msm_last_plan.restype = None

    # ------------------------------------------------------------------
    def check(self, rc: int):
        if rc != 0:
            raise ApbError(rc, self.c.apb_last_error().decode())

    def init(self, device: int = -1):
        self.check(self.c.apb_init(device))

    def set_stream(self, cuda_stream: int | None):
        """enqueue this thread's later calls on `cuda_stream` (a cudaStream_t as int, e.g.
        torch.cuda.current_stream().cuda_stream); None = the library's own stream"""
        self.check(self.c.apb_set_stream(cuda_stream))

    def version(self) -> str:
        return self.c.apb_version().decode()

    def kernel_launches(self) -> int:
        return int(self.c.apb_kernel_launches())

    def last_device_ms(self) -> float:
        return float(self.c.apb_last_device_ms())

    def set_profiling(self, on: bool):
        self.c.apb_set_profiling(1 if on else 0)

    def msm_work(self, reset: bool = False):
        """(model, issued) wide multiply-adds of the bucket accumulation stage while profiling"""
        a, b = C.c_double(0), C.c_double(0)
        self.c.apb_msm_work(C.byref(a), C.byref(b), 1 if reset else 0)
        return a.value, b.value

    def msm_call_ms(self, reset: bool = False) -> float:
        """device ms of whole MSM calls since the last reset (while profiling)"""
        v = C.c_double(0)
        self.c.apb_msm_call_ms(C.byref(v), 1 if reset else 0)
        return v.value

    def msm_last_plan(self):
        arr = (C.c_uint32 * 4)()
        self.c.apb_msm_last_plan(arr)
        return dict(zip(("digit_bits", "pair_levels", "slices", "unbalanced"), list(arr)))

    def msm_phase_ms(self):
        arr = (C.c_double * 4)()
        self.c.apb_msm_phase_ms(arr)
        return dict(zip(("sort", "accumulate", "stitch", "reduce"), list(arr)))

    def mul_bench(self, field, threads, blocks_per_sm, ilp, iters=2000):
        v = C.c_double()
        self.check(self.c.apb_mul_bench(field, threads, blocks_per_sm, ilp, iters, C.byref(v)))
        return v.value

    def msm_totals(self, reset: bool = False):
        ms, pts = C.c_double(), C.c_ulonglong()
        self.c.apb_msm_totals(C.byref(ms), C.byref(pts), 1 if reset else 0)
        return ms.value, pts.value

    def ntt_totals(self, reset: bool = False):
        ms, cnt = C.c_double(), C.c_ulonglong()
        self.c.apb_ntt_totals(C.byref(ms), C.byref(cnt), 1 if reset else 0)
        return ms.value, cnt.value

    def g1_add(self, curve: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros(18, dtype=np.uint64)
        self.check(self.c.apb_g1_add(curve, self._ptr(a), self._ptr(b), self._ptr(out)))
        return out

    def g1_fold(self, curve: int, pieces: np.ndarray, group, k: int) -> np.ndarray:
        """out[j] = sum of pieces[p] with group[p] == j: (npieces, 18) normalised points -> (k, 18), one inversion"""
        pieces = np.ascontiguousarray(pieces, dtype=np.uint64).reshape(-1, 18)
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        out = np.zeros((k, 18), dtype=np.uint64)
        self.check(self.c.apb_g1_fold(curve, pieces.shape[0], self._ptr(pieces), grp.ctypes.data, k, self._ptr(out)))
        return out

    def imad_peak(self):
        w, n = C.c_double(), C.c_double()
        self.check(self.c.apb_imad_peak(C.byref(w), C.byref(n)))
        return w.value, n.value

    @staticmethod
    def _ptr(a: np.ndarray):
        return a.ctypes.data_as(C.c_void_p)

    def field_op(self, field: int, op: int, a: np.ndarray, b: np.ndarray | None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        bb = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else None
        self.check(self.c.apb_field_op(field, op, self._ptr(a), self._ptr(bb) if bb is not None else None,
                                       self._ptr(out), a.shape[0]))
        return out

    def g1_compress(self, curve: int, xyz: np.ndarray) -> bytes:
        xyz = np.ascontiguousarray(xyz, dtype=np.uint64)
        out = np.empty(48, dtype=np.uint8)
        self.check(self.c.apb_g1_compress(curve, self._ptr(xyz), self._ptr(out)))
        return out.tobytes()


_LIB: Lib | None = None


def get_lib() -> Lib:
    """The product library (CUDA).  Raises if it is not built or no GPU is present."""
    global _LIB
    if _LIB is None:
        _LIB = Lib(DEFAULT_LIB)
    return _LIB
