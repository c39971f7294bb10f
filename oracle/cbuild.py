"""Builds oracle/c/apb_oracle.c into oracle/_build/libapb_oracle.so (gcc -O3 -march=native).

Oracle = test infrastructure / CPU baseline (see oracle/__init__.py).  The reference itself
(Rust + un-vendored arkworks crates) cannot be compiled in this image, so there is no
oracle/_ref; this C restatement is the `"kind": "port"` baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "apb_oracle.c")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libapb_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -march=x86-64-v3 rather than native: the .so is built here and travels to the GPU box
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_msm.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        _lib.oracle_ntt.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        _lib.oracle_fr_to_mont.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        _lib.oracle_fr_from_mont.argtypes = [C.c_int, C.c_void_p, C.c_size_t]
        _lib.oracle_fr_horner.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        _lib.oracle_init()
    return _lib


def msm(curve: int, bases_mont, scalars_canonical, threads: int = 1):
    """numpy (n,12) / (n,4) uint64 -> (x, y) Montgomery limbs as numpy (12,) or None for identity."""
    import numpy as np
    b = np.ascontiguousarray(bases_mont, dtype=np.uint64)
    s = np.ascontiguousarray(scalars_canonical, dtype=np.uint64)
    out = np.zeros(12, dtype=np.uint64)
    inf = lib().oracle_msm(curve, b.ctypes.data, s.ctypes.data, s.shape[0], threads, out.ctypes.data)
    return None if inf else out


def ntt(curve: int, kind: int, data_mont, log_n: int, threads: int = 1):
    """numpy (in_len,4) uint64 Montgomery -> (n,4) transformed."""
    import numpy as np
    n = 1 << log_n
    buf = np.zeros((n, 4), dtype=np.uint64)
    d = np.ascontiguousarray(data_mont, dtype=np.uint64).reshape(-1, 4)
    buf[: d.shape[0]] = d
    rc = lib().oracle_ntt(curve, kind, buf.ctypes.data, log_n, d.shape[0], threads)
    if rc != 0:
        raise ValueError("oracle_ntt failed")
    return buf


def fr_horner(curve: int, coeffs_mont, x_mont, threads: int = 1):
    """sum_i coeffs[i] x^i over Fr: numpy (n,4) and (4,) uint64 Montgomery -> (4,) uint64 Montgomery."""
    import numpy as np
    c = np.ascontiguousarray(coeffs_mont, dtype=np.uint64).reshape(-1, 4)
    x = np.ascontiguousarray(x_mont, dtype=np.uint64).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_horner(curve, c.ctypes.data, c.shape[0], x.ctypes.data, threads, out.ctypes.data)
    return out
