"""Build the CPU-emulation copy of the kernel sources (TEST INFRASTRUCTURE ONLY).

Compiles ark_plonk_b200/csrc/*.cu as plain C++ against tests/emu/cuda_emu.h into
tests/emu/_build/libapb_emu.so.  Used by the `-m "not gpu"` tests to check kernel logic
against the oracle on a box without a GPU.  The product package never loads this library.
"""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "ark_plonk_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libapb_emu.so")
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "msm_acc.cu", "msm_pairs_coop.cu", "msm_setup.cu", "poly.cu", "transcript.cu"]
FLAGS = ["-std=c++20", "-O2", "-DAPB_EMU", "-fPIC", "-pthread", "-I", HERE, "-I", CSRC, "-w"]


def build_asan() -> str:
    """AddressSanitizer variant (tools/emu_asan.sh): exact-size 'device' allocations, -fsanitize=address"""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libapb_emu_asan.so")
    objs = []
    for s in SOURCES:
        obj = os.path.join(OUT, s.replace(".cu", ".asan.o"))
        objs.append(obj)
        subprocess.check_call(["g++", *FLAGS, "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-DAPB_EMU_ASAN",
                               "-x", "c++", "-c", os.path.join(CSRC, s), "-o", obj])
    subprocess.check_call(["g++", "-shared", "-pthread", "-fsanitize=address", "-o", lib, *objs])
    return lib


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for d in (CSRC, HERE):
        for f in sorted(os.listdir(d)):
            p = os.path.join(d, f)
            if os.path.isfile(p) and not f.endswith(".py"):
                h.update(f.encode() + open(p, "rb").read())
    h.update(open(os.path.join(ROOT, "include", "apb.h"), "rb").read())
    stamp = os.path.join(OUT, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    procs, objs = [], []
    for s in SOURCES:
        obj = os.path.join(OUT, s.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", *FLAGS, "-x", "c++", "-c", os.path.join(CSRC, s), "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emu build failed")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", LIB, *objs])
    open(stamp, "w").write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    import sys
    print(build_asan() if "--asan" in sys.argv else build(force=True))
