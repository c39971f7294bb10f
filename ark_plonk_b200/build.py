"""Build ark_plonk_b200/lib/libapb.so (the C-ABI product library) with nvcc for sm_100a.

Explicit nvcc invocation, in-tree output (the .so travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libapb.so")
SOURCES = ["api.cu", "ntt.cu", "msm.cu", "msm_acc.cu", "msm_pairs_coop.cu", "msm_setup.cu", "poly.cu", "transcript.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]


def _digest() -> str:
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + fh.read())
    with open(os.path.join(HERE, "..", "include", "apb.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "libapb.stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    objs = []
    procs = []
    hdr = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            hdr.update(f.encode() + open(os.path.join(CSRC, f), "rb").read())
    hdr.update(open(os.path.join(HERE, "..", "include", "apb.h"), "rb").read() + " ".join(NVCC_FLAGS).encode())
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        ostamp = obj + ".stamp"
        odig = hashlib.sha256(hdr.digest() + open(os.path.join(CSRC, src), "rb").read()).hexdigest()
        if not force and not verbose and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == odig:
            continue                      # this translation unit and every header are unchanged
        cmd = ["nvcc", *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), ostamp, odig))
    for src, p, ostamp, odig in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
        with open(ostamp, "w") as fh:
            fh.write(odig)
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs])
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
