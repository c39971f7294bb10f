"""The C-ABI library loads and exports every symbol include/apb.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "apb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ("apb_init", "apb_ck_upload", "apb_msm", "apb_msm_batch", "apb_domain_new", "apb_ntt", "apb_g1_compress"):
        assert s in syms


def test_product_library_exports_all_symbols():
    from ark_plonk_b200 import build as apb_build
    lib = ctypes.CDLL(apb_build.build())
    for s in declared_symbols():
        assert hasattr(lib, s), "libapb.so does not export %s" % s
    assert b"sm_100a" in ctypes.c_char_p(ctypes.cast(lib.apb_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()).value


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ark_plonk_b200 import ApbError, Radix2EvaluationDomain
    with pytest.raises(ApbError) as ei:
        Radix2EvaluationDomain(0, 8)
    assert ei.value.code == 5
