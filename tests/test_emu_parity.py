"""Kernel-logic parity on the CPU: the product's kernel sources compiled against the CUDA
emulation shim (tests/emu), compared bit-exactly with the oracle.  Small sizes; the real
parity tests run the CUDA build on a B200 (test_gpu_parity.py)."""
import pytest

import parity_cases as pc


def test_field_ops(emu_lib):
    pc.check_field_ops(emu_lib, count=40)


@pytest.mark.parametrize("log_n,in_len,max_tile,log_cols", [
    (0, 1, 10, 4), (1, 2, 10, 4), (3, 5, 10, 4), (4, 16, 2, 1), (6, 64, 3, 2), (6, 17, 3, 1),
    (6, 0, 3, 1), (7, 100, 3, 2), (9, 300, 3, 2), (10, 256, 5, 3), (10, 1024, 10, 2),
])
def test_ntt_single_and_multi_pass(emu_lib, log_n, in_len, max_tile, log_cols):
    with pc.env(APB_NTT_MAX_LOG_TILE=max_tile, APB_NTT_LOG_COLS=log_cols):
        pc.check_ntt(emu_lib, 0, log_n, in_len)


def test_g1_fold(emu_lib):
    pc.check_g1_fold(emu_lib, 0)
    pc.check_g1_fold(emu_lib, 1, seed=4)


def test_ntt_bls12_377(emu_lib):
    with pc.env(APB_NTT_MAX_LOG_TILE=3, APB_NTT_LOG_COLS=2):
        pc.check_ntt(emu_lib, 1, 6, 40)
        pc.check_ntt_properties(emu_lib, 1, 7, 128)


@pytest.mark.parametrize("c,chunk,n,offset,mont", [(8, 8, 40, 0, False), (4, 3, 40, 0, False), (8, 8, 33, 3, True),
                                                    (8, 5, 1, 0, False)])
def test_msm_tau_identity(emu_lib, c, chunk, n, offset, mont):
    with pc.env(APB_MSM_C=c, APB_MSM_CHUNK=chunk):
        pc.check_msm_tau(emu_lib, 0, n, offset=offset, montgomery=mont)


def test_msm_edge_scalars(emu_lib):
    with pc.env(APB_MSM_C=8, APB_MSM_CHUNK=4):
        for scal in ([0] * 20, [1] * 20, [pc.FR[0].p - 1] * 20, pc.edge_scalars(0, 20)):
            pc.check_msm_tau(emu_lib, 0, 20, scalars=scal)


def test_msm_duplicate_points_and_cancellation(emu_lib):
    with pc.env(APB_MSM_C=8, APB_MSM_CHUNK=3):
        pc.check_msm_duplicates(emu_lib, 0)
        pc.check_msm_duplicates(emu_lib, 1)


def test_msm_bls12_377_and_batch(emu_lib):
    with pc.env(APB_MSM_C=8, APB_MSM_CHUNK=5):
        pc.check_msm_tau(emu_lib, 1, 30)
        pc.check_msm_progression(emu_lib, 0, 48, k=3)


def test_msm_batched_affine_pair_levels(emu_lib):
    """the batched-affine pair levels in front of the accumulate (narrow digits make buckets long enough
    for 3 levels at these sizes): random, edge scalars, equal points / P + (-P) / infinity bases inside
    a pair (doubling and cancellation branches), BLS12-377, batched commit, odd chunk sizes"""
    _pair_level_cases(emu_lib)


def _pair_level_cases(emu_lib):
    with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_C=4, APB_MSM_CHUNK=3):
        pc.check_msm_tau(emu_lib, 0, 40)
        pc.check_msm_tau(emu_lib, 0, 33, offset=3, montgomery=True)
        for scal in ([0] * 20, [1] * 20, [pc.FR[0].p - 1] * 20, pc.edge_scalars(0, 20)):
            pc.check_msm_tau(emu_lib, 0, 20, scalars=scal)
        pc.check_msm_tau(emu_lib, 1, 30)
        pc.check_msm_progression(emu_lib, 0, 48, k=3)
    with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_C=2, APB_MSM_CHUNK=5):
        pc.check_msm_duplicates(emu_lib, 0)
        pc.check_msm_duplicates(emu_lib, 1)
    with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_AFFINE_LEVELS=1, APB_MSM_C=4, APB_MSM_CHUNK=8):
        pc.check_msm_tau(emu_lib, 0, 40)
        pc.check_msm_duplicates(emu_lib, 0)


def test_msm_pair_levels_sliced_by_bucket_range(emu_lib):
    """a level workspace above the HBM budget: the stage runs over bucket ranges one after the other; a skewed
    digit distribution (every entry in one bucket) falls back to the plain accumulate"""
    import os
    seen = []
    with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_C=4, APB_MSM_CHUNK=3, APB_MSM_AFFINE_MAX_BYTES=190000):
        pc.check_msm_tau(emu_lib, 0, 40)
        pc.check_msm_tau(emu_lib, 1, 40)
        pc.check_msm_tau(emu_lib, 0, 40, scalars=[1] * 40)                     # unbalanced slices
        pc.check_msm_tau(emu_lib, 0, 40, scalars=pc.edge_scalars(0, 40))
        pc.check_msm_duplicates(emu_lib, 0)
        pc.check_msm_progression(emu_lib, 0, 48, k=2)
    with pc.env(APB_MSM_AFFINE_MIN=0, APB_MSM_C=8, APB_MSM_CHUNK=7, APB_MSM_AFFINE_MAX_BYTES=250000):
        same_digits = int("01" * 31, 16)                                       # every 8-bit digit is 1: two buckets hold everything
        pc.check_msm_tau(emu_lib, 0, 200, scalars=[same_digits] * 200)         # 8 slices, unbalanced -> fallback
        pc.check_msm_tau(emu_lib, 0, 200)                                      # 8 slices, balanced
    assert not seen and "APB_MSM_AFFINE_MAX_BYTES" not in os.environ


def test_msm_windowed_geometry(emu_lib):
    """step = 64 bits per precomputed copy -> 4 effective windows folded on the host"""
    with pc.env(APB_MSM_STEP=64, APB_MSM_C=8, APB_MSM_CHUNK=16):
        pc.check_msm_tau(emu_lib, 0, 24)


def test_msm_empty_and_errors(emu_lib):
    import numpy as np
    from ark_plonk_b200 import ApbError, encoding as enc, kzg, synth
    pts = synth.progression_bases(0, 3, 5, 4)
    ck = kzg.CommitterKey(0, enc.g1_affine_to_mont(0, pts), lib=emu_lib)
    out = kzg.multi_scalar_mul(ck, np.zeros((0, 4), dtype=np.uint64))
    assert enc.g1_from_xyz(0, out) is None and emu_lib.g1_compress(0, out)[-1] == 0x40
    zero_poly = np.zeros((3, 4), dtype=np.uint64)              # all-zero polynomial -> identity commitment
    assert enc.g1_from_xyz(0, kzg.commit(ck, [zero_poly])[0]) is None
    with pytest.raises(ApbError) as ei:
        kzg.commit(ck, [enc.fr_to_mont(0, [1, 2, 3, 4, 5])])
    assert ei.value.code == 2                                   # TooManyCoefficients
    ck.close()


def test_error_behaviour(emu_lib):
    """status codes mirror the reference's errors: Error::InvalidEvalDomainSize (prover.rs:169-173),
    TooManyCoefficients (kzg10 degree check), plus argument validation; messages are thread-local"""
    import ctypes as C
    import numpy as np
    from ark_plonk_b200 import ApbError, Radix2EvaluationDomain
    with pytest.raises(ApbError) as ei:
        Radix2EvaluationDomain(0, 1 << 33, lib=emu_lib)          # log_n 33 > two-adicity 32 of BLS12-381 Fr
    assert ei.value.code == 3
    h = C.c_void_p()
    assert emu_lib.c.apb_domain_new(7, 4, C.byref(h)) == 1          # bad curve id
    d = Radix2EvaluationDomain(0, 8, lib=emu_lib)
    with pytest.raises(ApbError) as ei:
        d.fft(np.zeros((9, 4), dtype=np.uint64))                    # more coefficients than the domain holds
    assert ei.value.code == 1
    out = np.zeros((8, 4), dtype=np.uint64)
    assert emu_lib.c.apb_ntt(None, 0, None, 0, out.ctypes.data) == 4 and b"handle" in emu_lib.c.apb_last_error()
    assert emu_lib.c.apb_ntt(d._h, 9, None, 0, out.ctypes.data) == 1  # bad transform kind
    z = d.fft(np.zeros((0, 4), dtype=np.uint64))                    # empty polynomial -> N zeros (coset_fft of q_range etc.)
    assert z.shape == (8, 4) and not z.any()
    d.close()


def test_msm_pass_split_at_the_entry_bound(emu_lib):
    # c = 8: 32 digit positions, 2 windows of 128 buckets per polynomial -> 48 x 32 + 256 entries each
    pc.check_msm_pass_split(emu_lib, 0, n=48, k=4, limit_split=4000, limit_fail=1000, APB_MSM_C=8, APB_MSM_CHUNK=5)
