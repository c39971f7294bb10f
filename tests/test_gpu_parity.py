"""Parity of the CUDA path (through the C ABI) against the oracle on a real B200. Bit-exact."""
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_field_ops(gpu_lib):
    pc.check_field_ops(gpu_lib, count=4096)


@pytest.mark.parametrize("log_n,in_len", [(0, 1), (1, 2), (5, 20), (10, 1024), (10, 0), (12, 1024), (13, 5000), (14, 1 << 14)])
def test_ntt_vs_oracle(gpu_lib, log_n, in_len):
    pc.check_ntt(gpu_lib, 0, log_n, in_len)


def test_ntt_vs_oracle_377(gpu_lib):
    pc.check_ntt(gpu_lib, 1, 12, 3000)


@pytest.mark.parametrize("max_tile,log_cols", [(4, 2), (6, 4), (7, 3)])
def test_ntt_three_pass_small(gpu_lib, max_tile, log_cols):
    with pc.env(APB_NTT_MAX_LOG_TILE=max_tile, APB_NTT_LOG_COLS=log_cols):
        pc.check_ntt(gpu_lib, 0, 12, 4096)


@pytest.mark.parametrize("curve,log_n,in_len", [(0, 16, 1 << 16), (0, 18, 1 << 18), (0, 20, 1 << 18), (0, 20, 1 << 20),
                                                 (1, 18, 1 << 16), (0, 22, 1 << 19), (0, 24, 1 << 18)])
def test_ntt_properties_large(gpu_lib, curve, log_n, in_len):
    """round trips + Horner spot checks of fft / coset_fft outputs (zero-extended inputs)"""
    pc.check_ntt_properties(gpu_lib, curve, log_n, in_len, spot=2 if log_n >= 20 else 6)


@pytest.mark.parametrize("curve,log_n", [(0, 20), (0, 22), (1, 21), (0, 24)])
def test_ntt_roundtrip_full(gpu_lib, curve, log_n):
    pc.check_ntt_roundtrip(gpu_lib, curve, log_n)


@pytest.mark.parametrize("curve,n,offset,mont", [(0, 1, 0, False), (0, 31, 0, False), (0, 1000, 0, True), (0, 4096, 5, False),
                                                  (0, 5000, 0, True), (1, 3000, 0, False)])
def test_msm_tau_identity(gpu_lib, curve, n, offset, mont):
    pc.check_msm_tau(gpu_lib, curve, n, offset=offset, montgomery=mont)


def test_msm_edge_scalars(gpu_lib):
    for n in (20, 5000):
        for scal in ([0] * n, [1] * n, [pc.FR[0].p - 1] * n, pc.edge_scalars(0, n)):
            pc.check_msm_tau(gpu_lib, 0, n, scalars=scal)


@pytest.mark.parametrize("curve,n,k", [(0, 1 << 14, 1), (0, 1 << 16, 4), (1, 1 << 15, 2), (0, (1 << 18) + 1, 1)])
def test_msm_progression_large(gpu_lib, curve, n, k):
    pc.check_msm_progression(gpu_lib, curve, n, k=k)


def test_msm_windowed_geometry(gpu_lib):
    with pc.env(APB_MSM_STEP=64):
        pc.check_msm_tau(gpu_lib, 0, 3000)
        pc.check_msm_progression(gpu_lib, 0, 1 << 14)
