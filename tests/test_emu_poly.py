"""Kernel-logic parity of the polynomial / prover kernels on the CPU emulation (small sizes)."""
import parity_cases as pc
import poly_cases


def test_lincomb_eval_divide(emu_lib):
    poly_cases.check_lincomb_eval_divide(emu_lib, 0, 300)
    poly_cases.check_lincomb_eval_divide(emu_lib, 1, 37, seed=5)
    poly_cases.check_lincomb_eval_divide(emu_lib, 0, 2100, seed=6)      # more than one 2048-coefficient eval block


def test_combine_split_matches_reference_semantics(emu_lib):
    poly_cases.check_combine_split(emu_lib, 0)
    poly_cases.check_combine_split(emu_lib, 1, seed=8)


def test_grand_products(emu_lib):
    with pc.env(APB_NTT_MAX_LOG_TILE=4):
        poly_cases.check_grand_products(emu_lib, 0, 5)
        poly_cases.check_grand_products(emu_lib, 1, 4, seed=4)
        poly_cases.check_grand_products(emu_lib, 0, 11, seed=7)         # several scan blocks


def test_kzg_open_and_domain_helpers(emu_lib):
    with pc.env(APB_MSM_C=8, APB_NTT_MAX_LOG_TILE=4):
        pc.check_kzg_open_and_domain_helpers(emu_lib, 0, 5)
        pc.check_kzg_open_and_domain_helpers(emu_lib, 1, 4, seed=13)


def test_quotient_range_matches_full(emu_lib):
    poly_cases.check_quotient_range(emu_lib, 0)
    poly_cases.check_quotient_range(emu_lib, 1, n4=32, seed=12)
