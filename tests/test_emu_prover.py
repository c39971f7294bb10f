"""Kernel-logic check of the whole device-resident prover on the CPU emulation: byte-identical
to the oracle's golden proof at 2^5 (BLS12-381) - the GPU tests repeat this at 2^10..2^16."""
import parity_cases as pc
import prover_cases


def test_prover_matches_golden_2p5(emu_lib):
    with pc.env(APB_MSM_C=8, APB_NTT_MAX_LOG_TILE=4):
        prover_cases.prove_case(emu_lib, prover_cases.golden_case(0, 5))


import gadget_cases  # noqa: E402
import pytest  # noqa: E402


@pytest.mark.parametrize("curve,kind", [(0, "range"), (1, "logic"), (0, "curve_add"), (1, "fixed_base")])
def test_prover_custom_gates_match_oracle(emu_lib, curve, kind):
    """range / logic / curve-addition / fixed-base gate terms and public inputs (quotient_poly.rs:231-264,
    linearisation_poly.rs:382-410): the proof equals the oracle's byte for byte and verifies"""
    gadget_cases.prove_gadget_case(emu_lib, curve, kind)


def test_one_key_two_public_input_assignments(emu_lib):
    gadget_cases.prove_two_public_input_assignments(emu_lib, 0)
