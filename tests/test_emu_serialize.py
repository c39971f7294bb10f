"""ark-serialize wire formats of the product (ark_plonk_b200/serialize.py) on the CPU emulation build."""
import parity_cases as pc
import prover_cases
import serialize_cases as sc


def test_points_scalars_and_proof_bytes():
    for curve in (0, 1):
        sc.check_points_and_scalars(curve)
    sc.check_proof_roundtrip(prover_cases.golden_case(0, 5))
    sc.check_proof_roundtrip(prover_cases.golden_case(1, 6))


def test_srs_powers_roundtrip(emu_lib):
    with pc.env(APB_MSM_C=8):
        sc.check_srs_roundtrip(emu_lib, 0, n=12)
        sc.check_srs_roundtrip(emu_lib, 1, n=9)


def test_verifier_and_prover_key_roundtrip(emu_lib):
    with pc.env(APB_MSM_C=8, APB_NTT_MAX_LOG_TILE=4):
        sc.check_keys_roundtrip(emu_lib, 0, 5)


def test_key_roundtrip_with_custom_gates_and_public_inputs(emu_lib):
    with pc.env(APB_MSM_C=8, APB_NTT_MAX_LOG_TILE=4):
        sc.check_keys_roundtrip(emu_lib, 1, kind="logic")
