#!/bin/bash
# Runs small MSM / NTT / prove cases of the kernel sources under AddressSanitizer on the CPU
# emulation (compute-sanitizer is closed on the GPU pool).  Usage: bash tools/emu_asan.sh
set -e
cd "$(dirname "$0")/.."
python tests/emu/build_emu.py --asan
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1
LD_PRELOAD=$(gcc -print-file-name=libasan.so) python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from ark_plonk_b200._lib import Lib
lib = Lib("tests/emu/_build/libapb_emu_asan.so"); lib.init()
import parity_cases as pc, poly_cases, prover_cases
with pc.env(APB_MSM_C=8, APB_MSM_CHUNK=5, APB_NTT_MAX_LOG_TILE=3, APB_NTT_LOG_COLS=1):
    pc.check_ntt(lib, 0, 7, 100); pc.check_ntt(lib, 0, 0, 1); pc.check_ntt(lib, 1, 5, 0)
    pc.check_msm_tau(lib, 0, 33, offset=3, montgomery=True); pc.check_msm_duplicates(lib, 0)
    pc.check_msm_progression(lib, 0, 48, k=3)
    poly_cases.check_lincomb_eval_divide(lib, 0, 300); poly_cases.check_combine_split(lib, 0)
    poly_cases.check_grand_products(lib, 0, 5)
with pc.env(APB_MSM_C=8, APB_NTT_MAX_LOG_TILE=4):
    prover_cases.prove_case(lib, prover_cases.golden_case(0, 5))
print("ASAN clean")
PY
