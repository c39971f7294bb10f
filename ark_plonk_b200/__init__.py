"""ark_plonk_b200: B200-native (sm_100a) proving hot path of heliaxdev/ark-plonk.

KZG10 commitment MSM over BLS12-381 / BLS12-377 G1 and radix-2 Fr NTTs behind the apb C ABI
(include/apb.h).  CUDA only - there is no CPU fallback.
"""
from ._lib import (CURVE_BLS12_377, CURVE_BLS12_381, NTT_COSET_FFT, NTT_COSET_IFFT, NTT_FFT, NTT_IFFT, ApbError,
                   Lib, get_lib)
from .domain import Radix2EvaluationDomain
from .kzg import CommitterKey, commit, compress, multi_scalar_mul, open as open_proof, to_affine

__all__ = [
    "CURVE_BLS12_381", "CURVE_BLS12_377", "NTT_FFT", "NTT_IFFT", "NTT_COSET_FFT", "NTT_COSET_IFFT",
    "ApbError", "Lib", "get_lib", "Radix2EvaluationDomain", "CommitterKey", "commit", "compress",
    "multi_scalar_mul", "open_proof", "to_affine",
]
