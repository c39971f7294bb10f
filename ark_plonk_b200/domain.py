"""Host-side mirror of ark_poly's Radix2EvaluationDomain over the apb C ABI.

Same method names and meaning as the reference's `EvaluationDomain` calls
(plonk-core/src/proof_system/prover.rs:197-203, quotient_poly.rs:72-120,176; ...): `fft`,
`ifft`, `coset_fft`, `coset_ifft` take <= size elements (zero-extended) and return `size`
elements in natural order.  Elements are (n, 4) uint64 Montgomery limb arrays (arkworks'
in-memory representation).  The twiddle tables live in HBM for the lifetime of the object.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import NTT_COSET_FFT, NTT_COSET_IFFT, NTT_FFT, NTT_IFFT, ApbError, Lib, get_lib


class Radix2EvaluationDomain:
    def __init__(self, curve: int, num_coeffs: int, lib: Lib | None = None):
        """`GeneralEvaluationDomain::new(num_coeffs)`: size = num_coeffs.next_power_of_two()."""
        self.lib = lib or get_lib()
        self.curve = curve
        self.log_size = max(num_coeffs - 1, 0).bit_length()
        self.size = 1 << self.log_size
        h = C.c_void_p()
        self.lib.check(self.lib.c.apb_domain_new(curve, self.log_size, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.c.apb_domain_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, kind: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        if x.shape[0] > self.size:
            raise ApbError(1, "input of %d elements exceeds domain size %d" % (x.shape[0], self.size))
        out = np.empty((self.size, 4), dtype=np.uint64)
        self.lib.check(self.lib.c.apb_ntt(self._h, kind, x.ctypes.data_as(C.c_void_p) if x.shape[0] else None,
                                          x.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def fft(self, coeffs):
        return self._run(NTT_FFT, coeffs)

    def ifft(self, evals):
        return self._run(NTT_IFFT, evals)

    def coset_fft(self, coeffs):
        return self._run(NTT_COSET_FFT, coeffs)

    def coset_ifft(self, evals):
        return self._run(NTT_COSET_IFFT, evals)

    # device-resident variants (pointers are ints / c_void_p of HBM buffers)
    def ntt_dev(self, kind: int, d_in, in_len: int, d_out, sync: bool = False):
        self.lib.check(self.lib.c.apb_ntt_dev(self._h, kind, d_in, in_len, d_out, 1 if sync else 0))

    def ntt_batch_dev(self, kind: int, d_in, in_len: int, in_stride: int, d_out, out_stride: int, batch: int,
                      sync: bool = False):
        self.lib.check(self.lib.c.apb_ntt_batch_dev(self._h, kind, d_in, in_len, in_stride, d_out, out_stride,
                                                    batch, 1 if sync else 0))
