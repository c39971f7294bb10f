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

from . import encoding as enc
from ._lib import NTT_COSET_FFT, NTT_COSET_IFFT, NTT_FFT, NTT_IFFT, ApbError, Lib, get_lib

FR_GENERATOR = (7, 22)            # Fr::multiplicative_generator() of BLS12-381 / BLS12-377
FR_TWO_ADICITY = (32, 47)


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

    # ---- domain constants and helper evaluations (host side, Python ints): the calls plonk-core makes at
    # prover.rs:169,613, preprocess.rs:284,443, permutation/mod.rs:144,692, proof_system/permutation.rs:181,292,
    # linearisation_poly.rs:268, util.rs:44-89 ------------------------------------------------------------
    @property
    def modulus(self) -> int:
        return enc.FR_MODULUS[self.curve]

    def group_gen(self) -> int:
        p = self.modulus
        adic = FR_TWO_ADICITY[self.curve]
        root = pow(FR_GENERATOR[self.curve], (p - 1) >> adic, p)
        return pow(root, 1 << (adic - self.log_size), p)

    def group_gen_inv(self) -> int:
        return pow(self.group_gen(), -1, self.modulus)

    def size_inv(self) -> int:
        return pow(self.size, -1, self.modulus)

    def element(self, i: int) -> int:
        return pow(self.group_gen(), i, self.modulus)

    def elements(self) -> np.ndarray:
        """all domain elements w^i as (size, 4) Montgomery limbs = fft of the polynomial X (device)"""
        x = enc.fr_to_mont(self.curve, [0, 1] if self.size > 1 else [1])
        return self.fft(x)

    def evaluate_vanishing_polynomial(self, tau: int) -> int:
        return (pow(tau, self.size, self.modulus) - 1) % self.modulus

    def evaluate_all_lagrange_coefficients(self, tau: int) -> list:
        """L_i(tau) for all i (ark_poly semantics incl. the tau-in-domain case)"""
        p, n, w = self.modulus, self.size, self.group_gen()
        z_h = self.evaluate_vanishing_polynomial(tau)
        if z_h == 0:
            out, cur = [0] * n, 1
            for i in range(n):
                if cur == tau % p:
                    out[i] = 1
                    break
                cur = cur * w % p
            return out
        # L_i(tau) = z_h * w^i / (n * (tau - w^i)), denominators inverted in one batch
        roots, cur = [], 1
        for _ in range(n):
            roots.append(cur)
            cur = cur * w % p
        dens = [n * (tau - r) % p for r in roots]
        prefix, acc = [], 1
        for d in dens:
            prefix.append(acc)
            acc = acc * d % p
        inv = pow(acc, -1, p)
        out = [0] * n
        for i in range(n - 1, -1, -1):
            out[i] = z_h * roots[i] % p * (inv * prefix[i] % p) % p
            inv = inv * dens[i] % p
        return out

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
