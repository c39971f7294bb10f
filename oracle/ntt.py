"""Radix-2 evaluation domains over Fr (ark-poly 0.3 `Radix2EvaluationDomain` semantics).

Oracle = test infrastructure (see oracle/__init__.py).

plonk-core reaches these through `GeneralEvaluationDomain` at
plonk-core/src/proof_system/prover.rs:197-203,241,282,303,305,
quotient_poly.rs:72-120,176,205,294,325, permutation/mod.rs:199-205,671-674,751,800,
proof_system/pi.rs:115, lookup/multiset.rs:201, preprocess.rs:145-210,304-340.

Semantics:
  fft(x)[i]        = sum_k x[k] w^(ik)            (input zero-extended to N, natural order out)
  ifft(y)[k]       = N^-1 sum_i y[i] w^(-ik)
  coset_fft(x)[i]  = sum_k x[k] g^k w^(ik),  g = Fr multiplicative generator (7 / 22)
  coset_ifft(y)[k] = g^-k * ifft(y)[k]
"""
from __future__ import annotations

from .fields import Field


class Domain:
    def __init__(self, field: Field, log_n: int):
        if log_n > field.two_adicity:
            raise ValueError("domain larger than 2^TWO_ADICITY")
        self.f = field
        self.log_n = log_n
        self.size = 1 << log_n
        p = field.p
        self.group_gen = pow(field.two_adic_root(), 1 << (field.two_adicity - log_n), p)
        self.group_gen_inv = pow(self.group_gen, -1, p)
        self.size_inv = pow(self.size, -1, p)
        self.coset_gen = field.generator
        self.coset_gen_inv = pow(field.generator, -1, p)

    @classmethod
    def for_size(cls, field: Field, m: int) -> "Domain":
        """GeneralEvaluationDomain::new(m): size = m.next_power_of_two()."""
        log_n = max(m - 1, 0).bit_length()
        return cls(field, log_n)

    # ---- helpers used by plonk-core (util.rs:44-89) -------------------------
    def element(self, i: int) -> int:
        return pow(self.group_gen, i, self.f.p)

    def elements(self):
        p, w, cur = self.f.p, self.group_gen, 1
        out = []
        for _ in range(self.size):
            out.append(cur)
            cur = cur * w % p
        return out

    def evaluate_vanishing_polynomial(self, tau: int) -> int:
        return (pow(tau, self.size, self.f.p) - 1) % self.f.p

    # ---- transforms ---------------------------------------------------------
    def _transform(self, a, root):
        """In-place iterative radix-2 DIT NTT (bit-reverse then butterflies)."""
        n, p = self.size, self.f.p
        a = list(a)
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            wlen = pow(root, n // length, p)
            half = length >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * wlen % p
            for start in range(0, n, length):
                for k in range(half):
                    u = a[start + k]
                    v = a[start + k + half] * tw[k] % p
                    a[start + k] = (u + v) % p
                    a[start + k + half] = (u - v) % p
            length <<= 1
        return a

    def _pad(self, x):
        x = [v % self.f.p for v in x]
        if len(x) > self.size:
            raise ValueError("input longer than domain")
        return x + [0] * (self.size - len(x))

    def fft(self, coeffs):
        return self._transform(self._pad(coeffs), self.group_gen)

    def ifft(self, evals):
        p, ninv = self.f.p, self.size_inv
        return [v * ninv % p for v in self._transform(self._pad(evals), self.group_gen_inv)]

    def coset_fft(self, coeffs):
        p, g = self.f.p, self.coset_gen
        out, cur = [], 1
        for c in coeffs:
            out.append(c * cur % p)
            cur = cur * g % p
        return self.fft(out)

    def coset_ifft(self, evals):
        p, gi = self.f.p, self.coset_gen_inv
        res = self.ifft(evals)
        cur = 1
        for i in range(len(res)):
            res[i] = res[i] * cur % p
            cur = cur * gi % p
        return res

    def dft_naive(self, coeffs, coset: bool = False):
        """O(N^2) definition, for pinning the fast path at small N."""
        p, n, w = self.f.p, self.size, self.group_gen
        x = self._pad(coeffs)
        shift = self.coset_gen if coset else 1
        out = []
        for i in range(n):
            pt = shift * pow(w, i, p) % p
            acc = 0
            for c in reversed(x):
                acc = (acc * pt + c) % p
            out.append(acc)
        return out


def poly_eval(field: Field, coeffs, x: int) -> int:
    acc = 0
    p = field.p
    for c in reversed(coeffs):
        acc = (acc * x + c) % p
    return acc
