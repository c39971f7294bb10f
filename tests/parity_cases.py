"""Parity checks of the apb C ABI against the oracle, shared by the CPU-emulation tests (small
sizes, kernel logic) and the GPU tests (the parity tests proper).  Bit-exact everywhere."""
import os
import random

import numpy as np

from ark_plonk_b200 import encoding as enc
from ark_plonk_b200 import kzg, synth
from ark_plonk_b200.domain import Radix2EvaluationDomain
from oracle.curves import CURVES, powers_of_tau_g1
from oracle.fields import FQ, FR
from oracle.ntt import Domain, poly_eval
from oracle.serialize import ser_g1


class env:
    """temporarily set tuning environment variables (tile sizes, digit width)"""

    def __init__(self, **kw):
        self.kw = {k: str(v) for k, v in kw.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def check_field_ops(lib, count=64, seed=1):
    rnd = random.Random(seed)
    for fid, f, nl in ((0, FR[0], 4), (1, FQ[0], 6), (2, FR[1], 4), (3, FQ[1], 6)):
        a = [rnd.randrange(f.p) for _ in range(count)] + [0, 1, f.p - 1, f.p - 1, 2]
        b = [rnd.randrange(f.p) for _ in range(count)] + [5, f.p - 1, f.p - 1, 1, (f.p + 1) // 2]
        A, B = enc.ints_to_limbs(a, nl), enc.ints_to_limbs(b, nl)
        rinv = pow(f.R, -1, f.p)
        assert enc.limbs_to_ints(lib.field_op(fid, 0, A, B)) == [x * y * rinv % f.p for x, y in zip(a, b)], f.name
        assert enc.limbs_to_ints(lib.field_op(fid, 1, A, B)) == [(x + y) % f.p for x, y in zip(a, b)], f.name
        assert enc.limbs_to_ints(lib.field_op(fid, 2, A, B)) == [(x - y) % f.p for x, y in zip(a, b)], f.name
        assert enc.limbs_to_ints(lib.field_op(fid, 3, A, None)) == [x * f.R % f.p for x in a], f.name
        assert enc.limbs_to_ints(lib.field_op(fid, 4, A, None)) == [x * rinv % f.p for x in a], f.name
        assert enc.limbs_to_ints(lib.field_op(fid, 5, A, None)) == [x * x * rinv % f.p for x in a], f.name
        # Montgomery inverse (binary extended Euclid): a = xR -> x^-1 R = a^-1 R^2; 0 -> 0
        assert enc.limbs_to_ints(lib.field_op(fid, 6, A, None)) == \
            [pow(x, -1, f.p) * f.R * f.R % f.p if x else 0 for x in a], f.name


def check_ntt(lib, curve, log_n, in_len, seed=2, kinds=("fft", "ifft", "coset_fft", "coset_ifft")):
    """all four transforms against the oracle's arkworks-semantics radix-2 transform"""
    f = FR[curve]
    rnd = random.Random(seed * 1000 + log_n * 37 + in_len)
    d = Radix2EvaluationDomain(curve, 1 << log_n, lib=lib)
    try:
        assert d.size == 1 << log_n
        od = Domain(f, log_n)
        x = [rnd.randrange(f.p) for _ in range(in_len)]
        X = enc.fr_to_mont(curve, x)
        for name in kinds:
            got = getattr(d, name)(X)
            assert got.shape == (d.size, 4)
            assert enc.fr_from_mont(curve, got) == getattr(od, name)(x), (name, log_n, in_len)
    finally:
        d.close()


def check_ntt_properties(lib, curve, log_n, in_len, seed=3, spot=6):
    """size-independent checks for sizes the Python oracle cannot transform quickly:
    round trips are the identity and fft(x)[i] == p(w^i) at a few points (Horner)."""
    f = FR[curve]
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.integers(0, 1 << 64, size=(in_len, 4), dtype=np.uint64)
    X[:, 3] >>= np.uint64(4)                           # < 2^252 < r: valid (if arbitrary) Montgomery residues
    d = Radix2EvaluationDomain(curve, n, lib=lib)
    try:
        od = Domain(f, log_n)
        Y = d.fft(X)
        back = d.ifft(Y)
        assert np.array_equal(back[:in_len], X) and not back[in_len:].any()
        Yc = d.coset_fft(X)
        backc = d.coset_ifft(Yc)
        assert np.array_equal(backc[:in_len], X) and not backc[in_len:].any()
        x = enc.fr_from_mont(curve, X)
        rnd = random.Random(seed)
        idx = [0, 1, n - 1] + [rnd.randrange(n) for _ in range(spot)]
        ys = enc.fr_from_mont(curve, Y[idx])
        ycs = enc.fr_from_mont(curve, Yc[idx])
        for i, y, yc in zip(idx, ys, ycs):
            pt = pow(od.group_gen, i, f.p)
            assert y == poly_eval(f, x, pt), ("fft", log_n, i)
            assert yc == poly_eval(f, x, od.coset_gen * pt % f.p), ("coset_fft", log_n, i)
    finally:
        d.close()


def check_ntt_roundtrip(lib, curve, log_n, seed=6):
    """full-length round trips (numpy only): ifft(fft(x)) == x and coset_ifft(coset_fft(x)) == x"""
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    X[:, 3] >>= np.uint64(4)
    d = Radix2EvaluationDomain(curve, n, lib=lib)
    try:
        assert np.array_equal(d.ifft(d.fft(X)), X)
        assert np.array_equal(d.coset_ifft(d.coset_fft(X)), X)
        assert np.array_equal(d.fft(d.ifft(X)), X)
    finally:
        d.close()


def check_msm_tau(lib, curve, n, seed=4, scalars=None, offset=0, montgomery=False):
    """KZG identity: MSM(tau^i G, s) == [sum s_i tau^i] G, plus ark-serialize bytes."""
    cv = CURVES[curve]
    rnd = random.Random(seed * 7919 + n)
    tau = rnd.randrange(1, cv.fr.p)
    pts = powers_of_tau_g1(cv, tau, n + offset)
    ck = kzg.CommitterKey(curve, enc.g1_affine_to_mont(curve, pts), lib=lib)
    try:
        s = scalars if scalars is not None else [rnd.randrange(cv.fr.p) for _ in range(n)]
        S = enc.fr_to_mont(curve, s) if montgomery else enc.ints_to_limbs(s, 4)
        out = kzg.multi_scalar_mul(ck, S, base_offset=offset, montgomery=montgomery)
        e = sum(si * pow(tau, i + offset, cv.fr.p) for i, si in enumerate(s)) % cv.fr.p
        exp = cv.mul(cv.G, e)
        assert enc.g1_from_xyz(curve, out) == exp, (curve, n, offset)
        assert lib.g1_compress(curve, out) == ser_g1(cv, exp)
    finally:
        ck.close()


def check_msm_progression(lib, curve, n, seed=5, k=1):
    """large-size KAT with bases P_i = [a + i b]G (closed-form expected value); k > 1 checks the
    batched entry point (PC::commit of several polynomials) incl. ragged lengths and offsets"""
    cv = CURVES[curve]
    rnd = random.Random(seed)
    a, b = rnd.randrange(1, cv.fr.p), rnd.randrange(1, cv.fr.p)
    pts = synth.progression_bases(curve, a, b, n)
    ck = kzg.CommitterKey(curve, enc.g1_affine_to_mont(curve, pts), lib=lib)
    try:
        polys, expect = [], []
        for j in range(k):
            ln = n if j == 0 else max(1, n - 3 * j)
            S = synth.seeded_scalars(curve, ln, seed=b"msm%d" % (seed + j))
            if j % 2 == 1:
                S[: min(5, ln - 1)] = 0                      # low zero coefficients shift the base offset
            s = synth.limbs_to_int_list(S)
            polys.append(enc.fr_to_mont(curve, s))
            expect.append(synth.progression_expected(curve, a, b, s))
        outs = kzg.commit(ck, polys)
        for out, exp in zip(outs, expect):
            assert enc.g1_from_xyz(curve, out) == exp
            assert lib.g1_compress(curve, out) == ser_g1(cv, exp)
    finally:
        ck.close()


def check_msm_duplicates(lib, curve, seed=9):
    """equal points meeting in one bucket (doubling branch of the mixed addition) and P + (-P) (identity)"""
    cv = CURVES[curve]
    r = cv.fr.p
    rnd = random.Random(seed)
    base = [cv.mul(cv.G, rnd.randrange(1, r)) for _ in range(3)]
    pts = [base[0], base[0], base[1], base[1], base[2], base[0], None, base[1]]
    cases = [
        [5, 5, 7, r - 7, 1, 5, 9, 0],                 # 3 x the same point into one bucket; P + (-P)
        [12345] * 8,
        [r - 1, 1, r - 2, 2, 0, r - 1, 3, 4],
        [1 << 16, 1 << 16, (1 << 32) + 1, (1 << 32) + 1, 0, 1 << 16, 1, 1 << 32],
    ]
    ck = kzg.CommitterKey(curve, enc.g1_affine_to_mont(curve, pts), lib=lib)
    try:
        for s in cases:
            out = kzg.multi_scalar_mul(ck, enc.ints_to_limbs(s, 4))
            assert enc.g1_from_xyz(curve, out) == cv.msm_naive(pts, s), s
    finally:
        ck.close()


def edge_scalars(curve, n):
    r = FR[curve].p
    base = [0, 1, 2, r - 1, r - 2, (r - 1) // 2, (r + 1) // 2, 5, 0, 0, 0, 7, 1 << 200, (1 << 250) + 5, 12345,
            (1 << 16) - 1, 1 << 15, (1 << 15) + 1, (1 << 255) % r, r - (1 << 15)]
    return (base * (n // len(base) + 1))[:n]


def check_kzg_open_and_domain_helpers(lib, curve, log_n, seed=12):
    """PC::open against the oracle's sonic_pc restatement, and the EvaluationDomain helper calls"""
    from oracle.plonk import Kzg
    cv = CURVES[curve]
    p = cv.fr.p
    n = 1 << log_n
    rnd = random.Random(seed)
    tau = rnd.randrange(1, p)
    ck = kzg.CommitterKey.from_tau(curve, tau, n, lib=lib)
    try:
        polys = [[rnd.randrange(p) for _ in range(ln)] for ln in (n, n - 3, 1, 0, n)]
        z, ch = rnd.randrange(p), rnd.randrange(2, p)
        got = kzg.open(ck, [enc.fr_to_mont(curve, q) for q in polys], z, ch)
        exp, _ = Kzg(cv, tau, n).open(polys, z, ch)
        assert enc.g1_from_xyz(curve, got) == exp
    finally:
        ck.close()
    d = Radix2EvaluationDomain(curve, n, lib=lib)
    try:
        od = Domain(FR[curve], log_n)
        assert d.group_gen() == od.group_gen and d.size_inv() == od.size_inv and d.element(3) == od.element(3)
        assert enc.fr_from_mont(curve, d.elements()) == od.elements()
        t = rnd.randrange(p)
        assert d.evaluate_vanishing_polynomial(t) == od.evaluate_vanishing_polynomial(t)
        lag = d.evaluate_all_lagrange_coefficients(t)
        x = [rnd.randrange(p) for _ in range(n)]                     # sum_i L_i(t) f(w^i) == f(t) for deg f < n
        evals = od.fft(x)
        assert sum(l * e for l, e in zip(lag, evals)) % p == poly_eval(FR[curve], x, t)
        assert d.evaluate_all_lagrange_coefficients(od.element(5 % n))[5 % n] == 1
    finally:
        d.close()


def check_msm_pass_split(lib, curve=0, n=48, k=4, limit_split=4000, limit_fail=1000, seed=21, **envkw):
    """the bucket-entry list of one pass has 32-bit positions: a batch above the bound is split into
    several passes (same results), a single polynomial above it is refused (APB_MSM_MAX_ENTRIES lowers
    the bound so that small inputs reach both paths)"""
    from ark_plonk_b200._lib import ApbError
    cv = CURVES[curve]
    rnd = random.Random(seed)
    a, b = rnd.randrange(1, cv.fr.p), rnd.randrange(1, cv.fr.p)
    pts = synth.progression_bases(curve, a, b, n)
    ck = kzg.CommitterKey(curve, enc.g1_affine_to_mont(curve, pts), lib=lib)
    try:
        polys, expect = [], []
        for j in range(k):
            s = [rnd.randrange(cv.fr.p) for _ in range(n - j)]
            polys.append(enc.fr_to_mont(curve, s))
            expect.append(synth.progression_expected(curve, a, b, s))
        with env(APB_MSM_MAX_ENTRIES=limit_split, **envkw):
            outs = kzg.commit(ck, polys)
        for out, exp in zip(outs, expect):
            assert enc.g1_from_xyz(curve, out) == exp
        with env(APB_MSM_MAX_ENTRIES=limit_fail, **envkw):
            try:
                kzg.commit(ck, polys)
            except ApbError as e:
                assert e.code == 1
            else:
                raise AssertionError("a pass above the 32-bit entry bound must be refused")
    finally:
        ck.close()


def check_g1_fold(lib, curve, seed=3):
    """apb_g1_fold (folding the per-GPU partial sums of a split commit batch): groups of 0, 1 and several pieces,
    identity pieces, a group that cancels to the identity - against the oracle's group law"""
    cv = CURVES[curve]
    rnd = random.Random(seed)
    pts = [cv.mul(cv.G, rnd.randrange(1, 1 << 64)) for _ in range(6)]
    pts.append(cv.neg(pts[5]))                                                                    # cancels pts[5]
    pts.append(None)                                                                              # identity piece
    group = [0, 0, 0, 2, 2, 3, 3, 2]                                                              # group 1 stays empty
    one = enc.fq_to_mont(curve, [1])[0]
    xyz = np.zeros((len(pts), 18), dtype=np.uint64)
    for i, P in enumerate(pts):
        if P is not None:
            xyz[i, :12] = enc.g1_affine_to_mont(curve, [P])[0]
            xyz[i, 12:] = one
    out = lib.g1_fold(curve, xyz, group, 4)
    exp = [None] * 4
    for P, g in zip(pts, group):
        exp[g] = cv.add(exp[g], P)
    assert [enc.g1_from_xyz(curve, o) for o in out] == exp
    assert exp[1] is None and exp[3] is None
