"""Validates the C restatement (oracle/c, the CPU baseline) against the Python big-int oracle."""
import random

from ark_plonk_b200 import encoding as enc
from oracle import cbuild
from oracle.curves import CURVES, powers_of_tau_g1
from oracle.fields import FR
from oracle.ntt import Domain


def test_c_ntt_matches_python_oracle():
    rnd = random.Random(3)
    for curve in (0, 1):
        f = FR[curve]
        for log_n, in_len in ((0, 1), (3, 5), (6, 64), (8, 100)):
            x = [rnd.randrange(f.p) for _ in range(in_len)]
            X = enc.fr_to_mont(curve, x)
            od = Domain(f, log_n)
            for kind, name in enumerate(("fft", "ifft", "coset_fft", "coset_ifft")):
                got = enc.fr_from_mont(curve, cbuild.ntt(curve, kind, X, log_n, threads=3))
                assert got == getattr(od, name)(x), (curve, log_n, name)


def test_c_msm_matches_tau_identity():
    rnd = random.Random(4)
    for curve in (0, 1):
        cv = CURVES[curve]
        for n in (1, 5, 40, 300):
            tau = rnd.randrange(cv.fr.p)
            pts = powers_of_tau_g1(cv, tau, n)
            s = [rnd.randrange(cv.fr.p) for _ in range(n)]
            if n > 3:
                s[0], s[1], s[2] = 0, 1, cv.fr.p - 1
            out = cbuild.msm(curve, enc.g1_affine_to_mont(curve, pts), enc.ints_to_limbs(s, 4), threads=4)
            e = sum(si * pow(tau, i, cv.fr.p) for i, si in enumerate(s)) % cv.fr.p
            exp = cv.mul(cv.G, e)
            got = None if out is None else tuple(enc.fq_from_mont(curve, out.reshape(2, 6)))
            assert got == exp, (curve, n)
            assert exp == cv.msm_pippenger(pts, s)


def test_c_horner_matches_python_oracle():
    from oracle.ntt import poly_eval
    rnd = random.Random(5)
    for curve in (0, 1):
        f = FR[curve]
        for n, threads in ((0, 1), (1, 1), (100, 1), (5000, 3), (10000, 8)):
            c = [rnd.randrange(f.p) for _ in range(n)]
            x = rnd.randrange(f.p)
            got = cbuild.fr_horner(curve, enc.fr_to_mont(curve, c) if n else [], enc.fr_to_mont(curve, [x])[0], threads=threads)
            assert enc.fr_from_mont(curve, got.reshape(1, 4)) == [poly_eval(f, c, x)], (curve, n)
