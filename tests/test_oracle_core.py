"""Pins the CPU oracle: constants, Merlin conformance vector, and the
mathematical identities every hot-path result must satisfy."""
import random

from oracle.curves import BLS12_377, BLS12_381, powers_of_tau_g1
from oracle.fields import FQ377, FQ381, FR377, FR381
from oracle.merlin import Transcript
from oracle.ntt import Domain, poly_eval
from oracle.serialize import deser_g1, ser_g1


def _is_probable_prime(n, rounds=16):
    if n % 2 == 0:
        return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rnd = random.Random(1)
    for _ in range(rounds):
        a = rnd.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def test_moduli_are_prime_and_sized():
    assert FR381.bits == 255 and FQ381.bits == 381 and FR377.bits == 253 and FQ377.bits == 377
    for f in (FR381, FQ381, FR377, FQ377):
        assert _is_probable_prime(f.p)


def test_two_adicity_and_generators():
    for f in (FR381, FR377):
        assert (f.p - 1) % (1 << f.two_adicity) == 0 and ((f.p - 1) >> f.two_adicity) % 2 == 1
        assert pow(f.generator, (f.p - 1) // 2, f.p) == f.p - 1          # non-residue
        root = f.two_adic_root()
        assert pow(root, 1 << f.two_adicity, f.p) == 1 and pow(root, 1 << (f.two_adicity - 1), f.p) == f.p - 1
    assert FR381.two_adic_root() == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert FR377.two_adic_root() == 8065159656716812877374967518403273466521432693661810619979959746626482506078


def test_generators_on_curve_and_order():
    for c in (BLS12_381, BLS12_377):
        assert c.on_curve(c.G)
        assert c.mul(c.G, c.fr.p - 1) == c.neg(c.G)
        assert c.jadd(c.to_jac(c.mul(c.G, c.fr.p - 1)), c.to_jac(c.G))[2] == 0   # r*G = identity


def test_merlin_conformance_vector():
    t = Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_ntt_matches_definition():
    rnd = random.Random(3)
    for f in (FR381, FR377):
        for log_n, in_len in ((0, 1), (1, 2), (4, 16), (5, 11), (6, 64)):
            d = Domain(f, log_n)
            x = [rnd.randrange(f.p) for _ in range(in_len)]
            assert d.fft(x) == d.dft_naive(x)
            assert d.coset_fft(x) == d.dft_naive(x, coset=True)
            assert d.ifft(d.fft(x))[:in_len] == x
            assert d.coset_ifft(d.coset_fft(x))[:in_len] == x
            # evaluation semantics: fft(x)[i] = p(w^i)
            assert d.fft(x)[-1] == poly_eval(f, x, d.element(d.size - 1))


def test_msm_pippenger_matches_naive_and_tau_identity():
    rnd = random.Random(4)
    for c in (BLS12_381, BLS12_377):
        tau = rnd.randrange(c.fr.p)
        pts = powers_of_tau_g1(c, tau, 40)
        sc = [rnd.randrange(c.fr.p) for _ in range(40)]
        sc[0], sc[1], sc[2] = 0, 1, c.fr.p - 1
        a = c.msm_naive(pts, sc)
        assert a == c.msm_pippenger(pts, sc)
        e = sum(s * pow(tau, i, c.fr.p) for i, s in enumerate(sc)) % c.fr.p
        assert a == c.mul(c.G, e)                 # commit(p) == [p(tau)]G


def test_g1_serialization_roundtrip_and_flags():
    rnd = random.Random(5)
    for c in (BLS12_381, BLS12_377):
        assert ser_g1(c, None)[-1] == 0x40 and deser_g1(c, ser_g1(c, None)) is None
        for _ in range(6):
            P = c.mul(c.G, rnd.randrange(1, c.fr.p))
            b = ser_g1(c, P)
            assert len(b) == 48 and deser_g1(c, b) == P
            assert deser_g1(c, ser_g1(c, c.neg(P))) == c.neg(P)
            assert (b[-1] & 0x80 != 0) != (ser_g1(c, c.neg(P))[-1] & 0x80 != 0)
