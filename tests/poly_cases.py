"""Unit parity of the device-resident polynomial / prover kernels (apb_fr_lincomb, apb_poly_eval,
apb_poly_divide_linear, apb_plonk_*) against their Python definitions - shared by emu and GPU tests."""
import ctypes as C
import random

import numpy as np

from ark_plonk_b200 import encoding as enc
from ark_plonk_b200.domain import Radix2EvaluationDomain
from ark_plonk_b200.plonk import Arena
from oracle import plonk as op
from oracle.fields import FR
from oracle.ntt import Domain, poly_eval


def _up(arena, vals, curve):
    off = arena.alloc(max(len(vals), 1))
    if vals:
        arena.upload(off, enc.fr_to_mont(curve, vals))
    return off


def _down(arena, off, n, curve):
    return enc.fr_from_mont(curve, arena.download(off, n))


def _m(curve, v):
    return np.ascontiguousarray(enc.fr_to_mont(curve, [v]))


def check_lincomb_eval_divide(lib, curve, n, seed=1):
    f = FR[curve]
    p = f.p
    rnd = random.Random(seed)
    arena = Arena(lib, 40 * n + 64)
    try:
        k = 19                                            # > 16: exercises the accumulate pass
        lens = [rnd.randrange(0, n + 1) for _ in range(k)]
        lens[0], lens[1] = n, 0
        polys = [[rnd.randrange(p) for _ in range(ln)] for ln in lens]
        offs = [_up(arena, q, curve) for q in polys]
        sc = [rnd.randrange(p) for _ in range(k)]
        out = arena.alloc(n)
        ptrs = (C.c_void_p * k)(*[arena.ptr(o) for o in offs])
        ln = (C.c_size_t * k)(*lens)
        scm = np.ascontiguousarray(enc.fr_to_mont(curve, sc))
        lib.check(lib.c.apb_fr_lincomb(curve, k, ptrs, ln, scm.ctypes.data, arena.ptr(out), n))
        exp = [sum(s * (q[i] if i < len(q) else 0) for s, q in zip(sc, polys)) % p for i in range(n)]
        assert _down(arena, out, n, curve) == exp
        # evaluations (DensePolynomial::evaluate) at several points, ragged lengths incl. empty
        pts = [rnd.randrange(p) for _ in range(k - 2)] + [0, 1]
        ptm = np.ascontiguousarray(enc.fr_to_mont(curve, pts))
        vals = np.zeros((k, 4), dtype=np.uint64)
        lib.check(lib.c.apb_poly_eval(curve, k, ptrs, ln, ptm.ctypes.data, vals.ctypes.data))
        assert enc.fr_from_mont(curve, vals) == [poly_eval(f, q, x) for q, x in zip(polys, pts)]
        # witness polynomial p / (X - z)
        for z in (rnd.randrange(p), 0, 1):
            q = polys[0]
            w = arena.alloc(n)
            zm = _m(curve, z)
            lib.check(lib.c.apb_poly_divide_linear(curve, arena.ptr(offs[0]), n, zm.ctypes.data, arena.ptr(w)))
            got = _down(arena, w, n - 1, curve)
            acc, expw = 0, [0] * (n - 1)
            for i in range(n - 1, 0, -1):
                acc = (q[i] + acc * z) % p
                expw[i - 1] = acc
            assert got == expw, z
    finally:
        arena.close()


def check_combine_split(lib, curve, seed=2):
    p = FR[curve].p
    rnd = random.Random(seed)
    cases = [([2, 4, 1, 3], [2, 3, 3, 2])]                 # lookup/multiset.rs:119-123 / test_combine_split
    for n in (8, 64, 256):
        distinct = [rnd.randrange(p) for _ in range(rnd.randrange(1, n))]
        t = [rnd.choice(distinct) for _ in range(n)]
        f = [rnd.choice(t) for _ in range(n)]
        cases.append((t, f))
    cases.append(([7] * 16, [7] * 16))
    for t, f in cases:
        n = len(t)
        arena = Arena(lib, 8 * n + 64)
        try:
            ot, of = _up(arena, t, curve), _up(arena, f, curve)
            h1, h2 = arena.alloc(n), arena.alloc(n)
            lib.check(lib.c.apb_plonk_combine_split(curve, arena.ptr(ot), arena.ptr(of), n, arena.ptr(h1), arena.ptr(h2)))
            e1, e2 = op.combine_split([v % p for v in t], [v % p for v in f])
            assert _down(arena, h1, n, curve) == e1 and _down(arena, h2, n, curve) == e2
        finally:
            arena.close()
    # an element of f that is not in t -> Error::ElementNotIndexed
    arena = Arena(lib, 64)
    try:
        ot, of = _up(arena, [1, 2, 3, 4], curve), _up(arena, [1, 2, 3, 5], curve)
        h1, h2 = arena.alloc(4), arena.alloc(4)
        rc = lib.c.apb_plonk_combine_split(curve, arena.ptr(ot), arena.ptr(of), 4, arena.ptr(h1), arena.ptr(h2))
        assert rc == 1 and b"ElementNotIndexed" in lib.c.apb_last_error()
    finally:
        arena.close()


def check_grand_products(lib, curve, log_n, seed=3):
    """z and z2 evaluations against permutation/mod.rs:652-822 restated; plus the reference's own properties
    (permutation/mod.rs:1243-1380 test_correct_permutation_poly): z[0] = 1 and the product closes to 1 for a
    valid permutation"""
    f = FR[curve]
    p = f.p
    n = 1 << log_n
    rnd = random.Random(seed)
    dom = Radix2EvaluationDomain(curve, n, lib=lib)
    od = Domain(f, log_n)
    arena = Arena(lib, 24 * n + 64)
    try:
        roots = od.elements()
        ks = (1, op.K1, op.K2, op.K3)
        # a random permutation of the 4n wire slots, wire values constant on each cycle
        slots = [(c, i) for c in range(4) for i in range(n)]
        perm = slots[:]
        rnd.shuffle(perm)
        sigma = {s: t for s, t in zip(slots, perm)}
        val = {}
        for s in slots:
            if s in val:
                continue
            v, cur = rnd.randrange(p), s
            while cur not in val:
                val[cur] = v
                cur = sigma[cur]
        wires = [[val[(c, i)] for i in range(n)] for c in range(4)]
        sig = [[ks[sigma[(c, i)][0]] * roots[sigma[(c, i)][1]] % p for i in range(n)] for c in range(4)]
        beta, gamma = rnd.randrange(p), rnd.randrange(p)
        w_off = [_up(arena, w, curve) for w in wires]
        s_off = [_up(arena, s, curve) for s in sig]
        z_off = arena.alloc(n)
        wp = (C.c_void_p * 4)(*[arena.ptr(o) for o in w_off])
        sp = (C.c_void_p * 4)(*[arena.ptr(o) for o in s_off])
        bm, gm = _m(curve, beta), _m(curve, gamma)
        lib.check(lib.c.apb_plonk_perm_z(dom._h, wp, sp, bm.ctypes.data, gm.ctypes.data, arena.ptr(z_off)))
        z = [1]
        ratios = []
        for i in range(n):
            num = den = 1
            for c in range(4):
                num = num * (wires[c][i] + beta * ks[c] * roots[i] + gamma) % p
                den = den * (wires[c][i] + beta * sig[c][i] + gamma) % p
            ratios.append(num * pow(den, -1, p) % p)
        for i in range(n - 1):
            z.append(z[-1] * ratios[i] % p)
        got = _down(arena, z_off, n, curve)
        assert got == z and got[0] == 1
        assert got[-1] * ratios[-1] % p == 1              # grand product of a valid permutation closes
        # lookup grand product
        t = [rnd.randrange(p) for _ in range(n)]
        fq = [rnd.choice(t) for _ in range(n)]
        h1, h2 = op.combine_split(t, fq)
        delta, eps = rnd.randrange(p), rnd.randrange(p)
        offs = [_up(arena, v, curve) for v in (fq, t, h1, h2)]
        z2_off = arena.alloc(n)
        dm, em = _m(curve, delta), _m(curve, eps)
        lib.check(lib.c.apb_plonk_lookup_z2(dom._h, arena.ptr(offs[0]), arena.ptr(offs[1]), arena.ptr(offs[2]), arena.ptr(offs[3]),
                                            dm.ctypes.data, em.ctypes.data, arena.ptr(z2_off)))
        opd = (1 + delta) % p
        eopd = eps * opd % p
        z2 = [1]
        for i in range(n - 1):
            num = opd * (eps + fq[i]) % p * (eopd + t[i] + delta * t[(i + 1) % n]) % p
            den = (eopd + h1[i] + h2[i] * delta) % p * ((eopd + h2[i] + h1[(i + 1) % n] * delta) % p) % p
            z2.append(z2[-1] * num % p * pow(den, -1, p) % p)
        assert _down(arena, z2_off, n, curve) == z2
    finally:
        arena.close()
        dom.close()


def check_quotient_range(lib, curve, n4=64, seed=11):
    """apb_plonk_quotient_range over ragged slices == apb_plonk_quotient_full over the whole 4n coset (the multi-GPU
    proof evaluates one slice per rank): random input vectors, all custom gate selectors present, public inputs;
    also the argument checks of the range entry point"""
    p = FR[curve].p
    rnd = random.Random(seed)
    arena = Arena(lib, 32 * n4 + 64)
    try:
        offs = [_up(arena, [rnd.randrange(p) for _ in range(n4)], curve) for _ in range(29)]
        ptrs = (C.c_void_p * 29)(*[arena.ptr(o) for o in offs])
        scal = np.ascontiguousarray(enc.fr_to_mont(curve, [rnd.randrange(p) for _ in range(16)]))
        vh = np.ascontiguousarray(enc.fr_to_mont(curve, [rnd.randrange(1, p) for _ in range(4)]))
        full, parts = arena.alloc(n4), arena.alloc(n4)
        lib.check(lib.c.apb_plonk_quotient_full(curve, ptrs, scal.ctypes.data, vh.ctypes.data, arena.ptr(full), n4))
        cuts = [0, 1, n4 // 3, n4 // 3, n4 - 5, n4]                 # incl. an empty slice and the wrap-around rows
        for lo, hi in zip(cuts, cuts[1:]):
            lib.check(lib.c.apb_plonk_quotient_range(curve, ptrs, scal.ctypes.data, vh.ctypes.data, arena.ptr(parts), n4, lo, hi - lo))
        assert np.array_equal(arena.download(full, n4), arena.download(parts, n4))
        assert lib.c.apb_plonk_quotient_range(curve, ptrs, scal.ctypes.data, vh.ctypes.data, arena.ptr(parts), n4, n4 - 2, 3) != 0
    finally:
        arena.close()
