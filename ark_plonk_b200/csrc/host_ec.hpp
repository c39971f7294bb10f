// Host-side (CPU) base-field and group arithmetic for the MSM epilogue.
//
// The device leaves a handful of partial sums per MSM (a few dozen points); folding them and the
// single field inversion of the final normalisation are a strictly sequential chain of a few
// hundred field multiplications - latency-bound on a GPU thread (~0.4 us per Fq product) but
// tens of ns each on a host core, and the result is needed in host memory anyway (it feeds the
// Fiat-Shamir transcript).  This is product code (the tail of the hot path), not a fallback.
#pragma once
#include <stdint.h>
#include <string.h>

namespace apb {
namespace host {

typedef unsigned __int128 u128;

struct Field {
    int L;                 // u64 limbs (4 or 6)
    uint64_t mod[6];
    uint64_t one[6];       // R mod p
    uint64_t r2[6];        // R^2 mod p
    uint64_t n0inv;        // -p^-1 mod 2^64

    template <class P>
    static Field make() {
        Field f;
        f.L = P::N / 2;
        for (int i = 0; i < 6; i++) f.mod[i] = f.one[i] = f.r2[i] = 0;
        for (int i = 0; i < f.L; i++) {
            f.mod[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
            f.one[i] = (uint64_t)P::one(2 * i) | ((uint64_t)P::one(2 * i + 1) << 32);
            f.r2[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32);
        }
        uint64_t inv = 1;                      // Newton: inv = p^-1 mod 2^64
        for (int i = 0; i < 6; i++) inv *= 2 - f.mod[0] * inv;
        f.n0inv = (uint64_t)0 - inv;
        return f;
    }

    bool is_zero(const uint64_t* a) const {
        uint64_t o = 0;
        for (int i = 0; i < L; i++) o |= a[i];
        return o == 0;
    }
    bool eq(const uint64_t* a, const uint64_t* b) const { return memcmp(a, b, 8 * L) == 0; }
    void set(uint64_t* r, const uint64_t* a) const { memcpy(r, a, 8 * L); }
    void set_zero(uint64_t* r) const { memset(r, 0, 8 * L); }

    bool geq_mod(const uint64_t* a) const {
        for (int i = L - 1; i >= 0; i--) {
            if (a[i] > mod[i]) return true;
            if (a[i] < mod[i]) return false;
        }
        return true;
    }
    void sub_mod_inplace(uint64_t* a) const {
        uint64_t borrow = 0;
        for (int i = 0; i < L; i++) {
            u128 d = (u128)a[i] - mod[i] - borrow;
            a[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
    }
    void add(uint64_t* r, const uint64_t* a, const uint64_t* b) const {
        uint64_t carry = 0;
        uint64_t t[6];
        for (int i = 0; i < L; i++) {
            u128 s = (u128)a[i] + b[i] + carry;
            t[i] = (uint64_t)s;
            carry = (uint64_t)(s >> 64);
        }
        if (carry || geq_mod(t)) sub_mod_inplace(t);
        set(r, t);
    }
    void sub(uint64_t* r, const uint64_t* a, const uint64_t* b) const {
        uint64_t borrow = 0;
        uint64_t t[6];
        for (int i = 0; i < L; i++) {
            u128 d = (u128)a[i] - b[i] - borrow;
            t[i] = (uint64_t)d;
            borrow = (uint64_t)(d >> 64) & 1;
        }
        if (borrow) {
            uint64_t carry = 0;
            for (int i = 0; i < L; i++) {
                u128 s = (u128)t[i] + mod[i] + carry;
                t[i] = (uint64_t)s;
                carry = (uint64_t)(s >> 64);
            }
        }
        set(r, t);
    }
    void neg(uint64_t* r, const uint64_t* a) const {
        if (is_zero(a)) { set_zero(r); return; }
        uint64_t z[6] = {0, 0, 0, 0, 0, 0};
        sub(r, z, a);
    }
    void dbl(uint64_t* r, const uint64_t* a) const { add(r, a, a); }

    // Montgomery product (CIOS, 64-bit limbs)
    void mul(uint64_t* r, const uint64_t* a, const uint64_t* b) const {
        uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < L; i++) {
            uint64_t carry = 0;
            for (int j = 0; j < L; j++) {
                u128 s = (u128)a[j] * b[i] + t[j] + carry;
                t[j] = (uint64_t)s;
                carry = (uint64_t)(s >> 64);
            }
            u128 s = (u128)t[L] + carry;
            t[L] = (uint64_t)s;
            t[L + 1] = (uint64_t)(s >> 64);
            uint64_t m = t[0] * n0inv;
            s = (u128)m * mod[0] + t[0];
            carry = (uint64_t)(s >> 64);
            for (int j = 1; j < L; j++) {
                s = (u128)m * mod[j] + t[j] + carry;
                t[j - 1] = (uint64_t)s;
                carry = (uint64_t)(s >> 64);
            }
            s = (u128)t[L] + carry;
            t[L - 1] = (uint64_t)s;
            t[L] = t[L + 1] + (uint64_t)(s >> 64);
        }
        if (t[L] || geq_mod(t)) sub_mod_inplace(t);
        set(r, t);
    }
    void sqr(uint64_t* r, const uint64_t* a) const { mul(r, a, a); }

    // a^-1 via Fermat (a != 0), Montgomery in / out
    void inv(uint64_t* r, const uint64_t* a) const {
        uint64_t e[6];
        for (int i = 0; i < L; i++) e[i] = mod[i];
        e[0] -= 2;                              // p - 2 (p odd, low limb >= 3)
        uint64_t acc[6], base[6];
        set(acc, one);
        set(base, a);
        for (int i = 0; i < 64 * L; i++) {
            if ((e[i / 64] >> (i % 64)) & 1) mul(acc, acc, base);
            sqr(base, base);
        }
        set(r, acc);
    }
    void to_canonical(uint64_t* r, const uint64_t* a) const {
        uint64_t o[6] = {1, 0, 0, 0, 0, 0};
        mul(r, a, o);
    }
};

// XYZZ point on a = 0 short Weierstrass curve (same coordinates as the device, ec.cuh)
struct Pt {
    uint64_t x[6], y[6], zz[6], zzz[6];
};

struct Group {
    Field f;
    bool is_identity(const Pt& p) const { return f.is_zero(p.zz); }
    void set_identity(Pt& p) const { memset(&p, 0, sizeof(Pt)); }

    void dbl(Pt& r, const Pt& p) const {
        if (is_identity(p)) { r = p; return; }
        uint64_t U[6], V[6], W[6], S[6], M[6], X2[6], t[6];
        Pt o;
        f.dbl(U, p.y);
        f.sqr(V, U);
        f.mul(W, U, V);
        f.mul(S, p.x, V);
        f.sqr(X2, p.x);
        f.dbl(M, X2);
        f.add(M, M, X2);
        f.sqr(o.x, M);
        f.dbl(t, S);
        f.sub(o.x, o.x, t);
        f.sub(t, S, o.x);
        f.mul(o.y, M, t);
        f.mul(t, W, p.y);
        f.sub(o.y, o.y, t);
        f.mul(o.zz, V, p.zz);
        f.mul(o.zzz, W, p.zzz);
        memset(o.x + f.L, 0, 8 * (6 - f.L)); memset(o.y + f.L, 0, 8 * (6 - f.L));
        memset(o.zz + f.L, 0, 8 * (6 - f.L)); memset(o.zzz + f.L, 0, 8 * (6 - f.L));
        r = o;
    }
    void add(Pt& r, const Pt& a, const Pt& b) const {
        if (is_identity(b)) { r = a; return; }
        if (is_identity(a)) { r = b; return; }
        uint64_t U1[6], U2[6], S1[6], S2[6], P[6], R[6], PP[6], PPP[6], Q[6], t[6];
        f.mul(U1, a.x, b.zz);
        f.mul(U2, b.x, a.zz);
        f.mul(S1, a.y, b.zzz);
        f.mul(S2, b.y, a.zzz);
        f.sub(P, U2, U1);
        f.sub(R, S2, S1);
        if (f.is_zero(P)) {
            if (f.is_zero(R)) dbl(r, a);
            else set_identity(r);
            return;
        }
        Pt o;
        memset(&o, 0, sizeof(o));
        f.sqr(PP, P);
        f.mul(PPP, P, PP);
        f.mul(Q, U1, PP);
        f.sqr(o.x, R);
        f.sub(o.x, o.x, PPP);
        f.dbl(t, Q);
        f.sub(o.x, o.x, t);
        f.sub(t, Q, o.x);
        f.mul(o.y, R, t);
        f.mul(t, S1, PPP);
        f.sub(o.y, o.y, t);
        f.mul(o.zz, a.zz, b.zz);
        f.mul(o.zz, o.zz, PP);
        f.mul(o.zzz, a.zzz, b.zzz);
        f.mul(o.zzz, o.zzz, PPP);
        r = o;
    }
    // affine (x, y) in Montgomery form; returns false for the identity
    bool to_affine(uint64_t* ax, uint64_t* ay, const Pt& p) const {
        if (is_identity(p)) return false;
        uint64_t zi[6], zi2[6], zi3[6];
        // 1/zzz, then 1/zz = (1/zzz)^2 * zz^2 ... simpler: invert both via one inversion of zz*zzz
        uint64_t prod[6], pinv[6];
        f.mul(prod, p.zz, p.zzz);
        f.inv(pinv, prod);
        f.mul(zi2, pinv, p.zzz);     // 1/zz
        f.mul(zi3, pinv, p.zz);      // 1/zzz
        (void)zi;
        f.mul(ax, p.x, zi2);
        f.mul(ay, p.y, zi3);
        return true;
    }
};

}  // namespace host
}  // namespace apb
