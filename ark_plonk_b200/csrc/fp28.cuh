// Reduced-radix base-field arithmetic for the MSM inner loop: 14 limbs of 28 bits.
//
// Why: on sm_100a the carry-chained IMAD.WIDE.U32.X that a saturated 32-bit-limb Montgomery
// product compiles to issues at HALF the rate of a plain IMAD.WIDE (measured,
// profiles/r01_mul_bench.json: 50.3 % of the plain-IMAD peak is that instruction mix's ceiling).
// With 28-bit limbs every partial product is < 2^58 (inputs may carry slack up to 3*2^28), so a
// whole column of the schoolbook product AND of the Montgomery reduction (28 terms) fits a 64-bit
// accumulator: the multiplier becomes 392 carry-free `acc += (u64)a*b` = plain IMAD.WIDE at full
// rate, plus ~170 ALU-pipe shifts/masks that issue in the FMA pipe's shadow.
//
// Domain: Montgomery with R' = 2^(28*14) = 2^392 (the resident MSM table is stored as x*2^392
// mod p; results are mapped back to the library-wide 2^384 domain when a bucket is flushed).
// Discipline ("tight" = limbs 0..12 < 2^28, the top limb holds the excess):
//   mul / sqr : inputs with limbs < 3*2^28 and value < 39p (< 2^386)  ->  tight, value < 2p
//   add       : limb-wise, no carry (lazy; result limbs < sum of the bounds)
//   sub(a, b) : b tight with value < 14p  ->  a - b + 16p, carry-normalised (tight), value < a + 16p
//   reduce    : tight value < 39p -> value < 2p (quotient estimate from the top limb, one k*p subtraction)
// A coordinate that is stored in the accumulator (and so becomes a subtrahend later) is reduced;
// differences that only feed a multiplication are not.  Zero tests compare against k*p.
#pragma once
#include "arith.cuh"

namespace apb {

template <class P28>
struct Fp28 {
    typedef typename P28::Base Base;
    static constexpr int L = 14;
    static constexpr uint32_t MASK = (1u << 28) - 1;
    uint32_t l[L];

    APB_HD static Fp28 zero() { Fp28 r; _Pragma("unroll") for (int i = 0; i < L; i++) r.l[i] = 0; return r; }
    APB_HD static Fp28 one() { Fp28 r; _Pragma("unroll") for (int i = 0; i < L; i++) r.l[i] = P28::one(i); return r; }

    // 12 x u32 packed integer (< 2^384) -> 14 x 28-bit limbs
    APB_HD static Fp28 from_words(const uint32_t* w) {
        Fp28 r;
        _Pragma("unroll") for (int i = 0; i < L; i++) {
            const int bit = 28 * i, wi = bit >> 5, sh = bit & 31;
            uint32_t lo = w[wi] >> sh;
            if (sh > 4 && wi + 1 < 12) lo |= w[wi + 1] << (32 - sh);
            r.l[i] = lo & MASK;
        }
        return r;
    }
    // tight limbs, value < 2^384 -> 12 x u32
    APB_HD void to_words(uint32_t* w) const {
        _Pragma("unroll") for (int j = 0; j < 12; j++) {
            const int bit = 32 * j, li = bit / 28, sh = bit - 28 * li;      // word j starts inside limb li
            uint32_t v = l[li] >> sh;
            int have = 28 - sh;
            if (li + 1 < L) { v |= l[li + 1] << have; have += 28; }
            if (have < 32 && li + 2 < L) v |= l[li + 2] << have;
            w[j] = v;
        }
    }

    APB_HD void normalize() {
        _Pragma("unroll") for (int i = 0; i < L - 1; i++) {
            l[i + 1] += l[i] >> 28;
            l[i] &= MASK;
        }
    }
    APB_HD friend Fp28 operator+(const Fp28& a, const Fp28& b) {
        Fp28 r;
        _Pragma("unroll") for (int i = 0; i < L; i++) r.l[i] = a.l[i] + b.l[i];
        return r;
    }
    APB_HD friend Fp28 operator-(const Fp28& a, const Fp28& b) {
        Fp28 r;
        _Pragma("unroll") for (int i = 0; i < L; i++) r.l[i] = a.l[i] + P28::sub_offset(i) - b.l[i];
        r.normalize();
        return r;
    }
    // a + a, normalised (usable as a subtrahend)
    APB_HD Fp28 dbl_norm() const {
        Fp28 r;
        _Pragma("unroll") for (int i = 0; i < L; i++) r.l[i] = l[i] << 1;
        r.normalize();
        return r;
    }
    // p - a for canonical 0 < a < p: canonical result (signed borrow propagation)
    APB_HD Fp28 neg_canonical() const {
        Fp28 r;
        int32_t carry = 0;
        _Pragma("unroll") for (int i = 0; i < L; i++) {
            int32_t t = (int32_t)P28::mod(i) - (int32_t)l[i] + carry;
            r.l[i] = (uint32_t)t & MASK;
            carry = t >> 28;
        }
        return r;
    }
    // tight value < 39p  ->  tight value < 2p
    APB_HD void reduce() {
        const uint32_t h = l[L - 1] >> P28::RED_SHIFT;
        const uint32_t q = (h * P28::RED_MUL) >> 8;             // q <= value / p, q < KP_COUNT
        int32_t carry = 0;
        _Pragma("unroll") for (int i = 0; i < L - 1; i++) {
            int32_t t = (int32_t)l[i] - (int32_t)P28::kp(q, i) + carry;
            l[i] = (uint32_t)t & MASK;
            carry = t >> 28;
        }
        l[L - 1] = (uint32_t)((int32_t)l[L - 1] - (int32_t)P28::kp(q, L - 1) + carry);
    }

    // value == 0 (mod p) for a tight value in (k0 * p - p, k1 * p): compares against k*p, k0 <= k < k1
    APB_HD bool is_zero_mod_p(int k0 = 0, int k1 = P28::KP_COUNT) const {
        bool hit = false;
        _Pragma("unroll 1") for (int k = k0; k < k1; k++) {
            if (l[0] != P28::kp(k, 0)) continue;
            bool same = true;
            for (int i = 1; i < L; i++) same = same && (l[i] == P28::kp(k, i));
            hit = hit || same;
        }
        return hit;
    }

    // Montgomery product, column by column; two independent 64-bit accumulators per column
    APB_HD friend Fp28 operator*(const Fp28& a, const Fp28& b) {
        uint32_t m[L];
        Fp28 r;
        uint64_t carry = 0;
        _Pragma("unroll") for (int k = 0; k < L; k++) {
            uint64_t ab = carry, mp = 0;
            _Pragma("unroll") for (int i = 0; i <= k; i++) ab += (uint64_t)a.l[i] * b.l[k - i];
            _Pragma("unroll") for (int i = 0; i < k; i++) mp += (uint64_t)m[i] * P28::mod(k - i);
            uint64_t col = ab + mp;
            m[k] = ((uint32_t)col * P28::N0INV28) & MASK;
            col += (uint64_t)m[k] * P28::mod(0);
            carry = col >> 28;
        }
        _Pragma("unroll") for (int k = L; k < 2 * L - 1; k++) {
            uint64_t ab = carry, mp = 0;
            _Pragma("unroll") for (int i = k - L + 1; i < L; i++) ab += (uint64_t)a.l[i] * b.l[k - i];
            _Pragma("unroll") for (int i = k - L + 1; i < L; i++) mp += (uint64_t)m[i] * P28::mod(k - i);
            uint64_t col = ab + mp;
            r.l[k - L] = (uint32_t)col & MASK;
            carry = col >> 28;
        }
        r.l[L - 1] = (uint32_t)carry;
        return r;
    }
    // square: off-diagonal products once, doubled
    APB_HD Fp28 sqr() const {
        uint32_t m[L], d[L];
        Fp28 r;
        _Pragma("unroll") for (int i = 0; i < L; i++) d[i] = l[i] << 1;
        uint64_t carry = 0;
        _Pragma("unroll") for (int k = 0; k < L; k++) {
            uint64_t ab = carry, mp = 0;
            _Pragma("unroll") for (int i = 0; 2 * i < k; i++) ab += (uint64_t)d[i] * l[k - i];
            if ((k & 1) == 0) ab += (uint64_t)l[k >> 1] * l[k >> 1];
            _Pragma("unroll") for (int i = 0; i < k; i++) mp += (uint64_t)m[i] * P28::mod(k - i);
            uint64_t col = ab + mp;
            m[k] = ((uint32_t)col * P28::N0INV28) & MASK;
            col += (uint64_t)m[k] * P28::mod(0);
            carry = col >> 28;
        }
        _Pragma("unroll") for (int k = L; k < 2 * L - 1; k++) {
            uint64_t ab = carry, mp = 0;
            _Pragma("unroll") for (int i = k - L + 1; 2 * i < k; i++) ab += (uint64_t)d[i] * l[k - i];
            if ((k & 1) == 0) ab += (uint64_t)l[k >> 1] * l[k >> 1];
            _Pragma("unroll") for (int i = k - L + 1; i < L; i++) mp += (uint64_t)m[i] * P28::mod(k - i);
            uint64_t col = ab + mp;
            r.l[k - L] = (uint32_t)col & MASK;
            carry = col >> 28;
        }
        r.l[L - 1] = (uint32_t)carry;
        return r;
    }

    // x * 2^392  ->  canonical x * 2^384 as an Fp<Base> (the library-wide Montgomery form)
    APB_HD Fp<Base> to_mont384() const {
        Fp28 c;
        _Pragma("unroll") for (int i = 0; i < L; i++) c.l[i] = P28::r384(i);
        Fp28 t = *this * c;                 // tight, < 2p
        Fp<Base> o;
        t.to_words(o.v);
        Fp<Base>::reduce_once(o.v);
        return o;
    }
};

// XYZZ accumulator in reduced radix (see ec.cuh for the formulas); `inf` marks the identity
template <class P28>
struct XYZZ28 {
    typedef Fp28<P28> F;
    F x, y, zz, zzz;
    bool inf;

    APB_HD void set_identity() { inf = true; x = F::zero(); y = F::zero(); zz = F::zero(); zzz = F::zero(); }

    __host__ __device__ __noinline__ void dbl_affine(const F& px, const F& py) {   // mdbl-2008-s-1; px, py canonical (rare)
        F U = py + py;                                           // lazy, < 2p
        F V = U.sqr();
        F W = U * V;
        F S = px * V;
        F X2 = px.sqr();
        F M = X2 + X2 + X2;                                      // limbs < 3*2^28, value < 6p
        F X3 = M.sqr() - S.dbl_norm();
        X3.reduce();
        F Y3 = M * (S - X3) - W * py;
        Y3.reduce();
        x = X3;
        y = Y3;
        zz = V;
        zzz = W;
        inf = false;
    }
    // this += (px, py); (px, py) canonical affine in the 2^392 domain, not the point at infinity
    APB_HD void add_affine(const F& px, const F& py) {          // madd-2008-s
        if (inf) {
            x = px; y = py; zz = F::one(); zzz = F::one();
            inf = false;
            return;
        }
        F U2 = px * zz;
        F S2 = py * zzz;
        F Pp = U2 - x;
        F R = S2 - y;
        if (Pp.is_zero_mod_p(15, 18)) {          // U2, x < 2p: Pp = U2 - x + 16p is in (14p, 18p)
            if (R.is_zero_mod_p(1, 32)) dbl_affine(px, py);
            else set_identity();
            return;
        }
        F PP = Pp.sqr();
        F PPP = Pp * PP;
        F Q = x * PP;
        F X3 = (R.sqr() - PPP) - Q.dbl_norm();                   // < 2p + 32p
        X3.reduce();
        F Y3 = R * (Q - X3) - y * PPP;
        Y3.reduce();
        x = X3;
        y = Y3;
        zz = zz * PP;
        zzz = zzz * PPP;
    }
};

}  // namespace apb
