// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces ark-ff 0.3 Fp256 / Fp384 (feature `asm`) under every hot-path call of the
// reference (plonk-core/Cargo.toml:28-29,50-59).  Representation is identical to arkworks':
// fully reduced a*R mod p, R = 2^(32*N), little-endian limbs - so the bytes that cross the
// C ABI are arkworks' in-memory limbs.
//
// The multiplier is a CIOS Montgomery product written as mad.lo.cc / madc.hi.cc carry
// chains.  Products a[j]*b_i with even j and with odd j are accumulated in two separate
// limb arrays so that every (lo, hi) pair lands on a 64-bit aligned pair of the same array;
// ptxas then fuses each pair into one IMAD.WIDE.U32(.X) and the two arrays give two
// independent carry chains for ILP.  After each 32-bit Montgomery shift the arrays swap
// roles (what was aligned to column 1 is now aligned to column 0).
//
// Cost: N*(2N+1) wide multiply-adds (Fr: 136, Fq: 300) = the figure SURVEY.md section 8(d)
// uses for the INT32 roofline.
#pragma once
#include "apb_cuda.h"
#include "params_gen.cuh"

namespace apb {

// ---------------------------------------------------------------------------------------
// carry-flag primitives.  Device: one PTX instruction each.  Host (tests/emu and the nvcc
// host pass): emulated with a thread-local carry so the exact same call sequence is checked
// on the CPU.
// ---------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
APB_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
APB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
APB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
APB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
APB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
inline thread_local uint32_t g_cc = 0;
inline uint32_t emu_add3(uint64_t a, uint64_t b, uint64_t cin, bool set) {
    uint64_t s = a + b + cin;
    if (set) g_cc = (uint32_t)(s >> 32);
    return (uint32_t)s;
}
inline uint32_t emu_sub3(uint64_t a, uint64_t b, uint64_t bin, bool set) {
    uint64_t s = a - b - bin;
    if (set) g_cc = (uint32_t)((s >> 32) & 1);      // borrow
    return (uint32_t)s;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add3(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add3(a, b, g_cc, true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add3(a, b, g_cc, false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub3(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub3(a, b, g_cc, true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub3(a, b, g_cc, false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_lo(a, b), c, g_cc, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_hi(a, b), c, g_cc, true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add3(mul_hi(a, b), c, g_cc, false); }
#endif

// ---------------------------------------------------------------------------------------
template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t v[N];

    APB_HD static Fp zero() { Fp r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
    APB_HD static Fp one() { Fp r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = P::one(i); return r; }
    APB_HD static Fp r2() { Fp r; _Pragma("unroll") for (int i = 0; i < N; i++) r.v[i] = P::r2(i); return r; }

    APB_HD bool is_zero() const {
        uint32_t o = 0;
        _Pragma("unroll") for (int i = 0; i < N; i++) o |= v[i];
        return o == 0;
    }
    APB_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
        _Pragma("unroll") for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    APB_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // r = a - p if a >= p else a      (a < 2p)
    APB_HD static void reduce_once(uint32_t* a) {
        uint32_t t[N];
        t[0] = sub_cc(a[0], P::mod(0));
        _Pragma("unroll") for (int i = 1; i < N; i++) t[i] = subc_cc(a[i], P::mod(i));
        uint32_t borrow = subc(0, 0);                 // 0 or 0xffffffff
        _Pragma("unroll") for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
    }

    APB_HD friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        r.v[0] = add_cc(a.v[0], b.v[0]);
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
        r.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);    // 2p < 2^(32N): no carry out
        reduce_once(r.v);
        return r;
    }
    APB_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        r.v[0] = sub_cc(a.v[0], b.v[0]);
        _Pragma("unroll") for (int i = 1; i < N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = subc(0, 0);                 // mask
        r.v[0] = add_cc(r.v[0], borrow & P::mod(0));
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(r.v[i], borrow & P::mod(i));
        r.v[N - 1] = addc(r.v[N - 1], borrow & P::mod(N - 1));
        return r;
    }
    APB_HD Fp neg() const {
        if (is_zero()) return *this;
        Fp r;
        r.v[0] = sub_cc(P::mod(0), v[0]);
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::mod(i), v[i]);
        r.v[N - 1] = subc(P::mod(N - 1), v[N - 1]);
        return r;
    }
    APB_HD Fp dbl() const { return *this + *this; }

    // ---- Montgomery product ------------------------------------------------------------
    // acc[0..n) += x[0], x[2], ... (every second limb starting at x[0]) * y, pairs (lo, hi)
    // on acc[j], acc[j+1]; returns with the carry flag of the chain live.
    template <int CNT>
    APB_HD static void chain_mad(uint32_t* acc, const uint32_t* x, uint32_t y) {
        acc[0] = mad_lo_cc(x[0], y, acc[0]);
        acc[1] = madc_hi_cc(x[0], y, acc[1]);
        _Pragma("unroll") for (int j = 2; j < CNT; j += 2) {
            acc[j] = madc_lo_cc(x[j], y, acc[j]);
            acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 1]);
        }
    }
    // same chain for the modulus (limbs come from constexpr so they fold to immediates)
    template <int START>
    APB_HD static void chain_mad_mod(uint32_t* acc, uint32_t y) {
        acc[0] = mad_lo_cc(P::mod(START), y, acc[0]);
        acc[1] = madc_hi_cc(P::mod(START), y, acc[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            acc[j] = madc_lo_cc(P::mod(START + j), y, acc[j]);
            acc[j + 1] = madc_hi_cc(P::mod(START + j), y, acc[j + 1]);
        }
    }

    // One CIOS step.  E is aligned to column 0, O to column 1 (of the current window).
    //   FIRST: both arrays are uninitialised and simply receive the products.
    //   else : O still holds last step's column -1.. layout (it was that step's E): O[0] == 0,
    //          O[1] belongs to column 0, O[k+2] to column k+1.
    template <bool FIRST>
    APB_HD static void cios_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {
        if (FIRST) {
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                E[j] = mul_lo(a[j], bi);
                E[j + 1] = mul_hi(a[j], bi);
                O[j] = mul_lo(a[j + 1], bi);
                O[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            E[0] = add_cc(E[0], O[1]);                       // column 0 <- stray limb
            _Pragma("unroll") for (int j = 1; j < N - 1; j += 2) {   // odd j: shift O down by 2 while accumulating
                O[j - 1] = madc_lo_cc(a[j], bi, O[j + 1]);
                O[j] = madc_hi_cc(a[j], bi, O[j + 2]);
            }
            O[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
            O[N - 1] = madc_hi(a[N - 1], bi, 0);
            chain_mad<N>(E, a, bi);                          // even j
            O[N - 1] = addc(O[N - 1], 0);
        }
        uint32_t m = mul_lo(E[0], P::N0INV);
        chain_mad_mod<1>(O, m);                              // odd limbs of p
        // (carry out of the O chain is provably 0: every partial sum fits N+1 limbs)
        chain_mad_mod<0>(E, m);                              // even limbs of p -> E[0] == 0
        O[N - 1] = addc(O[N - 1], 0);
    }

    // REDUCE = false leaves the result in [0, 2p) (no final conditional subtraction): allowed wherever the value is
    // only ever used as an OPERAND of further products - a CIOS product of operands a, b < 2p is again below
    // a b / R + p < 2p (4p < R for every field here), and the one reduce_once of the product that finally consumes it
    // makes the result canonical.  Never compare, add, subtract or store-as-output such a value.
    template <bool REDUCE>
    APB_HD static Fp mul_impl(const Fp& a, const Fp& b) {
        uint32_t X[N], Y[N];
        cios_step<true>(X, Y, a.v, b.v[0]);
        _Pragma("unroll") for (int i = 1; i < N; i += 2) {
            cios_step<false>(Y, X, a.v, b.v[i]);             // roles swap after every shift
            if (i + 1 < N) cios_step<false>(X, Y, a.v, b.v[i + 1]);
        }
        // N is even: after N steps the array aligned to (new) column 0 is X's partner ... i.e.
        // the last call had E = Y, so the current column-0 array is X and Y holds the stray layout.
        Fp r;
        r.v[0] = add_cc(X[0], Y[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
        r.v[N - 1] = addc(X[N - 1], 0);
        if (REDUCE) reduce_once(r.v);
        return r;
    }
    APB_HD friend Fp operator*(const Fp& a, const Fp& b) { return mul_impl<true>(a, b); }
    APB_HD Fp mul_lazy(const Fp& b) const { return mul_impl<false>(*this, b); }
    // ---- Montgomery square (base fields only: needs 3p < 2^(32N)) ---------------------------
    // CIOS where step i multiplies a_i by the vector [0,..,0, a_i, 2*a_{>i}]: every off-diagonal
    // product a_i*a_j is formed once (N(N+1)/2 = 78 instead of 144 for N = 12).  2*a_{>i} is read
    // limb-wise with funnel shifts: limb i+1 is a_{i+1} << 1, limb j > i+1 is (a_j << 1) | (a_{j-1} >> 31).
    template <int I>
    APB_HD static uint32_t sq_vec(const uint32_t* a, int j) {
        return j == I ? a[j] : (j == I + 1 ? (a[j] << 1) : ((a[j] << 1) | (a[j - 1] >> 31)));
    }
    template <int I>
    APB_HD static void cios_step_sq(uint32_t* E, uint32_t* O, const uint32_t* a) {
        const uint32_t bi = a[I];
        if (I == 0) {
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                const uint32_t ve = sq_vec<0>(a, j), vo = sq_vec<0>(a, j + 1);
                E[j] = mul_lo(ve, bi);
                E[j + 1] = mul_hi(ve, bi);
                O[j] = mul_lo(vo, bi);
                O[j + 1] = mul_hi(vo, bi);
            }
        } else {
            E[0] = add_cc(E[0], O[1]);
            _Pragma("unroll") for (int j = 1; j < N - 1; j += 2) {       // odd columns: shift O down by 2
                if (j >= I) {
                    const uint32_t v = sq_vec<I>(a, j);
                    O[j - 1] = madc_lo_cc(v, bi, O[j + 1]);
                    O[j] = madc_hi_cc(v, bi, O[j + 2]);
                } else {
                    O[j - 1] = addc_cc(O[j + 1], 0);
                    O[j] = addc_cc(O[j + 2], 0);
                }
            }
            {
                const uint32_t v = sq_vec<I>(a, N - 1);                  // N-1 >= I always
                O[N - 2] = madc_lo_cc(v, bi, 0);
                O[N - 1] = madc_hi(v, bi, 0);
            }
            constexpr int J0 = (I + 1) & ~1;                             // first even column >= I
            if (J0 < N) {
                const uint32_t v0 = sq_vec<I>(a, J0);
                E[J0] = mad_lo_cc(v0, bi, E[J0]);
                E[J0 + 1] = madc_hi_cc(v0, bi, E[J0 + 1]);
                _Pragma("unroll") for (int j = J0 + 2; j < N; j += 2) {
                    const uint32_t v = sq_vec<I>(a, j);
                    E[j] = madc_lo_cc(v, bi, E[j]);
                    E[j + 1] = madc_hi_cc(v, bi, E[j + 1]);
                }
                O[N - 1] = addc(O[N - 1], 0);
            }
        }
        uint32_t m = mul_lo(E[0], P::N0INV);
        chain_mad_mod<1>(O, m);
        chain_mad_mod<0>(E, m);
        O[N - 1] = addc(O[N - 1], 0);
    }
    template <int I>
    APB_HD static void sq_steps(uint32_t* X, uint32_t* Y, const uint32_t* a) {
        if (I < N) {
            if (I & 1) cios_step_sq<(I < N ? I : 0)>(Y, X, a);
            else cios_step_sq<(I < N ? I : 0)>(X, Y, a);
            sq_steps<(I < N ? I + 1 : N)>(X, Y, a);
        }
    }
    APB_HD Fp sqr() const {
        if (N != 12) return *this * *this;          // Fr: 3p does not fit 2^256 - use the generic product
        uint32_t X[N], Y[N];
        sq_steps<0>(X, Y, v);
        Fp r;
        r.v[0] = add_cc(X[0], Y[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.v[k] = addc_cc(X[k], Y[k + 1]);
        r.v[N - 1] = addc(X[N - 1], 0);
        reduce_once(r.v);
        return r;
    }

    // a*R mod p  <-  a          /   a  <-  a*R mod p
    APB_HD Fp to_mont() const { return *this * r2(); }
    APB_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }

    // 1/a in Montgomery form by the binary extended Euclidean algorithm (shifts, adds, compares on
    // the limbs: ~2 log2(p) halvings + log2(p) subtractions).  For ONE thread that must invert alone
    // (the batched-inversion root in msm.cu) this is ~10x shorter than the a^(p-2) chain, whose
    // ~570 dependent Montgomery products run at single-warp latency; it is data-dependent, so a
    // warp whose lanes all invert should keep using the exponentiation.  a != 0.
    APB_HD static void shr1(uint32_t* x) {
        _Pragma("unroll") for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
        x[N - 1] >>= 1;
    }
    APB_HD static void halve_mod(uint32_t* x) {          // x/2 mod p for x < p (p odd, p < 2^(32N-2))
        if (x[0] & 1) {
            x[0] = add_cc(x[0], P::mod(0));
            _Pragma("unroll") for (int i = 1; i < N - 1; i++) x[i] = addc_cc(x[i], P::mod(i));
            x[N - 1] = addc(x[N - 1], P::mod(N - 1));
        }
        shr1(x);
    }
    APB_HD static bool is_one_raw(const uint32_t* x) {
        uint32_t acc = x[0] ^ 1u;
        _Pragma("unroll") for (int i = 1; i < N; i++) acc |= x[i];
        return acc == 0;
    }
    APB_HD static bool geq_raw(const uint32_t* a, const uint32_t* b) {
        for (int i = N - 1; i >= 0; i--) {
            if (a[i] > b[i]) return true;
            if (a[i] < b[i]) return false;
        }
        return true;
    }
    APB_HD static void sub_raw(uint32_t* a, const uint32_t* b) {     // a -= b, a >= b
        a[0] = sub_cc(a[0], b[0]);
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) a[i] = subc_cc(a[i], b[i]);
        a[N - 1] = subc(a[N - 1], b[N - 1]);
    }
    APB_HD Fp inverse_binary() const {
        Fp u = *this, w, x1 = zero(), x2 = zero();
        _Pragma("unroll") for (int i = 0; i < N; i++) w.v[i] = P::mod(i);
        x1.v[0] = 1;
        while (!is_one_raw(u.v) && !is_one_raw(w.v)) {
            while (!(u.v[0] & 1)) { shr1(u.v); halve_mod(x1.v); }
            while (!(w.v[0] & 1)) { shr1(w.v); halve_mod(x2.v); }
            if (geq_raw(u.v, w.v)) { sub_raw(u.v, w.v); x1 = x1 - x2; }
            else { sub_raw(w.v, u.v); x2 = x2 - x1; }
        }
        // x = (aR)^-1 as a plain residue = a^-1 R^-1; two Montgomery products by R^2 give a^-1 R
        const Fp x = is_one_raw(u.v) ? x1 : x2;
        return (x * r2()) * r2();
    }
};

}  // namespace apb
