// Device-resident polynomial utilities and the PLONK prover's pointwise / scan kernels.
//
// SURVEY.md section 8(f) "next" rows: these keep every vector of a proof in HBM between the
// NTTs and MSMs of the hot path.  Each entry point cites the reference code it computes the
// same values as (paths relative to the reference repo, plonk-core/src/...):
//   apb_fr_lincomb            - DensePolynomial +/* scalar chains (linearisation_poly.rs:203-349,
//                               MultiSet::compress lookup/multiset.rs:207-213, sonic_pc open)
//   apb_plonk_lookup_f        - query vector f (proof_system/prover.rs:252-278)
//   apb_plonk_combine_split   - MultiSet::combine_split (lookup/multiset.rs:131-174)
//   apb_plonk_perm_z          - Permutation::compute_permutation_poly before its ifft
//                               (permutation/mod.rs:652-751)
//   apb_plonk_lookup_z2       - compute_lookup_permutation_poly before its ifft (:754-822)
//   apb_plonk_quotient        - quotient_poly::compute pointwise part (quotient_poly.rs:122-173,
//                               widget/arithmetic.rs:51-62, proof_system/permutation.rs:62-155,
//                               widget/lookup.rs:96-151)
//   apb_poly_eval             - DensePolynomial::evaluate (linearisation_poly.rs:203-261)
//   apb_poly_divide_linear    - kzg10 witness polynomial p / (X - z)
// All values are bit-identical to the reference's because field elements are unique.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "host_ec.hpp"
#include "scan_u32.cuh"

namespace apb {

static const int SCAN_T = 128;     // threads per scan block
static const int SCAN_E = 8;       // elements per thread

template <class FR>
APB_D Fp<FR> arg_fp(const uint64_t* a) {
    Fp<FR> r;
#pragma unroll
    for (int i = 0; i < 4; i++) { r.v[2 * i] = (uint32_t)a[i]; r.v[2 * i + 1] = (uint32_t)(a[i] >> 32); }
    return r;
}
struct Fr4 { uint64_t v[4]; };

// ---- out[i] = sum_j s_j * p_j[i] -----------------------------------------------------------
static const int LINCOMB_MAX = 16;
struct LincombArgs {
    uint32_t k;
    const void* p[LINCOMB_MAX];
    uint64_t len[LINCOMB_MAX];
    Fr4 s[LINCOMB_MAX];
};
template <class FR>
__global__ void k_lincomb(LincombArgs A, void* out, uint64_t n, int accumulate) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F acc = accumulate ? load_fp<FR>(out, i) : F::zero();
    for (uint32_t j = 0; j < A.k; j++) {
        if (i < A.len[j]) acc = acc + load_fp<FR>(A.p[j], i) * arg_fp<FR>(A.s[j].v);
    }
    store_fp<FR>(out, i, acc);
}

// ---- generic scans over Fr ------------------------------------------------------------------
template <class FR, int OP>
APB_D Fp<FR> scan_op(const Fp<FR>& a, const Fp<FR>& b) { return OP == 0 ? a * b : a + b; }
template <class FR, int OP>
APB_D Fp<FR> scan_id() { return OP == 0 ? Fp<FR>::one() : Fp<FR>::zero(); }

// exclusive scan inside each block of SCAN_T*SCAN_E logical elements; logical index L maps to
// physical n-1-L when reverse != 0.  block_tot[b] = reduction of block b.
template <class FR, int OP>
__global__ void __launch_bounds__(SCAN_T) k_scan_block(const void* in, void* out, void* block_tot, uint64_t n, int reverse) {
    typedef Fp<FR> F;
    __shared__ uint4 sm[SCAN_T * 2];
    const uint32_t tid = threadIdx.x;
    const uint64_t base = ((uint64_t)blockIdx.x * SCAN_T + tid) * SCAN_E;
    F pre[SCAN_E];
    F run = scan_id<FR, OP>();
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        pre[j] = run;
        uint64_t L = base + j;
        if (L < n) run = scan_op<FR, OP>(run, load_fp<FR>(in, reverse ? n - 1 - L : L));
    }
    store_fp<FR>(sm, tid, run);
    __syncthreads();
    for (uint32_t off = 1; off < SCAN_T; off <<= 1) {        // Hillis-Steele inclusive
        F v = scan_id<FR, OP>();
        if (tid >= off) v = load_fp<FR>(sm, tid - off);
        __syncthreads();
        if (tid >= off) {
            run = scan_op<FR, OP>(v, run);
            store_fp<FR>(sm, tid, run);
        }
        __syncthreads();
    }
    F excl = tid ? load_fp<FR>(sm, tid - 1) : scan_id<FR, OP>();
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        uint64_t L = base + j;
        if (L < n) store_fp<FR>(out, reverse ? n - 1 - L : L, scan_op<FR, OP>(excl, pre[j]));
    }
    if (tid == SCAN_T - 1) store_fp<FR>(block_tot, blockIdx.x, run);
}
// exclusive scan of the block totals, single block, in place; total[0] receives the grand total
template <class FR, int OP>
__global__ void __launch_bounds__(SCAN_T) k_scan_totals(void* block_tot, uint32_t nb, void* total) {
    typedef Fp<FR> F;
    __shared__ uint4 sm[SCAN_T * 2];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nb + SCAN_T - 1) / SCAN_T;
    const uint32_t lo = tid * per, hi = lo + per < nb ? lo + per : nb;
    F run = scan_id<FR, OP>();
    for (uint32_t i = lo; i < hi; i++) run = scan_op<FR, OP>(run, load_fp<FR>(block_tot, i));
    store_fp<FR>(sm, tid, run);
    __syncthreads();
    for (uint32_t off = 1; off < SCAN_T; off <<= 1) {
        F v = scan_id<FR, OP>();
        if (tid >= off) v = load_fp<FR>(sm, tid - off);
        __syncthreads();
        if (tid >= off) {
            run = scan_op<FR, OP>(v, run);
            store_fp<FR>(sm, tid, run);
        }
        __syncthreads();
    }
    F acc = tid ? load_fp<FR>(sm, tid - 1) : scan_id<FR, OP>();
    for (uint32_t i = lo; i < hi; i++) {
        F e = load_fp<FR>(block_tot, i);
        store_fp<FR>(block_tot, i, acc);
        acc = scan_op<FR, OP>(acc, e);
    }
    if (tid == SCAN_T - 1) store_fp<FR>(total, 0, run);
}
template <class FR, int OP>
__global__ void __launch_bounds__(SCAN_T) k_scan_apply(void* out, const void* block_tot, uint64_t n, int reverse) {
    typedef Fp<FR> F;
    const uint64_t base = ((uint64_t)blockIdx.x * SCAN_T + threadIdx.x) * SCAN_E;
    if (blockIdx.x == 0) return;
    F p = load_fp<FR>(block_tot, blockIdx.x);
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        uint64_t L = base + j;
        if (L < n) {
            uint64_t idx = reverse ? n - 1 - L : L;
            store_fp<FR>(out, idx, scan_op<FR, OP>(p, load_fp<FR>(out, idx)));
        }
    }
}

// domain element w^i from the half table
template <class FR>
APB_D Fp<FR> root_at(const void* tw, uint64_t i, uint64_t half) {
    return i < half ? load_fp<FR>(tw, i) : load_fp<FR>(tw, i - half).neg();
}

// ---- permutation grand product: numerators / denominators per row -----------------------
struct PermArgs {
    const void *w[4], *s[4];
    const void* tw;
    uint64_t n;
    Fr4 beta, gamma, k[4];      // k = 1, K1, K2, K3 (Montgomery)
};
template <class FR>
__global__ void k_perm_terms(PermArgs A, void* num, void* den) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const F beta = arg_fp<FR>(A.beta.v), gamma = arg_fp<FR>(A.gamma.v);
    const F root = A.n > 1 ? root_at<FR>(A.tw, i, A.n >> 1) : F::one();
    F nu = F::one(), de = F::one();
    const F br = beta * root;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        F w = load_fp<FR>(A.w[c], i);
        nu = nu * (w + br * arg_fp<FR>(A.k[c].v) + gamma);
        de = de * (w + beta * load_fp<FR>(A.s[c], i) + gamma);
    }
    store_fp<FR>(num, i, nu);
    store_fp<FR>(den, i, de);
}
// lookup grand product terms (permutation/mod.rs:803-822)
struct LookupZArgs {
    const void *f, *t, *h1, *h2;
    uint64_t n;
    Fr4 delta, epsilon;
};
template <class FR>
__global__ void k_lookup_terms(LookupZArgs A, void* num, void* den) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const F delta = arg_fp<FR>(A.delta.v), eps = arg_fp<FR>(A.epsilon.v);
    const F opd = F::one() + delta, eopd = eps * opd;
    const uint64_t nx = i + 1 == A.n ? 0 : i + 1;
    F f = load_fp<FR>(A.f, i), t = load_fp<FR>(A.t, i), tn = load_fp<FR>(A.t, nx);
    F h1 = load_fp<FR>(A.h1, i), h1n = load_fp<FR>(A.h1, nx), h2 = load_fp<FR>(A.h2, i);
    store_fp<FR>(num, i, opd * (eps + f) * (eopd + t + delta * tn));
    store_fp<FR>(den, i, (eopd + h1 + h2 * delta) * (eopd + h2 + h1n * delta));
}
// z[i] = PN[i] * SD[i] * inv_total, where PN = exclusive prefix product of num, SD[i] = product of den[k], k >= i
// (exclusive suffix product times den[i]); element n-1's own factors are not used by z.
template <class FR>
__global__ void k_grand_product_finish(const void* pn, const void* sd_excl, const void* den, Fr4 inv_total, void* z, uint64_t n) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F v = load_fp<FR>(pn, i) * load_fp<FR>(sd_excl, i) * load_fp<FR>(den, i) * arg_fp<FR>(inv_total.v);
    store_fp<FR>(z, i, v);
}

// ---- lookup query vector f ---------------------------------------------------------------------
template <class FR>
__global__ void k_lookup_f(const void* q_lookup, const void* wl, const void* wr, const void* wo, const void* w4,
                           const void* t_comp, Fr4 zeta, void* out, uint64_t n) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F q = load_fp<FR>(q_lookup, i);
    F r;
    if (q.is_zero()) {
        r = load_fp<FR>(t_comp, 0);
    } else {
        const F z = arg_fp<FR>(zeta.v);
        r = ((load_fp<FR>(w4, i) * z + load_fp<FR>(wo, i)) * z + load_fp<FR>(wr, i)) * z + load_fp<FR>(wl, i);
    }
    store_fp<FR>(out, i, r);
}

// ---- combine_split ---------------------------------------------------------------------------
// Hash table over the values of t: slot owner = index of the first inserting element.
APB_D uint32_t hash_fr(const uint4* p, uint64_t idx, uint32_t mask) {
    uint4 a = p[2 * idx], b = p[2 * idx + 1];
    uint32_t h = a.x * 0x9E3779B1u ^ a.y * 0x85EBCA77u ^ a.z * 0xC2B2AE3Du ^ a.w * 0x27D4EB2Fu ^ b.x * 0x165667B1u ^
                 b.y * 0xD3A2646Cu ^ b.z * 0xFD7046C5u ^ b.w * 0xB55A4F09u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    return h & mask;
}
APB_D bool eq_fr_raw(const uint4* p, uint64_t i, const uint4* q, uint64_t j) {
    uint4 a = p[2 * i], b = p[2 * i + 1], c = q[2 * j], d = q[2 * j + 1];
    return a.x == c.x && a.y == c.y && a.z == c.z && a.w == c.w && b.x == d.x && b.y == d.y && b.z == d.z && b.w == d.w;
}
// phase 0: insert t; phase 1: look up f.  owner[slot] = some index into t holding the slot's value,
// first[slot] = smallest such index, count[slot] = multiplicity in t (+ f).
__global__ void k_cs_hash(const void* tv, const void* fv, uint64_t n, uint32_t mask, uint32_t* owner, uint32_t* first,
                          uint32_t* count, int phase, uint32_t* error) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* t = reinterpret_cast<const uint4*>(tv);
    const uint4* f = reinterpret_cast<const uint4*>(fv);
    const uint4* src = phase == 0 ? t : f;
    uint32_t h = hash_fr(src, i, mask);
    for (uint32_t probe = 0; probe <= mask; probe++, h = (h + 1) & mask) {
        uint32_t o = owner[h];
        if (o == 0xffffffffu) {
            if (phase == 1) { atomicAdd(error, 1u); return; }      // element of f not in t
            o = atomicCAS(&owner[h], 0xffffffffu, (uint32_t)i);
            if (o == 0xffffffffu) o = (uint32_t)i;
        }
        if (eq_fr_raw(t, o, src, i)) {
            if (phase == 0) atomicMin(&first[h], (uint32_t)i);
            atomicAdd(&count[h], 1u);
            return;
        }
    }
    atomicAdd(error, 1u);
}
// per position i of t: if it is the first occurrence of its value, the bucket's (evens, odd-flag)
__global__ void k_cs_bucket_sizes(const void* tv, uint64_t n, uint32_t mask, const uint32_t* owner, const uint32_t* first,
                                  const uint32_t* count, uint32_t* half, uint32_t* odd) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* t = reinterpret_cast<const uint4*>(tv);
    uint32_t h = hash_fr(t, i, mask);
    for (;; h = (h + 1) & mask) {
        uint32_t o = owner[h];
        if (o != 0xffffffffu && eq_fr_raw(t, o, t, i)) break;
    }
    if (first[h] == (uint32_t)i) {
        half[i] = count[h] >> 1;
        odd[i] = count[h] & 1;
    } else {
        half[i] = 0;
        odd[i] = 0;
    }
}
// even_sz[i] = half + (odd && parity_before == 0), odd_sz[i] = half + (odd && parity_before == 1)
__global__ void k_cs_split_sizes(const uint32_t* half, const uint32_t* odd, const uint32_t* odd_prefix, uint32_t* even_sz,
                                 uint32_t* odd_sz, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t par = odd_prefix[i] & 1;
    even_sz[i] = half[i] + ((odd[i] && par == 0) ? 1 : 0);
    odd_sz[i] = half[i] + ((odd[i] && par == 1) ? 1 : 0);
}
// out[p] = t[bucket containing p], offsets[] = exclusive scan of the per-position sizes
__global__ void k_cs_fill(const void* tv, const uint32_t* offsets, uint64_t n, void* out, uint64_t out_len) {
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= out_len) return;
    uint64_t lo = 0, hi = n;                   // largest i with offsets[i] <= p
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid;
    }
    while (offsets[lo + 1] <= p) lo++;
    const uint4* t = reinterpret_cast<const uint4*>(tv);
    uint4* o = reinterpret_cast<uint4*>(out);
    o[2 * p] = t[2 * lo];
    o[2 * p + 1] = t[2 * lo + 1];
}
// ---- quotient -------------------------------------------------------------------------------
struct QuotientArgs {
    const void *wl, *wr, *wo, *w4, *z, *z2, *f, *table, *h1, *h2, *pi;
    const void *q_m, *q_l, *q_r, *q_o, *q_4, *q_c, *q_arith, *q_lookup;
    const void *s1, *s2, *s3, *s4, *linear, *l1;
    const void *q_range, *q_logic, *q_fixed, *q_var;          // custom gate selectors (NULL = identically zero)
    Fr4 alpha, beta, gamma, delta, epsilon, zeta, lookup_sep, k1, k2, k3;
    Fr4 range_sep, logic_sep, fixed_sep, var_sep, coeff_a, coeff_d;   // embedded curve: a x^2 + y^2 = 1 + d x^2 y^2
    Fr4 vh_inv[4];
    Fr4 alpha_sq, lsep_sq, lsep_cu, eps_opd;                  // alpha^2, lookup_sep^2, lookup_sep^3, epsilon (1 + delta): host-computed
};

// f (f-1) (f-2) (f-3)   (widget/range.rs:64-73, widget/logic.rs:100-108)
template <class FR>
APB_D Fp<FR> gate_delta(const Fp<FR>& f) {
    typedef Fp<FR> F;
    const F one = F::one(), two = one + one, three = two + one;
    return f * (f - one) * (f - two) * (f - three);
}
template <class FR>
APB_D Fp<FR> small_const(uint32_t v) {          // v as a field element (Montgomery), by repeated addition
    typedef Fp<FR> F;
    F acc = F::zero(), cur = F::one();
    for (; v; v >>= 1) {
        if (v & 1) acc = acc + cur;
        cur = cur + cur;
    }
    return acc;
}
// Range::constraints (widget/range.rs:46-62)
template <class FR>
APB_D Fp<FR> gate_range(const Fp<FR>& sep, const Fp<FR>& a, const Fp<FR>& b, const Fp<FR>& c, const Fp<FR>& d, const Fp<FR>& d_next) {
    typedef Fp<FR> F;
    const F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa;
    auto four = [](const F& x) { F t = x + x; return t + t; };
    F b1 = gate_delta<FR>(c - four(d));
    F b2 = gate_delta<FR>(b - four(c)) * kappa;
    F b3 = gate_delta<FR>(a - four(b)) * kappa_sq;
    F b4 = gate_delta<FR>(d_next - four(a)) * kappa_cu;
    return (b1 + b2 + b3 + b4) * sep;
}
// Logic::constraints (widget/logic.rs:66-98) with delta_xor_and (:119-141)
template <class FR>
APB_D Fp<FR> gate_logic(const Fp<FR>& sep, const Fp<FR>& av, const Fp<FR>& bv, const Fp<FR>& cv, const Fp<FR>& dv, const Fp<FR>& a_next,
                        const Fp<FR>& b_next, const Fp<FR>& d_next, const Fp<FR>& q_c) {
    typedef Fp<FR> F;
    const F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa, kappa_qu = kappa_cu * kappa;
    auto four = [](const F& x) { F t = x + x; return t + t; };
    const F a = a_next - four(av), b = b_next - four(bv), d = d_next - four(dv), w = cv;
    F c0 = gate_delta<FR>(a);
    F c1 = gate_delta<FR>(b) * kappa;
    F c2 = gate_delta<FR>(d) * kappa_sq;
    F c3 = (w - a * b) * kappa_cu;
    const F k3 = small_const<FR>(3), k9 = small_const<FR>(9), k18 = small_const<FR>(18), k81 = small_const<FR>(81),
            k83 = small_const<FR>(83);
    const F ab = a + b;
    F Fv = w * (w * (four(w) - k18 * ab + k81) + k18 * (a.sqr() + b.sqr()) - k81 * ab + k83);
    F E = k3 * (ab + d) - (Fv + Fv);
    F Bv = q_c * (k9 * d - k3 * ab);
    F c4 = (Bv + E) * kappa_qu;
    return (c0 + c1 + c2 + c3 + c4) * sep;
}
// FixedBaseScalarMul::constraints (widget/ecc/fixed_base_scalar_mul.rs:88-156)
template <class FR>
APB_D Fp<FR> gate_fixed_base(const Fp<FR>& sep, const Fp<FR>& acc_x, const Fp<FR>& acc_y, const Fp<FR>& xy_alpha, const Fp<FR>& acc_bit,
                             const Fp<FR>& acc_x_next, const Fp<FR>& acc_y_next, const Fp<FR>& acc_bit_next, const Fp<FR>& q_l,
                             const Fp<FR>& q_r, const Fp<FR>& q_c, const Fp<FR>& coeff_a, const Fp<FR>& coeff_d) {
    typedef Fp<FR> F;
    const F kappa = sep.sqr(), kappa_sq = kappa.sqr(), kappa_cu = kappa_sq * kappa, one = F::one();
    const F bit = acc_bit_next - acc_bit - acc_bit;
    F bit_consistency = bit * (bit - one) * (bit + one);
    const F y_alpha = bit.sqr() * (q_r - one) + one;
    const F x_alpha = q_l * bit;
    F xy_consistency = (bit * q_c - xy_alpha) * kappa;
    const F t = xy_alpha * acc_x * acc_y * coeff_d;
    F x_acc = ((acc_x_next + acc_x_next * t) - (x_alpha * acc_y + y_alpha * acc_x)) * kappa_sq;
    F y_acc = ((acc_y_next - acc_y_next * t) - (y_alpha * acc_y - coeff_a * x_alpha * acc_x)) * kappa_cu;
    return (bit_consistency + x_acc + y_acc + xy_consistency) * sep;
}
// CurveAddition::constraints (widget/ecc/curve_addition.rs:62-96)
template <class FR>
APB_D Fp<FR> gate_curve_add(const Fp<FR>& sep, const Fp<FR>& x1, const Fp<FR>& y1, const Fp<FR>& x2, const Fp<FR>& y2, const Fp<FR>& x3,
                            const Fp<FR>& y3, const Fp<FR>& x1_y2, const Fp<FR>& coeff_a, const Fp<FR>& coeff_d) {
    typedef Fp<FR> F;
    const F kappa = sep.sqr();
    F xy_consistency = x1 * y2 - x1_y2;
    const F y1_x2 = y1 * x2, y1_y2 = y1 * y2, x1_x2 = x1 * x2;
    const F t = coeff_d * x1_y2 * y1_x2;
    F x3_consistency = ((x1_y2 + y1_x2) - (x3 + x3 * t)) * kappa;
    F y3_consistency = ((y1_y2 - coeff_a * x1_x2) - (y3 - y3 * t)) * kappa.sqr();
    return (xy_consistency + x3_consistency + y3_consistency) * sep;
}
template <class FR>
__global__ void __launch_bounds__(128) k_quotient(QuotientArgs A, void* out, uint64_t n4, uint64_t first, uint64_t count) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    i += first;                                      // this call evaluates points [first, first + count) of the 4n coset
    const uint64_t j = i + 4 >= n4 ? i + 4 - n4 : i + 4;
    const F alpha = arg_fp<FR>(A.alpha.v), beta = arg_fp<FR>(A.beta.v), gamma = arg_fp<FR>(A.gamma.v);
    const F a = load_fp<FR>(A.wl, i), b = load_fp<FR>(A.wr, i), c = load_fp<FR>(A.wo, i), d = load_fp<FR>(A.w4, i);
    // arithmetic gate + public inputs
    F gate = a * b * load_fp<FR>(A.q_m, i) + a * load_fp<FR>(A.q_l, i) + b * load_fp<FR>(A.q_r, i) +
             c * load_fp<FR>(A.q_o, i) + d * load_fp<FR>(A.q_4, i) + load_fp<FR>(A.q_c, i);
    gate = gate * load_fp<FR>(A.q_arith, i);
    if (A.pi) gate = gate + load_fp<FR>(A.pi, i);
    // custom gates (quotient_poly.rs:231-264): selector * constraints(separation challenge, wires, next-row wires)
    if (A.q_range || A.q_logic || A.q_fixed || A.q_var) {
        const F an = load_fp<FR>(A.wl, j), bn = load_fp<FR>(A.wr, j), dn = load_fp<FR>(A.w4, j);
        if (A.q_range) gate = gate + load_fp<FR>(A.q_range, i) * gate_range<FR>(arg_fp<FR>(A.range_sep.v), a, b, c, d, dn);
        if (A.q_logic)
            gate = gate + load_fp<FR>(A.q_logic, i) *
                              gate_logic<FR>(arg_fp<FR>(A.logic_sep.v), a, b, c, d, an, bn, dn, load_fp<FR>(A.q_c, i));
        if (A.q_fixed)
            gate = gate + load_fp<FR>(A.q_fixed, i) *
                              gate_fixed_base<FR>(arg_fp<FR>(A.fixed_sep.v), a, b, c, d, an, bn, dn, load_fp<FR>(A.q_l, i),
                                                  load_fp<FR>(A.q_r, i), load_fp<FR>(A.q_c, i), arg_fp<FR>(A.coeff_a.v),
                                                  arg_fp<FR>(A.coeff_d.v));
        if (A.q_var)
            gate = gate + load_fp<FR>(A.q_var, i) * gate_curve_add<FR>(arg_fp<FR>(A.var_sep.v), a, b, c, d, an, bn, dn,
                                                                         arg_fp<FR>(A.coeff_a.v), arg_fp<FR>(A.coeff_d.v));
    }
    // permutation
    const F zi = load_fp<FR>(A.z, i), zn = load_fp<FR>(A.z, j);
    const F l1 = load_fp<FR>(A.l1, i);
    F perm;
    {
        const F bx = beta * load_fp<FR>(A.linear, i);
        F ident = (a + bx + gamma) * (b + bx * arg_fp<FR>(A.k1.v) + gamma);
        ident = ident * (c + bx * arg_fp<FR>(A.k2.v) + gamma) * (d + bx * arg_fp<FR>(A.k3.v) + gamma);
        ident = ident * zi * alpha;
        F copy = (a + beta * load_fp<FR>(A.s1, i) + gamma) * (b + beta * load_fp<FR>(A.s2, i) + gamma);
        copy = copy * (c + beta * load_fp<FR>(A.s3, i) + gamma) * (d + beta * load_fp<FR>(A.s4, i) + gamma);
        copy = copy * zn * alpha;
        perm = ident - copy + (zi - F::one()) * (l1 * arg_fp<FR>(A.alpha_sq.v));
    }
    // lookup
    F look;
    {
        const F delta = arg_fp<FR>(A.delta.v), eps = arg_fp<FR>(A.epsilon.v), zeta = arg_fp<FR>(A.zeta.v);
        const F ls = arg_fp<FR>(A.lookup_sep.v), lsq = arg_fp<FR>(A.lsep_sq.v), lcu = arg_fp<FR>(A.lsep_cu.v);
        const F opd = delta + F::one(), eopd = arg_fp<FR>(A.eps_opd.v);
        const F fi = load_fp<FR>(A.f, i), ti = load_fp<FR>(A.table, i), tn = load_fp<FR>(A.table, j);
        const F h1i = load_fp<FR>(A.h1, i), h1n = load_fp<FR>(A.h1, j), h2i = load_fp<FR>(A.h2, i);
        const F z2i = load_fp<FR>(A.z2, i), z2n = load_fp<FR>(A.z2, j);
        const F comp = ((d * zeta + c) * zeta + b) * zeta + a;
        F la = load_fp<FR>(A.q_lookup, i) * (comp - fi) * ls;
        F lb = z2i * opd * (eps + fi) * (eopd + ti + delta * tn) * lsq;
        F lc = z2n * (eopd + h1i + delta * h2i) * (eopd + h2i + delta * h1n) * lsq;
        F ld = (z2i - F::one()) * l1 * lcu;
        look = la + lb - lc + ld;
    }
    F q = (gate + perm + look) * arg_fp<FR>(A.vh_inv[i & 3].v);
    store_fp<FR>(out, i, q);
}

// ---- polynomial evaluation (block-tree Horner) -----------------------------------------------
// Level kernel: values v[0..m) of polynomial `poly` in the variable xs; thread t folds E
// consecutive values by Horner, the block combines 128 threads with a tree using
// pw[l] = xs^(E * 2^l).  One output per block.
static const int EVAL_T = 128;
struct EvalArgs {
    const void* in[32];
    uint64_t m[32];
    uint64_t in_stride_blocks;      // outputs per polynomial in `out`
    Fr4 xs[32];
    Fr4 pw[32][7];
    uint32_t E;
};
template <class FR>
__global__ void __launch_bounds__(EVAL_T) k_eval_level(const EvalArgs* Ap, void* out) {
    const EvalArgs& A = *Ap;
    typedef Fp<FR> F;
    __shared__ uint4 sm[EVAL_T * 2];
    const uint32_t q = blockIdx.y, tid = threadIdx.x;
    const uint64_t m = A.m[q];
    const uint64_t start = ((uint64_t)blockIdx.x * EVAL_T + tid) * A.E;
    F acc = F::zero();
    if (start < m) {
        const F xs = arg_fp<FR>(A.xs[q].v);
        uint64_t end = start + A.E < m ? start + A.E : m;
        for (uint64_t k = end; k-- > start;) acc = acc * xs + load_fp<FR>(A.in[q], k);
    }
    store_fp<FR>(sm, tid, acc);
    __syncthreads();
    for (uint32_t l = 0, s = 1; s < EVAL_T; l++, s <<= 1) {
        if ((tid & (2 * s - 1)) == 0) {
            F hi = load_fp<FR>(sm, tid + s);
            acc = acc + hi * arg_fp<FR>(A.pw[q][l].v);
            store_fp<FR>(sm, tid, acc);
        }
        __syncthreads();
    }
    if (tid == 0) store_fp<FR>(out, (uint64_t)q * A.in_stride_blocks + blockIdx.x, acc);
}

// ---- division by (X - z): w[i] = zinv^(i+1) * sum_{j>i} p_j z^j ---------------------------------
template <class FR>
__global__ void k_mul_pointwise(const void* a, const void* b, void* out, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fp<FR>(out, i, load_fp<FR>(a, i) * load_fp<FR>(b, i));
}
// out[i] = base^(i + shift_add) with pow2[k] = base^(2^k)
template <class FR>
__global__ void k_pow_seq(void* out, uint64_t count, const void* pow2, uint64_t add) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    F acc = F::one();
    uint64_t e = i + add;
    for (int k = 0; e != 0; k++, e >>= 1)
        if (e & 1) acc = acc * load_fp<FR>(pow2, k);
    store_fp<FR>(out, i, acc);
}

}  // namespace apb

using namespace apb;

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PolyWs {
    void* buf[6];
    size_t cap[6];
    uint32_t* u32;
    size_t u32_cap;
};
static PolyWs g_ws;

static int ws_get(int slot, size_t bytes, void** out) {
    if (g_ws.cap[slot] < bytes) {
        if (g_ws.buf[slot]) cudaFree(g_ws.buf[slot]);
        g_ws.buf[slot] = nullptr;
        g_ws.cap[slot] = 0;
        cudaError_t e = cudaMalloc(&g_ws.buf[slot], bytes + 256);
        if (e != cudaSuccess) return set_err(APB_ERR_OOM, "workspace cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        g_ws.cap[slot] = bytes;
    }
    *out = g_ws.buf[slot];
    return APB_OK;
}
static int ws_u32(size_t count, uint32_t** out) {
    if (g_ws.u32_cap < count) {
        if (g_ws.u32) cudaFree(g_ws.u32);
        g_ws.u32 = nullptr;
        g_ws.u32_cap = 0;
        cudaError_t e = cudaMalloc((void**)&g_ws.u32, count * 4 + 256);
        if (e != cudaSuccess) return set_err(APB_ERR_OOM, "workspace cudaMalloc: %s", cudaGetErrorString(e));
        g_ws.u32_cap = count;
    }
    *out = g_ws.u32;
    return APB_OK;
}
static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }
static inline Fr4 mk4(const uint64_t* p) { Fr4 r; memcpy(r.v, p, 32); return r; }
static bool bad_curve(int c) { return c != APB_CURVE_BLS12_381 && c != APB_CURVE_BLS12_377; }

#define DISPATCH_FR(curve, CALL381, CALL377) do { if ((curve) == APB_CURVE_BLS12_381) { CALL381; } else { CALL377; } } while (0)

extern "C" int apb_fr_lincomb(int curve, size_t k, const void* const* d_polys, const size_t* lens, const uint64_t* scalars,
                              void* d_out, size_t out_len) {
    APB_API_LOCK();
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_fr_lincomb: bad curve");
    if (!d_out || (k && (!d_polys || !lens || !scalars))) return set_err(APB_ERR_INVALID_ARG, "apb_fr_lincomb: null argument");
    APB_REQUIRE_INIT();
    if (out_len == 0) return APB_OK;
    int accumulate = 0;
    size_t done = 0;
    do {
        LincombArgs A;
        memset(&A, 0, sizeof(A));
        A.k = (uint32_t)(k - done < (size_t)LINCOMB_MAX ? k - done : (size_t)LINCOMB_MAX);
        for (uint32_t j = 0; j < A.k; j++) {
            A.p[j] = d_polys[done + j];
            A.len[j] = lens[done + j] < out_len ? lens[done + j] : out_len;
            if (A.len[j] && !A.p[j]) return set_err(APB_ERR_INVALID_ARG, "apb_fr_lincomb: null polynomial");
            A.s[j] = mk4(scalars + 4 * (done + j));
        }
        DISPATCH_FR(curve, APB_KLAUNCH(k_lincomb<Fr381>, nblk(out_len, 128), 128, 0, A, d_out, (uint64_t)out_len, accumulate),
                    APB_KLAUNCH(k_lincomb<Fr377>, nblk(out_len, 128), 128, 0, A, d_out, (uint64_t)out_len, accumulate));
        done += A.k;
        accumulate = 1;
    } while (done < k);
    APB_CHECK_LAUNCH();
    return APB_OK;
}

// exclusive scan of `in` (n elements) into `out`; total -> d_total (one element).  OP 0 = product, 1 = sum.
template <class FR, int OP>
static int fr_scan(const void* in, void* out, size_t n, int reverse, void* d_total) {
    const size_t per = (size_t)SCAN_T * SCAN_E;
    const unsigned nb = (unsigned)((n + per - 1) / per);
    void* tot = nullptr;
    int rc = ws_get(5, (size_t)nb * 32 + 64, &tot);
    if (rc != APB_OK) return rc;
    auto kb = k_scan_block<FR, OP>;
    auto kt = k_scan_totals<FR, OP>;
    auto ka = k_scan_apply<FR, OP>;
    APB_KLAUNCH(kb, nb, SCAN_T, 0, in, out, tot, (uint64_t)n, reverse);
    APB_KLAUNCH(kt, 1, SCAN_T, 0, tot, nb, d_total);
    if (nb > 1) APB_KLAUNCH(ka, nb, SCAN_T, 0, out, (const void*)tot, (uint64_t)n, reverse);
    APB_CHECK_LAUNCH();
    return APB_OK;
}

// z from per-row numerators / denominators already in ws slots 0 (num) and 1 (den)
template <class FR>
static int grand_product(size_t n, void* d_z) {
    void *num = g_ws.buf[0], *den = g_ws.buf[1], *pn = nullptr, *sd = nullptr, *tot = nullptr;
    int rc;
    if ((rc = ws_get(2, n * 32, &pn)) != APB_OK) return rc;
    if ((rc = ws_get(3, n * 32, &sd)) != APB_OK) return rc;
    if ((rc = ws_get(4, 256, &tot)) != APB_OK) return rc;
    // z[i] = prod_{k<i} num_k / den_k  for i < n  (row n-1's own ratio is dropped, mod.rs:741-747):
    //      = PN[i] * (prod_{k>=i} den_k) / (prod_all den_k)
    if ((rc = fr_scan<FR, 0>(num, pn, n, 0, tot)) != APB_OK) return rc;
    if ((rc = fr_scan<FR, 0>(den, sd, n, 1, (char*)tot + 32)) != APB_OK) return rc;
    uint64_t h_tot[4], h_inv[4];
    APB_CUDA_TRY(cudaMemcpyAsync(h_tot, (char*)tot + 32, 32, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    host::Field f = host::Field::make<FR>();
    if (f.is_zero(h_tot)) return set_err(APB_ERR_INVALID_ARG, "grand product: zero denominator (inverse().unwrap() panics in the reference)");
    f.inv(h_inv, h_tot);
    APB_KLAUNCH(k_grand_product_finish<FR>, nblk(n, 128), 128, 0, (const void*)pn, (const void*)sd, (const void*)den, mk4(h_inv), d_z, (uint64_t)n);
    APB_CHECK_LAUNCH();
    return APB_OK;
}

extern "C" int apb_plonk_perm_z(apb_domain_t dom, const void* const* d_wires, const void* const* d_sigmas, const uint64_t* beta,
                                const uint64_t* gamma, void* d_z);
extern "C" int apb_domain_info(apb_domain_t d, int* curve, uint32_t* log_n, const void** tw);

extern "C" int apb_plonk_perm_z(apb_domain_t dom, const void* const* d_wires, const void* const* d_sigmas, const uint64_t* beta,
                                const uint64_t* gamma, void* d_z) {
    APB_API_LOCK();
    int curve; uint32_t log_n; const void* tw;
    int rc = apb_domain_info(dom, &curve, &log_n, &tw);
    if (rc != APB_OK) return rc;
    if (!d_wires || !d_sigmas || !beta || !gamma || !d_z) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_perm_z: null argument");
    const size_t n = (size_t)1 << log_n;
    void *num, *den;
    if ((rc = ws_get(0, n * 32, &num)) != APB_OK) return rc;
    if ((rc = ws_get(1, n * 32, &den)) != APB_OK) return rc;
    PermArgs A;
    memset(&A, 0, sizeof(A));
    for (int c = 0; c < 4; c++) { A.w[c] = d_wires[c]; A.s[c] = d_sigmas[c]; }
    A.tw = tw;
    A.n = n;
    A.beta = mk4(beta);
    A.gamma = mk4(gamma);
    const uint64_t ks[4] = {1, 7, 13, 17};       // permutation/constants.rs:12-22
    host::Field f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fr381>() : host::Field::make<Fr377>();
    for (int c = 0; c < 4; c++) {
        uint64_t t[4] = {ks[c], 0, 0, 0}, m[4];
        f.mul(m, t, f.r2);
        A.k[c] = mk4(m);
    }
    DISPATCH_FR(curve, APB_KLAUNCH(k_perm_terms<Fr381>, nblk(n, 128), 128, 0, A, num, den),
                APB_KLAUNCH(k_perm_terms<Fr377>, nblk(n, 128), 128, 0, A, num, den));
    APB_CHECK_LAUNCH();
    return curve == APB_CURVE_BLS12_381 ? grand_product<Fr381>(n, d_z) : grand_product<Fr377>(n, d_z);
}

extern "C" int apb_plonk_lookup_z2(apb_domain_t dom, const void* d_f, const void* d_t, const void* d_h1, const void* d_h2,
                                   const uint64_t* delta, const uint64_t* epsilon, void* d_z2) {
    APB_API_LOCK();
    int curve; uint32_t log_n; const void* tw;
    int rc = apb_domain_info(dom, &curve, &log_n, &tw);
    if (rc != APB_OK) return rc;
    if (!d_f || !d_t || !d_h1 || !d_h2 || !delta || !epsilon || !d_z2) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_lookup_z2: null argument");
    const size_t n = (size_t)1 << log_n;
    void *num, *den;
    if ((rc = ws_get(0, n * 32, &num)) != APB_OK) return rc;
    if ((rc = ws_get(1, n * 32, &den)) != APB_OK) return rc;
    LookupZArgs A;
    A.f = d_f; A.t = d_t; A.h1 = d_h1; A.h2 = d_h2; A.n = n;
    A.delta = mk4(delta);
    A.epsilon = mk4(epsilon);
    DISPATCH_FR(curve, APB_KLAUNCH(k_lookup_terms<Fr381>, nblk(n, 128), 128, 0, A, num, den),
                APB_KLAUNCH(k_lookup_terms<Fr377>, nblk(n, 128), 128, 0, A, num, den));
    APB_CHECK_LAUNCH();
    return curve == APB_CURVE_BLS12_381 ? grand_product<Fr381>(n, d_z2) : grand_product<Fr377>(n, d_z2);
}

extern "C" int apb_plonk_lookup_f(int curve, const void* q_lookup, const void* wl, const void* wr, const void* wo, const void* w4,
                                  const void* t_comp, const uint64_t* zeta, void* d_out, size_t n) {
    APB_API_LOCK();
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_lookup_f: bad curve");
    if (!q_lookup || !wl || !wr || !wo || !w4 || !t_comp || !zeta || !d_out) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_lookup_f: null argument");
    APB_REQUIRE_INIT();
    DISPATCH_FR(curve, APB_KLAUNCH(k_lookup_f<Fr381>, nblk(n, 128), 128, 0, q_lookup, wl, wr, wo, w4, t_comp, mk4(zeta), d_out, (uint64_t)n),
                APB_KLAUNCH(k_lookup_f<Fr377>, nblk(n, 128), 128, 0, q_lookup, wl, wr, wo, w4, t_comp, mk4(zeta), d_out, (uint64_t)n));
    APB_CHECK_LAUNCH();
    return APB_OK;
}

extern "C" int apb_plonk_combine_split(int curve, const void* d_t, const void* d_f, size_t n, void* d_h1, void* d_h2) {
    APB_API_LOCK();
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_combine_split: bad curve");
    if (!d_t || !d_f || !d_h1 || !d_h2) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_combine_split: null argument");
    APB_REQUIRE_INIT();
    if (n == 0) return APB_OK;
    uint32_t cap = 1;
    while (cap < 2 * n) cap <<= 1;
    const uint32_t mask = cap - 1;
    const size_t nb = (n + 1023) / 1024;
    // layout: owner[cap] first[cap] count[cap] | half[n] odd[n] oddpre[n+1] esz[n+1] osz[n+1] eoff[n+1] ooff[n+1] | blk[nb+1] | misc[8]
    uint32_t* w;
    int rc = ws_u32((size_t)3 * cap + 7 * (n + 1) + nb + 16, &w);
    if (rc != APB_OK) return rc;
    uint32_t *owner = w, *first = owner + cap, *count = first + cap, *half = count + cap, *odd = half + (n + 1),
             *oddpre = odd + (n + 1), *esz = oddpre + (n + 1), *osz = esz + (n + 1), *eoff = osz + (n + 1),
             *ooff = eoff + (n + 1), *blk = ooff + (n + 1), *misc = blk + nb + 1;
    APB_CUDA_TRY(cudaMemsetAsync(owner, 0xff, (size_t)2 * cap * 4, cur_stream()));       // owner, first = 0xffffffff
    APB_CUDA_TRY(cudaMemsetAsync(count, 0, (size_t)cap * 4, cur_stream()));
    APB_CUDA_TRY(cudaMemsetAsync(misc, 0, 64, cur_stream()));
    APB_KLAUNCH(k_cs_hash, nblk(n, 128), 128, 0, d_t, d_f, (uint64_t)n, mask, owner, first, count, 0, misc);
    APB_KLAUNCH(k_cs_hash, nblk(n, 128), 128, 0, d_t, d_f, (uint64_t)n, mask, owner, first, count, 1, misc);
    APB_KLAUNCH(k_cs_bucket_sizes, nblk(n, 128), 128, 0, d_t, (uint64_t)n, mask, (const uint32_t*)owner, (const uint32_t*)first,
                (const uint32_t*)count, half, odd);
    if ((rc = u32_scan(odd, oddpre, n, blk, misc + 1)) != APB_OK) return rc;
    APB_KLAUNCH(k_cs_split_sizes, nblk(n, 128), 128, 0, (const uint32_t*)half, (const uint32_t*)odd, (const uint32_t*)oddpre, esz, osz, (uint64_t)n);
    if ((rc = u32_scan(esz, eoff, n, blk, misc + 2)) != APB_OK) return rc;
    if ((rc = u32_scan(osz, ooff, n, blk, misc + 3)) != APB_OK) return rc;
    uint32_t h_misc[4];
    APB_CUDA_TRY(cudaMemcpyAsync(h_misc, misc, 16, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    if (h_misc[0]) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_combine_split: ElementNotIndexed (an element of f is not in t)");
    const uint32_t n_even = h_misc[2], n_odd = h_misc[3];
    if (n_even != n || n_odd != n)
        return set_err(APB_ERR_INVALID_ARG, "apb_plonk_combine_split: halves of %u / %u elements (expected %zu each)", n_even, n_odd, n);
    // sentinel offsets[n] = total so the fill kernel's search terminates
    APB_CUDA_TRY(cudaMemcpyAsync(eoff + n, misc + 2, 4, cudaMemcpyDeviceToDevice, cur_stream()));
    APB_CUDA_TRY(cudaMemcpyAsync(ooff + n, misc + 3, 4, cudaMemcpyDeviceToDevice, cur_stream()));
    APB_KLAUNCH(k_cs_fill, nblk(n, 128), 128, 0, d_t, (const uint32_t*)eoff, (uint64_t)n, d_h1, (uint64_t)n);
    APB_KLAUNCH(k_cs_fill, nblk(n, 128), 128, 0, d_t, (const uint32_t*)ooff, (uint64_t)n, d_h2, (uint64_t)n);
    APB_CHECK_LAUNCH();
    return APB_OK;
}

// args mirror struct QuotientArgs: 29 pointers (wl wr wo w4 z z2 f table h1 h2 pi | q_m q_l q_r q_o q_4 q_c q_arith
// q_lookup | s1 s2 s3 s4 linear l1 | q_range q_logic q_fixed q_var), 16 scalars (alpha beta gamma delta epsilon zeta
// lookup_sep K1 K2 K3 | range_sep logic_sep fixed_sep var_sep coeff_a coeff_d), 4 inverse vanishing values
extern "C" int apb_plonk_quotient_full(int curve, const void* const* ptrs29, const uint64_t* scalars16, const uint64_t* vh_inv4,
                                       void* d_out, size_t n4) {
    return apb_plonk_quotient_range(curve, ptrs29, scalars16, vh_inv4, d_out, n4, 0, n4);
}

// points [first, first + count) only (all vectors are still the full 4n-point ones: "next row" values wrap around):
// one proof split over several GPUs evaluates a slice per rank
extern "C" int apb_plonk_quotient_range(int curve, const void* const* ptrs29, const uint64_t* scalars16, const uint64_t* vh_inv4,
                                        void* d_out, size_t n4, size_t first, size_t count) {
    APB_API_LOCK();
    if (first > n4 || count > n4 - first) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_quotient: range outside the domain");
    if (count == 0) return APB_OK;
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_quotient: bad curve");
    if (!ptrs29 || !scalars16 || !vh_inv4 || !d_out) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_quotient: null argument");
    APB_REQUIRE_INIT();
    QuotientArgs A;
    const void** dst = &A.wl;
    for (int i = 0; i < 29; i++) {
        dst[i] = ptrs29[i];
        if (!ptrs29[i] && i != 10 && i < 25) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_quotient: null vector %d", i);
    }
    Fr4* sc = &A.alpha;
    for (int i = 0; i < 16; i++) sc[i] = mk4(scalars16 + 4 * i);
    for (int i = 0; i < 4; i++) A.vh_inv[i] = mk4(vh_inv4 + 4 * i);
    {   // challenge powers every point needs: once on the host instead of once per thread
        host::Field f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fr381>() : host::Field::make<Fr377>();
        uint64_t opd[4];
        f.sqr(A.alpha_sq.v, A.alpha.v);
        f.sqr(A.lsep_sq.v, A.lookup_sep.v);
        f.mul(A.lsep_cu.v, A.lsep_sq.v, A.lookup_sep.v);
        f.add(opd, A.delta.v, f.one);
        f.mul(A.eps_opd.v, A.epsilon.v, opd);
    }
    DISPATCH_FR(curve, APB_KLAUNCH(k_quotient<Fr381>, nblk(count, 128), 128, 0, A, d_out, (uint64_t)n4, (uint64_t)first, (uint64_t)count),
                APB_KLAUNCH(k_quotient<Fr377>, nblk(count, 128), 128, 0, A, d_out, (uint64_t)n4, (uint64_t)first, (uint64_t)count));
    APB_CHECK_LAUNCH();
    return APB_OK;
}

// arithmetic + permutation + lookup terms only (all custom gate selectors identically zero)
extern "C" int apb_plonk_quotient(int curve, const void* const* ptrs25, const uint64_t* scalars10, const uint64_t* vh_inv4, void* d_out,
                                  size_t n4) {
    if (!ptrs25 || !scalars10) return set_err(APB_ERR_INVALID_ARG, "apb_plonk_quotient: null argument");
    const void* p[29];
    uint64_t sc[16 * 4];
    for (int i = 0; i < 25; i++) p[i] = ptrs25[i];
    for (int i = 25; i < 29; i++) p[i] = nullptr;
    memset(sc, 0, sizeof(sc));
    memcpy(sc, scalars10, 10 * 32);
    return apb_plonk_quotient_full(curve, p, sc, vh_inv4, d_out, n4);
}

// k evaluations p_j(x_j); results (Montgomery) in host memory; blocking
template <class FR>
static int poly_eval_impl(size_t k, const void* const* d_polys, const size_t* lens, const uint64_t* points, uint64_t* out_vals) {
    host::Field f = host::Field::make<FR>();
    const uint32_t E1 = 16;
    size_t max_len = 0;
    for (size_t j = 0; j < k; j++) if (lens[j] > max_len) max_len = lens[j];
    if (max_len == 0) { memset(out_vals, 0, k * 32); return APB_OK; }
    for (size_t done = 0; done < k; done += 32) {
        const uint32_t kk = (uint32_t)(k - done < 32 ? k - done : 32);
        // level 1
        const size_t per1 = (size_t)EVAL_T * E1;
        const size_t nb1 = (max_len + per1 - 1) / per1;
        void *lvl1 = nullptr, *lvl2 = nullptr, *dargs = nullptr;
        int rc;
        if ((rc = ws_get(2, (size_t)kk * nb1 * 32 + 64, &lvl1)) != APB_OK) return rc;
        if ((rc = ws_get(3, (size_t)kk * 32 + 64, &lvl2)) != APB_OK) return rc;
        if ((rc = ws_get(0, 2 * sizeof(EvalArgs) + 64, &dargs)) != APB_OK) return rc;   // argument blocks live in HBM
        EvalArgs A;
        memset(&A, 0, sizeof(A));
        auto fill_powers = [&](EvalArgs& B, uint32_t q, const uint64_t* xs, uint32_t E) {
            B.xs[q] = mk4(xs);
            uint64_t cur[4];
            f.set(cur, f.one);
            // cur = xs^E
            uint64_t base[4];
            f.set(base, xs);
            for (uint32_t e = E; e; e >>= 1) {
                if (e & 1) f.mul(cur, cur, base);
                f.sqr(base, base);
            }
            for (int l = 0; l < 7; l++) {
                B.pw[q][l] = mk4(cur);
                f.sqr(cur, cur);
            }
        };
        std::vector<uint64_t> xs2(4 * kk);
        for (uint32_t q = 0; q < kk; q++) {
            A.in[q] = d_polys[done + q];
            A.m[q] = lens[done + q];
            fill_powers(A, q, points + 4 * (done + q), E1);
            // xs2 = x^(EVAL_T * E1)  = pw[6]^2
            uint64_t t[4];
            memcpy(t, A.pw[q][6].v, 32);
            f.sqr(t, t);
            memcpy(&xs2[4 * q], t, 32);
        }
        A.E = E1;
        A.in_stride_blocks = nb1;
        APB_CUDA_TRY(cudaMemcpyAsync(dargs, &A, sizeof(EvalArgs), cudaMemcpyHostToDevice, cur_stream()));
        APB_KLAUNCH(k_eval_level<FR>, dim3((unsigned)nb1, kk), EVAL_T, 0, (const EvalArgs*)dargs, lvl1);
        // level 2: nb1 partials per polynomial in the variable x^(2048); one block each
        if (nb1 > 1) {
            const uint32_t E2 = (uint32_t)((nb1 + EVAL_T - 1) / EVAL_T);
            EvalArgs B;
            memset(&B, 0, sizeof(B));
            for (uint32_t q = 0; q < kk; q++) {
                B.in[q] = (const char*)lvl1 + (size_t)q * nb1 * 32;
                B.m[q] = nb1;
                fill_powers(B, q, &xs2[4 * q], E2);
            }
            B.E = E2;
            B.in_stride_blocks = 1;
            EvalArgs* dB = (EvalArgs*)dargs + 1;
            APB_CUDA_TRY(cudaMemcpyAsync(dB, &B, sizeof(EvalArgs), cudaMemcpyHostToDevice, cur_stream()));
            APB_KLAUNCH(k_eval_level<FR>, dim3(1, kk), EVAL_T, 0, (const EvalArgs*)dB, lvl2);
        }
        APB_CHECK_LAUNCH();
        APB_CUDA_TRY(cudaMemcpyAsync(out_vals + 4 * done, nb1 > 1 ? lvl2 : lvl1, (size_t)kk * 32, cudaMemcpyDeviceToHost, cur_stream()));
        APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    }
    return APB_OK;
}

extern "C" int apb_poly_eval(int curve, size_t k, const void* const* d_polys, const size_t* lens, const uint64_t* points,
                             uint64_t* out_vals) {
    APB_API_LOCK();
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_poly_eval: bad curve");
    if (k == 0) return APB_OK;
    if (!d_polys || !lens || !points || !out_vals) return set_err(APB_ERR_INVALID_ARG, "apb_poly_eval: null argument");
    APB_REQUIRE_INIT();
    return curve == APB_CURVE_BLS12_381 ? poly_eval_impl<Fr381>(k, d_polys, lens, points, out_vals)
                                        : poly_eval_impl<Fr377>(k, d_polys, lens, points, out_vals);
}

// w = p / (X - z) (remainder dropped): len-1 coefficients
template <class FR>
static int divide_impl(const void* d_p, size_t len, const uint64_t* z, void* d_out) {
    host::Field f = host::Field::make<FR>();
    if (len <= 1) return APB_OK;
    if (f.is_zero(z)) {      // p / X: shift
        APB_CUDA_TRY(cudaMemcpyAsync(d_out, (const char*)d_p + 32, (len - 1) * 32, cudaMemcpyDeviceToDevice, cur_stream()));
        return APB_OK;
    }
    uint64_t zinv[4];
    f.inv(zinv, z);
    uint64_t h_pow2[2 * 64 * 4];
    uint64_t cur[4], curi[4];
    f.set(cur, z);
    f.set(curi, zinv);
    for (int k = 0; k < 64; k++) {
        memcpy(h_pow2 + 4 * k, cur, 32);
        memcpy(h_pow2 + 4 * (64 + k), curi, 32);
        f.sqr(cur, cur);
        f.sqr(curi, curi);
    }
    void *pw = nullptr, *a = nullptr, *b = nullptr, *tot = nullptr;
    int rc;
    if ((rc = ws_get(4, 2 * 64 * 32 + 256, &pw)) != APB_OK) return rc;
    if ((rc = ws_get(0, len * 32, &a)) != APB_OK) return rc;
    if ((rc = ws_get(1, len * 32, &b)) != APB_OK) return rc;
    tot = (char*)pw + 2 * 64 * 32;
    APB_CUDA_TRY(cudaMemcpyAsync(pw, h_pow2, sizeof(h_pow2), cudaMemcpyHostToDevice, cur_stream()));
    // a[j] = z^j ; b[j] = p_j z^j ; a <- exclusive suffix sums of b: a[i] = sum_{j>i} p_j z^j
    APB_KLAUNCH(k_pow_seq<FR>, nblk(len, 128), 128, 0, a, (uint64_t)len, (const void*)pw, (uint64_t)0);
    APB_KLAUNCH(k_mul_pointwise<FR>, nblk(len, 128), 128, 0, d_p, (const void*)a, b, (uint64_t)len);
    if ((rc = fr_scan<FR, 1>(b, a, len, 1, tot)) != APB_OK) return rc;
    // b[i] = zinv^(i+1) ; out[i] = a[i] * b[i], i < len-1
    APB_KLAUNCH(k_pow_seq<FR>, nblk(len - 1, 128), 128, 0, b, (uint64_t)(len - 1), (const void*)((char*)pw + 64 * 32), (uint64_t)1);
    APB_KLAUNCH(k_mul_pointwise<FR>, nblk(len - 1, 128), 128, 0, (const void*)a, (const void*)b, d_out, (uint64_t)(len - 1));
    APB_CHECK_LAUNCH();
    return APB_OK;
}

extern "C" int apb_poly_divide_linear(int curve, const void* d_p, size_t len, const uint64_t* z, void* d_out) {
    APB_API_LOCK();
    if (bad_curve(curve)) return set_err(APB_ERR_INVALID_ARG, "apb_poly_divide_linear: bad curve");
    if (!z || (len > 1 && (!d_p || !d_out))) return set_err(APB_ERR_INVALID_ARG, "apb_poly_divide_linear: null argument");
    APB_REQUIRE_INIT();
    return curve == APB_CURVE_BLS12_381 ? divide_impl<Fr381>(d_p, len, z, d_out) : divide_impl<Fr377>(d_p, len, z, d_out);
}
