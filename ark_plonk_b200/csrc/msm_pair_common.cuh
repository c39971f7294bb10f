// Helpers of the pair-level kernel (msm_pairs_coop.cu).
#pragma once
#include "msm_kernels.cuh"

namespace apb {

static const uint32_t NO_PARTNER = 0xffffffffu;

// One input of a pair level.  FIRST: `e` is an entry of the bucket-sorted list (table index, sign in
// bit 31; (0,0) in the table = infinity).  Otherwise `e` is an index into the previous level's array
// (infinity = all-ones top limb of x).
template <class FQ, int FIRST>
APB_D void pair_load_xy(const void* src, uint32_t e, Fp<FQ>& x, Fp<FQ>& y) {
    const uint64_t idx = FIRST ? (e & 0x7fffffffu) : e;
    x = load_fp_early<FQ, 1>(src, 2 * idx);
    y = load_fp_early<FQ, 1>(src, 2 * idx + 1);
}
template <class FQ, int FIRST>
APB_D bool pair_fix(uint32_t e, const Fp<FQ>& x, Fp<FQ>& y) {      // applies the sign, returns "is infinity"
    if (FIRST) {
        const bool inf = x.is_zero() && y.is_zero();
        if (e >> 31) y = y.neg();
        return inf;
    }
    return x.v[FQ::N - 1] == 0xffffffffu;
}
// rare path of the denominator pass: an operand at infinity, or equal x (doubling / inverse pair)
template <class FQ, int FIRST>
__device__ __noinline__ Fp<FQ> pair_den_special(const void* src, uint32_t e1, uint32_t e2) {
    Fp<FQ> x1, y1, x2, y2;
    pair_load_xy<FQ, FIRST>(src, e1, x1, y1);
    pair_load_xy<FQ, FIRST>(src, e2, x2, y2);
    const bool inf1 = pair_fix<FQ, FIRST>(e1, x1, y1), inf2 = pair_fix<FQ, FIRST>(e2, x2, y2);
    if (inf1 || inf2) return Fp<FQ>::one();
    if (x1 != x2) return x2 - x1;
    if (y1 == y2 && !y1.is_zero()) return y1 + y1;       // doubling: lambda = 3 x^2 / 2 y
    return Fp<FQ>::one();                                 // P + (-P) (or a 2-torsion point): infinity
}


}  // namespace apb
