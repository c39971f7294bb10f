// MSM kernels: commitment-key setup (precomputed multiples, powers of tau).
#include "msm_kernels.cuh"

namespace apb {

// a^(p-2)
template <class FQ>
__device__ __noinline__ Fp<FQ> fp_inverse(const Fp<FQ>& a) {
    Fp<FQ> acc = Fp<FQ>::one(), base = a;
    uint32_t e[FQ::N];
#pragma unroll
    for (int i = 0; i < FQ::N; i++) e[i] = FQ::mod(i);
    {   // e = p - 2 with borrow propagation (the low limb of the BLS12-377 modulus is 1)
        uint32_t borrow = 2;
        for (int i = 0; i < FQ::N && borrow; i++) {
            uint32_t nb = e[i] < borrow ? 1u : 0u;
            e[i] -= borrow;
            borrow = nb;
        }
    }
    for (int i = 0; i < 32 * FQ::N; i++) {
        if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * base;
        base = base.sqr();
    }
    return acc;
}

// copies[f*n + i] = 2^(step*f) * P_i as affine points, f = 0..F-1 (copy 0 is the input itself)
template <class FQ>
__global__ void __launch_bounds__(128) k_ck_precompute(void* bases, uint64_t n, uint32_t F, uint32_t step) {
    typedef Fp<FQ> Fe;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe px, py;
    load_affine<FQ>(bases, i, px, py);
    if (px.is_zero() && py.is_zero()) {
        for (uint32_t f = 1; f < F; f++) {
            store_fp<FQ>(bases, 2 * (f * n + i), px);
            store_fp<FQ>(bases, 2 * (f * n + i) + 1, py);
        }
        return;
    }
    XYZZ<FQ> pts[MAX_COPIES];
    Fe prefix[MAX_COPIES];
    XYZZ<FQ> cur;
    cur.x = px; cur.y = py; cur.zz = Fe::one(); cur.zzz = Fe::one();
    Fe run = Fe::one();
    for (uint32_t f = 1; f < F; f++) {
        for (uint32_t s = 0; s < step; s++) cur = cur.dbl();
        pts[f] = cur;
        prefix[f] = run;             // product of zzz of copies 1..f-1
        run = run * cur.zzz;
    }
    Fe inv = fp_inverse<FQ>(run);
    for (uint32_t f = F - 1; f >= 1; f--) {
        Fe zinv = inv * prefix[f];   // 1 / zzz_f
        inv = inv * pts[f].zzz;
        Fe r = pts[f].zz * zinv;     // zz/zzz = 1/z
        Fe ax = pts[f].x * r.sqr();
        Fe ay = pts[f].y * zinv;
        store_fp<FQ>(bases, 2 * (f * n + i), ax);
        store_fp<FQ>(bases, 2 * (f * n + i) + 1, ay);
    }
}

// bases[i] = [tau^i] G as affine points: powers[i] holds tau^i (Fr, Montgomery)
template <class CV>
__global__ void __launch_bounds__(128) k_srs_powers(void* bases, const void* powers, uint64_t n, const void* gen_xy) {
    typedef typename CV::FQ FQ;
    typedef typename CV::FR FR;
    typedef Fp<FQ> Fe;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<FR> s = load_fp<FR>(powers, i).from_mont();
    Fe gx = load_fp<FQ>(gen_xy, 0), gy = load_fp<FQ>(gen_xy, 1);
    XYZZ<FQ> acc = XYZZ<FQ>::identity();
    for (int bit = FR::BITS - 1; bit >= 0; bit--) {
        acc = acc.dbl();
        if ((s.v[bit >> 5] >> (bit & 31)) & 1) acc.add_affine(gx, gy);
    }
    Fe ax = Fe::zero(), ay = Fe::zero();
    if (!acc.is_identity()) {
        Fe zinv = fp_inverse<FQ>(acc.zzz);
        Fe r = acc.zz * zinv;
        ax = acc.x * r.sqr();
        ay = acc.y * zinv;
    }
    store_fp<FQ>(bases, 2 * i, ax);
    store_fp<FQ>(bases, 2 * i + 1, ay);
}
template <class FR>
__global__ void k_tau_powers(void* out, uint64_t count, const void* pow2) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    F acc = F::one();
    uint64_t e = i;
    for (int k = 0; e != 0; k++, e >>= 1)
        if (e & 1) acc = acc * load_fp<FR>(pow2, k);
    store_fp<FR>(out, i, acc);
}

void msm_launch_ck_precompute(int curve, void* bases, uint64_t n, uint32_t F, uint32_t step) {
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (curve == APB_CURVE_BLS12_381) APB_KLAUNCH(k_ck_precompute<Fq381>, blocks, 128, 0, bases, n, F, step);
    else APB_KLAUNCH(k_ck_precompute<Fq377>, blocks, 128, 0, bases, n, F, step);
}
void msm_launch_srs_powers(int curve, void* bases, void* d_powers, uint64_t n, const void* d_pow2, const void* d_gen) {
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (curve == APB_CURVE_BLS12_381) {
        APB_KLAUNCH(k_tau_powers<Fr381>, blocks, 128, 0, d_powers, n, d_pow2);
        APB_KLAUNCH(k_srs_powers<Curve381>, blocks, 128, 0, bases, (const void*)d_powers, n, d_gen);
    } else {
        APB_KLAUNCH(k_tau_powers<Fr377>, blocks, 128, 0, d_powers, n, d_pow2);
        APB_KLAUNCH(k_srs_powers<Curve377>, blocks, 128, 0, bases, (const void*)d_powers, n, d_gen);
    }
}

}  // namespace apb
