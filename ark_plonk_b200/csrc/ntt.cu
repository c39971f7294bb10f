// Radix-2^k number-theoretic transforms over Fr (BLS12-381 / BLS12-377) for sm_100a.
//
// Replaces ark_poly 0.3 Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft} as called
// by plonk-core (prover.rs:197-203,241,282,303,305; quotient_poly.rs:72-120,176,205,294,325;
// permutation/mod.rs:199-205,671-674,751,800; pi.rs:115; lookup/multiset.rs:201;
// preprocess.rs:145-210,304-340).  Same semantics: input zero-extended to N, natural order in
// and out, coset shift g = Fr::multiplicative_generator(), inverse scaled by 1/N.
//
// Algorithm (not arkworks' in-place radix-2 + bit-reversal): an autosorting mixed-radix
// decomposition N = N1*N2(*N3).  Pass p transforms digit p of the index inside shared memory
// (tile = N_p points x C adjacent columns so that every global access is a run of C*32 B),
// multiplies by the inter-pass twiddle w_N^(lo * k_p * M_p) from the resident table, and the
// last pass writes to the digit-reversed position, so the result lands in natural order
// without a separate permutation pass.  The coset pre-scale (g^i, only on the in_len supplied
// coefficients), zero-extension (no loads of implied zeros), 1/N and g^-i post-scales are
// fused into the first / last pass.  HBM traffic: P reads + P writes of the vector (P = 2 up
// to 2^20, 3 up to 2^30); for N <= 2^21 the vector is L2-resident between passes.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "host_ec.hpp"

namespace apb {

extern int g_num_sms;
extern int g_profile;
double g_ntt_ms_total = 0.0;              // device time of transforms while profiling (apb_set_profiling)
unsigned long long g_ntt_count = 0;
struct NttPending { cudaEvent_t a, b; unsigned long long batch; };
static std::vector<NttPending> g_ntt_pending;        // recorded, not yet read (see apb_ntt_batch_dev)
static std::mutex g_ntt_pending_mu;
static void ntt_resolve_pending() {                  // caller holds g_ntt_pending_mu
    for (NttPending& p : g_ntt_pending) {
        float ms = 0;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            g_ntt_ms_total += ms;
            g_ntt_count += p.batch;
        }
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_ntt_pending.clear();
}
static const int COSET_LO_BITS = 10;

struct NttPassArgs {
    const void* in;
    void* out;
    uint64_t in_batch_stride, out_batch_stride;   // elements
    uint32_t log_n, log_t, log_c;
    uint64_t na, in_a, in_b, out_a, out_b;
    uint64_t in_sd, in_sc, out_sk, out_sc;
    const void* roots;        // w^i (or w^-i), i < N/2
    uint64_t tw_mult;         // inter-pass twiddle exponent multiplier; 0 = none
    uint64_t in_len;          // first pass: elements >= in_len are zero (no load); else ~0
    const void* pre_lo;       // first pass coset pre-scale tables (g^j, g^(j<<10)) or null
    const void* pre_hi;
    const void* post_lo;      // last pass post-scale: lo/hi tables (coset_ifft) ...
    const void* post_hi;
    int post_const;           // ... or the constant 1/N (ifft)
    uint32_t size_inv[8];
    // k_ntt_pass_reg: the tile's log_t butterfly stages grouped into rounds of up to LR.  In a round every thread holds
    // the 2^LR elements whose flat tile index (d * C + c) differs in bits [lb, lb + LR); bit ob of `mask` set = own bit
    // ob is a butterfly stage of this round (high to low).
    uint32_t nrounds;
    uint32_t round_lb[8], round_mask[8];
};

template <class FR>
APB_D Fp<FR> smem_get(const uint4* lo, const uint4* hi, uint32_t i) {
    Fp<FR> r;
    uint4 a = lo[i], b = hi[i];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class FR>
APB_D void smem_put(uint4* lo, uint4* hi, uint32_t i, const Fp<FR>& a) {
    lo[i] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    hi[i] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// One pass over one tile.  Shared memory: data planes (2 x T*C uint4) + root planes (2 x T/2).
template <class FR>
__global__ void __launch_bounds__(512) k_ntt_pass(NttPassArgs A) {
    typedef Fp<FR> F;
    APB_DYN_SMEM(smem);
    const uint32_t T = 1u << A.log_t, C = 1u << A.log_c, TC = T << A.log_c;
    uint4* xlo = reinterpret_cast<uint4*>(smem);
    uint4* xhi = xlo + TC;
    uint4* wlo = xhi + TC;
    uint4* whi = wlo + (T >> 1);
    const uint32_t tid = threadIdx.x, nth = blockDim.x;
    const uint64_t tile = blockIdx.x;
    const uint64_t ta = tile % A.na, tb = tile / A.na;
    const uint64_t base_in = ta * A.in_a + tb * A.in_b;
    const uint64_t base_out = ta * A.out_a + tb * A.out_b;
    const uint4* in = reinterpret_cast<const uint4*>(A.in) + 2 * A.in_batch_stride * blockIdx.y;
    uint4* out = reinterpret_cast<uint4*>(A.out) + 2 * A.out_batch_stride * blockIdx.y;

    // roots of the tile transform: w_T^j = w_N^(j * N/T)
    for (uint32_t j = tid; j < (T >> 1); j += nth) {
        F w = load_fp<FR>(A.roots, (uint64_t)j << (A.log_n - A.log_t));
        smem_put<FR>(wlo, whi, j, w);
    }
    // load (zero-extension and coset pre-scale fused)
    for (uint32_t e = tid; e < TC; e += nth) {
        uint32_t c = e & (C - 1), d = e >> A.log_c;
        uint64_t idx = base_in + d * A.in_sd + c * A.in_sc;
        F v;
        if (idx < A.in_len) {
            v = load_fp<FR>(in, idx);
            if (A.pre_lo) {
                F s = load_fp<FR>(A.pre_lo, idx & ((1u << COSET_LO_BITS) - 1)) * load_fp<FR>(A.pre_hi, idx >> COSET_LO_BITS);
                v = v * s;
            }
        } else {
            v = F::zero();
        }
        smem_put<FR>(xlo, xhi, e, v);
    }
    __syncthreads();

    // decimation-in-frequency stages; result index k sits at bit-reversed row
    for (uint32_t lh = A.log_t; lh-- > 0;) {
        const uint32_t h = 1u << lh;
        for (uint32_t b = tid; b < (TC >> 1); b += nth) {
            uint32_t c = b & (C - 1), r = b >> A.log_c;
            uint32_t j = r & (h - 1), blk = r >> lh;
            uint32_t i0 = (((blk << (lh + 1)) + j) << A.log_c) + c;
            uint32_t i1 = i0 + (h << A.log_c);
            F u = smem_get<FR>(xlo, xhi, i0), v = smem_get<FR>(xlo, xhi, i1);
            smem_put<FR>(xlo, xhi, i0, u + v);
            F t = u - v;
            if (lh > 0) t = t * smem_get<FR>(wlo, whi, j << (A.log_t - 1 - lh));
            smem_put<FR>(xlo, xhi, i1, t);
        }
        __syncthreads();
    }

    // store (inter-pass twiddle / final scaling fused)
    const uint64_t half_n = (uint64_t)1 << (A.log_n - 1);
    for (uint32_t e = tid; e < TC; e += nth) {
        uint32_t c = e & (C - 1), k = e >> A.log_c;
        uint32_t row = A.log_t ? (__brev(k) >> (32 - A.log_t)) : 0;
        F v = smem_get<FR>(xlo, xhi, (row << A.log_c) + c);
        uint64_t oidx = base_out + k * A.out_sk + c * A.out_sc;
        if (A.tw_mult) {
            uint64_t ex = (ta * C + c) * k * A.tw_mult;
            if (ex != 0) {
                F w = (ex < half_n) ? load_fp<FR>(A.roots, ex) : load_fp<FR>(A.roots, ex - half_n).neg();
                v = v * w;
            }
        }
        if (A.post_lo) {
            F s = load_fp<FR>(A.post_lo, oidx & ((1u << COSET_LO_BITS) - 1)) * load_fp<FR>(A.post_hi, oidx >> COSET_LO_BITS);
            v = v * s;
        } else if (A.post_const) {
            F s;
#pragma unroll
            for (int i = 0; i < 8; i++) s.v[i] = A.size_inv[i];
            v = v * s;
        }
        store_fp<FR>(out, oidx, v);
    }
}

// ---- register-resident pass: R = 4 elements per thread, two stages between shared-memory exchanges ----------------------
// The radix-2 kernel above (kept for tiles of fewer than 4 elements) does one stage per shared-memory round trip and CTA
// barrier.  Here a thread keeps R elements in registers and runs log2(R) stages on them before the tile is exchanged
// through shared memory once (one barrier per ROUND, not per stage); the first round reads global memory straight into
// registers and the last one writes straight from them.  Shared-memory positions are XOR-swizzled (the three 3-bit
// fields above bit 3 folded into the low three bits), so the 8 lanes of every LDS.128 / STS.128 phase hit 8 different
// 16-byte bank groups whatever stride the round uses - data planes and the tile's root table alike.  In the LAST round of
// a tile (own bits = the lowest tile bits) the twiddles depend only on the register index: w_4 is loaded once and three
// of the four butterflies need no product.
APB_D uint32_t ntt_swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u) ^ ((i >> 9) & 7u); }

template <class FR>
APB_D void ntt_bfly(Fp<FR>& u, Fp<FR>& v) {          // (u, v) <- (u + v, u - v)
    const Fp<FR> t = u - v;
    u = u + v;
    v = t;
}

template <class FR, int LR>
__global__ void __launch_bounds__(256) k_ntt_pass_reg(NttPassArgs A) {
    constexpr int R = 1 << LR;                       // elements per thread
    typedef Fp<FR> F;
    APB_DYN_SMEM(smem);
    const uint32_t T = 1u << A.log_t, C = 1u << A.log_c, TC = T << A.log_c;
    uint4* xlo = reinterpret_cast<uint4*>(smem);
    uint4* xhi = xlo + TC;
    uint4* wlo = xhi + TC;
    uint4* whi = wlo + (T >> 1);
    const uint32_t tid = threadIdx.x, nth = blockDim.x;             // nth == TC / R
    const uint64_t tile = blockIdx.x;
    const uint64_t ta = tile % A.na, tb = tile / A.na;
    const uint64_t base_in = ta * A.in_a + tb * A.in_b;
    const uint64_t base_out = ta * A.out_a + tb * A.out_b;
    const uint4* in = reinterpret_cast<const uint4*>(A.in) + 2 * A.in_batch_stride * blockIdx.y;
    uint4* out = reinterpret_cast<uint4*>(A.out) + 2 * A.out_batch_stride * blockIdx.y;

    // roots of the tile transform: w_T^j = w_N^(j * N/T)
    for (uint32_t j = tid; j < (T >> 1); j += nth) {
        F w = load_fp<FR>(A.roots, (uint64_t)j << (A.log_n - A.log_t));
        smem_put<FR>(wlo, whi, ntt_swz(j), w);
    }
    F e[R];
    uint32_t lb = A.round_lb[0], mask = A.round_mask[0];
    uint32_t i0 = ((tid >> lb) << (lb + LR)) | (tid & ((1u << lb) - 1));
    // first round: global -> registers (zero-extension and coset pre-scale fused)
#pragma unroll
    for (int k = 0; k < R; k++) {
        const uint32_t i = i0 + ((uint32_t)k << lb), c = i & (C - 1), d = i >> A.log_c;
        const uint64_t idx = base_in + d * A.in_sd + c * A.in_sc;
        if (idx < A.in_len) {
            e[k] = load_fp<FR>(in, idx);
            if (A.pre_lo) {
                F s = load_fp<FR>(A.pre_lo, idx & ((1u << COSET_LO_BITS) - 1)) * load_fp<FR>(A.pre_hi, idx >> COSET_LO_BITS);
                e[k] = e[k] * s;
            }
        } else {
            e[k] = F::zero();
        }
    }
    __syncthreads();                                 // root table complete
    for (uint32_t r = 0;; r++) {
        // butterflies of this round, own bits high to low
        if (lb == A.log_c) {
            // bottom-aligned window (always all LR stages, halves R/2 .. 1): twiddle index = (k mod half) * T / (2 half),
            // the same in every thread
            static_assert(LR == 2, "constant twiddles of the last round are written out for 4 elements per thread");
            const F w2 = smem_get<FR>(wlo, whi, ntt_swz(T >> 2));
            ntt_bfly<FR>(e[0], e[2]);
            ntt_bfly<FR>(e[1], e[3]); e[3] = e[3] * w2;
            ntt_bfly<FR>(e[0], e[1]);
            ntt_bfly<FR>(e[2], e[3]);
        } else {
#pragma unroll
            for (int ob = LR - 1; ob >= 0; ob--) {
                if (!((mask >> ob) & 1u)) continue;
                const uint32_t lh = lb + ob - A.log_c;                   // stage half = 2^lh (in tile indices)
#pragma unroll
                for (int k = 0; k < R; k++) {
                    if (k & (1 << ob)) continue;
                    ntt_bfly<FR>(e[k], e[k | (1 << ob)]);
                    if (lh > 0) {
                        const uint32_t j = ((i0 + ((uint32_t)k << lb)) >> A.log_c) & ((1u << lh) - 1);
                        e[k | (1 << ob)] = e[k | (1 << ob)] * smem_get<FR>(wlo, whi, ntt_swz(j << (A.log_t - 1 - lh)));
                    }
                }
            }
        }
        if (r + 1 == A.nrounds) break;
        // exchange: every thread writes its 8 positions, then picks up the 8 positions of the next window
#pragma unroll
        for (int k = 0; k < R; k++) smem_put<FR>(xlo, xhi, ntt_swz(i0 + ((uint32_t)k << lb)), e[k]);
        __syncthreads();
        lb = A.round_lb[r + 1];
        mask = A.round_mask[r + 1];
        i0 = ((tid >> lb) << (lb + LR)) | (tid & ((1u << lb) - 1));
#pragma unroll
        for (int k = 0; k < R; k++) e[k] = smem_get<FR>(xlo, xhi, ntt_swz(i0 + ((uint32_t)k << lb)));
    }

    // registers -> global: position d of the in-place transform holds output bitrev(d) (inter-pass twiddle / final scaling fused)
    const uint64_t half_n = (uint64_t)1 << (A.log_n - 1);
#pragma unroll
    for (int k = 0; k < R; k++) {
        const uint32_t i = i0 + ((uint32_t)k << lb), c = i & (C - 1), d = i >> A.log_c;
        const uint32_t ko = A.log_t ? (__brev(d) >> (32 - A.log_t)) : 0;
        F v = e[k];
        const uint64_t oidx = base_out + ko * A.out_sk + c * A.out_sc;
        if (A.tw_mult) {
            const uint64_t ex = (ta * C + c) * ko * A.tw_mult;
            if (ex != 0) {
                F w = (ex < half_n) ? load_fp<FR>(A.roots, ex) : load_fp<FR>(A.roots, ex - half_n).neg();
                v = v * w;
            }
        }
        if (A.post_lo) {
            F s = load_fp<FR>(A.post_lo, oidx & ((1u << COSET_LO_BITS) - 1)) * load_fp<FR>(A.post_hi, oidx >> COSET_LO_BITS);
            v = v * s;
        } else if (A.post_const) {
            F s;
#pragma unroll
            for (int q = 0; q < 8; q++) s.v[q] = A.size_inv[q];
            v = v * s;
        }
        store_fp<FR>(out, oidx, v);
    }
}

// out[i] = scale * base^(i << shift), base^(2^k) given in pow2[k]
template <class FR>
__global__ void k_pow_table(void* out, uint64_t count, const void* pow2, uint32_t shift, const void* scale) {
    typedef Fp<FR> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    F acc = load_fp<FR>(scale, 0);
    uint64_t e = i << shift;
    for (int k = 0; e != 0; k++, e >>= 1)
        if (e & 1) acc = acc * load_fp<FR>(pow2, k);
    store_fp<FR>(out, i, acc);
}

}  // namespace apb

using namespace apb;

struct apb_domain_s {
    uint32_t magic;
    std::recursive_mutex mu;              // entry points on one domain serialise here (APB_HANDLE_LOCK)
    int curve;
    uint32_t log_n;
    size_t n;
    void *tw, *itw;                       // N/2 each
    void *clo, *chi, *iclo, *ichi;        // coset tables
    void* scratch;
    size_t scratch_elems;
    void *io_in, *io_out;                 // staging for the host-buffer entry point (apb_ntt)
    uint32_t size_inv[8];
    int npass;
    uint32_t lbits[3];
};
static const uint32_t DOMAIN_MAGIC = 0x444f4d31;

template <class FR>
static int build_tables(apb_domain_s* d) {
    host::Field f = host::Field::make<FR>();
    const uint32_t L = d->log_n;
    // w = two_adic_root ^ (2^(adicity - L))
    uint64_t w[4], winv[4], g[4], ginv[4], ninv[4], nval[4];
    for (int i = 0; i < 4; i++) {
        w[i] = (uint64_t)FR::two_adic_root_mont(2 * i) | ((uint64_t)FR::two_adic_root_mont(2 * i + 1) << 32);
        g[i] = (uint64_t)FR::generator_mont(2 * i) | ((uint64_t)FR::generator_mont(2 * i + 1) << 32);
        ginv[i] = (uint64_t)FR::generator_inv_mont(2 * i) | ((uint64_t)FR::generator_inv_mont(2 * i + 1) << 32);
    }
    for (uint32_t i = L; i < (uint32_t)FR::TWO_ADICITY; i++) f.sqr(w, w);
    f.inv(winv, w);
    // 1/N in Montgomery form: N as field element = 2^L
    f.set(nval, f.one);
    for (uint32_t i = 0; i < L; i++) f.dbl(nval, nval);
    f.inv(ninv, nval);
    for (int i = 0; i < 4; i++) { d->size_inv[2 * i] = (uint32_t)ninv[i]; d->size_inv[2 * i + 1] = (uint32_t)(ninv[i] >> 32); }

    std::vector<uint64_t> h_pow2(4 * 64 * 4);                  // 4 bases x 64 powers
    const uint64_t* bases[4] = {w, winv, g, ginv};
    for (int b = 0; b < 4; b++) {
        uint64_t cur[4];
        f.set(cur, bases[b]);
        for (int k = 0; k < 64; k++) {
            memcpy(h_pow2.data() + (b * 64 + k) * 4, cur, 32);
            f.sqr(cur, cur);
        }
    }
    DevBuf b_pow2, b_scale;
    APB_CUDA_TRY(b_pow2.alloc(4 * 64 * 32));
    APB_CUDA_TRY(b_scale.alloc(64));
    void *d_pow2 = b_pow2.p, *d_scale = b_scale.p;
    APB_CUDA_TRY(cudaMemcpyAsync(d_pow2, h_pow2.data(), 4 * 64 * 32, cudaMemcpyHostToDevice, cur_stream()));
    uint64_t scales[8];
    memcpy(scales, f.one, 32);
    memcpy(scales + 4, ninv, 32);
    APB_CUDA_TRY(cudaMemcpyAsync(d_scale, scales, 64, cudaMemcpyHostToDevice, cur_stream()));

    const size_t half = d->n > 1 ? d->n / 2 : 1;
    const size_t nhi = (d->n >> COSET_LO_BITS) + 1, nlo = (size_t)1 << COSET_LO_BITS;
    APB_CUDA_TRY(cudaMalloc(&d->tw, half * 32));
    APB_CUDA_TRY(cudaMalloc(&d->itw, half * 32));
    APB_CUDA_TRY(cudaMalloc(&d->clo, nlo * 32));
    APB_CUDA_TRY(cudaMalloc(&d->iclo, nlo * 32));
    APB_CUDA_TRY(cudaMalloc(&d->chi, nhi * 32));
    APB_CUDA_TRY(cudaMalloc(&d->ichi, nhi * 32));
    const char* p2 = (const char*)d_pow2;
    const char* sc = (const char*)d_scale;
    auto blocks = [](size_t c) { return (unsigned)((c + 127) / 128); };
    APB_KLAUNCH(k_pow_table<FR>, blocks(half), 128, 0, d->tw, (uint64_t)half, (const void*)(p2 + 0 * 64 * 32), 0u, (const void*)sc);
    APB_KLAUNCH(k_pow_table<FR>, blocks(half), 128, 0, d->itw, (uint64_t)half, (const void*)(p2 + 1 * 64 * 32), 0u, (const void*)sc);
    APB_KLAUNCH(k_pow_table<FR>, blocks(nlo), 128, 0, d->clo, (uint64_t)nlo, (const void*)(p2 + 2 * 64 * 32), 0u, (const void*)sc);
    APB_KLAUNCH(k_pow_table<FR>, blocks(nhi), 128, 0, d->chi, (uint64_t)nhi, (const void*)(p2 + 2 * 64 * 32), (uint32_t)COSET_LO_BITS, (const void*)sc);
    APB_KLAUNCH(k_pow_table<FR>, blocks(nlo), 128, 0, d->iclo, (uint64_t)nlo, (const void*)(p2 + 3 * 64 * 32), 0u, (const void*)sc);
    // 1/N folded into the high table of the inverse coset scale
    APB_KLAUNCH(k_pow_table<FR>, blocks(nhi), 128, 0, d->ichi, (uint64_t)nhi, (const void*)(p2 + 3 * 64 * 32), (uint32_t)COSET_LO_BITS, (const void*)(sc + 32));
    APB_CHECK_LAUNCH();
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

static void plan_passes(apb_domain_s* d) {
    // tiles of up to 2^10 points: two passes up to 2^20 (every extra pass costs one more twiddle product per element
    // and one more trip through L2 / HBM); the register-resident pass kernel exchanges a tile once per three stages
    int maxbits = 10;
    if (const char* e = getenv("APB_NTT_MAX_LOG_TILE")) maxbits = atoi(e);
    if (maxbits < 1) maxbits = 1;
    if (maxbits > 10) maxbits = 10;
    uint32_t L = d->log_n;
    int P = L == 0 ? 1 : (int)((L + maxbits - 1) / maxbits);
    if (P > 3) P = 3;     // log_n <= 30 enforced by apb_domain_new
    d->npass = P;
    for (int i = 0; i < P; i++) d->lbits[i] = L / P + ((uint32_t)i < L % P ? 1 : 0);
}

extern "C" int apb_domain_new(int curve, uint32_t log_n, apb_domain_t* out) {
    APB_API_LOCK();
    if (!out) return set_err(APB_ERR_INVALID_ARG, "apb_domain_new: null out");
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_domain_new: bad curve %d", curve);
    uint32_t adicity = curve == APB_CURVE_BLS12_381 ? (uint32_t)Fr381::TWO_ADICITY : (uint32_t)Fr377::TWO_ADICITY;
    if (log_n > adicity || log_n > 30) return set_err(APB_ERR_DOMAIN_TOO_LARGE, "apb_domain_new: log_n %u exceeds supported size", log_n);
    APB_REQUIRE_INIT();
    *out = nullptr;
    apb_domain_s* d = new apb_domain_s();      // value-initialised: all members zero
    d->magic = DOMAIN_MAGIC;
    d->curve = curve;
    d->log_n = log_n;
    d->n = (size_t)1 << log_n;
    plan_passes(d);
    int rc = curve == APB_CURVE_BLS12_381 ? build_tables<Fr381>(d) : build_tables<Fr377>(d);
    if (rc != APB_OK) { apb_domain_free(d); return rc; }
    *out = d;
    return APB_OK;
}

extern "C" int apb_domain_size(apb_domain_t d, size_t* n) {
    if (!d || d->magic != DOMAIN_MAGIC || !n) return set_err(APB_ERR_BAD_HANDLE, "apb_domain_size: bad handle");
    *n = d->n;
    return APB_OK;
}

// internal: lets the polynomial kernels reach the resident root table w^i, i < N/2
extern "C" int apb_domain_info(apb_domain_t d, int* curve, uint32_t* log_n, const void** tw) {
    if (!d || d->magic != DOMAIN_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "bad domain handle");
    *curve = d->curve;
    *log_n = d->log_n;
    *tw = d->tw;
    return APB_OK;
}

extern "C" void apb_domain_free(apb_domain_t d) {
    if (!d || d->magic != DOMAIN_MAGIC) return;
    {   // wait for a call that is still using this domain's scratch, then retire the handle
        APB_HANDLE_LOCK(d);
        cudaStreamSynchronize(cur_stream());
        d->magic = 0;
    }
    cudaFree(d->tw); cudaFree(d->itw); cudaFree(d->clo); cudaFree(d->chi);
    cudaFree(d->iclo); cudaFree(d->ichi); cudaFree(d->scratch);
    cudaFree(d->io_in); cudaFree(d->io_out);
    delete d;
}

template <class FR>
static int run_ntt(apb_domain_s* d, int kind, const void* d_in, size_t in_len, uint64_t in_stride, void* d_out,
                   uint64_t out_stride, size_t batch) {
    const bool inverse = kind == APB_NTT_IFFT || kind == APB_NTT_COSET_IFFT;
    const uint32_t L = d->log_n;
    const int P = d->npass;
    int log_cols_max = 2;
    if (const char* e = getenv("APB_NTT_LOG_COLS")) log_cols_max = atoi(e);
    if (P > 1 && d->scratch_elems < d->n * batch) {
        cudaFree(d->scratch);
        d->scratch = nullptr;
        d->scratch_elems = 0;
        APB_CUDA_TRY(cudaMalloc(&d->scratch, d->n * batch * 32));
        d->scratch_elems = d->n * batch;
    }
    uint32_t sbits = L;      // log2(S_p) running
    uint32_t mbits = 0;      // log2(M_p)
    for (int p = 0; p < P; p++) {
        const uint32_t lt = d->lbits[p];
        sbits -= lt;
        NttPassArgs A;
        memset(&A, 0, sizeof(A));
        A.log_n = L;
        A.log_t = lt;
        const bool first = p == 0, last = p == P - 1;
        A.in = first ? d_in : d->scratch;
        A.out = last ? d_out : d->scratch;
        A.in_batch_stride = first ? in_stride : d->n;
        A.out_batch_stride = last ? out_stride : d->n;
        A.roots = inverse ? d->itw : d->tw;
        A.in_len = first ? in_len : ~(uint64_t)0;
        // columns: keep T*C <= 1024 elements (the register-resident pass kernel runs T*C/4 <= 256 threads per tile; 48 KB of
        // shared memory -> 3-4 CTAs per SM) and enough tiles to fill the chip a few times over
        uint32_t lc = 10 > lt ? 10 - lt : 0;
        if (getenv("APB_NTT_LOG_RADIX") && atoi(getenv("APB_NTT_LOG_RADIX")) == 0) lc = 11 > lt ? 11 - lt : 0;   // radix-2 kernel: up to 2048
        if ((int)lc > log_cols_max) lc = log_cols_max;
        while (lc > 0 && ((d->n >> (lt + lc)) * batch) < (uint64_t)4 * g_num_sms) lc--;
        if (!last) {
            if (lc > sbits) lc = sbits;
            A.log_c = lc;
            A.na = (uint64_t)1 << (sbits - lc);
            A.in_a = A.out_a = (uint64_t)1 << lc;
            A.in_b = A.out_b = (uint64_t)1 << (sbits + lt);
            A.in_sd = A.out_sk = (uint64_t)1 << sbits;
            A.in_sc = A.out_sc = 1;
            A.tw_mult = (uint64_t)1 << mbits;
        } else if (P == 1) {
            A.log_c = 0;
            A.na = 1;
            A.in_sd = A.out_sk = 1;
            A.in_sc = A.out_sc = 1;
        } else {
            const uint32_t l1 = d->lbits[0];
            const uint32_t s1 = L - l1;              // log2(S_1)
            if (lc > l1) lc = l1;
            A.log_c = lc;
            A.na = (uint64_t)1 << (l1 - lc);
            A.in_a = (uint64_t)1 << (lc + s1);
            A.in_b = (uint64_t)1 << lt;
            A.in_sd = 1;
            A.in_sc = (uint64_t)1 << s1;
            A.out_a = (uint64_t)1 << lc;
            A.out_b = (uint64_t)1 << l1;
            A.out_sk = (uint64_t)1 << mbits;
            A.out_sc = 1;
        }
        if (first && kind == APB_NTT_COSET_FFT) { A.pre_lo = d->clo; A.pre_hi = d->chi; }
        if (last && kind == APB_NTT_COSET_IFFT) { A.post_lo = d->iclo; A.post_hi = d->ichi; }
        if (last && kind == APB_NTT_IFFT) { A.post_const = 1; memcpy(A.size_inv, d->size_inv, 32); }
        const uint64_t tiles = d->n >> (lt + A.log_c);
        const uint32_t tc = 1u << (lt + A.log_c);
        const size_t smem = (size_t)tc * 32 + (size_t)(1u << lt) * 16 + 64;
        static size_t smem_set = 0;
        if (smem > 48 * 1024 && smem > smem_set) {
            APB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass<Fr381>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            APB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass<Fr377>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            APB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass_reg<Fr381, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            APB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass_reg<Fr377, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            smem_set = 200 * 1024;
        }
        // 4 elements per thread, two stages per round (profiles/r02_ntt_tune.txt: 8 elements per thread at 134 registers
        // were 25 % SLOWER than the radix-2 kernel, 4 elements at 72 registers are 8-9 % faster); APB_NTT_LOG_RADIX=0
        // selects the radix-2 kernel (also used for tiles of fewer than 4 elements)
        int lr = 2;
        if (const char* e = getenv("APB_NTT_LOG_RADIX")) lr = atoi(e) == 0 ? 0 : 2;
        if (lr == 2 && tc >= 4 && (tc >> 2) <= 256 && lt >= 2) {
            // rounds of up to `lr` stages; the first round takes the remainder so that the last window is bottom-aligned
            uint32_t hi = lt + A.log_c, remaining = lt, take = lt % lr ? lt % lr : lr;
            A.nrounds = 0;
            while (remaining) {
                const uint32_t lb = hi - lr;
                uint32_t mask = 0;
                for (uint32_t b = hi - take; b < hi; b++) mask |= 1u << (b - lb);
                A.round_lb[A.nrounds] = lb;
                A.round_mask[A.nrounds] = mask;
                A.nrounds++;
                hi -= take;
                remaining -= take;
                take = lr;
            }
            APB_KLAUNCH((k_ntt_pass_reg<FR, 2>), dim3((unsigned)tiles, (unsigned)batch), tc >> 2, smem, A);
        } else {
            uint32_t threads = tc / 2;
            if (threads < 32) threads = 32;
            if (threads > 512) threads = 512;
            APB_KLAUNCH(k_ntt_pass<FR>, dim3((unsigned)tiles, (unsigned)batch), threads, smem, A);
        }
        mbits += lt;
    }
    APB_CHECK_LAUNCH();
    return APB_OK;
}

extern "C" int apb_ntt_batch_dev(apb_domain_t d, int kind, const void* d_in, size_t in_len, size_t in_stride, void* d_out,
                                 size_t out_stride, size_t batch, int sync) {
    if (!d || d->magic != DOMAIN_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_ntt: bad domain handle");
    APB_HANDLE_LOCK(d);
    if (kind < 0 || kind > 3) return set_err(APB_ERR_INVALID_ARG, "apb_ntt: bad kind %d", kind);
    if (in_len > d->n) return set_err(APB_ERR_INVALID_ARG, "apb_ntt: in_len %zu > domain size %zu", in_len, d->n);
    if (batch == 0) return APB_OK;
    if (!d_out || (!d_in && in_len)) return set_err(APB_ERR_INVALID_ARG, "apb_ntt: null buffer");
    // profiling (apb_set_profiling): the two events are only RECORDED here and read when the totals are asked for -
    // a cudaEventSynchronize per transform would put a host round trip behind every NTT of a proof
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (g_profile) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); cudaEventRecord(pe0, cur_stream()); }
    int rc = d->curve == APB_CURVE_BLS12_381
                 ? run_ntt<Fr381>(d, kind, d_in, in_len, in_stride, d_out, out_stride, batch)
                 : run_ntt<Fr377>(d, kind, d_in, in_len, in_stride, d_out, out_stride, batch);
    if (pe0) {
        cudaEventRecord(pe1, cur_stream());
        std::lock_guard<std::mutex> lk(g_ntt_pending_mu);
        g_ntt_pending.push_back(NttPending{pe0, pe1, (unsigned long long)batch});
        if (g_ntt_pending.size() > 8192) ntt_resolve_pending();
    }
    if (rc != APB_OK) return rc;
    if (sync) APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

extern "C" void apb_ntt_totals(double* ms, unsigned long long* transforms, int reset) {
    {
        std::lock_guard<std::mutex> lk(g_ntt_pending_mu);
        ntt_resolve_pending();
    }
    if (ms) *ms = g_ntt_ms_total;
    if (transforms) *transforms = g_ntt_count;
    if (reset) { g_ntt_ms_total = 0.0; g_ntt_count = 0; }
}

extern "C" int apb_ntt_dev(apb_domain_t d, int kind, const void* d_in, size_t in_len, void* d_out, int sync) {
    if (!d || d->magic != DOMAIN_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_ntt: bad domain handle");
    return apb_ntt_batch_dev(d, kind, d_in, in_len, d->n, d_out, d->n, 1, sync);
}

extern "C" int apb_ntt(apb_domain_t d, int kind, const uint64_t* in, size_t in_len, uint64_t* out) {
    if (!d || d->magic != DOMAIN_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_ntt: bad domain handle");
    APB_HANDLE_LOCK(d);
    if (in_len > d->n) return set_err(APB_ERR_INVALID_ARG, "apb_ntt: in_len %zu > domain size %zu", in_len, d->n);
    if (!out || (!in && in_len)) return set_err(APB_ERR_INVALID_ARG, "apb_ntt: null buffer");
    if (!d->io_in) {
        APB_CUDA_TRY(cudaMalloc(&d->io_in, d->n * 32));
        APB_CUDA_TRY(cudaMalloc(&d->io_out, d->n * 32));
    }
    void *d_in = d->io_in, *d_out = d->io_out;
    if (in_len) APB_CUDA_TRY(cudaMemcpyAsync(d_in, in, in_len * 32, cudaMemcpyHostToDevice, cur_stream()));
    EventPair ev;
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, cur_stream());
    int rc = apb_ntt_dev(d, kind, d_in, in_len, d_out, 0);
    cudaEventRecord(e1, cur_stream());
    if (rc == APB_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, d->n * 32, cudaMemcpyDeviceToHost, cur_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(cur_stream());
        if (e != cudaSuccess) rc = set_err(APB_ERR_CUDA, "apb_ntt: %s", cudaGetErrorString(e));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        g_last_ms = ms;
    }
    return rc;
}
