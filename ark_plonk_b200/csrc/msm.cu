// Variable-base multi-scalar multiplication over G1 (BLS12-381 / BLS12-377) for sm_100a.
//
// Replaces ark_ec::msm::VariableBaseMSM::multi_scalar_mul (plonk-core/src/commitment.rs:45 and,
// through SonicKZG10::commit/open, plonk-core/src/proof_system/prover.rs:213,290,313,316,362,388,
// 459,579,582,606,609; preprocess.rs:351; lookup/preprocess.rs:63).  Same function
// (sum_i s_i * P_i), different algorithm; the result is unique as a group element and leaves
// the library as a normalised affine point.
//
// Design (B200-first, see DESIGN.md "MSM"):
//   * The commitment key is resident in HBM together with F-1 precomputed multiples
//     2^(step*f) * P_i (affine).  A signed c-bit digit at position w = f*G + g of scalar i is
//     then "add +-copy_f[i] to bucket |d|-1 of effective window g": with step = 16, c = 16,
//     G = 1 all 16 digit positions share ONE bucket set and no window doublings remain.
//   * digits -> histogram -> exclusive scan -> scatter gives the (bucket, point) pairs sorted
//     by bucket (a counting sort; keys are bucket ids).
//   * large lists first go through batched-affine PAIR LEVELS (k_msm_pairs): the points of every
//     bucket are added pairwise as affine points with all denominators of a CTA inverted together
//     (5M+1S per addition instead of 8M+2S), two or three times, which halves the list each time.
//   * accumulation walks the (remaining) sorted list in equal chunks per thread (perfect balance
//     for any scalar distribution); runs that cross chunk borders are stitched by a second kernel.
//   * the weighted bucket sum  sum_b (b+1) * S_b  is evaluated with log-depth trees only:
//     buckets are viewed as a 2^a x 2^b matrix, row/column sums are trees, and the two small
//     weighted sums are split by weight bit (V_k = sum of entries whose weight has bit k set).
//   * the last ~40 group operations (Horner over the V_k, window fold, one inversion) run on
//     the host in 64-bit limbs: they are a strictly sequential chain and the result is needed
//     in host memory for the transcript.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "host_ec.hpp"
#include "scan_u32.cuh"

namespace apb {

int g_num_sms = 1;
int g_profile = 0;
double g_acc_ms_total = 0.0;            // accumulated k_msm_accumulate time while profiling
double g_msm_call_ms_total = 0.0;       // whole MSM calls (sort + accumulation + reduction + copy + host epilogue) while profiling
unsigned long long g_points_total = 0;  // scalars processed while profiling
double g_madds_model = 0.0;             // while profiling: wide multiply-adds by the XYZZ cost model (entries x 10 x 300)
double g_madds_issued = 0.0;            // ... and as issued: 6 x 300 per batched-affine pair addition, 10 x 300 per XYZZ one
double g_phase_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // sort, accumulate, stitch, trees, copy+host epilogue
uint32_t g_last_plan[4] = {0, 0, 0, 0};            // of the last pass: digit bits, pair levels, bucket-range slices, fell back (unbalanced)

static const int MAX_BATCH = 16;
static const int MAX_COPIES = 16;

struct MsmBatch {
    uint32_t k;
    uint64_t scal_off[MAX_BATCH];    // element offset into the concatenated scalar buffer
    uint64_t base_off[MAX_BATCH];
    uint64_t len[MAX_BATCH];
};

struct MsmGeom {
    uint32_t c;          // digit bits
    uint32_t G;          // effective windows (bucket groups)
    uint32_t W;          // digit positions
    uint32_t hb;         // buckets per window = 2^(c-1)
    uint64_t ck_n;       // points per copy in the resident table
};

// ---- scalar recoding ---------------------------------------------------------------------
template <class FR>
APB_D bool geq_mod(const Fp<FR>& a) {
#pragma unroll
    for (int i = FR::N - 1; i >= 0; i--) {
        if (a.v[i] > FR::mod(i)) return true;
        if (a.v[i] < FR::mod(i)) return false;
    }
    return true;
}
template <class FR>
APB_D bool gt_half(const Fp<FR>& a) {
#pragma unroll
    for (int i = FR::N - 1; i >= 0; i--) {
        if (a.v[i] > FR::half_mod(i)) return true;
        if (a.v[i] < FR::half_mod(i)) return false;
    }
    return false;
}
template <class FR>
APB_D void sub_mod_raw(Fp<FR>& a) {
    a.v[0] = sub_cc(a.v[0], FR::mod(0));
#pragma unroll
    for (int i = 1; i < FR::N - 1; i++) a.v[i] = subc_cc(a.v[i], FR::mod(i));
    a.v[FR::N - 1] = subc(a.v[FR::N - 1], FR::mod(FR::N - 1));
}

// canonical |s| <= (r-1)/2 and the sign that was factored out
template <class FR>
APB_D Fp<FR> load_scalar(const void* scalars, uint64_t idx, int mont, bool& negative) {
    Fp<FR> s = load_fp<FR>(scalars, idx);
    if (mont) s = s.from_mont();
    else while (geq_mod<FR>(s)) sub_mod_raw<FR>(s);
    negative = gt_half<FR>(s);
    if (negative) s = s.neg();            // r - s
    return s;
}

// digit at position w (c bits from bit w*c) plus incoming carry; returns signed digit, updates carry
template <class FR>
APB_D int take_digit(const Fp<FR>& s, uint32_t w, uint32_t c, uint32_t& carry) {
    uint32_t lo = w * c;
    uint32_t limb = lo >> 5, sh = lo & 31;
    uint64_t window = limb < (uint32_t)FR::N ? s.v[limb] : 0;
    if (limb + 1 < (uint32_t)FR::N) window |= (uint64_t)s.v[limb + 1] << 32;
    uint32_t raw = (uint32_t)((window >> sh) & ((1u << c) - 1)) + carry;
    if (raw > (1u << (c - 1))) {
        carry = 1;
        return (int)raw - (int)(1u << c);
    }
    carry = 0;
    return (int)raw;
}

// pass 0: histogram; pass 1: scatter (digits are recomputed instead of stored)
template <class FR, int PASS>
__global__ void k_msm_digits(const void* scalars, MsmBatch B, MsmGeom g, int mont, uint32_t* counts,
                             const uint32_t* offsets, uint32_t* cursors, uint32_t* entries) {
    const uint32_t j = blockIdx.y;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.len[j]) return;
    bool negative;
    Fp<FR> s = load_scalar<FR>(scalars, B.scal_off[j] + i, mont, negative);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < g.W; w++) {
        int d = take_digit<FR>(s, w, g.c, carry);
        if (d == 0) continue;
        bool neg = negative != (d < 0);
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint32_t f = w / g.G, gw = w % g.G;
        uint32_t bucket = (j * g.G + gw) * g.hb + (mag - 1);
        if (PASS == 0) {
            atomicAdd(&counts[bucket], 1u);
        } else {
            uint32_t pos = offsets[bucket] + atomicAdd(&cursors[bucket], 1u);
            uint64_t pidx = (uint64_t)f * g.ck_n + B.base_off[j] + i;
            entries[pos] = (uint32_t)pidx | (neg ? 0x80000000u : 0u);
        }
    }
}

template <class FQ>
APB_D void load_affine(const void* bases, uint64_t idx, Fp<FQ>& x, Fp<FQ>& y) {
    x = load_fp<FQ>(bases, 2 * idx);
    y = load_fp<FQ>(bases, 2 * idx + 1);
}
template <class FQ>
APB_D XYZZ<FQ> load_xyzz(const void* arr, uint64_t idx) {
    XYZZ<FQ> p;
    p.x = load_fp<FQ>(arr, 4 * idx);
    p.y = load_fp<FQ>(arr, 4 * idx + 1);
    p.zz = load_fp<FQ>(arr, 4 * idx + 2);
    p.zzz = load_fp<FQ>(arr, 4 * idx + 3);
    return p;
}
template <class FQ>
APB_D void store_xyzz(void* arr, uint64_t idx, const XYZZ<FQ>& p) {
    store_fp<FQ>(arr, 4 * idx, p.x);
    store_fp<FQ>(arr, 4 * idx + 1, p.y);
    store_fp<FQ>(arr, 4 * idx + 2, p.zz);
    store_fp<FQ>(arr, 4 * idx + 3, p.zzz);
}

// Each thread owns entries [t*E, (t+1)*E) of the bucket-sorted list.  The next point is
// fetched (entry id, then the 96-byte affine record) while the current mixed add runs.
// SRC 0: entry ids into the resident table (sign in bit 31; (0,0) = infinity).  SRC 1: the list is
// itself an array of affine partial sums (output of the batched-affine pair levels below; infinity is
// marked by an all-ones top limb of x), entry p is point p.
template <class FQ, int MINB, int SRC>
__global__ void __launch_bounds__(128, MINB) k_msm_accumulate(const uint32_t* entries, const uint32_t* offsets, uint32_t nbuckets,
                                                              const void* bases, uint32_t E, void* bucket_sums, void* partials,
                                                              int32_t* part_bucket) {
    typedef Fp<FQ> F;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t M = offsets[nbuckets];
    part_bucket[2 * t] = -1;
    part_bucket[2 * t + 1] = -1;
    uint64_t pos = t * E;
    if (pos >= M) return;
    const uint64_t end = pos + E < M ? pos + E : M;
    // largest b with offsets[b] <= pos
    uint32_t lo = 0, hi = nbuckets;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= pos) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    while (offsets[b + 1] <= pos) b++;     // skip empty buckets sharing the same offset
    // One flat loop over the chunk: every lane performs its mixed add in the same iteration
    // (bucket borders fall at different positions in different lanes; a nested run loop lets
    // the lanes drift apart and the warp then executes the add twice at half occupancy).
    uint32_t e_cur = SRC == 0 ? entries[pos] : 0u;
    F px, py;
    load_affine<FQ>(bases, SRC == 0 ? (uint64_t)(e_cur & 0x7fffffffu) : pos, px, py);
    uint64_t bstart = offsets[b], bend = offsets[b + 1], run_start = pos;
    XYZZ<FQ> acc = XYZZ<FQ>::identity();
    while (pos < end) {
        uint32_t e_nxt = 0;
        F nx, ny;
        const bool more = pos + 1 < end;
        if (more) {
            if (SRC == 0) e_nxt = entries[pos + 1];
            load_affine<FQ>(bases, SRC == 0 ? (uint64_t)(e_nxt & 0x7fffffffu) : pos + 1, nx, ny);
        }
        const bool inf = SRC == 0 ? (px.is_zero() && py.is_zero()) : (px.v[FQ::N - 1] == 0xffffffffu);
        if (!inf) {                                          // skip the point at infinity
            if (e_cur >> 31) py = py.neg();
            acc.add_affine(px, py);
        }
        pos++;
        if (pos == bend || pos == end) {                     // run finished: flush
            const bool head = run_start == bstart, tail = pos == bend;
            if (head && tail) {
                store_xyzz<FQ>(bucket_sums, b, acc);
            } else if (head) {          // bucket continues in the next chunk(s)
                store_xyzz<FQ>(partials, 2 * t + 1, acc);
                part_bucket[2 * t + 1] = (int32_t)b;
            } else {                    // bucket began in an earlier chunk
                store_xyzz<FQ>(partials, 2 * t, acc);
                part_bucket[2 * t] = (int32_t)b;
            }
            if (pos < end) {
                b++;
                while (offsets[b + 1] <= pos) b++;
                bstart = offsets[b];
                bend = offsets[b + 1];
                run_start = pos;
                acc = XYZZ<FQ>::identity();
            }
        }
        __syncwarp();
        if (more) { e_cur = e_nxt; px = nx; py = ny; }
    }
}

// stitch buckets that straddle chunk borders: the chunk holding the head piece sums the rest
template <class FQ>
__global__ void __launch_bounds__(128) k_msm_stitch(const uint32_t* offsets, uint32_t E, uint64_t nthreads, void* bucket_sums,
                                                    const void* partials, const int32_t* part_bucket) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const int32_t b = part_bucket[2 * t + 1];
    if (b < 0) return;
    XYZZ<FQ> acc = load_xyzz<FQ>(partials, 2 * t + 1);
    const uint64_t bend = offsets[b + 1];
    for (uint64_t u = t + 1; u * E < bend; u++) {
        XYZZ<FQ> p = load_xyzz<FQ>(partials, 2 * u);
        acc.add(p);
    }
    store_xyzz<FQ>(bucket_sums, b, acc);
}

// Tree-sum jobs: out[job] = sum over e < m of in[base + e*stride], restricted (selbit >= 0) to
// entries whose weight (e + woff) has bit `selbit` set.
struct TreeJob {
    uint32_t base, stride, m, woff;
    int32_t selbit;
    uint32_t out;
};

template <class FQ>
APB_D XYZZ<FQ> shfl_down_xyzz(const XYZZ<FQ>& a, uint32_t delta) {
    XYZZ<FQ> r;
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int i = 0; i < FQ::N; i++) {
        r.x.v[i] = __shfl_down_sync(0xffffffffu, a.x.v[i], delta);
        r.y.v[i] = __shfl_down_sync(0xffffffffu, a.y.v[i], delta);
        r.zz.v[i] = __shfl_down_sync(0xffffffffu, a.zz.v[i], delta);
        r.zzz.v[i] = __shfl_down_sync(0xffffffffu, a.zzz.v[i], delta);
    }
#else
    (void)delta;
    r = a;
#endif
    return r;
}

template <class FQ>
__global__ void __launch_bounds__(128) k_msm_tree(const void* in, void* out, const TreeJob* jobs, uint32_t njobs) {
    // one job per warp (4 per CTA): every lane first folds m/32 strided elements sequentially, then a
    // 5-level tree inside the warp.  On the device the tree exchanges points with register shuffles: no
    // shared memory and no CTA barrier (round 1 went through shared memory with two __syncthreads per
    // level, which cost 10 barrier-stall cycles per issued instruction: profiles/r01_ncu_prove_kernels.json).
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t job = blockIdx.x * 4 + (tid >> 5);
    const bool active = job < njobs;
    TreeJob J;
    if (active) J = jobs[job];
    else { J.base = 0; J.stride = 0; J.m = 0; J.woff = 0; J.selbit = -1; J.out = 0; }
    XYZZ<FQ> acc = XYZZ<FQ>::identity();
    for (uint32_t e = lane; e < J.m; e += 32) {
        if (J.selbit >= 0 && !(((e + J.woff) >> J.selbit) & 1)) continue;
        XYZZ<FQ> p = load_xyzz<FQ>(in, (uint64_t)J.base + (uint64_t)e * J.stride);
        acc.add(p);
    }
#ifndef APB_EMU
    for (uint32_t s = 16; s >= 1; s >>= 1) {
        const XYZZ<FQ> p = shfl_down_xyzz<FQ>(acc, s);
        if (lane < s) acc.add(p);
    }
#else
    // CPU emulation (one OS thread per CUDA thread, no lockstep warps): the same tree through shared memory
    __shared__ uint4 sm[128 * 12];       // 128 XYZZ points (4 * 48 bytes), 32 per warp
    store_xyzz<FQ>(sm, tid, acc);
    __syncthreads();
    for (uint32_t s = 16; s >= 1; s >>= 1) {
        if (lane < s) {
            XYZZ<FQ> p = load_xyzz<FQ>(sm, tid + s);
            acc.add(p);
        }
        __syncthreads();
        if (lane < s) store_xyzz<FQ>(sm, tid, acc);
        __syncthreads();
    }
#endif
    if (active && lane == 0) store_xyzz<FQ>(out, J.out, acc);
}

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

// ---- batched-affine pair levels -----------------------------------------------------------------
// A bucket that holds m points needs m-1 additions whatever the order.  Adding the points of a
// bucket PAIRWISE (level r+1 holds ceil(m_r / 2) partial sums per bucket) makes every addition an
// affine + affine -> affine one, whose only expensive part is 1/(x2 - x1): all the denominators a
// CTA handles are inverted together (Montgomery's trick: per-thread prefix products, a product
// tree over the 128 thread totals in shared memory, ONE Fermat inversion per CTA), so an addition
// costs 5 M + 1 S instead of the 8 M + 2 S of the XYZZ mixed addition.  After a few levels the
// buckets are short and the remaining list goes through k_msm_accumulate<SRC = 1>.
//
// Level-r lists are sorted by bucket like the entry list; off_r = exclusive scan of the per-bucket
// counts.  Output j of bucket b (local index jl) adds inputs off_r[b] + 2 jl and + 2 jl + 1; a
// trailing odd element is passed through.

__global__ void k_msm_level_counts(const uint32_t* offsets0, uint32_t nbuckets, uint32_t levels, uint32_t* cnt) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    uint32_t c = offsets0[b + 1] - offsets0[b];
    for (uint32_t r = 0; r < levels; r++) {
        c = (c + 1) >> 1;
        cnt[(size_t)r * (nbuckets + 1) + b] = c;
    }
}

static const uint32_t NO_PARTNER = 0xffffffffu;

// One input of a pair level.  FIRST: `e` is an entry of the bucket-sorted list (table index, sign in
// bit 31; (0,0) in the table = infinity).  Otherwise `e` is an index into the previous level's array
// (infinity = all-ones top limb of x).
template <class FQ, int FIRST>
APB_D Fp<FQ> pair_load_x(const void* src, uint32_t e) {
    return load_fp_early<FQ, 1>(src, 2 * (uint64_t)(FIRST ? (e & 0x7fffffffu) : e));
}
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
APB_D uint2 load_u2_early(const uint2* p) {      // see load_fp_early: keeps its position among the volatile chains
#ifdef __CUDA_ARCH__
    uint2 r;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
#else
    return *p;
#endif
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

// walks the outputs of a level in order and yields the ids of the (one or two) inputs each one adds
struct PairWalker {
    const uint32_t *off_in, *off_out;
    uint32_t b, ipos, in_end;
    uint64_t out_end;
    APB_D void init(const uint32_t* oin, const uint32_t* oout, uint32_t nbuckets, uint64_t j) {
        off_in = oin;
        off_out = oout;
        uint32_t lo = 0, hi = nbuckets;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (off_out[mid] <= j) lo = mid; else hi = mid;
        }
        b = lo;
        while (off_out[b + 1] <= j) b++;
        out_end = off_out[b + 1];
        ipos = off_in[b] + 2 * (uint32_t)(j - off_out[b]);
        in_end = off_in[b + 1];
    }
    template <int FIRST>
    APB_D uint2 next(const uint32_t* entries, uint64_t j) {
        if (j >= out_end) {
            b++;
            while (off_out[b + 1] <= j) b++;
            out_end = off_out[b + 1];
            ipos = off_in[b];
            in_end = off_in[b + 1];
        }
        uint2 r;
        const bool pair = ipos + 1 < in_end;
        r.x = FIRST ? entries[ipos] : ipos;
        r.y = pair ? (FIRST ? entries[ipos + 1] : ipos + 1) : NO_PARTNER;
        ipos += pair ? 2 : 1;
        return r;
    }
};

// (No __syncwarp() in the loops: lanes whose range is shorter wait at the CTA barrier below, and a
// warp-wide sync that names them would never complete.)
// Pass 2 is software pipelined two deep (ids for output j-2 and the 96-byte records for j-1 are in
// flight while the multiplications of output j run): a thread's inputs are consecutive in the
// level's list, but the table records behind the first level's ids are random 96-byte gathers.
// Measured alternatives (B200, 2^18-point commits; table in profiles/r01_msm_pair_levels.md): no
// prefetch - same time; three CTAs per SM at 168 registers - slower (spills); two interleaved
// batches per thread for instruction-level parallelism - slower (245 registers); two explicit
// register sets instead of rotating one - slower; the denominator pass as its own kernel at twice
// the occupancy - slower; L2 prefetch of the next group's records - slower; a^(p-2) for the CTA's one
// inversion - the single-lane dependent chain took 0.45 ms per launch, the binary Euclid inverse
// takes ~0.04 ms.  What remains (ncu): a strictly sequential product chain per thread at 2 warps
// per scheduler (27 % issue-active against 31 % for the XYZZ kernel, which has two independent
// products in flight) and long-scoreboard stalls in the denominator pass.
template <class FQ, int FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_pairs(const uint32_t* entries, const void* src, const uint32_t* off_in,
                                                         const uint32_t* off_out, uint32_t nbuckets, uint32_t E, void* dst,
                                                         void* prefix, uint2* stash) {
    typedef Fp<FQ> F;
    __shared__ uint4 sm[256 * (FQ::N / 4)];              // product tree: node i at sm[i], leaves 128..255
    const uint32_t tid = threadIdx.x;
    const uint64_t t = (uint64_t)blockIdx.x * 128 + tid;
    const uint64_t Mout = off_out[nbuckets];
    const uint64_t j0 = t * E < Mout ? t * E : Mout;
    const uint64_t j1 = j0 + E < Mout ? j0 + E : Mout;

    // pass 1 (forward): denominators and running prefix products; the ids each output reads are stashed.
    // One multiplication per output is too short to hide a gather behind, so outputs go in groups of
    // four: 8 ids, then 8 x-coordinates in flight together, then the 4 dependent products.
    F run = F::one();
    if (j0 < j1) {
        PairWalker W;
        W.init(off_in, off_out, nbuckets, j0);
        for (uint64_t j = j0; j < j1; j += 4) {
            uint2 id[4];
            F xa[4], xb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) id[q] = j + q < j1 ? W.template next<FIRST>(entries, j + q) : make_uint2(0, NO_PARTNER);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (id[q].y != NO_PARTNER) {
                    xa[q] = pair_load_x<FQ, FIRST>(src, id[q].x);
                    xb[q] = pair_load_x<FQ, FIRST>(src, id[q].y);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (j + q < j1) {
                    stash[j + q] = id[q];
                    store_fp<FQ>(prefix, j + q, run);
                    if (id[q].y != NO_PARTNER) {
                        F d = xb[q] - xa[q];
                        const bool special = FIRST ? (xa[q].is_zero() || xb[q].is_zero() || d.is_zero())
                                                   : (xa[q].v[FQ::N - 1] == 0xffffffffu || xb[q].v[FQ::N - 1] == 0xffffffffu || d.is_zero());
                        if (special) d = pair_den_special<FQ, FIRST>(src, id[q].x, id[q].y);
                        run = run * d;
                    }
                }
            }
        }
    }

    // 1 / (this thread's product) through a product tree over the CTA and one inversion
    store_fp<FQ>(sm, 128 + tid, run);
    __syncthreads();
    for (uint32_t s = 64; s >= 1; s >>= 1) {
        if (tid < s) {
            F a = load_fp<FQ>(sm, 2 * (s + tid)), c = load_fp<FQ>(sm, 2 * (s + tid) + 1);
            store_fp<FQ>(sm, s + tid, a * c);
        }
        __syncthreads();
    }
    if (tid == 0) store_fp<FQ>(sm, 1, load_fp<FQ>(sm, 1).inverse_binary());   // one thread: latency matters, not throughput
    __syncthreads();
    for (uint32_t s = 1; s <= 64; s <<= 1) {
        if (tid < s) {
            F inv = load_fp<FQ>(sm, s + tid);
            F a = load_fp<FQ>(sm, 2 * (s + tid)), c = load_fp<FQ>(sm, 2 * (s + tid) + 1);
            store_fp<FQ>(sm, 2 * (s + tid), inv * c);
            store_fp<FQ>(sm, 2 * (s + tid) + 1, inv * a);
        }
        __syncthreads();
    }
    F rinv = load_fp<FQ>(sm, 128 + tid);

    // pass 2 (backward): 1/d_j = rinv * prefix_j, then the affine addition
    if (j0 < j1) {
        uint2 cur = stash[j1 - 1], nxt = make_uint2(0, NO_PARTNER);
        F x1, y1, x2 = F::zero(), y2 = F::zero(), pre = load_fp<FQ>(prefix, j1 - 1);
        F nx1 = F::zero(), ny1 = F::zero(), nx2 = F::zero(), ny2 = F::zero(), npre = F::zero();
        pair_load_xy<FQ, FIRST>(src, cur.x, x1, y1);
        if (cur.y != NO_PARTNER) pair_load_xy<FQ, FIRST>(src, cur.y, x2, y2);
        if (j1 - 1 > j0) nxt = stash[j1 - 2];
        for (uint64_t j = j1; j-- > j0;) {
            uint2 nn = make_uint2(0, NO_PARTNER);
            if (j > j0) {
                pair_load_xy<FQ, FIRST>(src, nxt.x, nx1, ny1);
                if (nxt.y != NO_PARTNER) pair_load_xy<FQ, FIRST>(src, nxt.y, nx2, ny2);
                npre = load_fp_early<FQ, 0>(prefix, j - 1);
            }
            if (j > j0 + 1) nn = load_u2_early(stash + (j - 2));
            bool inf1 = pair_fix<FQ, FIRST>(cur.x, x1, y1);
            if (cur.y != NO_PARTNER) {
                const bool inf2 = pair_fix<FQ, FIRST>(cur.y, x2, y2);
                if (inf1 || inf2) {
                    if (inf1) { x1 = x2; y1 = y2; inf1 = inf2; }
                } else if (x1 != x2) {
                    const F dinv = rinv * pre;
                    rinv = rinv * (x2 - x1);
                    const F lam = (y2 - y1) * dinv;
                    const F x3 = lam.sqr() - x1 - x2;
                    y1 = lam * (x1 - x3) - y1;
                    x1 = x3;
                } else if (y1 == y2 && !y1.is_zero()) {
                    const F dinv = rinv * pre;
                    rinv = rinv * (y1 + y1);
                    const F xx = x1.sqr();
                    const F lam = (xx + xx + xx) * dinv;
                    const F x3 = lam.sqr() - x1 - x1;
                    y1 = lam * (x1 - x3) - y1;
                    x1 = x3;
                } else {
                    inf1 = true;
                }
            }
            if (inf1) {
                x1 = F::zero();
                y1 = F::zero();
                x1.v[FQ::N - 1] = 0xffffffffu;
            }
            store_fp<FQ>(dst, 2 * j, x1);
            store_fp<FQ>(dst, 2 * j + 1, y1);
            cur = nxt; nxt = nn; x1 = nx1; y1 = ny1; x2 = nx2; y2 = ny2; pre = npre;
        }
    }
}

// ---- the same pair level with asynchronous shared-memory staging (cp.async) ----------------------------------
// k_msm_pairs keeps its software pipeline in REGISTERS (the operands of the next output are loaded into a
// second register set while the current addition runs): 214 registers, 2 CTAs = 8 warps per SM, and ncu showed
// the kernel latency-bound (issue-active 28 %, long-scoreboard stalls on the 96-byte gathers).  Here the
// operands of the next step are fetched by cp.async (LDGSTS: global -> shared memory, no register in between)
// into a per-thread slot of a two-stage ring, and read back with LDS right where they are used: the prefetch
// costs no registers, so three CTAs per SM fit without spills, and the gather latency is covered by a whole
// step of arithmetic of 12 warps instead of 8.  A slot is private to its thread (only the thread that issued
// the copies reads them), so cp.async.wait_group orders everything and no CTA barrier is needed in the loops.
// Slot layout: chunk c (16 bytes) of thread t at uint4 index (stage * PAIR_CHUNKS + c) * 128 + t: consecutive
// lanes touch consecutive 16-byte words, i.e. conflict-free LDS.128 / LDGSTS.128.
static const int PAIR_CHUNKS = 15;           // x1 y1 x2 y2 prefix: 5 x 48 bytes
static const int PAIR_STAGES = 2;
static const size_t PAIR_SMEM_BYTES = (size_t)PAIR_STAGES * PAIR_CHUNKS * 128 * 16;

APB_D void cp_async16(uint4* smem_dst, const uint4* gsrc) {
#ifdef __CUDA_ARCH__
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
#else
    *smem_dst = *gsrc;
#endif
}
APB_D void cp_async_commit() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
APB_D void cp_async_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
// 48-byte field element: global -> chunks [c0, c0 + 3) of this thread's slot
template <class FQ>
APB_D void stage_fp(uint4* slot, int c0, const void* base, uint64_t idx) {
    const uint4* g = reinterpret_cast<const uint4*>(base) + idx * (FQ::N / 4);
#pragma unroll
    for (int i = 0; i < FQ::N / 4; i++) cp_async16(slot + (size_t)(c0 + i) * 128, g + i);
}
template <class FQ>
APB_D Fp<FQ> unstage_fp(const uint4* slot, int c0) {
    Fp<FQ> r;
#pragma unroll
    for (int i = 0; i < FQ::N / 4; i++) {
        const uint4 t = slot[(size_t)(c0 + i) * 128];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}

template <class FQ, int FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_pairs2(const uint32_t* entries, const void* src, const uint32_t* off_in,
                                                          const uint32_t* off_out, uint32_t nbuckets, uint32_t E, void* dst,
                                                          void* prefix, uint2* stash) {
    typedef Fp<FQ> F;
    APB_DYN_SMEM(smem_raw);
    uint4* sm = reinterpret_cast<uint4*>(smem_raw);      // staging ring; its head doubles as the product tree
    const uint32_t tid = threadIdx.x;
    const uint64_t t = (uint64_t)blockIdx.x * 128 + tid;
    const uint64_t Mout = off_out[nbuckets];
    const uint64_t j0 = t * E < Mout ? t * E : Mout;
    const uint64_t j1 = j0 + E < Mout ? j0 + E : Mout;
    uint4* const ring = sm + tid;
    const size_t stage_words = (size_t)PAIR_CHUNKS * 128;

    // pass 1 (forward): denominators x2 - x1 and running prefix products, two outputs per stage
    F run = F::one();
    if (j0 < j1) {
        PairWalker W;
        W.init(off_in, off_out, nbuckets, j0);
        const uint64_t ngroups = (j1 - j0 + 1) / 2;
        uint2 idc[2], idn[2];                             // ids of the group being staged / of the one after it
        auto fetch_ids = [&](uint2* id, uint64_t g) {
            const uint64_t j = j0 + 2 * g;
#pragma unroll
            for (int q = 0; q < 2; q++) id[q] = j + q < j1 ? W.template next<FIRST>(entries, j + q) : make_uint2(0, NO_PARTNER);
        };
        auto issue = [&](const uint2* id, uint32_t stage) {
            uint4* slot = ring + stage * stage_words;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (id[q].y != NO_PARTNER) {
                    stage_fp<FQ>(slot, 6 * q, src, 2 * (uint64_t)(FIRST ? (id[q].x & 0x7fffffffu) : id[q].x));
                    stage_fp<FQ>(slot, 6 * q + 3, src, 2 * (uint64_t)(FIRST ? (id[q].y & 0x7fffffffu) : id[q].y));
                }
            }
            cp_async_commit();
        };
        fetch_ids(idc, 0);
        issue(idc, 0);
        if (ngroups > 1) fetch_ids(idn, 1);
        for (uint64_t g = 0; g < ngroups; g++) {
            uint2 cur[2] = {idc[0], idc[1]};
            if (g + 1 < ngroups) {
                issue(idn, (uint32_t)((g + 1) & 1));
                idc[0] = idn[0]; idc[1] = idn[1];
                if (g + 2 < ngroups) fetch_ids(idn, g + 2);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            const uint4* slot = ring + (g & 1) * stage_words;
            const uint64_t j = j0 + 2 * g;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (j + q < j1) {
                    stash[j + q] = cur[q];
                    store_fp<FQ>(prefix, j + q, run);
                    if (cur[q].y != NO_PARTNER) {
                        const F xa = unstage_fp<FQ>(slot, 6 * q), xb = unstage_fp<FQ>(slot, 6 * q + 3);
                        F d = xb - xa;
                        const bool special = FIRST ? (xa.is_zero() || xb.is_zero() || d.is_zero())
                                                   : (xa.v[FQ::N - 1] == 0xffffffffu || xb.v[FQ::N - 1] == 0xffffffffu || d.is_zero());
                        if (special) d = pair_den_special<FQ, FIRST>(src, cur[q].x, cur[q].y);
                        run = run * d;
                    }
                }
            }
        }
    }

    // 1 / (this thread's product) through a product tree over the CTA and one inversion
    __syncthreads();                                     // every thread is done with its staging slots
    store_fp<FQ>(sm, 128 + tid, run);
    __syncthreads();
    for (uint32_t s = 64; s >= 1; s >>= 1) {
        if (tid < s) {
            F a = load_fp<FQ>(sm, 2 * (s + tid)), c = load_fp<FQ>(sm, 2 * (s + tid) + 1);
            store_fp<FQ>(sm, s + tid, a * c);
        }
        __syncthreads();
    }
    if (tid == 0) store_fp<FQ>(sm, 1, load_fp<FQ>(sm, 1).inverse_binary());
    __syncthreads();
    for (uint32_t s = 1; s <= 64; s <<= 1) {
        if (tid < s) {
            F inv = load_fp<FQ>(sm, s + tid);
            F a = load_fp<FQ>(sm, 2 * (s + tid)), c = load_fp<FQ>(sm, 2 * (s + tid) + 1);
            store_fp<FQ>(sm, 2 * (s + tid), inv * c);
            store_fp<FQ>(sm, 2 * (s + tid) + 1, inv * a);
        }
        __syncthreads();
    }
    F rinv = load_fp<FQ>(sm, 128 + tid);
    __syncthreads();                                     // the tree is dead: the ring may be overwritten again

    // pass 2 (backward): 1/d_j = rinv * prefix_j, then the affine addition; one output per stage
    if (j0 < j1) {
        auto issue2 = [&](const uint2 id, uint64_t j, uint32_t stage) {
            uint4* slot = ring + stage * stage_words;
            const uint64_t i1 = FIRST ? (id.x & 0x7fffffffu) : id.x;
            stage_fp<FQ>(slot, 0, src, 2 * i1);
            stage_fp<FQ>(slot, 3, src, 2 * i1 + 1);
            if (id.y != NO_PARTNER) {
                const uint64_t i2 = FIRST ? (id.y & 0x7fffffffu) : id.y;
                stage_fp<FQ>(slot, 6, src, 2 * i2);
                stage_fp<FQ>(slot, 9, src, 2 * i2 + 1);
                stage_fp<FQ>(slot, 12, prefix, j);
            }
            cp_async_commit();
        };
        uint2 cur = stash[j1 - 1], nxt = make_uint2(0, NO_PARTNER);
        issue2(cur, j1 - 1, 0);
        if (j1 - 1 > j0) nxt = stash[j1 - 2];
        uint32_t it = 0;
        for (uint64_t j = j1; j-- > j0; it++) {
            uint2 nn = make_uint2(0, NO_PARTNER);
            if (j > j0) {
                issue2(nxt, j - 1, (it + 1) & 1);
                if (j > j0 + 1) nn = load_u2_early(stash + (j - 2));
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            const uint4* slot = ring + (it & 1) * stage_words;
            F x1 = unstage_fp<FQ>(slot, 0), y1 = unstage_fp<FQ>(slot, 3);
            bool inf1 = pair_fix<FQ, FIRST>(cur.x, x1, y1);
            if (cur.y != NO_PARTNER) {
                F x2 = unstage_fp<FQ>(slot, 6), y2 = unstage_fp<FQ>(slot, 9);
                const bool inf2 = pair_fix<FQ, FIRST>(cur.y, x2, y2);
                if (inf1 || inf2) {
                    if (inf1) { x1 = x2; y1 = y2; inf1 = inf2; }
                } else if (x1 != x2) {
                    const F dinv = rinv * unstage_fp<FQ>(slot, 12);
                    rinv = rinv * (x2 - x1);
                    const F lam = (y2 - y1) * dinv;
                    const F x3 = lam.sqr() - x1 - x2;
                    y1 = lam * (x1 - x3) - y1;
                    x1 = x3;
                } else if (y1 == y2 && !y1.is_zero()) {
                    const F dinv = rinv * unstage_fp<FQ>(slot, 12);
                    rinv = rinv * (y1 + y1);
                    const F xx = x1.sqr();
                    const F lam = (xx + xx + xx) * dinv;
                    const F x3 = lam.sqr() - x1 - x1;
                    y1 = lam * (x1 - x3) - y1;
                    x1 = x3;
                } else {
                    inf1 = true;
                }
            }
            if (inf1) {
                x1 = F::zero();
                y1 = F::zero();
                x1.v[FQ::N - 1] = 0xffffffffu;
            }
            store_fp<FQ>(dst, 2 * j, x1);
            store_fp<FQ>(dst, 2 * j + 1, y1);
            cur = nxt;
            nxt = nn;
        }
    }
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

}  // namespace apb

using namespace apb;

static const uint32_t CK_MAGIC = 0x434b3031;

struct apb_ck_s {
    uint32_t magic;
    std::recursive_mutex mu;       // entry points on one key serialise here (APB_HANDLE_LOCK)
    int curve;
    size_t n;
    uint32_t F, step;
    void* bases;                   // F * n affine points: copy f holds 2^(step*f) * P_i
    // workspace (grown on demand)
    void* d_scalars; size_t scalars_cap;
    uint32_t *counts, *offsets, *cursors; size_t buckets_cap;
    uint32_t* scan_tmp; size_t scan_tmp_cap;
    uint32_t* entries; size_t entries_cap;
    void* bucket_sums; size_t sums_cap;
    void* partials; int32_t* part_bucket; size_t partial_cap;
    void *stage_a, *stage_b; size_t stage_cap;
    TreeJob* jobs; size_t jobs_cap;
    uint64_t* h_out; size_t h_out_cap;   // pinned
    // batched-affine pair levels
    uint32_t* lvl_words; size_t lvl_words_cap;      // per level: counts, offsets (nbuckets + 1 each)
    void* lvl_pts[2]; size_t lvl_pts_cap[2];        // ping-pong arrays of affine partial sums
    void* lvl_prefix; size_t lvl_prefix_cap;
    uint2* lvl_stash; size_t lvl_stash_cap;
    // cached job table key
    uint32_t jobs_c, jobs_windows;
    std::vector<TreeJob>* h_jobs_a;
    std::vector<TreeJob>* h_jobs_b;
};

template <class T>
static int grow(T** p, size_t* cap, size_t need_bytes) {
    if (*cap >= need_bytes) return APB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = need_bytes + need_bytes / 8 + 256;
    cudaError_t e = cudaMalloc((void**)p, want);
    if (e != cudaSuccess) return set_err(APB_ERR_OOM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    *cap = want;
    return APB_OK;
}

static int ck_alloc(int curve, size_t n, apb_ck_s** out) {
    apb_ck_s* ck = new apb_ck_s();           // value-initialised: all members zero
    ck->magic = CK_MAGIC;
    ck->curve = curve;
    ck->n = n;
    size_t big = (size_t)1 << 24;
    if (const char* e = getenv("APB_MSM_FULL_PRECOMP_MAX")) big = (size_t)atoll(e);
    if (n <= big) { ck->step = 16; ck->F = 16; } else { ck->step = 64; ck->F = 4; }
    if (const char* e = getenv("APB_MSM_STEP")) {
        ck->step = (uint32_t)atoi(e);
        ck->F = (256 + ck->step - 1) / ck->step;
        if (ck->F > (uint32_t)MAX_COPIES) { delete ck; return set_err(APB_ERR_INVALID_ARG, "APB_MSM_STEP too small"); }
    }
    size_t bytes = (n ? n : 1) * ck->F * 96;
    cudaError_t e = cudaMalloc(&ck->bases, bytes);
    if (e != cudaSuccess) { delete ck; return set_err(APB_ERR_OOM, "commitment key: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    *out = ck;
    return APB_OK;
}
struct CkOwner {                 // unique ownership of a key under construction
    apb_ck_s* ck;
    explicit CkOwner(apb_ck_s* c) : ck(c) {}
    CkOwner(const CkOwner&) = delete;
    CkOwner& operator=(const CkOwner&) = delete;
    ~CkOwner() { if (ck) apb_ck_free(ck); }
    apb_ck_s* release() { apb_ck_s* c = ck; ck = nullptr; return c; }
};

static int ck_precompute(apb_ck_s* ck) {
    if (!ck->n) return APB_OK;
    unsigned blocks = (unsigned)((ck->n + 127) / 128);
    if (ck->curve == APB_CURVE_BLS12_381) APB_KLAUNCH(k_ck_precompute<Fq381>, blocks, 128, 0, ck->bases, (uint64_t)ck->n, ck->F, ck->step);
    else APB_KLAUNCH(k_ck_precompute<Fq377>, blocks, 128, 0, ck->bases, (uint64_t)ck->n, ck->F, ck->step);
    APB_CHECK_LAUNCH();
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

// KZG10 setup with a caller-supplied tau: powers_of_g[i] = [tau^i] G computed on the device and
// kept resident (the reference samples tau from OsRng: benches/plonk.rs:98, PC::setup).
// generator_xy: 12 u64 affine generator (Montgomery); tau: 4 u64 Montgomery.
extern "C" int apb_ck_from_tau(int curve, const uint64_t* generator_xy, const uint64_t* tau, size_t n, apb_ck_t* out) {
    APB_API_LOCK();
    if (!out || !generator_xy || !tau) return set_err(APB_ERR_INVALID_ARG, "apb_ck_from_tau: null argument");
    *out = nullptr;
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_ck_from_tau: bad curve %d", curve);
    APB_REQUIRE_INIT();
    apb_ck_s* ck = nullptr;
    int rc = ck_alloc(curve, n, &ck);
    if (rc != APB_OK) return rc;
    CkOwner owner(ck);                          // frees the half-built key on any early return
    if (n) {
        host::Field f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fr381>() : host::Field::make<Fr377>();
        uint64_t h_pow2[64 * 4], cur[4];
        memcpy(cur, tau, 32);
        for (int k = 0; k < 64; k++) { memcpy(h_pow2 + 4 * k, cur, 32); f.sqr(cur, cur); }
        DevBuf b_pow2, b_powers, b_gen;
        APB_CUDA_TRY(b_pow2.alloc(sizeof(h_pow2)));
        APB_CUDA_TRY(b_powers.alloc(n * 32));
        APB_CUDA_TRY(b_gen.alloc(96));
        void *d_pow2 = b_pow2.p, *d_powers = b_powers.p, *d_gen = b_gen.p;
        APB_CUDA_TRY(cudaMemcpyAsync(d_pow2, h_pow2, sizeof(h_pow2), cudaMemcpyHostToDevice, cur_stream()));
        APB_CUDA_TRY(cudaMemcpyAsync(d_gen, generator_xy, 96, cudaMemcpyHostToDevice, cur_stream()));
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (curve == APB_CURVE_BLS12_381) {
            APB_KLAUNCH(k_tau_powers<Fr381>, blocks, 128, 0, d_powers, (uint64_t)n, (const void*)d_pow2);
            APB_KLAUNCH(k_srs_powers<Curve381>, blocks, 128, 0, ck->bases, (const void*)d_powers, (uint64_t)n, (const void*)d_gen);
        } else {
            APB_KLAUNCH(k_tau_powers<Fr377>, blocks, 128, 0, d_powers, (uint64_t)n, (const void*)d_pow2);
            APB_KLAUNCH(k_srs_powers<Curve377>, blocks, 128, 0, ck->bases, (const void*)d_powers, (uint64_t)n, (const void*)d_gen);
        }
        APB_CHECK_LAUNCH();
        APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
        if ((rc = ck_precompute(ck)) != APB_OK) return rc;
    }
    *out = owner.release();
    return APB_OK;
}

// copies `count` resident powers starting at `first` to host memory (12 u64 each)
extern "C" int apb_ck_download(apb_ck_t ck, size_t first, size_t count, uint64_t* out_xy) {
    if (!ck || ck->magic != CK_MAGIC || !out_xy) return set_err(APB_ERR_BAD_HANDLE, "apb_ck_download: bad handle");
    APB_HANDLE_LOCK(ck);
    if (first + count > ck->n) return set_err(APB_ERR_INVALID_ARG, "apb_ck_download: range exceeds key size");
    const char* src = (const char*)ck->bases;
    APB_CUDA_TRY(cudaMemcpyAsync(out_xy, src + first * 96, count * 96, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    return APB_OK;
}

extern "C" int apb_ck_upload(int curve, const uint64_t* xy, size_t n, apb_ck_t* out) {
    APB_API_LOCK();
    if (!out || (!xy && n)) return set_err(APB_ERR_INVALID_ARG, "apb_ck_upload: null argument");
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_ck_upload: bad curve %d", curve);
    APB_REQUIRE_INIT();
    *out = nullptr;
    apb_ck_s* ck = nullptr;
    int rc = ck_alloc(curve, n, &ck);
    if (rc != APB_OK) return rc;
    CkOwner owner(ck);
    if (n) {
        APB_CUDA_TRY(cudaMemcpyAsync(ck->bases, xy, n * 96, cudaMemcpyHostToDevice, cur_stream()));
        if ((rc = ck_precompute(ck)) != APB_OK) return rc;
    }
    *out = owner.release();
    return APB_OK;
}

extern "C" int apb_ck_size(apb_ck_t ck, size_t* n) {
    if (!ck || ck->magic != CK_MAGIC || !n) return set_err(APB_ERR_BAD_HANDLE, "apb_ck_size: bad handle");
    *n = ck->n;
    return APB_OK;
}

extern "C" void apb_ck_free(apb_ck_t ck) {
    if (!ck || ck->magic != CK_MAGIC) return;
    {   // wait for a call that is still using this key's workspaces, then retire the handle
        APB_HANDLE_LOCK(ck);
        cudaStreamSynchronize(cur_stream());
        ck->magic = 0;
    }
    // offsets / cursors / stage_b are interior pointers of counts / stage_a
    cudaFree(ck->bases); cudaFree(ck->d_scalars); cudaFree(ck->counts);
    cudaFree(ck->scan_tmp);
    cudaFree(ck->entries); cudaFree(ck->bucket_sums); cudaFree(ck->partials);
    cudaFree(ck->part_bucket); cudaFree(ck->stage_a); cudaFree(ck->jobs);
    cudaFree(ck->lvl_words); cudaFree(ck->lvl_pts[0]); cudaFree(ck->lvl_pts[1]); cudaFree(ck->lvl_prefix); cudaFree(ck->lvl_stash);
    if (ck->h_out) cudaFreeHost(ck->h_out);
    delete ck->h_jobs_a;
    delete ck->h_jobs_b;
    delete ck;
}

// choose digit width: c*G == step.  Small inputs use narrow digits (few buckets to reduce).
static void choose_geom(const apb_ck_s* ck, size_t max_len, MsmGeom& g, uint32_t scalar_bits) {
    uint32_t c = ck->step >= 16 ? 16 : ck->step;
    if (ck->step % 16 != 0) c = ck->step;
    if (max_len <= 4096 && ck->step % 8 == 0) c = 8;
    if (const char* e = getenv("APB_MSM_C")) {
        uint32_t v = (uint32_t)atoi(e);
        if (v >= 2 && v <= 20 && ck->step % v == 0) c = v;
    }
    g.c = c;
    g.G = ck->step / c;
    g.W = (scalar_bits + c - 1) / c;       // |s| < 2^(bits-1): top digit cannot carry out
    g.hb = 1u << (c - 1);
    g.ck_n = ck->n;
}

// job tables for the two tree stages, for `windows` bucket windows of 2^(c-1) buckets
static bool build_jobs(apb_ck_s* ck, uint32_t c, uint32_t windows, uint32_t& a_bits, uint32_t& b_bits) {
    b_bits = (c - 1) / 2;
    a_bits = (c - 1) - b_bits;
    if (ck->h_jobs_a && ck->jobs_c == c && ck->jobs_windows == windows) return false;
    if (!ck->h_jobs_a) { ck->h_jobs_a = new std::vector<TreeJob>(); ck->h_jobs_b = new std::vector<TreeJob>(); }
    ck->h_jobs_a->clear();
    ck->h_jobs_b->clear();
    const uint32_t R = 1u << a_bits, Cc = 1u << b_bits, hb = 1u << (c - 1);
    const uint32_t per_a = R + Cc, per_b = a_bits + b_bits + 1;
    for (uint32_t w = 0; w < windows; w++) {
        for (uint32_t r = 0; r < R; r++) ck->h_jobs_a->push_back(TreeJob{w * hb + r * Cc, 1, Cc, 0, -1, w * per_a + r});
        for (uint32_t cc = 0; cc < Cc; cc++) ck->h_jobs_a->push_back(TreeJob{w * hb + cc, Cc, R, 0, -1, w * per_a + R + cc});
        // stage B reads stage A's output: rows weighted by hi (bits 0..a-1), columns by lo+1 (bits 0..b)
        for (uint32_t k = 0; k < a_bits; k++) ck->h_jobs_b->push_back(TreeJob{w * per_a, 1, R, 0, (int32_t)k, w * per_b + k});
        for (uint32_t k = 0; k <= b_bits; k++) ck->h_jobs_b->push_back(TreeJob{w * per_a + R, 1, Cc, 1, (int32_t)k, w * per_b + a_bits + k});
    }
    ck->jobs_c = c;
    ck->jobs_windows = windows;
    return true;
}

// bucket entries (scalars x digit positions) one pass can sort: positions are uint32
static uint64_t max_entries_per_pass() {
    uint64_t lim = ((uint64_t)1 << 32) - 1;
    if (const char* e = getenv("APB_MSM_MAX_ENTRIES")) lim = (uint64_t)atoll(e);      // tests force the split path at small sizes
    return lim;
}

template <class CV>
static int run_msm(apb_ck_s* ck, const MsmBatch& B, const void* d_scalars, int mont, uint64_t* out_xyz) {
    typedef typename CV::FR FR;
    typedef typename CV::FQ FQ;
    host::Group grp;
    grp.f = host::Field::make<FQ>();
    const host::Field& f = grp.f;

    size_t max_len = 0, total = 0;
    for (uint32_t j = 0; j < B.k; j++) { max_len = std::max<size_t>(max_len, B.len[j]); total += B.len[j]; }
    auto write_identity = [&](uint32_t j) { memset(out_xyz + 18 * j, 0, 18 * 8); };
    if (total == 0) {
        for (uint32_t j = 0; j < B.k; j++) write_identity(j);
        return APB_OK;
    }
    MsmGeom g;
    choose_geom(ck, max_len, g, FR::BITS);
    const uint32_t windows = B.k * g.G;
    const uint32_t nbuckets = windows * g.hb;
    const uint64_t Mmax = (uint64_t)total * g.W;
    // list positions, bucket offsets and cursors are 32-bit: the batch entry points split larger requests
    if (Mmax + nbuckets >= max_entries_per_pass()) return set_err(APB_ERR_INVALID_ARG, "apb_msm: %llu bucket entries exceed the 32-bit list of one pass", (unsigned long long)Mmax);
    if ((uint64_t)ck->F * ck->n >= ((uint64_t)1 << 31)) return set_err(APB_ERR_INVALID_ARG, "apb_msm: key too large for 31-bit point ids");

    // batched-affine pair levels in front of the XYZZ accumulate: worth it when buckets are long
    // (each level halves them) and the list is large enough to amortise one inversion per CTA
    uint32_t levels = 0, slices = 1;
    {
        // measured on B200 (2^18-point commits, profiles/r01_msm_pair_levels.md): 2 levels for 2-4 polynomials
        // per call, a third one pays from ~2^25 entries; the level arrays (152 B per first-level output)
        // must fit a fixed HBM budget next to the resident table
        uint32_t max_levels = Mmax >= ((uint64_t)1 << 25) ? 3 : 2;
        uint64_t min_entries = (uint64_t)6 << 20, max_bytes = (uint64_t)32 << 30;      // one 2^18-point MSM: no gain
        if (const char* e = getenv("APB_MSM_AFFINE_LEVELS")) max_levels = (uint32_t)atoi(e);
        if (const char* e = getenv("APB_MSM_AFFINE_MIN")) min_entries = (uint64_t)atoll(e);
        if (const char* e = getenv("APB_MSM_AFFINE_MAX_BYTES")) max_bytes = (uint64_t)atoll(e);
        if (max_levels > 6) max_levels = 6;
        const uint64_t need = (Mmax / 2 + nbuckets) * (96 + 48 + 8) + (Mmax / 4 + 2 * (uint64_t)nbuckets) * 96;
        bool want = Mmax >= min_entries && Mmax < ((uint64_t)1 << 31);
        if (want && need > max_bytes) {
            // too large for the budget in one piece: run the stage over equal BUCKET RANGES one after the other
            // (the list is sorted by bucket, so a range of buckets is a contiguous piece of it)
            want = false;
            for (uint32_t S = 2; S <= 64 && nbuckets % S == 0 && nbuckets / S >= 8; S *= 2) {
                const uint64_t Mb = Mmax / S + Mmax / S / 4 + 1024, nbS = nbuckets / S;
                if ((Mb / 2 + nbS) * (96 + 48 + 8) + (Mb / 4 + 2 * nbS) * 96 <= max_bytes) { slices = S; want = true; break; }
            }
        }
        if (want) {
            const uint64_t avg = Mmax / nbuckets;
            while (levels < max_levels && (avg >> (levels + 1)) >= 8) levels++;
        }
        if (!levels) slices = 1;
    }
    if (getenv("APB_MSM_DEBUG")) fprintf(stderr, "apb_msm: k=%u c=%u buckets=%u entries<=%llu pair levels=%u slices=%u\n", B.k, g.c, nbuckets, (unsigned long long)Mmax, levels, slices);
    const uint32_t nbS = nbuckets / slices;                     // buckets per slice
    // entries per slice: exact sizes are read back after the sort; a slice above this bound (skewed
    // scalars) makes the whole call fall back to the plain accumulate
    const uint64_t Mslice = slices == 1 ? Mmax : Mmax / slices + Mmax / slices / 4 + 1024;
    uint64_t U[8];                       // upper bounds of the list length per level (per slice)
    U[0] = Mslice;
    for (uint32_t r = 0; r < levels; r++) U[r + 1] = U[r] / 2 + nbS;

    // chunk size for the accumulate pass: exactly one resident wave of threads
    static int occupancy_known = 0, resident_blocks[3] = {2, 2, 3};
    if (!occupancy_known) {
        occupancy_known = 1;
#ifndef APB_EMU
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_accumulate<FQ, 2, 0>, 128, 0) == cudaSuccess && nb > 0) resident_blocks[0] = nb;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_pairs<FQ, 1, 2>, 128, 0) == cudaSuccess && nb > 0) resident_blocks[1] = nb;
        cudaFuncSetAttribute(k_msm_pairs2<Fq381, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM_BYTES);
        cudaFuncSetAttribute(k_msm_pairs2<Fq381, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM_BYTES);
        cudaFuncSetAttribute(k_msm_pairs2<Fq377, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM_BYTES);
        cudaFuncSetAttribute(k_msm_pairs2<Fq377, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM_BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_pairs2<FQ, 1, 3>, 128, PAIR_SMEM_BYTES) == cudaSuccess && nb > 0) resident_blocks[2] = nb;
#endif
    }
    // pair-level kernel: 2 = cp.async staging through shared memory (default), 1 = register software pipeline
    int pairs_variant = 2;
    if (const char* e = getenv("APB_MSM_PAIRS")) pairs_variant = atoi(e) == 1 ? 1 : 2;
    uint64_t target_threads = (uint64_t)g_num_sms * resident_blocks[0] * 128;
    const uint64_t Macc = U[levels];
    uint32_t E = (uint32_t)((Macc + target_threads - 1) / target_threads);
    if (E < 8) E = 8;
    if (const char* e = getenv("APB_MSM_CHUNK")) E = (uint32_t)atoi(e);
    const uint64_t acc_threads = (Macc + E - 1) / E;
    const uint64_t acc_blocks = (acc_threads + 127) / 128;
    const uint64_t acc_slots = acc_blocks * 128;
    // plan B (only for a sliced call whose slices turn out unbalanced): plain accumulate over the whole list
    uint32_t E_fb = (uint32_t)((Mmax + target_threads - 1) / target_threads);
    if (E_fb < 8) E_fb = 8;
    const uint64_t fb_blocks = ((Mmax + E_fb - 1) / E_fb + 127) / 128;
    const uint64_t slots_alloc = slices > 1 ? std::max<uint64_t>(acc_slots, fb_blocks * 128) : acc_slots;

    int rc;
    if ((rc = grow(&ck->counts, &ck->buckets_cap, (size_t)(nbuckets + 1) * 4 * 3)) != APB_OK) return rc;
    if ((rc = grow(&ck->scan_tmp, &ck->scan_tmp_cap, ((size_t)nbuckets / 1024 + 8) * 4)) != APB_OK) return rc;
    ck->offsets = ck->counts + (nbuckets + 1);
    ck->cursors = ck->offsets + (nbuckets + 1);
    if ((rc = grow(&ck->entries, &ck->entries_cap, (size_t)Mmax * 4)) != APB_OK) return rc;
    if ((rc = grow(&ck->bucket_sums, &ck->sums_cap, (size_t)nbuckets * 192)) != APB_OK) return rc;
    {
        size_t need = slots_alloc * 2 * 192;
        size_t cap2 = ck->partial_cap;
        if ((rc = grow(&ck->partials, &ck->partial_cap, need)) != APB_OK) return rc;
        if (cap2 != ck->partial_cap) {
            if (ck->part_bucket) cudaFree(ck->part_bucket);
            ck->part_bucket = nullptr;
            APB_CUDA_TRY(cudaMalloc((void**)&ck->part_bucket, ck->partial_cap / 192 * 4 + 64));
        }
    }
    const size_t lvl_stride = (size_t)nbS + 1;
    if (levels) {
        if ((rc = grow(&ck->lvl_words, &ck->lvl_words_cap, (size_t)levels * 2 * lvl_stride * 4)) != APB_OK) return rc;
        if ((rc = grow(&ck->lvl_pts[0], &ck->lvl_pts_cap[0], (size_t)U[1] * 96)) != APB_OK) return rc;
        if (levels > 1 && (rc = grow(&ck->lvl_pts[1], &ck->lvl_pts_cap[1], (size_t)U[2] * 96)) != APB_OK) return rc;
        if ((rc = grow(&ck->lvl_prefix, &ck->lvl_prefix_cap, (size_t)U[1] * 48)) != APB_OK) return rc;
        if ((rc = grow(&ck->lvl_stash, &ck->lvl_stash_cap, (size_t)U[1] * 8)) != APB_OK) return rc;
    }
    uint32_t a_bits, b_bits;
    const bool jobs_new = build_jobs(ck, g.c, windows, a_bits, b_bits);
    const uint32_t per_a = (1u << a_bits) + (1u << b_bits), per_b = a_bits + b_bits + 1;
    {
        size_t need_a = (size_t)windows * per_a * 192, need_b = (size_t)windows * per_b * 192;
        size_t cap_a = ck->stage_cap;
        if (cap_a < need_a + need_b) {
            if (ck->stage_a) cudaFree(ck->stage_a);
            ck->stage_a = nullptr;
            ck->stage_cap = 0;
            APB_CUDA_TRY(cudaMalloc(&ck->stage_a, need_a + need_b + 256));
            ck->stage_cap = need_a + need_b;
        }
        ck->stage_b = (char*)ck->stage_a + need_a;
    }
    const size_t njobs_a = ck->h_jobs_a->size(), njobs_b = ck->h_jobs_b->size();
    const size_t jobs_cap_before = ck->jobs_cap;
    if ((rc = grow(&ck->jobs, &ck->jobs_cap, (njobs_a + njobs_b) * sizeof(TreeJob))) != APB_OK) return rc;
    if (jobs_new || jobs_cap_before != ck->jobs_cap) {
        APB_CUDA_TRY(cudaMemcpyAsync(ck->jobs, ck->h_jobs_a->data(), njobs_a * sizeof(TreeJob), cudaMemcpyHostToDevice, cur_stream()));
        APB_CUDA_TRY(cudaMemcpyAsync(ck->jobs + njobs_a, ck->h_jobs_b->data(), njobs_b * sizeof(TreeJob), cudaMemcpyHostToDevice, cur_stream()));
    }
    const size_t out_bytes = (size_t)windows * per_b * 192;
    if (ck->h_out_cap < out_bytes) {
        if (ck->h_out) cudaFreeHost(ck->h_out);
        ck->h_out = nullptr;
        APB_CUDA_TRY(cudaMallocHost((void**)&ck->h_out, out_bytes + 256));
        ck->h_out_cap = out_bytes;
    }

    struct PhaseEvents {               // per-phase timers while profiling; destroyed on every return path
        cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        bool on = false;
        ~PhaseEvents() { if (on) for (int i = 0; i < 6; i++) cudaEventDestroy(e[i]); }
    } pev;
    cudaEvent_t* ev = pev.e;
    if (g_profile) { pev.on = true; for (int i = 0; i < 6; i++) cudaEventCreate(&ev[i]); cudaEventRecord(ev[0], cur_stream()); }
    // 1. histogram  2. scan  3. scatter
    APB_CUDA_TRY(cudaMemsetAsync(ck->counts, 0, (size_t)(nbuckets + 1) * 4 * 3, cur_stream()));
    APB_CUDA_TRY(cudaMemsetAsync(ck->bucket_sums, 0, (size_t)nbuckets * 192, cur_stream()));
    dim3 dgrid((unsigned)((max_len + 255) / 256), B.k);
    auto k_hist = k_msm_digits<FR, 0>;
    auto k_scatter = k_msm_digits<FR, 1>;
    APB_KLAUNCH(k_hist, dgrid, 256, 0, d_scalars, B, g, mont, ck->counts, (const uint32_t*)ck->offsets, ck->cursors, ck->entries);
    {   // offsets = exclusive scan of counts; offsets[nbuckets] = number of sorted entries
        int rc2 = u32_scan(ck->counts, ck->offsets, nbuckets, ck->scan_tmp, ck->offsets + nbuckets);
        if (rc2 != APB_OK) return rc2;
    }
    APB_KLAUNCH(k_scatter, dgrid, 256, 0, d_scalars, B, g, mont, ck->counts, (const uint32_t*)ck->offsets, ck->cursors, ck->entries);
    if (g_profile) cudaEventRecord(ev[1], cur_stream());
    // 4. pair levels (batched-affine)  5. accumulate  6. stitch - over the whole bucket range, or slice by slice
    auto k_acc2 = k_msm_accumulate<FQ, 2, 0>;
    auto k_acc_lvl = k_msm_accumulate<FQ, 2, 1>;
    auto k_pairs_first = k_msm_pairs<FQ, 1, 2>;
    auto k_pairs_next = k_msm_pairs<FQ, 0, 2>;
    auto k_pairs2_first = k_msm_pairs2<FQ, 1, 3>;
    auto k_pairs2_next = k_msm_pairs2<FQ, 0, 3>;
    bool unbalanced = false;
    std::vector<uint32_t> bound(slices + 1);
    if (slices > 1) {
        for (uint32_t sl = 0; sl <= slices; sl++)
            APB_CUDA_TRY(cudaMemcpyAsync(&bound[sl], ck->offsets + (size_t)sl * nbS, 4, cudaMemcpyDeviceToHost, cur_stream()));
        APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
        for (uint32_t sl = 0; sl < slices; sl++) unbalanced = unbalanced || (uint64_t)(bound[sl + 1] - bound[sl]) > Mslice;
    }
    g_last_plan[0] = g.c; g_last_plan[1] = levels; g_last_plan[2] = slices; g_last_plan[3] = unbalanced ? 1u : 0u;
    if (unbalanced) {
        if (getenv("APB_MSM_DEBUG")) fprintf(stderr, "apb_msm: unbalanced slices, plain accumulate\n");
        APB_KLAUNCH(k_acc2, (unsigned)fb_blocks, 128, 0, (const uint32_t*)ck->entries, (const uint32_t*)ck->offsets, nbuckets,
                    (const void*)ck->bases, E_fb, ck->bucket_sums, ck->partials, ck->part_bucket);
        APB_KLAUNCH(k_msm_stitch<FQ>, (unsigned)fb_blocks, 128, 0, (const uint32_t*)ck->offsets, E_fb, (uint64_t)(fb_blocks * 128), ck->bucket_sums,
                    (const void*)ck->partials, (const int32_t*)ck->part_bucket);
    }
    for (uint32_t sl = 0; sl < slices && !unbalanced; sl++) {
        const uint32_t* offsets0 = ck->offsets + (size_t)sl * nbS;          // absolute positions in the entry list
        void* sums = (char*)ck->bucket_sums + (size_t)sl * nbS * 192;
        const uint32_t* acc_offsets = offsets0;
        // grids from the exact size of a slice (the bounds U[] above size the workspace)
        uint64_t Ue[8];
        uint32_t Es = E;
        uint64_t blocks_s = acc_blocks;
        for (uint32_t r = 0; r <= levels; r++) Ue[r] = U[r];
        if (slices > 1) {
            Ue[0] = bound[sl + 1] - bound[sl];
            for (uint32_t r = 0; r < levels; r++) Ue[r + 1] = Ue[r] / 2 + nbS;
            Es = (uint32_t)((Ue[levels] + target_threads - 1) / target_threads);
            if (Es < 8) Es = 8;
            blocks_s = ((Ue[levels] + Es - 1) / Es + 127) / 128;
            if (blocks_s > slots_alloc / 128) { blocks_s = slots_alloc / 128; Es = (uint32_t)((Ue[levels] + blocks_s * 128 - 1) / (blocks_s * 128)); }
        }
        if (levels) {
            uint32_t* cnt = ck->lvl_words;
            uint32_t* off = ck->lvl_words + (size_t)levels * lvl_stride;
            APB_KLAUNCH(k_msm_level_counts, (nbS + 255) / 256, 256, 0, offsets0, nbS, levels, cnt);
            for (uint32_t r = 0; r < levels; r++) {
                uint32_t* off_r = off + (size_t)r * lvl_stride;
                int rc2 = u32_scan(cnt + (size_t)r * lvl_stride, off_r, nbS, ck->scan_tmp, off_r + nbS);
                if (rc2 != APB_OK) return rc2;
            }
            const uint64_t pair_threads = (uint64_t)g_num_sms * resident_blocks[pairs_variant] * 128;
            for (uint32_t r = 0; r < levels; r++) {
                uint32_t Ep = (uint32_t)((Ue[r + 1] + pair_threads - 1) / pair_threads);
                if (Ep < 4) Ep = 4;
                const unsigned blocks = (unsigned)(((Ue[r + 1] + Ep - 1) / Ep + 127) / 128);
                const uint32_t* off_in = r == 0 ? offsets0 : off + (size_t)(r - 1) * lvl_stride;
                const uint32_t* off_out = off + (size_t)r * lvl_stride;
                if (pairs_variant == 2) {
                    if (r == 0)
                        APB_KLAUNCH(k_pairs2_first, blocks, 128, PAIR_SMEM_BYTES, (const uint32_t*)ck->entries, (const void*)ck->bases, off_in,
                                    off_out, nbS, Ep, ck->lvl_pts[0], ck->lvl_prefix, ck->lvl_stash);
                    else
                        APB_KLAUNCH(k_pairs2_next, blocks, 128, PAIR_SMEM_BYTES, (const uint32_t*)nullptr, (const void*)ck->lvl_pts[(r - 1) & 1],
                                    off_in, off_out, nbS, Ep, ck->lvl_pts[r & 1], ck->lvl_prefix, ck->lvl_stash);
                } else if (r == 0)
                    APB_KLAUNCH(k_pairs_first, blocks, 128, 0, (const uint32_t*)ck->entries, (const void*)ck->bases, off_in, off_out, nbS, Ep,
                                ck->lvl_pts[0], ck->lvl_prefix, ck->lvl_stash);
                else
                    APB_KLAUNCH(k_pairs_next, blocks, 128, 0, (const uint32_t*)nullptr, (const void*)ck->lvl_pts[(r - 1) & 1], off_in, off_out,
                                nbS, Ep, ck->lvl_pts[r & 1], ck->lvl_prefix, ck->lvl_stash);
            }
            acc_offsets = off + (size_t)(levels - 1) * lvl_stride;          // slice-relative from here on
        }
        if (levels)
            APB_KLAUNCH(k_acc_lvl, (unsigned)blocks_s, 128, 0, (const uint32_t*)nullptr, acc_offsets, nbS,
                        (const void*)ck->lvl_pts[(levels - 1) & 1], Es, sums, ck->partials, ck->part_bucket);
        else
            APB_KLAUNCH(k_acc2, (unsigned)acc_blocks, 128, 0, (const uint32_t*)ck->entries, (const uint32_t*)ck->offsets, nbuckets,
                        (const void*)ck->bases, E, ck->bucket_sums, ck->partials, ck->part_bucket);
        APB_KLAUNCH(k_msm_stitch<FQ>, (unsigned)blocks_s, 128, 0, acc_offsets, Es, (uint64_t)(blocks_s * 128), sums,
                    (const void*)ck->partials, (const int32_t*)ck->part_bucket);
    }
    if (g_profile) cudaEventRecord(ev[2], cur_stream());
    if (g_profile) cudaEventRecord(ev[3], cur_stream());
    // 6. bucket reduction trees
    APB_KLAUNCH(k_msm_tree<FQ>, (unsigned)((njobs_a + 3) / 4), 128, 0, (const void*)ck->bucket_sums, ck->stage_a, (const TreeJob*)ck->jobs,
                (uint32_t)njobs_a);
    APB_KLAUNCH(k_msm_tree<FQ>, (unsigned)((njobs_b + 3) / 4), 128, 0, (const void*)ck->stage_a, ck->stage_b,
                (const TreeJob*)(ck->jobs + njobs_a), (uint32_t)njobs_b);
    APB_CHECK_LAUNCH();
    if (g_profile) cudaEventRecord(ev[4], cur_stream());
    APB_CUDA_TRY(cudaMemcpyAsync(ck->h_out, ck->stage_b, out_bytes, cudaMemcpyDeviceToHost, cur_stream()));
    APB_CUDA_TRY(cudaStreamSynchronize(cur_stream()));
    if (g_profile) {
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            g_phase_ms[i] = ms;
        }
        g_acc_ms_total += g_phase_ms[1];
        g_points_total += total;
        {
            double m = (double)Mmax, issued = 0.0;
            for (uint32_t r = 0; r < levels; r++) { issued += (m / 2) * 1800.0; m /= 2; }
            g_madds_model += (double)Mmax * 3000.0;
            g_madds_issued += issued + m * 3000.0;
        }
    }

    // 7. host epilogue: Horner over weight bits, fold windows, then normalise all k results with ONE
    //    field inversion (Montgomery's trick over the zz*zzz products)
    const host::Pt* V = reinterpret_cast<const host::Pt*>(ck->h_out);
    std::vector<host::Pt> totals(B.k);
    for (uint32_t j = 0; j < B.k; j++) {
        host::Pt total_pt;
        grp.set_identity(total_pt);
        for (int gw = (int)g.G - 1; gw >= 0; gw--) {
            const host::Pt* v = V + (size_t)(j * g.G + gw) * per_b;
            host::Pt rows, cols;
            grp.set_identity(rows);
            grp.set_identity(cols);
            for (int k = (int)a_bits - 1; k >= 0; k--) { grp.dbl(rows, rows); grp.add(rows, rows, v[k]); }
            for (int k = (int)b_bits; k >= 0; k--) { grp.dbl(cols, cols); grp.add(cols, cols, v[a_bits + k]); }
            for (uint32_t s = 0; s < b_bits; s++) grp.dbl(rows, rows);      // hi * 2^b
            grp.add(rows, rows, cols);
            for (uint32_t s = 0; s < g.c; s++) grp.dbl(total_pt, total_pt);
            grp.add(total_pt, total_pt, rows);
        }
        totals[j] = total_pt;
    }
    std::vector<uint64_t> prod(6 * B.k), prefix(6 * B.k);
    uint64_t run[6], inv[6];
    f.set(run, f.one);
    for (uint32_t j = 0; j < B.k; j++) {
        f.set(&prefix[6 * j], run);
        if (grp.is_identity(totals[j])) continue;
        f.mul(&prod[6 * j], totals[j].zz, totals[j].zzz);
        f.mul(run, run, &prod[6 * j]);
    }
    f.inv(inv, run);
    for (int j = (int)B.k - 1; j >= 0; j--) {
        uint64_t* o = out_xyz + 18 * j;
        if (grp.is_identity(totals[j])) { write_identity(j); continue; }
        uint64_t pinv[6], zi2[6], zi3[6], ax[6], ay[6];
        f.mul(pinv, inv, &prefix[6 * j]);            // 1 / (zz * zzz)
        f.mul(inv, inv, &prod[6 * j]);
        f.mul(zi2, pinv, totals[j].zzz);             // 1 / zz
        f.mul(zi3, pinv, totals[j].zz);              // 1 / zzz
        f.mul(ax, totals[j].x, zi2);
        f.mul(ay, totals[j].y, zi3);
        memcpy(o, ax, 48);
        memcpy(o + 6, ay, 48);
        memcpy(o + 12, f.one, 48);
    }
    return APB_OK;
}

static int msm_dispatch(apb_ck_s* ck, const MsmBatch& B, const void* d_scalars, int mont, uint64_t* out) {
    EventPair ev;
    cudaEventRecord(ev.a, cur_stream());
    int rc = ck->curve == APB_CURVE_BLS12_381 ? run_msm<Curve381>(ck, B, d_scalars, mont, out)
                                               : run_msm<Curve377>(ck, B, d_scalars, mont, out);
    cudaEventRecord(ev.b, cur_stream());
    cudaEventSynchronize(ev.b);
    g_last_ms = ev.ms();
    if (g_profile) g_msm_call_ms_total += g_last_ms;
    return rc;
}

// how many of the polynomials lens[0..avail) go into one pass: at most MAX_BATCH, and the bucket-entry
// list (scalars x digit positions) of the pass must stay below the 32-bit position limit
static uint32_t batch_take(const apb_ck_s* ck, const size_t* lens, size_t avail) {
    const uint32_t bits = ck->curve == APB_CURVE_BLS12_381 ? (uint32_t)Fr381::BITS : (uint32_t)Fr377::BITS;
    const uint64_t limit = max_entries_per_pass();
    size_t max_len = 0, total = 0;
    uint32_t take = 0;
    while (take < (uint32_t)MAX_BATCH && take < avail) {
        const size_t ml = std::max(max_len, lens[take]), tot = total + lens[take];
        MsmGeom g;
        choose_geom(ck, ml, g, bits);
        if (take > 0 && (uint64_t)tot * g.W + (uint64_t)(take + 1) * g.G * g.hb >= limit) break;
        max_len = ml;
        total = tot;
        take++;
    }
    return take;
}

extern "C" int apb_msm_batch(apb_ck_t ck, size_t k, const uint64_t* const* scalars, const size_t* base_offsets,
                             const size_t* lens, int mont, uint64_t* out_xyz) {
    if (!ck || ck->magic != CK_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_msm: bad key handle");
    APB_HANDLE_LOCK(ck);
    if (k == 0) return APB_OK;
    if (!scalars || !lens || !out_xyz) return set_err(APB_ERR_INVALID_ARG, "apb_msm: null argument");
    for (size_t done = 0; done < k;) {
        MsmBatch B;
        memset(&B, 0, sizeof(B));
        B.k = batch_take(ck, lens + done, k - done);
        size_t total = 0;
        for (uint32_t j = 0; j < B.k; j++) {
            size_t off = base_offsets ? base_offsets[done + j] : 0, len = lens[done + j];
            if (off + len > ck->n) return set_err(APB_ERR_TOO_MANY_COEFFS, "apb_msm: %zu scalars at base offset %zu exceed the %zu resident powers", len, off, ck->n);
            if (len && !scalars[done + j]) return set_err(APB_ERR_INVALID_ARG, "apb_msm: null scalars");
            B.scal_off[j] = total;
            B.base_off[j] = off;
            B.len[j] = len;
            total += len;
        }
        int rc = grow(&ck->d_scalars, &ck->scalars_cap, (total ? total : 1) * 32);
        if (rc != APB_OK) return rc;
        for (uint32_t j = 0; j < B.k; j++)
            if (B.len[j])
                APB_CUDA_TRY(cudaMemcpyAsync((char*)ck->d_scalars + B.scal_off[j] * 32, scalars[done + j], B.len[j] * 32, cudaMemcpyHostToDevice, cur_stream()));
        rc = msm_dispatch(ck, B, ck->d_scalars, mont, out_xyz + 18 * done);
        if (rc != APB_OK) return rc;
        done += B.k;
    }
    return APB_OK;
}

extern "C" int apb_msm(apb_ck_t ck, size_t base_offset, const uint64_t* scalars, size_t n, int mont, uint64_t out_xyz[18]) {
    const uint64_t* sp[1] = {scalars};
    size_t off[1] = {base_offset}, len[1] = {n};
    return apb_msm_batch(ck, 1, sp, off, len, mont, out_xyz);
}

extern "C" int apb_msm_dev(apb_ck_t ck, size_t base_offset, const void* d_scalars, size_t n, int mont, uint64_t out_xyz[18]) {
    if (!ck || ck->magic != CK_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_msm_dev: bad key handle");
    APB_HANDLE_LOCK(ck);
    if (!out_xyz || (n && !d_scalars)) return set_err(APB_ERR_INVALID_ARG, "apb_msm_dev: null argument");
    if (base_offset + n > ck->n) return set_err(APB_ERR_TOO_MANY_COEFFS, "apb_msm_dev: %zu scalars at base offset %zu exceed the %zu resident powers", n, base_offset, ck->n);
    MsmBatch B;
    memset(&B, 0, sizeof(B));
    B.k = 1;
    B.base_off[0] = base_offset;
    B.len[0] = n;
    return msm_dispatch(ck, B, d_scalars, mont, out_xyz);
}

// d_scalars: one device buffer; scal_offs in elements
extern "C" int apb_msm_batch_dev(apb_ck_t ck, size_t k, const void* d_scalars, const size_t* scal_offs, const size_t* base_offsets,
                                 const size_t* lens, int mont, uint64_t* out_xyz) {
    if (!ck || ck->magic != CK_MAGIC) return set_err(APB_ERR_BAD_HANDLE, "apb_msm_batch_dev: bad key handle");
    APB_HANDLE_LOCK(ck);
    if (k == 0) return APB_OK;
    if (!d_scalars || !scal_offs || !lens || !out_xyz) return set_err(APB_ERR_INVALID_ARG, "apb_msm_batch_dev: null argument");
    for (size_t done = 0; done < k;) {
        MsmBatch B;
        memset(&B, 0, sizeof(B));
        B.k = batch_take(ck, lens + done, k - done);
        for (uint32_t j = 0; j < B.k; j++) {
            size_t off = base_offsets ? base_offsets[done + j] : 0, len = lens[done + j];
            if (off + len > ck->n) return set_err(APB_ERR_TOO_MANY_COEFFS, "apb_msm: %zu scalars at base offset %zu exceed the %zu resident powers", len, off, ck->n);
            B.scal_off[j] = scal_offs[done + j];
            B.base_off[j] = off;
            B.len[j] = len;
        }
        int rc = msm_dispatch(ck, B, d_scalars, mont, out_xyz + 18 * done);
        if (rc != APB_OK) return rc;
        done += B.k;
    }
    return APB_OK;
}

extern "C" void apb_set_profiling(int on) { g_profile = on; }
// diagnostics (tests assert which path a call took): {digit bits, pair levels, bucket-range slices, unbalanced fallback}
extern "C" void apb_msm_last_plan(uint32_t out[4]) { for (int i = 0; i < 4; i++) out[i] = g_last_plan[i]; }
extern "C" void apb_msm_call_ms(double* whole_calls_ms, int reset) {
    if (whole_calls_ms) *whole_calls_ms = g_msm_call_ms_total;
    if (reset) g_msm_call_ms_total = 0.0;
}
extern "C" void apb_msm_totals(double* accumulate_ms, unsigned long long* points, int reset) {
    if (accumulate_ms) *accumulate_ms = g_acc_ms_total;
    if (points) *points = g_points_total;
    if (reset) { g_acc_ms_total = 0.0; g_points_total = 0; }
}

// host-side group addition of two normalised / Jacobian points (folding per-GPU partial sums)
extern "C" void apb_msm_work(double* model_madds, double* issued_madds, int reset) {
    if (model_madds) *model_madds = g_madds_model;
    if (issued_madds) *issued_madds = g_madds_issued;
    if (reset) { g_madds_model = 0.0; g_madds_issued = 0.0; }
}
extern "C" int apb_g1_add(int curve, const uint64_t a_xyz[18], const uint64_t b_xyz[18], uint64_t out_xyz[18]) {
    APB_API_LOCK();
    if (!a_xyz || !b_xyz || !out_xyz) return set_err(APB_ERR_INVALID_ARG, "apb_g1_add: null argument");
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_g1_add: bad curve");
    host::Group grp;
    grp.f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fq381>() : host::Field::make<Fq377>();
    const host::Field& f = grp.f;
    auto load = [&](const uint64_t* j, host::Pt& p) {      // Jacobian (X, Y, Z) -> XYZZ (X, Y, Z^2, Z^3)
        memcpy(p.x, j, 48);
        memcpy(p.y, j + 6, 48);
        f.sqr(p.zz, j + 12);
        f.mul(p.zzz, p.zz, j + 12);
    };
    host::Pt a, b, r;
    load(a_xyz, a);
    load(b_xyz, b);
    grp.add(r, a, b);
    uint64_t ax[6], ay[6];
    memset(out_xyz, 0, 18 * 8);
    if (!grp.to_affine(ax, ay, r)) return APB_OK;
    memcpy(out_xyz, ax, 48);
    memcpy(out_xyz + 6, ay, 48);
    memcpy(out_xyz + 12, f.one, 48);
    return APB_OK;
}
extern "C" void apb_msm_phase_ms(double out[4]) {
    for (int i = 0; i < 4; i++) out[i] = g_phase_ms[i];
}

extern "C" int apb_g1_compress(int curve, const uint64_t xyz[18], uint8_t out[48]) {
    APB_API_LOCK();
    if (!xyz || !out) return set_err(APB_ERR_INVALID_ARG, "apb_g1_compress: null argument");
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_g1_compress: bad curve");
    host::Field f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fq381>() : host::Field::make<Fq377>();
    memset(out, 0, 48);
    if (f.is_zero(xyz + 12)) { out[47] |= 0x40; return APB_OK; }
    uint64_t x[6], y[6], ny[6];
    if (f.eq(xyz + 12, f.one)) {
        f.to_canonical(x, xyz);
        f.to_canonical(y, xyz + 6);
    } else {       // general Jacobian input: x = X/Z^2, y = Y/Z^3
        uint64_t zi[6], zi2[6], zi3[6], t[6];
        f.inv(zi, xyz + 12);
        f.sqr(zi2, zi);
        f.mul(zi3, zi2, zi);
        f.mul(t, xyz, zi2);
        f.to_canonical(x, t);
        f.mul(t, xyz + 6, zi3);
        f.to_canonical(y, t);
    }
    // ny = p - y (canonical); flag = y > ny
    uint64_t borrow = 0;
    for (int i = 0; i < 6; i++) {
        host::u128 d = (host::u128)f.mod[i] - y[i] - borrow;
        ny[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    bool y_is_larger = false;
    for (int i = 5; i >= 0; i--) {
        if (y[i] != ny[i]) { y_is_larger = y[i] > ny[i]; break; }
    }
    memcpy(out, x, 48);
    if (y_is_larger) out[47] |= 0x80;
    return APB_OK;
}
