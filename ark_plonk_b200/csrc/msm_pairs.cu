// MSM kernels: batched-affine pair levels (see msm.cu for the design).
#include "msm_kernels.cuh"

namespace apb {

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
        uint2 idn[4];                                     // ids of the NEXT group: fetched one group ahead, so the
#pragma unroll                                            // gathers below never wait for an id load first
        for (int q = 0; q < 4; q++) idn[q] = j0 + q < j1 ? W.template next<FIRST>(entries, j0 + q) : make_uint2(0, NO_PARTNER);
        for (uint64_t j = j0; j < j1; j += 4) {
            uint2 id[4];
            F xa[4], xb[4];
#pragma unroll
            for (int q = 0; q < 4; q++) id[q] = idn[q];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (id[q].y != NO_PARTNER) {
                    xa[q] = pair_load_x<FQ, FIRST>(src, id[q].x);
                    xb[q] = pair_load_x<FQ, FIRST>(src, id[q].y);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) idn[q] = j + 4 + q < j1 ? W.template next<FIRST>(entries, j + 4 + q) : make_uint2(0, NO_PARTNER);
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

void msm_launch_level_counts(const uint32_t* offsets0, uint32_t nbuckets, uint32_t levels, uint32_t* cnt) {
    APB_KLAUNCH(k_msm_level_counts, (nbuckets + 255) / 256, 256, 0, offsets0, nbuckets, levels, cnt);
}
template <class FQ>
static void launch_pairs(int first, unsigned blocks, const uint32_t* entries, const void* src, const uint32_t* off_in,
                         const uint32_t* off_out, uint32_t nbuckets, uint32_t E, void* dst, void* prefix, uint2* stash) {
    if (first) APB_KLAUNCH((k_msm_pairs<FQ, 1, 2>), blocks, 128, 0, entries, src, off_in, off_out, nbuckets, E, dst, prefix, stash);
    else APB_KLAUNCH((k_msm_pairs<FQ, 0, 2>), blocks, 128, 0, entries, src, off_in, off_out, nbuckets, E, dst, prefix, stash);
}
void msm_launch_pairs(int curve, int first, unsigned blocks, const uint32_t* entries, const void* src, const uint32_t* off_in,
                      const uint32_t* off_out, uint32_t nbuckets, uint32_t E, void* dst, void* prefix, uint2* stash) {
    if (curve == APB_CURVE_BLS12_381) launch_pairs<Fq381>(first, blocks, entries, src, off_in, off_out, nbuckets, E, dst, prefix, stash);
    else launch_pairs<Fq377>(first, blocks, entries, src, off_in, off_out, nbuckets, E, dst, prefix, stash);
}
int msm_resident_blocks_pairs(int curve) {
    int nb = 0;
#ifndef APB_EMU
    cudaError_t e = curve == APB_CURVE_BLS12_381 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_pairs<Fq381, 1, 2>, 128, 0)
                                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_pairs<Fq377, 1, 2>, 128, 0);
    if (e != cudaSuccess) nb = 0;
#endif
    return nb > 0 ? nb : 2;
}

}  // namespace apb
