// MSM kernels: one batched-affine pair level (see msm.cu for where it sits in the MSM).
//
// A bucket that holds m points needs m-1 additions whatever the order.  Adding the points of a bucket PAIRWISE (level r+1
// holds ceil(m_r / 2) partial sums per bucket) makes every addition an affine + affine -> affine one, whose only expensive
// part is 1/(x2 - x1): all the denominators a CTA handles are inverted together (Montgomery's trick: per-thread prefix
// products, a product tree over the 128 thread totals in shared memory, ONE binary-Euclid inversion per CTA), so an
// addition costs 5 M + 1 S instead of the 8 M + 2 S of the XYZZ mixed addition.  After a few levels the buckets are short
// and the remaining list goes through k_msm_accumulate<SRC = 1>.  Level-r lists are sorted by bucket like the entry list;
// off_r = exclusive scan of the per-bucket counts.  Output j of bucket b (local index jl) adds inputs off_r[b] + 2 jl and
// + 2 jl + 1; a trailing odd element is passed through.  Pass 1 walks a thread's outputs forwards (denominators, prefix
// products), pass 2 backwards (inverse, lambda, sum).  Infinity, doubling and P + (-P) inside a pair take a rare slow path.
//
// Data movement is WARP-COOPERATIVE.  The round-1 kernel gave every thread a contiguous range of outputs, so the 32 lanes
// of a warp always touched 32 different 128-byte lines and every LDG.128 / STG.128 cost ~64 cycles of the SM's single
// L1TEX wavefront queue.  Here a warp owns a contiguous range and its lanes take the outputs round-robin (output j of
// iteration i belongs to lane j mod 32), so everything the level itself produces moves as contiguous 512-byte runs:
//   * prefix products live in three 16-byte planes (plane c, output j): one coalesced STG / cp.async per chunk;
//   * the input ids of an output (8 bytes, written by k_msm_pair_ids) are copied as 16 contiguous chunks per iteration;
//   * results are staged in shared memory and leave as 6 fully contiguous 512-byte stores per 32 outputs;
//   * operands are fetched by the whole warp with cp.async: chunk q = lane + 32 m of the iteration's 64 records (96-byte
//     table / level records), so one instruction touches ~6 records instead of 32, and the records of a later level are
//     contiguous anyway.  The copies land in a two-stage ring in shared memory: the fetch of the next iteration costs no
//     registers and overlaps the arithmetic of the current one; each lane reads its own two records back with LDS.128.
// Measured on B200 (profiles/r02_msm_pairs_variants.md): 3.3 -> 3.5 M additions/ms at 2^18 x 8, 3.2 -> 3.7 at 2^22; the
// per-thread-range kernel, a cp.async variant of it, and this kernel compiled for 4 CTAs per SM (128 registers) are all
// within +-3 % of each other - ncu shows the same picture for all of them: 34 % issue-active, `wait` (fixed-latency
// dependency) is the dominant stall, i.e. the kernels run at ~70 % of what the carry-chained IMAD.WIDE.X stream allows and
// neither occupancy (8 / 12 / 16 warps) nor gather coalescing moves that.
#include <stdlib.h>

#include "msm_pair_common.cuh"

namespace apb {

static const int PC_RS = 6;                          // record stride in shared memory, in 16-byte chunks (x, y)
static const int PC_REC = 64 * PC_RS;                // one stage: the 64 input records of 32 outputs
static const int PC_WARP = 2 * PC_REC + 3 * 32 + 16; // per warp: two record stages, 32 prefix products (3 planes of 32), 64 ids
static const size_t PAIRS_COOP_SMEM = (size_t)4 * PC_WARP * 16;      // 4 warps per CTA: 56320 bytes

APB_D void warp_barrier() {
#ifdef APB_EMU
    apb_emu::warp_sync();        // the CPU emulation runs lanes as free-running OS threads
#else
    __syncwarp();
#endif
}
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
// field element <- CH consecutive 16-byte chunks at `p` (stride `step` chunks)
template <class FQ>
APB_D Fp<FQ> unstage(const uint4* p, int step = 1) {
    Fp<FQ> r;
#pragma unroll
    for (int i = 0; i < FQ::N / 4; i++) {
        const uint4 t = p[i * step];
        r.v[4 * i] = t.x; r.v[4 * i + 1] = t.y; r.v[4 * i + 2] = t.z; r.v[4 * i + 3] = t.w;
    }
    return r;
}
template <class FQ>
APB_D void stage(uint4* p, const Fp<FQ>& a, size_t step = 1) {
#pragma unroll
    for (int i = 0; i < FQ::N / 4; i++) p[i * step] = make_uint4(a.v[4 * i], a.v[4 * i + 1], a.v[4 * i + 2], a.v[4 * i + 3]);
}

// Everything that is not "two finite points with different x": a single input passed through, an operand at infinity,
// doubling, P + (-P) - and the (2^-64) false positives of the caller's cheap filter.  Writes the result into the lane's
// slot and returns the updated running inverse.
template <class FQ, int FIRST>
__device__ __noinline__ Fp<FQ> pair_add_slow(uint4* st, uint32_t lane, uint2 own, Fp<FQ> rinv, Fp<FQ> pre) {
    typedef Fp<FQ> F;
    F x1 = unstage<FQ>(st + lane * PC_RS), y1 = unstage<FQ>(st + lane * PC_RS + 3);
    bool inf1 = pair_fix<FQ, FIRST>(own.x, x1, y1);
    if (own.y != NO_PARTNER) {
        F x2 = unstage<FQ>(st + (32 + lane) * PC_RS), y2 = unstage<FQ>(st + (32 + lane) * PC_RS + 3);
        const bool inf2 = pair_fix<FQ, FIRST>(own.y, x2, y2);
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
    stage<FQ>(st + lane * PC_RS, x1);
    stage<FQ>(st + lane * PC_RS + 3, y1);
    return rinv;
}

__global__ void k_msm_level_counts(const uint32_t* offsets0, uint32_t nbuckets, uint32_t levels, uint32_t* cnt) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    uint32_t c = offsets0[b + 1] - offsets0[b];
    for (uint32_t r = 0; r < levels; r++) {
        c = (c + 1) >> 1;
        cnt[(size_t)r * (nbuckets + 1) + b] = c;
    }
}
void msm_launch_level_counts(const uint32_t* offsets0, uint32_t nbuckets, uint32_t levels, uint32_t* cnt) {
    APB_KLAUNCH(k_msm_level_counts, (nbuckets + 255) / 256, 256, 0, offsets0, nbuckets, levels, cnt);
}

// ids[j] = the (one or two) inputs output j of the level adds: entries of the sorted list (FIRST) or positions in the
// previous level's array; second = NO_PARTNER for the odd element a bucket passes through.  One WARP per bucket (strided
// over the buckets): the lanes take the bucket's outputs round-robin, so the entry reads and the id writes are
// contiguous (a thread per output with a binary search over the offsets took 5 % of the level's time).
template <int FIRST>
__global__ void __launch_bounds__(256) k_msm_pair_ids(const uint32_t* entries, const uint32_t* off_in, const uint32_t* off_out,
                                                      uint32_t nbuckets, uint2* ids) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nbuckets; b += warps) {
        const uint32_t o0 = off_out[b], cnt = off_out[b + 1] - o0, i0 = off_in[b], iend = off_in[b + 1];
        for (uint32_t jl = lane; jl < cnt; jl += 32) {
            const uint32_t ipos = i0 + 2 * jl;
            const bool pair = ipos + 1 < iend;
            uint2 r;
            r.x = FIRST ? entries[ipos] : ipos;
            r.y = pair ? (FIRST ? entries[ipos + 1] : ipos + 1) : NO_PARTNER;
            ids[o0 + jl] = r;
        }
    }
}

template <class FQ, int FIRST, int MINB>
__global__ void __launch_bounds__(128, MINB) k_msm_pairs_coop(const uint2* ids, const void* src, const uint32_t* off_out, uint32_t nbuckets,
                                                              uint32_t E, void* dst, uint4* prefix, uint64_t pstride) {
    typedef Fp<FQ> F;
    static_assert(FQ::N == 12, "records are 6 chunks of 16 bytes");
    APB_DYN_SMEM(smem_raw);
    uint4* sm = reinterpret_cast<uint4*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint4* const wsm = sm + (size_t)warp * PC_WARP;
    const uint64_t Mout = off_out[nbuckets];
    const uint64_t w0 = ((uint64_t)blockIdx.x * 4 + warp) * 32 * E;      // this warp owns outputs [w0, w0 + 32 E)
    uint32_t n_it = 0;                                                   // iterations of 32 outputs (warp-uniform)
    if (w0 < Mout) {
        const uint64_t rem = (Mout - w0 + 31) / 32;
        n_it = rem < E ? (uint32_t)rem : E;
    }
    const uint32_t* ids32 = reinterpret_cast<const uint32_t*>(ids);
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    const uint2 none = make_uint2(0, NO_PARTNER);

    // ---- pass 1 (forward): denominators x2 - x1 and this lane's running prefix products -------------------------------
    F run = F::one();
    if (n_it) {
        uint32_t ge[6];                              // ids behind this lane's 6 chunk copies of the iteration after next
        auto load_gids = [&](uint32_t it) {
            const uint64_t jb = w0 + 32ull * it;
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const uint32_t r = (lane + 32 * m) / 3;                  // record: output r >> 1, input r & 1
                ge[m] = jb + (r >> 1) < Mout ? ids32[2 * jb + r] : NO_PARTNER;
            }
        };
        auto issue = [&](uint32_t it) {
            uint4* st = wsm + (it & 1) * 192;       // pass 1 stages only x: 64 records of 3 chunks
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const uint32_t q = lane + 32 * m, r = q / 3, c = q - 3 * r, e = ge[m];
                if (e != NO_PARTNER)
                    cp_async16(st + ((r & 1) * 32 + (r >> 1)) * 3 + c, src4 + (uint64_t)(FIRST ? (e & 0x7fffffffu) : e) * 6 + c);
            }
            cp_async_commit();
        };
        uint2 own = w0 + lane < Mout ? ids[w0 + lane] : none;
        load_gids(0);
        issue(0);
        if (n_it > 1) load_gids(1);
        for (uint32_t it = 0; it < n_it; it++) {
            const uint64_t j = w0 + 32ull * it + lane;
            uint2 own_next = none;
            if (it + 1 < n_it) {
                issue(it + 1);
                if (it + 2 < n_it) load_gids(it + 2);
                if (j + 32 < Mout) own_next = ids[j + 32];
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            warp_barrier();                          // every lane's copies of this stage have landed
            const uint4* st = wsm + (it & 1) * 192;
            if (own.y != NO_PARTNER) {               // (lanes past the end of the level carry `none`)
                const F xa = unstage<FQ>(st + lane * 3), xb = unstage<FQ>(st + (32 + lane) * 3);
                F d = xb - xa;
                const bool special = FIRST ? (xa.is_zero() || xb.is_zero() || d.is_zero())
                                           : (xa.v[FQ::N - 1] == 0xffffffffu || xb.v[FQ::N - 1] == 0xffffffffu || d.is_zero());
                if (special) d = pair_den_special<FQ, FIRST>(src, own.x, own.y);
                stage<FQ>(prefix + j, run, pstride);
                run = run.mul_lazy(d);               // only ever a product operand again (see Fp::mul_impl)
            }
            warp_barrier();                          // this stage may be overwritten by the fetch after next
            own = own_next;
        }
    }

    // ---- 1 / (this thread's product): product tree over the CTA, one inversion ------------------------------------------
    __syncthreads();                                 // all warps are done with their staging rings
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
    // (the leaves may be in [0, 2p); every tree node is a full product of two of them, hence canonical)
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
    __syncthreads();                                 // the tree is dead: the rings may be overwritten again

    // ---- pass 2 (backward): 1/d = rinv * prefix, the affine addition, coalesced store ---------------------------------------
    // One fetch is in flight at a time: it is issued right after the current iteration's operands were found in shared
    // memory and has the whole arithmetic of the iteration to land (records into the other stage, this lane's prefix
    // product into its private slot, which was read into registers just before).  The ids the NEXT fetch needs are copied
    // into shared memory one iteration earlier still, so no id is ever waited for and none lives in a register.
    if (n_it) {
        uint4* dst4 = reinterpret_cast<uint4*>(dst);
        uint4* const pbuf = wsm + 2 * PC_REC;                            // [3][32] prefix chunks, slot = lane
        uint32_t* const idbuf = reinterpret_cast<uint32_t*>(wsm + 2 * PC_REC + 96);      // ids32 of one iteration (64 words)
        auto fetch_ids = [&](uint32_t it) {                              // 64 ids = 16 chunks, contiguous in `ids`
            const uint64_t jb = w0 + 32ull * it;
            if (lane < 16 && jb + 2 * lane < Mout)                       // (a chunk = 2 outputs; the tail chunk may be half valid)
                cp_async16(reinterpret_cast<uint4*>(idbuf) + lane, reinterpret_cast<const uint4*>(ids + jb) + lane);
            cp_async_commit();
        };
        auto issue = [&](uint32_t it, uint32_t stg) {                    // ids of iteration `it` are in idbuf
            uint4* st = wsm + stg * PC_REC;
            const uint64_t jb = w0 + 32ull * it;
#pragma unroll
            for (int m = 0; m < 12; m++) {
                const uint32_t q = lane + 32 * m, r = q / 6, c = q - 6 * r;
                const uint32_t e = jb + (r >> 1) < Mout ? idbuf[r] : NO_PARTNER;
                if (e != NO_PARTNER)
                    cp_async16(st + ((r & 1) * 32 + (r >> 1)) * PC_RS + c, src4 + (uint64_t)(FIRST ? (e & 0x7fffffffu) : e) * 6 + c);
            }
            if (jb + lane < Mout) {                                      // (unused when the output has no partner)
#pragma unroll
                for (int c = 0; c < 3; c++) cp_async16(pbuf + c * 32 + lane, prefix + c * pstride + jb + lane);
            }
            cp_async_commit();
        };
        auto own_ids = [&](uint32_t it) {
            return w0 + 32ull * it + lane < Mout ? make_uint2(idbuf[2 * lane], idbuf[2 * lane + 1]) : none;
        };
        fetch_ids(n_it - 1);
        cp_async_wait<0>();
        warp_barrier();
        issue(n_it - 1, 0);
        uint2 own = own_ids(n_it - 1);
        warp_barrier();                              // every lane has read the ids
        if (n_it > 1) fetch_ids(n_it - 2);
        uint32_t stg = 0;
        for (uint32_t it = n_it; it-- > 0; stg ^= 1) {
            const uint64_t jb = w0 + 32ull * it, j = jb + lane;
            cp_async_wait<0>();
            warp_barrier();                          // operands of `it` (and the ids of `it - 1`) have landed
            uint4* st = wsm + stg * PC_REC;
            const F pre = unstage<FQ>(pbuf + lane, 32);
            uint2 own_next = none;
            if (it > 0) {
                issue(it - 1, stg ^ 1);
                own_next = own_ids(it - 1);
            }
            const uint4 *r1 = st + lane * PC_RS, *r2 = st + (32 + lane) * PC_RS;
            if (j < Mout) {
                bool fast = own.y != NO_PARTNER;
                F dx;
                if (fast) {
                    const F x1 = unstage<FQ>(r1), x2 = unstage<FQ>(r2);
                    dx = x2 - x1;
                    // cheap supersets of "x is the infinity marker / zero" and "dx == 0": rare, the slow path decides exactly
                    const bool odd = FIRST ? (((x1.v[0] | x1.v[FQ::N - 1]) == 0) || ((x2.v[0] | x2.v[FQ::N - 1]) == 0))
                                           : (x1.v[FQ::N - 1] == 0xffffffffu || x2.v[FQ::N - 1] == 0xffffffffu);
                    fast = !odd && (dx.v[0] | dx.v[FQ::N - 1]) != 0;
                }
                if (fast) {
                    const F dinv = rinv.mul_lazy(pre);                   // dinv, rinv, lam: product operands only
                    rinv = rinv.mul_lazy(dx);
                    F y1 = unstage<FQ>(r1 + 3), y2 = unstage<FQ>(r2 + 3);
                    if (FIRST) {
                        if (own.x >> 31) y1 = y1.neg();
                        if (own.y >> 31) y2 = y2.neg();
                    }
                    const F lam = (y2 - y1).mul_lazy(dinv);
                    const F x3 = lam.sqr() - unstage<FQ>(r1) - unstage<FQ>(r2);
                    const F y3 = lam * (unstage<FQ>(r1) - x3) - y1;
                    stage<FQ>(st + lane * PC_RS, x3);    // this lane's own first-input slot: nobody else reads it
                    stage<FQ>(st + lane * PC_RS + 3, y3);
                } else {
                    rinv = pair_add_slow<FQ, FIRST>(st, lane, own, rinv, pre);
                }
            }
            warp_barrier();                          // results staged; every lane is past its id reads
            if (it > 1) fetch_ids(it - 2);
            // the 32 results of the iteration are 3072 contiguous bytes of the level array
#pragma unroll
            for (int m = 0; m < 6; m++) {
                const uint32_t q = lane + 32 * m;
                if (jb + q / 6 < Mout) dst4[jb * 6 + q] = st[q];
            }
            warp_barrier();                          // stage free for the next fetch
            own = own_next;
        }
    }
}

template <class FQ, int FIRST, int MINB>
static void launch_pairs_coop_k(unsigned blocks, const uint2* ids, const void* src, const uint32_t* off_out, uint32_t nbuckets, uint32_t E,
                                void* dst, void* prefix, uint64_t pstride) {
    static bool attr_set = false;
    if (!attr_set) {
        attr_set = true;
        cudaFuncSetAttribute(k_msm_pairs_coop<FQ, FIRST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIRS_COOP_SMEM);
    }
    APB_KLAUNCH((k_msm_pairs_coop<FQ, FIRST, MINB>), blocks, 128, PAIRS_COOP_SMEM, ids, src, off_out, nbuckets, E, dst, (uint4*)prefix, pstride);
}
template <class FQ>
static void launch_pairs_coop(int first, unsigned blocks, const uint32_t* entries, const void* src, const uint32_t* off_in,
                              const uint32_t* off_out, uint32_t nbuckets, uint64_t max_out, uint32_t E, void* dst, void* prefix,
                              uint64_t pstride, uint2* ids) {
    (void)max_out;
    unsigned idb = (nbuckets + 7) / 8;                // 8 warps per CTA, one bucket per warp and round
    if (idb > 148 * 8 * 4) idb = 148 * 8 * 4;
    if (first) {
        APB_KLAUNCH(k_msm_pair_ids<1>, idb, 256, 0, entries, off_in, off_out, nbuckets, ids);
        launch_pairs_coop_k<FQ, 1, 3>(blocks, ids, src, off_out, nbuckets, E, dst, prefix, pstride);
    } else {
        APB_KLAUNCH(k_msm_pair_ids<0>, idb, 256, 0, entries, off_in, off_out, nbuckets, ids);
        launch_pairs_coop_k<FQ, 0, 3>(blocks, ids, src, off_out, nbuckets, E, dst, prefix, pstride);
    }
}
void msm_launch_pairs_coop(int curve, int first, unsigned blocks, const uint32_t* entries, const void* src, const uint32_t* off_in,
                           const uint32_t* off_out, uint32_t nbuckets, uint64_t max_out, uint32_t E, void* dst, void* prefix,
                           uint64_t pstride, uint2* ids) {
    if (curve == APB_CURVE_BLS12_381) launch_pairs_coop<Fq381>(first, blocks, entries, src, off_in, off_out, nbuckets, max_out, E, dst, prefix, pstride, ids);
    else launch_pairs_coop<Fq377>(first, blocks, entries, src, off_in, off_out, nbuckets, max_out, E, dst, prefix, pstride, ids);
}
int msm_resident_blocks_pairs_coop(int curve) {
    (void)curve;
    return 3;        // compiled for 3 CTAs per SM (168 registers per thread; 56 KB of shared memory per CTA)
}

}  // namespace apb
