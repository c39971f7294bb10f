// MSM kernels: scalar recoding + counting sort, XYZZ bucket accumulation, stitch, reduction trees.
// (See msm.cu for the design; this file only holds the kernels and their launchers.)
#include "msm_kernels.cuh"

namespace apb {

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


template <class FR>
static void launch_digits(int pass, dim3 grid, const void* scalars, const MsmBatch& B, const MsmGeom& g, int mont, uint32_t* counts,
                          const uint32_t* offsets, uint32_t* cursors, uint32_t* entries) {
    if (pass == 0) APB_KLAUNCH((k_msm_digits<FR, 0>), grid, 256, 0, scalars, B, g, mont, counts, offsets, cursors, entries);
    else APB_KLAUNCH((k_msm_digits<FR, 1>), grid, 256, 0, scalars, B, g, mont, counts, offsets, cursors, entries);
}
void msm_launch_digits(int curve, int pass, dim3 grid, const void* scalars, const MsmBatch& B, const MsmGeom& g, int mont,
                       uint32_t* counts, const uint32_t* offsets, uint32_t* cursors, uint32_t* entries) {
    if (curve == APB_CURVE_BLS12_381) launch_digits<Fr381>(pass, grid, scalars, B, g, mont, counts, offsets, cursors, entries);
    else launch_digits<Fr377>(pass, grid, scalars, B, g, mont, counts, offsets, cursors, entries);
}
template <class FQ>
static void launch_accumulate(int src, unsigned blocks, const uint32_t* entries, const uint32_t* offsets, uint32_t nbuckets,
                              const void* bases, uint32_t E, void* bucket_sums, void* partials, int32_t* part_bucket) {
    if (src == 0) APB_KLAUNCH((k_msm_accumulate<FQ, 2, 0>), blocks, 128, 0, entries, offsets, nbuckets, bases, E, bucket_sums, partials, part_bucket);
    else APB_KLAUNCH((k_msm_accumulate<FQ, 2, 1>), blocks, 128, 0, entries, offsets, nbuckets, bases, E, bucket_sums, partials, part_bucket);
}
void msm_launch_accumulate(int curve, int src, unsigned blocks, const uint32_t* entries, const uint32_t* offsets, uint32_t nbuckets,
                           const void* bases, uint32_t E, void* bucket_sums, void* partials, int32_t* part_bucket) {
    if (curve == APB_CURVE_BLS12_381) launch_accumulate<Fq381>(src, blocks, entries, offsets, nbuckets, bases, E, bucket_sums, partials, part_bucket);
    else launch_accumulate<Fq377>(src, blocks, entries, offsets, nbuckets, bases, E, bucket_sums, partials, part_bucket);
}
void msm_launch_stitch(int curve, unsigned blocks, const uint32_t* offsets, uint32_t E, uint64_t nthreads, void* bucket_sums,
                       const void* partials, const int32_t* part_bucket) {
    if (curve == APB_CURVE_BLS12_381) APB_KLAUNCH(k_msm_stitch<Fq381>, blocks, 128, 0, offsets, E, nthreads, bucket_sums, partials, part_bucket);
    else APB_KLAUNCH(k_msm_stitch<Fq377>, blocks, 128, 0, offsets, E, nthreads, bucket_sums, partials, part_bucket);
}
void msm_launch_tree(int curve, const void* in, void* out, const TreeJob* jobs, uint32_t njobs) {
    const unsigned blocks = (njobs + 3) / 4;
    if (curve == APB_CURVE_BLS12_381) APB_KLAUNCH(k_msm_tree<Fq381>, blocks, 128, 0, in, out, jobs, njobs);
    else APB_KLAUNCH(k_msm_tree<Fq377>, blocks, 128, 0, in, out, jobs, njobs);
}
int msm_resident_blocks_accumulate(int curve) {
    int nb = 0;
#ifndef APB_EMU
    cudaError_t e = curve == APB_CURVE_BLS12_381 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_accumulate<Fq381, 2, 0>, 128, 0)
                                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_msm_accumulate<Fq377, 2, 0>, 128, 0);
    if (e != cudaSuccess) nb = 0;
#endif
    return nb > 0 ? nb : 2;
}

}  // namespace apb
