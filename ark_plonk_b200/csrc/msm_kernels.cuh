// Shared declarations of the MSM translation units (msm.cu: host logic; msm_acc.cu: sort / accumulate / reduce
// kernels; msm_pairs_coop.cu: batched-affine pair levels; msm_setup.cu: key precomputation).  The kernels are split
// over several files only so that nvcc compiles them in parallel; each file exports plain launcher functions.
#pragma once
#include "common.cuh"

namespace apb {

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

// Tree-sum jobs: out[job] = sum over e < m of in[base + e*stride], restricted (selbit >= 0) to
// entries whose weight (e + woff) has bit `selbit` set.
struct TreeJob {
    uint32_t base, stride, m, woff;
    int32_t selbit;
    uint32_t out;
};

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

// ---- launchers (curve: APB_CURVE_*; every launch goes to cur_stream() and is counted in g_launches) ----
// msm_acc.cu
void msm_launch_digits(int curve, int pass, dim3 grid, const void* scalars, const MsmBatch& B, const MsmGeom& g, int mont,
                       uint32_t* counts, const uint32_t* offsets, uint32_t* cursors, uint32_t* entries);
// src 0: `entries` index the resident table `bases`; src 1: `bases` is a level array of affine partial sums
void msm_launch_accumulate(int curve, int src, unsigned blocks, const uint32_t* entries, const uint32_t* offsets, uint32_t nbuckets,
                           const void* bases, uint32_t E, void* bucket_sums, void* partials, int32_t* part_bucket);
void msm_launch_stitch(int curve, unsigned blocks, const uint32_t* offsets, uint32_t E, uint64_t nthreads, void* bucket_sums,
                       const void* partials, const int32_t* part_bucket);
void msm_launch_tree(int curve, const void* in, void* out, const TreeJob* jobs, uint32_t njobs);
int msm_resident_blocks_accumulate(int curve);
// msm_pairs_coop.cu: `ids` = 8 bytes per output, `prefix` = 3 planes of `pstride` 16-byte chunks
void msm_launch_level_counts(const uint32_t* offsets0, uint32_t nbuckets, uint32_t levels, uint32_t* cnt);
void msm_launch_pairs_coop(int curve, int first, unsigned blocks, const uint32_t* entries, const void* src, const uint32_t* off_in,
                           const uint32_t* off_out, uint32_t nbuckets, uint64_t max_out, uint32_t E, void* dst, void* prefix,
                           uint64_t pstride, uint2* ids);
int msm_resident_blocks_pairs_coop(int curve);
// msm_setup.cu
void msm_launch_ck_precompute(int curve, void* bases, uint64_t n, uint32_t F, uint32_t step);
void msm_launch_srs_powers(int curve, void* bases, void* d_powers, uint64_t n, const void* d_pow2, const void* d_gen);

}  // namespace apb
