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
//     then "add +-copy_f[i] to bucket |d|-1 of effective window g": with step = c = 16 (up to 2^21
//     points; step = c = 20 from 2^22 points) and G = 1 all digit positions share ONE bucket set and
//     no window doublings remain.
//   * digits -> histogram -> exclusive scan -> scatter gives the (bucket, point) pairs sorted
//     by bucket (a counting sort; keys are bucket ids).
//   * large lists first go through batched-affine PAIR LEVELS (k_msm_pairs_coop): the points of every
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
#include <thread>
#include <vector>

#include "msm_kernels.cuh"
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
    // Digit width of the resident table (copy f holds 2^(step f) P_i).  16-bit digits (16 copies, 2^15 buckets) up to
    // 2^21 points; from 2^22 points 20-bit digits (13 copies, 2^19 buckets): 13 instead of 16 additions per point, and the
    // larger bucket reduction (1.45 ms instead of 0.36 ms on B200) is repaid from ~2^21.5 points (2^22: 183 -> 200 Mpts/s,
    // profiles/r02_msm_step20.json).  Keys whose 13-copy table would not fit the HBM budget fall back to 4 copies with
    // 64-bit steps (4 bucket windows of 16-bit digits).
    size_t wide_min = (size_t)1 << 22, table_budget = (size_t)112 << 30;
    if (const char* e = getenv("APB_MSM_WIDE_DIGITS_MIN")) wide_min = (size_t)atoll(e);
    if (const char* e = getenv("APB_MSM_TABLE_BUDGET")) table_budget = (size_t)atoll(e);
    if (n < wide_min && n * 16 * 96 <= table_budget) { ck->step = 16; ck->F = 16; }
    else if (n * 13 * 96 <= table_budget && (uint64_t)n * 13 < ((uint64_t)1 << 31)) { ck->step = 20; ck->F = 13; }
    else { ck->step = 64; ck->F = 4; }
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
    msm_launch_ck_precompute(ck->curve, ck->bases, (uint64_t)ck->n, ck->F, ck->step);
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
        msm_launch_srs_powers(curve, ck->bases, d_powers, (uint64_t)n, (const void*)d_pow2, (const void*)d_gen);
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
        uint32_t max_levels = Mmax >= ((uint64_t)1 << 25) ? 4 : 2;      // r02 sweep: profiles/r02_msm_tune_pairs_variants.json
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
    static int resident_blocks[2][2] = {{0, 0}, {0, 0}};          // [curve][accumulate, pair level]
    const int cv = ck->curve == APB_CURVE_BLS12_381 ? 0 : 1;
    if (!resident_blocks[cv][0]) {
        resident_blocks[cv][0] = msm_resident_blocks_accumulate(ck->curve);
        resident_blocks[cv][1] = msm_resident_blocks_pairs_coop(ck->curve);
    }
    uint64_t target_threads = (uint64_t)g_num_sms * resident_blocks[cv][0] * 128;
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
        if ((rc = grow(&ck->lvl_prefix, &ck->lvl_prefix_cap, (size_t)(U[1] + 32) * 48)) != APB_OK) return rc;
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
    msm_launch_digits(ck->curve, 0, dgrid, d_scalars, B, g, mont, ck->counts, (const uint32_t*)ck->offsets, ck->cursors, ck->entries);
    {   // offsets = exclusive scan of counts; offsets[nbuckets] = number of sorted entries
        int rc2 = u32_scan(ck->counts, ck->offsets, nbuckets, ck->scan_tmp, ck->offsets + nbuckets);
        if (rc2 != APB_OK) return rc2;
    }
    msm_launch_digits(ck->curve, 1, dgrid, d_scalars, B, g, mont, ck->counts, (const uint32_t*)ck->offsets, ck->cursors, ck->entries);
    if (g_profile) cudaEventRecord(ev[1], cur_stream());
    // 4. pair levels (batched-affine)  5. accumulate  6. stitch - over the whole bucket range, or slice by slice
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
        msm_launch_accumulate(ck->curve, 0, (unsigned)fb_blocks, (const uint32_t*)ck->entries, (const uint32_t*)ck->offsets, nbuckets,
                              (const void*)ck->bases, E_fb, ck->bucket_sums, ck->partials, ck->part_bucket);
        msm_launch_stitch(ck->curve, (unsigned)fb_blocks, (const uint32_t*)ck->offsets, E_fb, (uint64_t)(fb_blocks * 128), ck->bucket_sums,
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
            msm_launch_level_counts(offsets0, nbS, levels, cnt);
            for (uint32_t r = 0; r < levels; r++) {
                uint32_t* off_r = off + (size_t)r * lvl_stride;
                int rc2 = u32_scan(cnt + (size_t)r * lvl_stride, off_r, nbS, ck->scan_tmp, off_r + nbS);
                if (rc2 != APB_OK) return rc2;
            }
            const uint64_t pair_threads = (uint64_t)g_num_sms * resident_blocks[cv][1] * 128;
            for (uint32_t r = 0; r < levels; r++) {
                uint32_t Ep = (uint32_t)((Ue[r + 1] + pair_threads - 1) / pair_threads);
                if (Ep < 4) Ep = 4;
                const unsigned blocks = (unsigned)(((Ue[r + 1] + Ep - 1) / Ep + 127) / 128);
                const uint32_t* off_in = r == 0 ? offsets0 : off + (size_t)(r - 1) * lvl_stride;
                const uint32_t* off_out = off + (size_t)r * lvl_stride;
                msm_launch_pairs_coop(ck->curve, r == 0, blocks, r == 0 ? (const uint32_t*)ck->entries : (const uint32_t*)nullptr,
                                          r == 0 ? (const void*)ck->bases : (const void*)ck->lvl_pts[(r - 1) & 1], off_in, off_out, nbS, Ue[r + 1],
                                          Ep, ck->lvl_pts[r & 1], ck->lvl_prefix, (Ue[r + 1] + 31) / 32 * 32, ck->lvl_stash);
            }
            acc_offsets = off + (size_t)(levels - 1) * lvl_stride;          // slice-relative from here on
        }
        if (levels)
            msm_launch_accumulate(ck->curve, 1, (unsigned)blocks_s, (const uint32_t*)nullptr, acc_offsets, nbS,
                                  (const void*)ck->lvl_pts[(levels - 1) & 1], Es, sums, ck->partials, ck->part_bucket);
        else
            msm_launch_accumulate(ck->curve, 0, (unsigned)acc_blocks, (const uint32_t*)ck->entries, (const uint32_t*)ck->offsets, nbuckets,
                                  (const void*)ck->bases, E, ck->bucket_sums, ck->partials, ck->part_bucket);
        msm_launch_stitch(ck->curve, (unsigned)blocks_s, acc_offsets, Es, (uint64_t)(blocks_s * 128), sums,
                          (const void*)ck->partials, (const int32_t*)ck->part_bucket);
    }
    if (g_profile) cudaEventRecord(ev[2], cur_stream());
    if (g_profile) cudaEventRecord(ev[3], cur_stream());
    // 6. bucket reduction trees
    msm_launch_tree(ck->curve, (const void*)ck->bucket_sums, ck->stage_a, (const TreeJob*)ck->jobs, (uint32_t)njobs_a);
    msm_launch_tree(ck->curve, (const void*)ck->stage_a, ck->stage_b, (const TreeJob*)(ck->jobs + njobs_a), (uint32_t)njobs_b);
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
    auto fold_poly = [&](uint32_t j) {
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
    };
    // ~50 group operations per window (~40 us): the windows of a batch are independent, so larger batches are folded by
    // a few host threads (the 16-polynomial batch of a proof: 0.6 -> 0.1 ms; everything here is read-only or per-j)
    const uint32_t fold_threads = B.k >= 4 ? std::min<uint32_t>(B.k, 8) : 1;
    if (fold_threads > 1) {
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < fold_threads; t++)
            pool.emplace_back([&, t]() { for (uint32_t j = t; j < B.k; j += fold_threads) fold_poly(j); });
        for (std::thread& th : pool) th.join();
    } else {
        for (uint32_t j = 0; j < B.k; j++) fold_poly(j);
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
// Folds per-GPU partial sums: out[j] = sum of the pieces p with group[p] == j (normalised Jacobian in, normalised out),
// all k results normalised with ONE field inversion.  The N-GPU commit split hands every rank (k + N - 1) pieces per batch.
extern "C" int apb_g1_fold(int curve, size_t npieces, const uint64_t* pieces_xyz, const uint32_t* group, size_t k, uint64_t* out_xyz) {
    if ((npieces && (!pieces_xyz || !group)) || (k && !out_xyz)) return set_err(APB_ERR_INVALID_ARG, "apb_g1_fold: null argument");
    if (curve != APB_CURVE_BLS12_381 && curve != APB_CURVE_BLS12_377) return set_err(APB_ERR_INVALID_ARG, "apb_g1_fold: bad curve");
    host::Group grp;
    grp.f = curve == APB_CURVE_BLS12_381 ? host::Field::make<Fq381>() : host::Field::make<Fq377>();
    const host::Field& f = grp.f;
    std::vector<host::Pt> tot(k);
    for (size_t j = 0; j < k; j++) grp.set_identity(tot[j]);
    for (size_t p = 0; p < npieces; p++) {
        if (group[p] >= k) return set_err(APB_ERR_INVALID_ARG, "apb_g1_fold: group index out of range");
        const uint64_t* j = pieces_xyz + 18 * p;
        if (f.is_zero(j + 12)) continue;                     // identity
        host::Pt a;
        memcpy(a.x, j, 48);
        memcpy(a.y, j + 6, 48);
        f.sqr(a.zz, j + 12);
        f.mul(a.zzz, a.zz, j + 12);
        grp.add(tot[group[p]], tot[group[p]], a);
    }
    std::vector<uint64_t> prod(6 * k), prefix(6 * k);
    uint64_t run[6], inv[6];
    f.set(run, f.one);
    for (size_t j = 0; j < k; j++) {
        f.set(&prefix[6 * j], run);
        if (grp.is_identity(tot[j])) continue;
        f.mul(&prod[6 * j], tot[j].zz, tot[j].zzz);
        f.mul(run, run, &prod[6 * j]);
    }
    f.inv(inv, run);
    for (size_t jj = k; jj-- > 0;) {
        uint64_t* o = out_xyz + 18 * jj;
        memset(o, 0, 18 * 8);
        if (grp.is_identity(tot[jj])) continue;
        uint64_t pinv[6], zi2[6], zi3[6];
        f.mul(pinv, inv, &prefix[6 * jj]);               // 1 / (zz * zzz)
        f.mul(inv, inv, &prod[6 * jj]);
        f.mul(zi2, pinv, tot[jj].zzz);                   // 1 / zz
        f.mul(zi3, pinv, tot[jj].zz);                    // 1 / zzz
        f.mul(o, tot[jj].x, zi2);
        f.mul(o + 6, tot[jj].y, zi3);
        memcpy(o + 12, f.one, 48);
    }
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
