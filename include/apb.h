/*
 * apb.h - C ABI of the B200-native proving hot path of heliaxdev/ark-plonk.
 *
 * This is the boundary a Rust `apb-sys` crate would bind.  Each entry
 * point cites the reference interface it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - Field elements cross the boundary as little-endian u64 limbs in MONTGOMERY form
 *     (R = 2^256 for Fr, 2^384 for Fq): byte-identical to arkworks 0.3 `Fp256` / `Fp384`
 *     in-memory limbs.  MSM scalars may also be passed canonical (what `into_repr()` yields).
 *   - G1 affine bases are packed 96-byte records x || y (6+6 u64, Montgomery).  The point at
 *     infinity is the record of all zero bytes (arkworks keeps an out-of-band `infinity` flag).
 *   - Every function returns an `apb_status`; nothing unwinds across the boundary;
 *     `apb_last_error()` gives a thread-local message.
 *   - Handles own device memory (bases, twiddles) which stays resident in HBM until freed.
 *   - `_dev` variants take device pointers and enqueue on the CALLING THREAD's stream - the one
 *     given to `apb_set_stream` (a cudaStream_t), else the library's own - without a host
 *     synchronisation; buffers produced on that stream need no external sync.  The plain variants
 *     block until results are in host memory (the semantics of the reference's synchronous Rust
 *     calls).  Entry points on one handle serialise on that handle; different handles (their
 *     workspaces are private) can be driven concurrently from different threads / streams.
 *   - One process drives one GPU (`apb_init(device)`); multi-GPU sharding is done by the
 *     host layer with one process per GPU.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with
 *     APB_ERR_CUDA.
 */
#ifndef APB_H
#define APB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    APB_OK = 0,
    APB_ERR_INVALID_ARG = 1,     /* null pointer, bad enum, bad length */
    APB_ERR_TOO_MANY_COEFFS = 2, /* ark-poly-commit Error::TooManyCoefficients (plonk-core/src/error.rs:96-107) */
    APB_ERR_DOMAIN_TOO_LARGE = 3,/* log_n > TWO_ADICITY (plonk-core/src/proof_system/prover.rs:169-173 -> Error::InvalidEvalDomainSize) */
    APB_ERR_BAD_HANDLE = 4,
    APB_ERR_CUDA = 5,            /* CUDA runtime error or no device */
    APB_ERR_OOM = 6
} apb_status;

/* curve / field selectors: the test matrix of plonk-core/src/test.rs:84-115 */
#define APB_CURVE_BLS12_381 0
#define APB_CURVE_BLS12_377 1

/* transform kinds: ark_poly::EvaluationDomain::{fft, ifft, coset_fft, coset_ifft} */
#define APB_NTT_FFT 0
#define APB_NTT_IFFT 1
#define APB_NTT_COSET_FFT 2
#define APB_NTT_COSET_IFFT 3

typedef struct apb_ck_s* apb_ck_t;          /* resident CommitterKey powers (SonicKZG10 CommitterKey::powers_of_g) */
typedef struct apb_domain_s* apb_domain_t;  /* resident Radix2EvaluationDomain (twiddles) */

/* ---- library ---------------------------------------------------------------------------- */
int apb_init(int device);                   /* idempotent; binds this process to one GPU (every entry point re-binds its calling thread) */
/* stream (cudaStream_t, as void*) that the calling host thread's later calls enqueue on; NULL = the
 * library's own stream.  SURVEY 8(b): "apb_*_dev(...): same, device pointers + cudaStream_t, no host sync". */
int apb_set_stream(void* cuda_stream);
void* apb_stream(void);                     /* the stream this thread's calls currently use */
const char* apb_last_error(void);
const char* apb_version(void);

/* ---- commitment key: replaces PC::trim + CommitterKey (plonk-core/src/circuit.rs:236,276,310) */
/* xy: n packed affine points (12 u64 each, Montgomery).  Uploaded once; a table of 2^(c*k)
 * multiples may be precomputed on the device (see DESIGN.md "MSM"). */
int apb_ck_upload(int curve, const uint64_t* xy, size_t n, apb_ck_t* out);
/* KZG10 setup from a caller-supplied tau (benches/plonk.rs:98 `PC::setup`, there with OsRng):
 * powers_of_g[i] = [tau^i] G computed on the device and kept resident.  generator_xy: 12 u64,
 * tau: 4 u64, both Montgomery. */
int apb_ck_from_tau(int curve, const uint64_t* generator_xy, const uint64_t* tau, size_t n, apb_ck_t* out);
int apb_ck_download(apb_ck_t ck, size_t first, size_t count, uint64_t* out_xy);
int apb_ck_size(apb_ck_t ck, size_t* n);
void apb_ck_free(apb_ck_t ck);

/* ---- MSM: replaces ark_ec::msm::VariableBaseMSM::multi_scalar_mul
 *      (plonk-core/src/commitment.rs:45; inside KZG10::commit/open reached from
 *       plonk-core/src/proof_system/prover.rs:213,290,313,316,362,388,459,579,582,606,609) */
/* result = sum_{i<n} scalars[i] * bases[base_offset + i], returned as a normalised Jacobian
 * point X,Y,Z (Z = 1 in Montgomery form, or X=Y=Z=0.. with Z == 0 for the identity).
 * scalars: n x 4 u64; `scalars_are_montgomery` != 0 -> Fr Montgomery limbs (polynomial
 * coefficients as stored), else canonical integers < r.  n == 0 is valid (identity). */
int apb_msm(apb_ck_t ck, size_t base_offset, const uint64_t* scalars, size_t n,
            int scalars_are_montgomery, uint64_t out_xyz[18]);
/* k independent MSMs over the same key in one pass (PC::commit of several polynomials,
 * e.g. the 4 wire polynomials at prover.rs:213).  out: k x 18 u64. */
int apb_msm_batch(apb_ck_t ck, size_t k, const uint64_t* const* scalars, const size_t* base_offsets,
                  const size_t* lens, int scalars_are_montgomery, uint64_t* out_xyz);
/* device-resident scalars (e.g. an ifft result that never left HBM) */
int apb_msm_dev(apb_ck_t ck, size_t base_offset, const void* d_scalars, size_t n,
                int scalars_are_montgomery, uint64_t out_xyz[18]);

/* k MSMs whose scalars already sit in ONE device buffer; scal_offs are element offsets into it */
int apb_msm_batch_dev(apb_ck_t ck, size_t k, const void* d_scalars, const size_t* scal_offs,
                      const size_t* base_offsets, const size_t* lens, int scalars_are_montgomery,
                      uint64_t* out_xyz);

/* out = a + b for two points in the (X, Y, Z) form apb_msm returns; folds per-GPU partial sums of a
 * point-split MSM (host side, a handful of field operations) */
int apb_g1_add(int curve, const uint64_t a_xyz[18], const uint64_t b_xyz[18], uint64_t out_xyz[18]);
/* out[j] = sum of the pieces p with group[p] == j, j < k (npieces normalised Jacobian points in, k normalised points out, one
 * shared field inversion): folds the per-GPU partial sums of a commit batch that was split over several GPUs */
int apb_g1_fold(int curve, size_t npieces, const uint64_t* pieces_xyz, const uint32_t* group, size_t k, uint64_t* out_xyz);

/* ark-serialize compressed G1Affine (48 bytes; flags in the top bits of the last byte):
 * what `Commitment` contributes to the transcript (plonk-core/src/transcript.rs:27-33). */
int apb_g1_compress(int curve, const uint64_t xyz[18], uint8_t out[48]);

/* ---- evaluation domain: replaces ark_poly::GeneralEvaluationDomain / Radix2EvaluationDomain
 *      (constructed at prover.rs:169, preprocess.rs:284, quotient_poly.rs:43-47) */
int apb_domain_new(int curve, uint32_t log_n, apb_domain_t* out);
int apb_domain_size(apb_domain_t d, size_t* n);
void apb_domain_free(apb_domain_t d);

/* fft / ifft / coset_fft / coset_ifft (call sites: prover.rs:197-203,241,282,303,305;
 * quotient_poly.rs:72-120,176,205,294,325; permutation/mod.rs:199-205,671-674,751,800;
 * pi.rs:115; lookup/multiset.rs:201; preprocess.rs:145-210,304-340).
 * in: in_len <= N elements (4 u64 each, Montgomery), implicitly zero-extended to N;
 * out: N elements, natural order.  in_len == 0 gives N zeros. */
int apb_ntt(apb_domain_t d, int kind, const uint64_t* in, size_t in_len, uint64_t* out);
/* device pointers; d_in may equal d_out; asynchronous on the domain's stream unless sync != 0 */
int apb_ntt_dev(apb_domain_t d, int kind, const void* d_in, size_t in_len, void* d_out, int sync);

/* `batch` transforms of the same kind in one launch sequence (e.g. the 13 coset FFTs of
 * quotient_poly.rs:72-120); vector b starts at d_in + b*in_stride / d_out + b*out_stride elements */
int apb_ntt_batch_dev(apb_domain_t d, int kind, const void* d_in, size_t in_len, size_t in_stride,
                      void* d_out, size_t out_stride, size_t batch, int sync);

/* ---- device-resident polynomial utilities and prover kernels (SURVEY.md 8f "next" rows) ----
 * All vectors are device pointers to Fr elements (32 B, Montgomery); scalars are 4 x u64 Montgomery
 * in host memory.  Work is enqueued on the library stream; functions that return values to the
 * host block. */
/* out[i] = sum_j scalars[j] * polys[j][i], i < out_len (polys zero-extended): DensePolynomial
 * linear combinations of linearisation_poly.rs:203-349, MultiSet::compress (lookup/multiset.rs:207-213),
 * the opening combination of sonic_pc::open */
int apb_fr_lincomb(int curve, size_t k, const void* const* d_polys, const size_t* lens, const uint64_t* scalars,
                   void* d_out, size_t out_len);
/* compressed lookup query vector f (proof_system/prover.rs:252-278) */
int apb_plonk_lookup_f(int curve, const void* q_lookup, const void* wl, const void* wr, const void* wo, const void* w4,
                       const void* t_comp, const uint64_t* zeta, void* d_out, size_t n);
/* MultiSet::combine_split (lookup/multiset.rs:131-174): h1, h2 of n elements each; blocking;
 * APB_ERR_INVALID_ARG when an element of f is not in t (Error::ElementNotIndexed) */
int apb_plonk_combine_split(int curve, const void* d_t, const void* d_f, size_t n, void* d_h1, void* d_h2);
/* evaluations of the permutation grand product z over the domain (permutation/mod.rs:652-751, before its ifft) */
int apb_plonk_perm_z(apb_domain_t dom, const void* const* d_wires4, const void* const* d_sigmas4, const uint64_t* beta,
                     const uint64_t* gamma, void* d_z);
/* evaluations of the lookup grand product z2 (permutation/mod.rs:754-822) */
int apb_plonk_lookup_z2(apb_domain_t dom, const void* d_f, const void* d_t, const void* d_h1, const void* d_h2,
                        const uint64_t* delta, const uint64_t* epsilon, void* d_z2);
/* quotient evaluations on the 4n coset (quotient_poly.rs:122-173): ptrs25 = wl wr wo w4 z z2 f table h1 h2 pi(NULL ok)
 * q_m q_l q_r q_o q_4 q_c q_arith q_lookup s1 s2 s3 s4 linear l1 ; scalars10 = alpha beta gamma delta epsilon zeta
 * lookup_sep K1 K2 K3 ; vh_inv4 = inverses of the 4-periodic vanishing-polynomial values */
int apb_plonk_quotient(int curve, const void* const* ptrs25, const uint64_t* scalars10, const uint64_t* vh_inv4,
                       void* d_out, size_t n4);
/* same with the custom gate terms of quotient_poly.rs:231-264 (widget/range.rs, widget/logic.rs,
 * widget/ecc/fixed_base_scalar_mul.rs, widget/ecc/curve_addition.rs): ptrs29 = the 25 above + q_range q_logic
 * q_fixed_group_add q_variable_group_add (NULL = identically zero); scalars16 = the 10 above + the range / logic /
 * fixed-base / variable-base separation challenges + the embedded curve's COEFF_A, COEFF_D */
int apb_plonk_quotient_full(int curve, const void* const* ptrs29, const uint64_t* scalars16, const uint64_t* vh_inv4,
                            void* d_out, size_t n4);
/* the same for the points [first, first + count) of the 4n coset only (inputs are the full vectors; d_out is the full
 * output vector, only that slice is written): lets the ranks of a multi-GPU proof evaluate one slice each */
int apb_plonk_quotient_range(int curve, const void* const* ptrs29, const uint64_t* scalars16, const uint64_t* vh_inv4,
                             void* d_out, size_t n4, size_t first, size_t count);
/* k evaluations polys[j](points[j]) -> out_vals (host, Montgomery); DensePolynomial::evaluate */
int apb_poly_eval(int curve, size_t k, const void* const* d_polys, const size_t* lens, const uint64_t* points,
                  uint64_t* out_vals);
/* witness polynomial p / (X - z) (len - 1 coefficients): kzg10 compute_witness_polynomial */
int apb_poly_divide_linear(int curve, const void* d_p, size_t len, const uint64_t* z, void* d_out);

/* ---- Fiat-Shamir transcript (merlin 3.0 / STROBE-128), host side: plonk-core/src/transcript.rs:16-50 */
typedef struct apb_transcript_s* apb_transcript_t;
int apb_transcript_new(const uint8_t* label, size_t label_len, apb_transcript_t* out);
int apb_transcript_append(apb_transcript_t t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len);
int apb_transcript_challenge(apb_transcript_t t, const uint8_t* label, size_t label_len, uint8_t* out, size_t out_len);
void apb_transcript_free(apb_transcript_t t);

/* ---- device memory helpers for host layers that keep polynomials resident -------------- */
int apb_dev_alloc(size_t bytes, void** d_ptr);
int apb_dev_free(void* d_ptr);
int apb_dev_upload(void* d_dst, const void* h_src, size_t bytes);
int apb_dev_download(void* h_dst, const void* d_src, size_t bytes);
int apb_dev_sync(void);
void* apb_stream(void);                     /* the cudaStream_t all work is enqueued on */

/* ---- diagnostics used by the parity tests and bench.py ---------------------------------- */
/* element-wise field op on the device: op 0 mul, 1 add, 2 sub, 3 to_mont, 4 from_mont, 5 sqr,
 * 6 inverse (0 -> 0).
 * field: 0 Fr381, 1 Fq381, 2 Fr377, 3 Fq377.  a, b, out: count x (4 or 6) u64. */
int apb_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
uint64_t apb_kernel_launches(void);
/* timed microbenchmark of dependent IMAD.WIDE chains: returns multiply-adds per second */
int apb_imad_peak(double* wide_imad_per_s, double* imad32_per_s);
/* throughput of dependent Montgomery products (ilp independent chains per thread) at a given occupancy */
int apb_mul_bench(int field, int threads, int blocks_per_sm, int ilp, uint32_t iters, double* muls_per_s);
/* per-phase CUDA-event timing of the last MSM: [sort, accumulate, stitch, reduce trees] in ms
 * (recorded only after apb_set_profiling(1)) */
void apb_set_profiling(int on);
void apb_msm_phase_ms(double out[4]);
/* plan of the last MSM pass: {digit bits, batched-affine pair levels, bucket-range slices, 1 if it fell
 * back to the plain accumulate because the slices were unbalanced} */
void apb_msm_last_plan(uint32_t out[4]);
/* totals since the last reset while profiling: k_msm_accumulate milliseconds and scalars processed */
void apb_msm_totals(double* accumulate_ms, unsigned long long* points, int reset);
/* device time of whole MSM calls (sort + accumulation + bucket reduction + copy-out + host epilogue) */
void apb_msm_call_ms(double* whole_calls_ms, int reset);
/* wide multiply-adds of the bucket accumulation stage since the last reset while profiling: by the
 * XYZZ cost model of SURVEY 8(d) (entries x 10 Fq products x 300) and as issued (6 products per
 * batched-affine pair addition, 10 per XYZZ mixed addition) */
void apb_msm_work(double* model_madds, double* issued_madds, int reset);
/* totals since the last reset while profiling: device milliseconds and number of transforms */
void apb_ntt_totals(double* ms, unsigned long long* transforms, int reset);
/* milliseconds of device time of the last blocking apb_msm / apb_ntt call (CUDA events) */
double apb_last_device_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* APB_H */
