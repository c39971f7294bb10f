/*
 * apb_oracle.c - CPU restatement (C, pthreads) of the arkworks 0.3 algorithms on the hot path
 * of heliaxdev/ark-plonk.  TEST INFRASTRUCTURE / CPU BASELINE ONLY: used by tests/ as a fast
 * checker, by bench.py's cpu_baseline leg and by `bench.py --impl reference`.  The product
 * (ark_plonk_b200) never links or loads it.
 *
 * PARITY UNPINNED BY REFERENCE BYTES (see oracle/__init__.py): arkworks is not vendored in the
 * reference and there is no Rust toolchain here, so this file restates the published
 * algorithms and is pinned against oracle/*.py (big-int definitions) in
 * tests/test_oracle_c.py.
 *
 *  - oracle_msm: ark_ec::msm::VariableBaseMSM::multi_scalar_mul as called at
 *    plonk-core/src/commitment.rs:45: c = 3 if n < 32 else ceil_log2(n)*69/100+2,
 *    unsigned c-bit windows over MODULUS_BITS, one task per window (rayon -> pthreads),
 *    2^c - 1 Jacobian buckets with mixed additions, running-sum, high->low fold with c
 *    doublings.
 *  - oracle_ntt: ark_poly Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}
 *    (call sites plonk-core/src/proof_system/prover.rs:197-203, quotient_poly.rs:72-120,176):
 *    in-place radix-2, roots table of n/2 built per call, layers split across threads.
 *  - field arithmetic: Montgomery, 64-bit limbs, R = 2^256 / 2^384 (ark-ff Fp256 / Fp384).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

typedef struct {
    int L;
    u64 mod[6], one[6], r2[6], n0inv;
    int bits;
} field_t;

/* moduli, little-endian u64 limbs */
static const u64 FR381_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const u64 FQ381_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 FR377_MOD[4] = {0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL, 0x12ab655e9a2ca556ULL};
static const u64 FQ377_MOD[6] = {0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL, 0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL};

static int geq(const u64* a, const u64* b, int L) {
    for (int i = L - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static void sub_n(u64* r, const u64* a, const u64* b, int L) {
    u64 borrow = 0;
    for (int i = 0; i < L; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
}
static void f_add(const field_t* f, u64* r, const u64* a, const u64* b) {
    u64 t[6], carry = 0;
    for (int i = 0; i < f->L; i++) {
        u128 s = (u128)a[i] + b[i] + carry;
        t[i] = (u64)s;
        carry = (u64)(s >> 64);
    }
    if (carry || geq(t, f->mod, f->L)) sub_n(t, t, f->mod, f->L);
    memcpy(r, t, 8 * f->L);
}
static void f_sub(const field_t* f, u64* r, const u64* a, const u64* b) {
    u64 t[6], borrow = 0;
    for (int i = 0; i < f->L; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        t[i] = (u64)d;
        borrow = (u64)(d >> 64) & 1;
    }
    if (borrow) {
        u64 carry = 0;
        for (int i = 0; i < f->L; i++) {
            u128 s = (u128)t[i] + f->mod[i] + carry;
            t[i] = (u64)s;
            carry = (u64)(s >> 64);
        }
    }
    memcpy(r, t, 8 * f->L);
}
#define DEF_MUL(NAME, LL)                                                     \
    static void NAME(const field_t* f, u64* r, const u64* a, const u64* b) {  \
        u64 t[LL + 2];                                                        \
        for (int i = 0; i < LL + 2; i++) t[i] = 0;                            \
        for (int i = 0; i < LL; i++) {                                        \
            u64 carry = 0;                                                    \
            for (int j = 0; j < LL; j++) {                                    \
                u128 s = (u128)a[j] * b[i] + t[j] + carry;                    \
                t[j] = (u64)s;                                                \
                carry = (u64)(s >> 64);                                       \
            }                                                                 \
            u128 s = (u128)t[LL] + carry;                                     \
            t[LL] = (u64)s;                                                   \
            t[LL + 1] = (u64)(s >> 64);                                       \
            u64 m = t[0] * f->n0inv;                                          \
            s = (u128)m * f->mod[0] + t[0];                                   \
            carry = (u64)(s >> 64);                                           \
            for (int j = 1; j < LL; j++) {                                    \
                s = (u128)m * f->mod[j] + t[j] + carry;                       \
                t[j - 1] = (u64)s;                                            \
                carry = (u64)(s >> 64);                                       \
            }                                                                 \
            s = (u128)t[LL] + carry;                                          \
            t[LL - 1] = (u64)s;                                               \
            t[LL] = t[LL + 1] + (u64)(s >> 64);                               \
        }                                                                     \
        if (t[LL] || geq(t, f->mod, LL)) sub_n(t, t, f->mod, LL);             \
        memcpy(r, t, 8 * LL);                                                 \
    }
DEF_MUL(f_mul4, 4)
DEF_MUL(f_mul6, 6)
static inline void f_mul(const field_t* f, u64* r, const u64* a, const u64* b) {
    if (f->L == 4) f_mul4(f, r, a, b); else f_mul6(f, r, a, b);
}
static int f_is_zero(const field_t* f, const u64* a) {
    u64 o = 0;
    for (int i = 0; i < f->L; i++) o |= a[i];
    return o == 0;
}
static void f_pow(const field_t* f, u64* r, const u64* a, const u64* e, int elimbs) {
    u64 acc[6], base[6];
    memcpy(acc, f->one, 48);
    memcpy(base, a, 8 * f->L);
    for (int i = 0; i < 64 * elimbs; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(f, acc, acc, base);
        f_mul(f, base, base, base);
    }
    memcpy(r, acc, 8 * f->L);
}
static void f_inv(const field_t* f, u64* r, const u64* a) {
    u64 e[6];
    memcpy(e, f->mod, 8 * f->L);
    e[0] -= 2;
    f_pow(f, r, a, e, f->L);
}
static void field_init(field_t* f, const u64* mod, int L) {
    memset(f, 0, sizeof(*f));
    f->L = L;
    memcpy(f->mod, mod, 8 * L);
    u64 inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - mod[0] * inv;
    f->n0inv = (u64)0 - inv;
    /* R mod p by repeated doubling of 1; R^2 by 64*L more doublings */
    u64 x[6] = {1, 0, 0, 0, 0, 0};
    f->bits = 0;
    for (int i = 64 * L - 1; i >= 0; i--)
        if ((mod[i / 64] >> (i % 64)) & 1) { f->bits = i + 1; break; }
    for (int i = 0; i < 128 * L; i++) {
        f_add(f, x, x, x);
        if (i == 64 * L - 1) memcpy(f->one, x, 8 * L);
    }
    memcpy(f->r2, x, 8 * L);
}

static field_t FR[2], FQ[2];
static u64 FR_ROOT[2][4], FR_GEN[2][4];       /* 2-adic root / multiplicative generator (Montgomery) */
static const int FR_ADICITY[2] = {32, 47};
static const u64 FR_GENERATOR[2] = {7, 22};
static int g_init = 0;

static void from_u64(const field_t* f, u64* r, u64 v) {
    u64 t[6] = {v, 0, 0, 0, 0, 0};
    f_mul(f, r, t, f->r2);
}
void oracle_init(void) {
    if (g_init) return;
    field_init(&FR[0], FR381_MOD, 4);
    field_init(&FQ[0], FQ381_MOD, 6);
    field_init(&FR[1], FR377_MOD, 4);
    field_init(&FQ[1], FQ377_MOD, 6);
    for (int c = 0; c < 2; c++) {
        from_u64(&FR[c], FR_GEN[c], FR_GENERATOR[c]);
        u64 e[4];
        memcpy(e, FR[c].mod, 32);
        e[0] -= 1;                                  /* p - 1 */
        /* e >>= adicity */
        int sh = FR_ADICITY[c];
        u64 t[4] = {0, 0, 0, 0};
        for (int i = 0; i < 256 - sh; i++) {
            int src = i + sh;
            if ((e[src / 64] >> (src % 64)) & 1) t[i / 64] |= (u64)1 << (i % 64);
        }
        f_pow(&FR[c], FR_ROOT[c], FR_GEN[c], t, 4);
    }
    g_init = 1;
}

/* ---------------- Jacobian G1, a = 0 (ark-ec short_weierstrass_jacobian) ---------------- */
typedef struct { u64 x[6], y[6], z[6]; } jac_t;

static void j_set_inf(jac_t* p) { memset(p, 0, sizeof(*p)); }
static int j_is_inf(const field_t* f, const jac_t* p) { return f_is_zero(f, p->z); }

static void j_double(const field_t* f, jac_t* r, const jac_t* p) {
    if (j_is_inf(f, p)) { *r = *p; return; }
    u64 A[6], B[6], C[6], D[6], E[6], F[6], t[6];
    jac_t o;
    f_mul(f, A, p->x, p->x);
    f_mul(f, B, p->y, p->y);
    f_mul(f, C, B, B);
    f_add(f, t, p->x, B);
    f_mul(f, t, t, t);
    f_sub(f, t, t, A);
    f_sub(f, t, t, C);
    f_add(f, D, t, t);
    f_add(f, E, A, A);
    f_add(f, E, E, A);
    f_mul(f, F, E, E);
    f_mul(f, o.z, p->y, p->z);
    f_add(f, o.z, o.z, o.z);
    f_sub(f, o.x, F, D);
    f_sub(f, o.x, o.x, D);
    f_sub(f, t, D, o.x);
    f_mul(f, o.y, E, t);
    f_add(f, C, C, C); f_add(f, C, C, C); f_add(f, C, C, C);
    f_sub(f, o.y, o.y, C);
    *r = o;
}
/* r = p + (qx, qy) affine, add_assign_mixed */
static void j_add_mixed(const field_t* f, jac_t* r, const jac_t* p, const u64* qx, const u64* qy) {
    if (f_is_zero(f, qx) && f_is_zero(f, qy)) { *r = *p; return; }      /* infinity */
    if (j_is_inf(f, p)) {
        memcpy(r->x, qx, 48); memcpy(r->y, qy, 48); memcpy(r->z, f->one, 48);
        return;
    }
    u64 Z1Z1[6], U2[6], S2[6], H[6], HH[6], I[6], J[6], rr[6], V[6], t[6];
    jac_t o;
    f_mul(f, Z1Z1, p->z, p->z);
    f_mul(f, U2, qx, Z1Z1);
    f_mul(f, S2, qy, p->z);
    f_mul(f, S2, S2, Z1Z1);
    if (memcmp(U2, p->x, 48) == 0) {
        if (memcmp(S2, p->y, 48) == 0) { j_double(f, r, p); return; }
        j_set_inf(r);
        return;
    }
    f_sub(f, H, U2, p->x);
    f_mul(f, HH, H, H);
    f_add(f, I, HH, HH); f_add(f, I, I, I);
    f_mul(f, J, H, I);
    f_sub(f, rr, S2, p->y); f_add(f, rr, rr, rr);
    f_mul(f, V, p->x, I);
    f_mul(f, o.x, rr, rr);
    f_sub(f, o.x, o.x, J);
    f_sub(f, o.x, o.x, V); f_sub(f, o.x, o.x, V);
    f_sub(f, t, V, o.x);
    f_mul(f, o.y, rr, t);
    f_mul(f, t, p->y, J); f_add(f, t, t, t);
    f_sub(f, o.y, o.y, t);
    f_add(f, o.z, p->z, H);
    f_mul(f, o.z, o.z, o.z);
    f_sub(f, o.z, o.z, Z1Z1);
    f_sub(f, o.z, o.z, HH);
    *r = o;
}
static void j_add(const field_t* f, jac_t* r, const jac_t* p, const jac_t* q) {
    if (j_is_inf(f, p)) { *r = *q; return; }
    if (j_is_inf(f, q)) { *r = *p; return; }
    u64 Z1Z1[6], Z2Z2[6], U1[6], U2[6], S1[6], S2[6], H[6], I[6], J[6], rr[6], V[6], t[6];
    jac_t o;
    f_mul(f, Z1Z1, p->z, p->z);
    f_mul(f, Z2Z2, q->z, q->z);
    f_mul(f, U1, p->x, Z2Z2);
    f_mul(f, U2, q->x, Z1Z1);
    f_mul(f, S1, p->y, q->z); f_mul(f, S1, S1, Z2Z2);
    f_mul(f, S2, q->y, p->z); f_mul(f, S2, S2, Z1Z1);
    if (memcmp(U1, U2, 48) == 0) {
        if (memcmp(S1, S2, 48) == 0) { j_double(f, r, p); return; }
        j_set_inf(r);
        return;
    }
    f_sub(f, H, U2, U1);
    f_add(f, I, H, H); f_mul(f, I, I, I);
    f_mul(f, J, H, I);
    f_sub(f, rr, S2, S1); f_add(f, rr, rr, rr);
    f_mul(f, V, U1, I);
    f_mul(f, o.x, rr, rr);
    f_sub(f, o.x, o.x, J);
    f_sub(f, o.x, o.x, V); f_sub(f, o.x, o.x, V);
    f_sub(f, t, V, o.x);
    f_mul(f, o.y, rr, t);
    f_mul(f, t, S1, J); f_add(f, t, t, t);
    f_sub(f, o.y, o.y, t);
    f_add(f, o.z, p->z, q->z);
    f_mul(f, o.z, o.z, o.z);
    f_sub(f, o.z, o.z, Z1Z1);
    f_sub(f, o.z, o.z, Z2Z2);
    f_mul(f, o.z, o.z, H);
    *r = o;
}

/* ---------------- VariableBaseMSM ---------------- */
typedef struct {
    const field_t* fq;
    const u64* bases;     /* n x 12 */
    const u64* scalars;   /* n x 4 canonical */
    size_t n;
    int c, nwin;
    jac_t* window_sums;
    int next;             /* work queue over windows */
    pthread_mutex_t mu;
} msm_job_t;

static void msm_window(msm_job_t* J, int w) {
    const field_t* f = J->fq;
    const int c = J->c, w_start = w * c;
    const size_t nb = ((size_t)1 << c) - 1;
    jac_t* buckets = (jac_t*)calloc(nb, sizeof(jac_t));
    jac_t res;
    j_set_inf(&res);
    for (size_t i = 0; i < J->n; i++) {
        const u64* s = J->scalars + 4 * i;
        if ((s[0] | s[1] | s[2] | s[3]) == 0) continue;
        const u64* bx = J->bases + 12 * i;
        if (s[0] == 1 && (s[1] | s[2] | s[3]) == 0) {
            if (w_start == 0) j_add_mixed(f, &res, &res, bx, bx + 6);
            continue;
        }
        int limb = w_start / 64, sh = w_start % 64;
        u64 d = s[limb] >> sh;
        if (sh && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
        d &= ((u64)1 << c) - 1;
        if (d) j_add_mixed(f, &buckets[d - 1], &buckets[d - 1], bx, bx + 6);
    }
    jac_t running;
    j_set_inf(&running);
    for (size_t b = nb; b-- > 0;) {
        j_add(f, &running, &running, &buckets[b]);
        j_add(f, &res, &res, &running);
    }
    J->window_sums[w] = res;
    free(buckets);
}
static void* msm_worker(void* arg) {
    msm_job_t* J = (msm_job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int w = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (w >= J->nwin) break;
        msm_window(J, w);
    }
    return NULL;
}
static int ceil_log2(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}
/* out: affine x, y (Montgomery, 6+6 u64); returns 1 if the result is the identity */
int oracle_msm(int curve, const u64* bases, const u64* scalars, size_t n, int threads, u64* out_xy) {
    oracle_init();
    const field_t* f = &FQ[curve];
    memset(out_xy, 0, 96);
    if (n == 0) return 1;
    msm_job_t J;
    memset(&J, 0, sizeof(J));
    J.fq = f; J.bases = bases; J.scalars = scalars; J.n = n;
    J.c = n < 32 ? 3 : ceil_log2(n) * 69 / 100 + 2;
    J.nwin = (FR[curve].bits + J.c - 1) / J.c;
    J.window_sums = (jac_t*)calloc(J.nwin, sizeof(jac_t));
    pthread_mutex_init(&J.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads > J.nwin) threads = J.nwin;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, msm_worker, &J);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    jac_t total;
    j_set_inf(&total);
    for (int w = J.nwin - 1; w >= 1; w--) {
        j_add(f, &total, &total, &J.window_sums[w]);
        for (int k = 0; k < J.c; k++) j_double(f, &total, &total);
    }
    j_add(f, &total, &total, &J.window_sums[0]);
    free(J.window_sums);
    if (j_is_inf(f, &total)) return 1;
    u64 zi[6], zi2[6], zi3[6];
    f_inv(f, zi, total.z);
    f_mul(f, zi2, zi, zi);
    f_mul(f, zi3, zi2, zi);
    f_mul(f, out_xy, total.x, zi2);
    f_mul(f, out_xy + 6, total.y, zi3);
    return 0;
}

/* ---------------- Radix2EvaluationDomain ---------------- */
typedef struct {
    const field_t* f;
    u64* a;
    const u64* roots;
    size_t n, half, lo, hi, step;
} layer_job_t;

static void* layer_worker(void* arg) {
    layer_job_t* J = (layer_job_t*)arg;
    const field_t* f = J->f;
    /* butterflies b in [lo, hi): block = b / half, k = b % half */
    for (size_t b = J->lo; b < J->hi; b++) {
        size_t blk = b / J->half, k = b % J->half;
        u64* u = J->a + 4 * (blk * 2 * J->half + k);
        u64* v = u + 4 * J->half;
        u64 t[4];
        f_mul4(f, t, v, J->roots + 4 * (k * J->step));
        f_sub(f, v, u, t);
        f_add(f, u, u, t);
    }
    return NULL;
}
/* kind: 0 fft, 1 ifft, 2 coset_fft, 3 coset_ifft.  data: n x 4 u64 Montgomery, first in_len valid */
int oracle_ntt(int curve, int kind, u64* data, size_t log_n, size_t in_len, int threads) {
    oracle_init();
    const field_t* f = &FR[curve];
    const size_t n = (size_t)1 << log_n;
    if ((int)log_n > FR_ADICITY[curve] || in_len > n) return -1;
    memset(data + 4 * in_len, 0, 32 * (n - in_len));
    const int inverse = kind == 1 || kind == 3;
    u64 w[4], g[4];
    memcpy(w, FR_ROOT[curve], 32);
    for (size_t i = log_n; i < (size_t)FR_ADICITY[curve]; i++) f_mul4(f, w, w, w);
    if (inverse) f_inv(f, w, w);
    if (kind == 2) {                            /* coset_fft: coeffs[i] *= g^i */
        u64 cur[4];
        memcpy(cur, f->one, 32);
        for (size_t i = 0; i < in_len; i++) {
            f_mul4(f, data + 4 * i, data + 4 * i, cur);
            f_mul4(f, cur, cur, FR_GEN[curve]);
        }
    }
    /* roots table of n/2 (arkworks recomputes it per call) */
    size_t half_n = n > 1 ? n / 2 : 1;
    u64* roots = (u64*)malloc(32 * half_n);
    memcpy(roots, f->one, 32);
    for (size_t i = 1; i < half_n; i++) f_mul4(f, roots + 4 * i, roots + 4 * (i - 1), w);
    /* bit reversal then DIT layers */
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (size_t b = 0; b < log_n; b++) if (i >> b & 1) j |= (size_t)1 << (log_n - 1 - b);
        if (i < j) {
            u64 t[4];
            memcpy(t, data + 4 * i, 32);
            memcpy(data + 4 * i, data + 4 * j, 32);
            memcpy(data + 4 * j, t, 32);
        }
    }
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    layer_job_t* jobs = (layer_job_t*)malloc(sizeof(layer_job_t) * threads);
    for (size_t half = 1; half < n; half <<= 1) {
        size_t total = n / 2;
        int nt = total < 4096 ? 1 : threads;
        for (int t = 0; t < nt; t++) {
            jobs[t].f = f; jobs[t].a = data; jobs[t].roots = roots; jobs[t].n = n; jobs[t].half = half;
            jobs[t].step = n / (2 * half);
            jobs[t].lo = total * t / nt; jobs[t].hi = total * (t + 1) / nt;
            if (nt > 1) pthread_create(&th[t], NULL, layer_worker, &jobs[t]);
        }
        if (nt == 1) layer_worker(&jobs[0]);
        else for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
    free(th); free(jobs); free(roots);
    if (inverse) {
        u64 ninv[4], nval[4], cur[4], gi[4];
        from_u64(f, nval, 1);
        for (size_t i = 0; i < log_n; i++) f_add(f, nval, nval, nval);
        f_inv(f, ninv, nval);
        if (kind == 3) {                        /* coset_ifft: res[i] *= g^-i / n */
            f_inv(f, gi, FR_GEN[curve]);
            memcpy(cur, ninv, 32);
            for (size_t i = 0; i < n; i++) {
                f_mul4(f, data + 4 * i, data + 4 * i, cur);
                f_mul4(f, cur, cur, gi);
            }
        } else {
            for (size_t i = 0; i < n; i++) f_mul4(f, data + 4 * i, data + 4 * i, ninv);
        }
    }
    (void)g;
    return 0;
}

/* Fr helpers for tests: to/from Montgomery of n elements */
void oracle_fr_to_mont(int curve, u64* data, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) f_mul4(&FR[curve], data + 4 * i, data + 4 * i, FR[curve].r2);
}
void oracle_fr_from_mont(int curve, u64* data, size_t n) {
    oracle_init();
    u64 o[4] = {1, 0, 0, 0};
    for (size_t i = 0; i < n; i++) f_mul4(&FR[curve], data + 4 * i, data + 4 * i, o);
}

/* ---- polynomial evaluation (test checker): sum_i coeffs[i] * x^i over Fr, all values Montgomery ----
 * DensePolynomial::evaluate of ark-poly (Horner); used by the tests as the size-independent check of
 * large transforms (fft(p)[i] == p(w^i)) and of large MSMs over powers of tau (MSM == [p(tau)]G).
 * Threads evaluate contiguous chunks by Horner and the chunk values are combined with x^(chunk start). */
typedef struct {
    const field_t* f;
    const u64* c;
    size_t lo, hi;
    const u64* x;
    u64 val[4];
} horner_job_t;
static void* horner_worker(void* arg) {
    horner_job_t* J = (horner_job_t*)arg;
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t i = J->hi; i-- > J->lo;) {
        f_mul4(J->f, acc, acc, J->x);
        f_add(J->f, acc, acc, J->c + 4 * i);
    }
    memcpy(J->val, acc, 32);
    return NULL;
}
/* coeffs: n x 4 u64 (Montgomery), x: 4 u64 (Montgomery) -> out: 4 u64 (Montgomery) */
int oracle_fr_horner(int curve, const u64* coeffs, size_t n, const u64* x, int threads, u64* out) {
    oracle_init();
    const field_t* f = &FR[curve];
    if (threads < 1) threads = 1;
    if (n < 4096) threads = 1;
    horner_job_t* jobs = (horner_job_t*)malloc(sizeof(horner_job_t) * threads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].f = f; jobs[t].c = coeffs; jobs[t].x = x;
        jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
        if (threads > 1) pthread_create(&th[t], NULL, horner_worker, &jobs[t]);
    }
    if (threads == 1) horner_worker(&jobs[0]);
    else for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    u64 acc[4] = {0, 0, 0, 0};
    for (int t = threads; t-- > 0;) {           /* acc = acc * x^(len of chunk t) + val_t */
        u64 e[4] = {jobs[t].hi - jobs[t].lo, 0, 0, 0}, xp[4];
        f_pow(f, xp, x, e, 1);
        f_mul4(f, acc, acc, xp);
        f_add(f, acc, acc, jobs[t].val);
    }
    memcpy(out, acc, 32);
    free(jobs); free(th);
    return 0;
}
