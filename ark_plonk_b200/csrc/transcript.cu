// Merlin 3.0 transcripts (STROBE-128 over Keccak-f[1600]) - host code.
//
// plonk-core drives merlin through plonk-core/src/transcript.rs:16-50: `append` is
// append_message(label, compressed bytes) and `challenge_scalar` reads size_in_bits/8 = 31
// challenge bytes.  The prover's Fiat-Shamir chain is strictly sequential host work between the
// device phases; it lives in the library so that the host layer does not pay an interpreter per
// permutation.  Written from the published STROBE/Merlin specifications.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace {

const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
const int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
const int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

inline uint64_t rol(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

void keccak_f1600(uint8_t* bytes) {
    uint64_t st[25];
    memcpy(st, bytes, 200);          // little-endian host
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rol(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = PILN[i];
            uint64_t b = st[j];
            st[j] = rol(t, ROTC[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= RC[round];
    }
    memcpy(bytes, st, 200);
}

const int STROBE_R = 166;
enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

}  // namespace

struct apb_transcript_s {
    uint32_t magic;
    uint8_t state[200];
    uint8_t pos, pos_begin, cur_flags;

    void run_f() {
        state[pos] ^= pos_begin;
        state[pos + 1] ^= 0x04;
        state[STROBE_R + 1] ^= 0x80;
        keccak_f1600(state);
        pos = 0;
        pos_begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            state[pos++] ^= d[i];
            if (pos == STROBE_R) run_f();
        }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) {
            d[i] = state[pos];
            state[pos++] = 0;
            if (pos == STROBE_R) run_f();
        }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        uint8_t old_begin = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
    }
    void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_M | FLAG_A, more); absorb(d, n); }
    void ad(const uint8_t* d, size_t n, bool more) { begin_op(FLAG_A, more); absorb(d, n); }
    void prf(uint8_t* d, size_t n) { begin_op(FLAG_I | FLAG_A | FLAG_C, false); squeeze(d, n); }
    void append_message(const uint8_t* label, size_t ll, const uint8_t* msg, size_t ml) {
        uint8_t len[4] = {(uint8_t)ml, (uint8_t)(ml >> 8), (uint8_t)(ml >> 16), (uint8_t)(ml >> 24)};
        meta_ad(label, ll, false);
        meta_ad(len, 4, true);
        ad(msg, ml, false);
    }
};
static const uint32_t TR_MAGIC = 0x54524e31;

extern "C" int apb_transcript_new(const uint8_t* label, size_t label_len, apb_transcript_t* out) {
    if (!out || (!label && label_len)) return apb::set_err(APB_ERR_INVALID_ARG, "apb_transcript_new: null argument");
    apb_transcript_s* t = new apb_transcript_s();
    memset(t, 0, sizeof(*t));
    t->magic = TR_MAGIC;
    const uint8_t init[6] = {1, (uint8_t)(STROBE_R + 2), 1, 0, 1, 96};
    memcpy(t->state, init, 6);
    memcpy(t->state + 6, "STROBEv1.0.2", 12);
    keccak_f1600(t->state);
    t->meta_ad((const uint8_t*)"Merlin v1.0", 11, false);
    t->append_message((const uint8_t*)"dom-sep", 7, label, label_len);
    *out = t;
    return APB_OK;
}
extern "C" int apb_transcript_append(apb_transcript_t t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t msg_len) {
    if (!t || t->magic != TR_MAGIC) return apb::set_err(APB_ERR_BAD_HANDLE, "apb_transcript_append: bad handle");
    t->append_message(label, label_len, msg, msg_len);
    return APB_OK;
}
extern "C" int apb_transcript_challenge(apb_transcript_t t, const uint8_t* label, size_t label_len, uint8_t* out, size_t out_len) {
    if (!t || t->magic != TR_MAGIC || !out) return apb::set_err(APB_ERR_BAD_HANDLE, "apb_transcript_challenge: bad handle");
    uint8_t len[4] = {(uint8_t)out_len, (uint8_t)(out_len >> 8), (uint8_t)(out_len >> 16), (uint8_t)(out_len >> 24)};
    t->meta_ad(label, label_len, false);
    t->meta_ad(len, 4, true);
    t->prf(out, out_len);
    return APB_OK;
}
extern "C" void apb_transcript_free(apb_transcript_t t) {
    if (!t || t->magic != TR_MAGIC) return;
    t->magic = 0;
    delete t;
}
