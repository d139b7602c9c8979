/* oracle/poseidon2.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Poseidon2 over KoalaBear, width 16, S-box x^3, 8 external + 13 internal rounds, with the
 * reference's constant carve-out, plus the sponge and the 2-to-1 compression built on it.
 *
 * Follows:
 *   reference crates/stark/src/kb31_poseidon2.rs:35-50   (my_perm: which rows of RC_16_30 go where)
 *   reference crates/primitives/src/lib.rs:13-554          (RC_16_30 data, see rc_16_30.h)
 *   reference crates/stark/src/kb31_poseidon2.rs:23-26     (MyHash = PaddingFreeSponge<Perm,16,8,8>,
 *                                                           MyCompress = TruncatedPermutation<Perm,2,8,16>)
 *   Plonky3 (un-vendored, rev 93967fce, v0.1.0): p3-poseidon2 `Poseidon2::permute_mut`,
 *   `mds_light_permutation` with MDSMat4, p3-koala-bear internal diagonal for width 16,
 *   p3-symmetric `PaddingFreeSponge::hash_iter` / `TruncatedPermutation::compress`
 *   — restated from the published algorithm.
 *
 * PARITY UNPINNED (no reference golden vectors; see bf_oracle.h).
 */
#include "bf_oracle.h"
#include "kb31.h"
#include "rc_16_30.h"
#include <string.h>

#define ROUNDS_F 8
#define ROUNDS_P 13

/* kb31_poseidon2.rs:39-48: internal constants = column 0 of rows 4..16 (drained out of the
   table); external-initial = rows 0..3; external-terminal = rows 4..7 of what is LEFT after the
   drain, i.e. original rows 17..20.  Rows 21..29 are unused. */
void bfo_poseidon2_constants(uint32_t ext_initial[4][16], uint32_t internal[13], uint32_t ext_terminal[4][16]) {
    for (int r = 0; r < ROUNDS_F / 2; r++)
        for (int i = 0; i < 16; i++) {
            ext_initial[r][i] = BF_RC_16_30[r][i];
            ext_terminal[r][i] = BF_RC_16_30[ROUNDS_F / 2 + ROUNDS_P + r][i];
        }
    for (int r = 0; r < ROUNDS_P; r++) internal[r] = BF_RC_16_30[ROUNDS_F / 2 + r][0];
}

static inline uint32_t sbox(uint32_t x) { return kb_mul(kb_mul(x, x), x); }

/* M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] applied to one 4-chunk */
static inline void mat4(uint32_t* x) {
    uint32_t a = x[0], b = x[1], c = x[2], d = x[3];
    uint32_t s = kb_add(kb_add(a, b), kb_add(c, d));
    /* row i = s + x_i + 2 x_{i+1} */
    x[0] = kb_add(kb_add(s, a), kb_dbl(b));
    x[1] = kb_add(kb_add(s, b), kb_dbl(c));
    x[2] = kb_add(kb_add(s, c), kb_dbl(d));
    x[3] = kb_add(kb_add(s, d), kb_dbl(a));
}

/* "MDS light" external linear layer for width 16: M4 on each chunk, then add the
   column-wise sum of the four chunks to every chunk. */
static inline void external_linear(uint32_t s[16]) {
    for (int k = 0; k < 4; k++) mat4(s + 4 * k);
    for (int i = 0; i < 4; i++) {
        uint32_t t = kb_add(kb_add(s[i], s[4 + i]), kb_add(s[8 + i], s[12 + i]));
        for (int k = 0; k < 4; k++) s[4 * k + i] = kb_add(s[4 * k + i], t);
    }
}

/* internal linear layer 1 + Diag(V),
   V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/8, 1/2^24, -1/2^8, -1/8, -1/16, -1/2^24] */
/* 2^-k mod p for the diagonal entries, filled on first use */
static uint32_t INV2[25];
static void init_inv2(void) {
    if (INV2[0]) return;
    uint32_t v = 1;
    for (int k = 1; k <= 24; k++) { v = kb_halve(v); INV2[k] = v; }
    __atomic_store_n(&INV2[0], 1u, __ATOMIC_RELEASE);
}

static inline void internal_linear(uint32_t s[16]) {
    uint32_t sum = 0;
    for (int i = 0; i < 16; i++) sum = kb_add(sum, s[i]);
    uint32_t t[16];
    t[0] = kb_neg(kb_dbl(s[0]));
    t[1] = s[1];
    t[2] = kb_dbl(s[2]);
    t[3] = kb_halve(s[3]);
    t[4] = kb_add(kb_dbl(s[4]), s[4]);
    t[5] = kb_dbl(kb_dbl(s[5]));
    t[6] = kb_neg(kb_halve(s[6]));
    t[7] = kb_neg(kb_add(kb_dbl(s[7]), s[7]));
    t[8] = kb_neg(kb_dbl(kb_dbl(s[8])));
    t[9] = kb_mul(s[9], INV2[8]);
    t[10] = kb_mul(s[10], INV2[3]);
    t[11] = kb_mul(s[11], INV2[24]);
    t[12] = kb_neg(kb_mul(s[12], INV2[8]));
    t[13] = kb_neg(kb_mul(s[13], INV2[3]));
    t[14] = kb_neg(kb_mul(s[14], INV2[4]));
    t[15] = kb_neg(kb_mul(s[15], INV2[24]));
    for (int i = 0; i < 16; i++) s[i] = kb_add(t[i], sum);
}

void bfo_poseidon2_permute(uint32_t s[16]) {
    init_inv2();
    external_linear(s);
    for (int r = 0; r < ROUNDS_F / 2; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(kb_add(s[i], BF_RC_16_30[r][i]));
        external_linear(s);
    }
    for (int r = 0; r < ROUNDS_P; r++) {
        s[0] = sbox(kb_add(s[0], BF_RC_16_30[ROUNDS_F / 2 + r][0]));
        internal_linear(s);
    }
    for (int r = 0; r < ROUNDS_F / 2; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(kb_add(s[i], BF_RC_16_30[ROUNDS_F / 2 + ROUNDS_P + r][i]));
        external_linear(s);
    }
}

void bfo_poseidon2_permute_many(uint32_t* states, uint64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) bfo_poseidon2_permute(states + 16 * i);
}

/* PaddingFreeSponge<_,16,8,8>::hash_iter: overwrite-mode absorption, rate 8; a trailing partial
   block overwrites only its prefix; an empty tail block does not trigger a permutation. */
void bfo_sponge_hash(const uint32_t* in, uint64_t n, uint32_t out[8]) {
    uint32_t st[16];
    memset(st, 0, sizeof st);
    uint64_t pos = 0;
    while (pos < n) {
        uint64_t take = n - pos < 8 ? n - pos : 8;
        for (uint64_t i = 0; i < take; i++) st[i] = in[pos + i];
        bfo_poseidon2_permute(st);
        pos += take;
    }
    memcpy(out, st, 8 * sizeof(uint32_t));
}

void bfo_compress(const uint32_t left[8], const uint32_t right[8], uint32_t out[8]) {
    uint32_t st[16];
    memcpy(st, left, 32);
    memcpy(st + 8, right, 32);
    bfo_poseidon2_permute(st);
    memcpy(out, st, 32);
}
