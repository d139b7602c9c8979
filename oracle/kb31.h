/* oracle/kb31.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * KoalaBear prime field p = 2^31 - 2^24 + 1 and its degree-4 binomial extension
 * F_p[X]/(X^4 - 3), in CANONICAL representation (plain residues in [0,p)).
 *
 * Restates the arithmetic the reference obtains from the un-vendored Plonky3
 * dependency (p3-koala-bear / p3-monty-31 / p3-field v0.1.0, git rev
 * 93967fce8949d2275c06fd91e9f495a35418d68d, Cargo.lock:2518-2824) behind the
 * aliases `Val = KoalaBear`, `Challenge = BinomialExtensionField<Val, 4>`
 * (reference crates/stark/src/kb31_poseidon2.rs:20-21).
 *
 * PARITY UNPINNED: the reference ships no known-answer vectors for this layer
 * (SURVEY.md §4/§8c); constants were re-derived numerically (tests/test_oracle_field.py).
 */
#ifndef BF_ORACLE_KB31_H
#define BF_ORACLE_KB31_H
#include <stdint.h>
#include <stddef.h>

#define KB_P 2130706433u /* 0x7f000001 */
#define KB_GENERATOR 3u  /* multiplicative generator, also the PCS coset shift */
#define KB_TWO_ADICITY 24

static inline uint32_t kb_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b; /* < 2^32 since a,b < 2^31 */
    return s >= KB_P ? s - KB_P : s;
}
static inline uint32_t kb_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + KB_P - b; }
static inline uint32_t kb_neg(uint32_t a) { return a ? KB_P - a : 0; }
static inline uint32_t kb_mul(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % KB_P); }
static inline uint32_t kb_dbl(uint32_t a) { return kb_add(a, a); }
static inline uint32_t kb_pow(uint32_t a, uint64_t e) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = kb_mul(r, a);
        a = kb_mul(a, a);
        e >>= 1;
    }
    return r;
}
static inline uint32_t kb_inv(uint32_t a) { return kb_pow(a, KB_P - 2); }
static inline uint32_t kb_halve(uint32_t a) { return (a & 1) ? (uint32_t)(((uint64_t)a + KB_P) >> 1) : a >> 1; }
/* a / 2^k */
static inline uint32_t kb_div_2exp(uint32_t a, unsigned k) {
    while (k--) a = kb_halve(a);
    return a;
}
/* generator of the multiplicative subgroup of order 2^bits: 3^((p-1)/2^bits) */
static inline uint32_t kb_two_adic_generator(unsigned bits) { return kb_pow(KB_GENERATOR, (uint64_t)(KB_P - 1) >> bits); }

/* ---- F_p^4 = F_p[X]/(X^4 - 3); element = coefficients of X^0..X^3 ------------------- */
typedef struct { uint32_t c[4]; } kb4;
#define KB4_W 3u

static inline kb4 kb4_zero(void) { kb4 r = {{0, 0, 0, 0}}; return r; }
static inline kb4 kb4_one(void) { kb4 r = {{1, 0, 0, 0}}; return r; }
static inline kb4 kb4_from_base(uint32_t a) { kb4 r = {{a, 0, 0, 0}}; return r; }
static inline int kb4_eq(kb4 a, kb4 b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3]; }
static inline kb4 kb4_add(kb4 a, kb4 b) { kb4 r; for (int i = 0; i < 4; i++) r.c[i] = kb_add(a.c[i], b.c[i]); return r; }
static inline kb4 kb4_sub(kb4 a, kb4 b) { kb4 r; for (int i = 0; i < 4; i++) r.c[i] = kb_sub(a.c[i], b.c[i]); return r; }
static inline kb4 kb4_neg(kb4 a) { kb4 r; for (int i = 0; i < 4; i++) r.c[i] = kb_neg(a.c[i]); return r; }
static inline kb4 kb4_scale(kb4 a, uint32_t s) { kb4 r; for (int i = 0; i < 4; i++) r.c[i] = kb_mul(a.c[i], s); return r; }
static inline kb4 kb4_add_base(kb4 a, uint32_t s) { a.c[0] = kb_add(a.c[0], s); return a; }
static inline kb4 kb4_sub_base(kb4 a, uint32_t s) { a.c[0] = kb_sub(a.c[0], s); return a; }
static inline kb4 kb4_mul(kb4 a, kb4 b) {
    /* schoolbook, X^4 = 3 */
    uint64_t t[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) t[i + j] = (t[i + j] + (uint64_t)a.c[i] * b.c[j]) % KB_P;
    kb4 r;
    for (int i = 0; i < 4; i++) {
        uint64_t v = t[i];
        if (i < 3) v += KB4_W * t[i + 4];
        r.c[i] = (uint32_t)(v % KB_P);
    }
    return r;
}
static inline kb4 kb4_sqr(kb4 a) { return kb4_mul(a, a); }
static inline kb4 kb4_inv(kb4 a) {
    /* view as A + B*X over K = F_p[Y]/(Y^2-3), Y = X^2: A = a0 + a2 Y, B = a1 + a3 Y.
       a^{-1} = (A - B X) / (A^2 - Y B^2), the denominator N in K is inverted via its K/F_p norm. */
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    /* A^2 = (a0^2 + 3 a2^2) + (2 a0 a2) Y */
    uint32_t A2_0 = kb_add(kb_mul(a0, a0), kb_mul(KB4_W, kb_mul(a2, a2)));
    uint32_t A2_1 = kb_dbl(kb_mul(a0, a2));
    /* B^2 = (a1^2 + 3 a3^2) + (2 a1 a3) Y ; Y*B^2 = 3*(2 a1 a3) + (a1^2 + 3 a3^2) Y */
    uint32_t B2_0 = kb_add(kb_mul(a1, a1), kb_mul(KB4_W, kb_mul(a3, a3)));
    uint32_t B2_1 = kb_dbl(kb_mul(a1, a3));
    uint32_t n0 = kb_sub(A2_0, kb_mul(KB4_W, B2_1));
    uint32_t n1 = kb_sub(A2_1, B2_0);
    /* N^{-1} = (n0 - n1 Y) / (n0^2 - 3 n1^2) */
    uint32_t d = kb_sub(kb_mul(n0, n0), kb_mul(KB4_W, kb_mul(n1, n1)));
    uint32_t di = kb_inv(d);
    uint32_t m0 = kb_mul(n0, di), m1 = kb_neg(kb_mul(n1, di));
    /* (A - B X) * (m0 + m1 Y): A*M = (a0 m0 + 3 a2 m1) + (a0 m1 + a2 m0) Y ; B*M likewise */
    kb4 r;
    r.c[0] = kb_add(kb_mul(a0, m0), kb_mul(KB4_W, kb_mul(a2, m1)));
    r.c[2] = kb_add(kb_mul(a0, m1), kb_mul(a2, m0));
    r.c[1] = kb_neg(kb_add(kb_mul(a1, m0), kb_mul(KB4_W, kb_mul(a3, m1))));
    r.c[3] = kb_neg(kb_add(kb_mul(a1, m1), kb_mul(a3, m0)));
    return r;
}
static inline kb4 kb4_pow(kb4 a, uint64_t e) {
    kb4 r = kb4_one();
    while (e) {
        if (e & 1) r = kb4_mul(r, a);
        a = kb4_sqr(a);
        e >>= 1;
    }
    return r;
}

static inline unsigned bfo_log2(uint64_t n) { unsigned l = 0; while ((1ull << l) < n) l++; return l; }
static inline uint64_t bfo_bitrev(uint64_t x, unsigned bits) {
    uint64_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1ull) << (bits - 1 - i);
    return r;
}
#endif
