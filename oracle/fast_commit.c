/* oracle/fast_commit.c — TEST INFRASTRUCTURE / CPU BASELINE, not product code.
 *
 * A tuned host implementation of the SAME computation as bfo_pcs_commit (oracle/pcs_commit.c): `TwoAdicFriPcs::commit` of
 * one matrix = coset LDE (blow-up 2, shift GENERATOR) + bit-reversed rows + Poseidon2 `MerkleTreeMmcs::commit`
 * (reference call sites crates/stark/src/prover.rs:227,334,411; aliases crates/stark/src/kb31_poseidon2.rs:22-32),
 * written the way Plonky3's own CPU code is fast (p3-monty-31 `PackedMontyField31AVX512`, p3-dft `Radix2DitParallel`,
 * p3-merkle-tree's vertically packed leaf hashing — un-vendored, rev 93967fce; restated, not copied):
 *   - Montgomery arithmetic (R = 2^32), 16 lanes per AVX-512 register, products through vpmuludq on even/odd lanes;
 *   - the batch NTT applies every butterfly to whole ROWS (vectorised across the columns, no shuffles), in passes of up to 11
 *     stages over row sets that are closed under those stages and fit the L2 cache (2 passes over memory per transform);
 *     inverse = decimation in frequency (natural in, bit-reversed out), forward = decimation in time (bit-reversed in,
 *     natural out), so neither needs a bit-reversal pass, and the zero padding is the trivial first forward stage;
 *   - Poseidon2 on 16 independent states at a time, word i of the 16 states in one register ("vertical" packing): 16 leaves
 *     or 16 tree nodes per permutation call;
 *   - OpenMP over row sets / leaf groups.
 * This is what bench.py times as the CPU arm (`--impl reference`, `cpu_baseline`): a scalar `% p` port says nothing about
 * a rayon + AVX-512 prover.  The slow restatement (oracle/dft.c, merkle.c, poseidon2.c) stays the CHECKER; this file is
 * itself checked against it (tests/test_oracle_fast_commit.py) and produces the golden root of the bench workload.
 *
 * Needs AVX-512F/DQ/BW (bfo_fast_available()); without it callers use the slow oracle.
 */
#include "bf_oracle.h"
#include "kb31.h"
#include "rc_16_30.h"
#include <immintrin.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define TGT __attribute__((target("avx512f,avx512dq,avx512bw,avx512vl")))
#define FP 2130706433u
#define FPINV 0x81000001u /* p^-1 mod 2^32 */
#define FR2 402124772u    /* 2^64 mod p */
#define FONE 0x01fffffeu  /* 2^32 mod p */

int bfo_fast_available(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512bw") &&
           __builtin_cpu_supports("avx512vl");
}

/* ---- scalar Montgomery helpers (tables, conversions) ---------------------------------------------------------- */
static inline uint32_t m_mul(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t m = (uint32_t)t * FPINV;
    uint32_t u = (uint32_t)(((uint64_t)m * FP) >> 32);
    uint32_t hi = (uint32_t)(t >> 32);
    return hi >= u ? hi - u : hi - u + FP;
}
static inline uint32_t m_to(uint32_t x) { return m_mul(x, FR2); }
static inline uint32_t m_from(uint32_t x) { return m_mul(x, 1u); }
static uint32_t m_pow(uint32_t a, uint64_t e) {
    uint32_t r = FONE;
    while (e) {
        if (e & 1) r = m_mul(r, a);
        a = m_mul(a, a);
        e >>= 1;
    }
    return r;
}

/* ---- packed field arithmetic: 16 Montgomery residues in [0, p) per register ------------------------------------ */
typedef __m512i V;
TGT static inline V v_add(V a, V b) {
    V t = _mm512_add_epi32(a, b);
    return _mm512_min_epu32(t, _mm512_sub_epi32(t, _mm512_set1_epi32((int)FP)));
}
TGT static inline V v_sub(V a, V b) {
    V t = _mm512_sub_epi32(a, b);
    return _mm512_min_epu32(t, _mm512_add_epi32(t, _mm512_set1_epi32((int)FP)));
}
TGT static inline V v_dbl(V a) { return v_add(a, a); }
/* Montgomery product: six 32x32->64 multiplies (even and odd lanes), result corrected into [0, p) */
TGT static inline V v_mul(V a, V b) {
    const V P = _mm512_set1_epi32((int)FP), MU = _mm512_set1_epi32((int)FPINV);
    V ao = _mm512_srli_epi64(a, 32), bo = _mm512_srli_epi64(b, 32);
    V pe = _mm512_mul_epu32(a, b), po = _mm512_mul_epu32(ao, bo);
    V qe = _mm512_mul_epu32(pe, MU), qo = _mm512_mul_epu32(po, MU);
    V de = _mm512_sub_epi64(pe, _mm512_mul_epu32(qe, P)), dv = _mm512_sub_epi64(po, _mm512_mul_epu32(qo, P));
    V r = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(de, 32), dv); /* high halves: values in (-p, p) */
    return _mm512_min_epu32(r, _mm512_add_epi32(r, P));
}
TGT static inline V v_halve(V a) { /* a/2: (a + (a odd ? p : 0)) >> 1 */
    __mmask16 odd = _mm512_test_epi32_mask(a, _mm512_set1_epi32(1));
    return _mm512_srli_epi32(_mm512_mask_add_epi32(a, odd, a, _mm512_set1_epi32((int)FP)), 1);
}
/* sum +- a / 2^k: p = 1 mod 2^k for k <= 24, so with lo = a mod 2^k: a / 2^k = (a >> k) - lo * 127 * 2^(24-k) (mod p) */
TGT static inline V v_div2k_addto(V a, int k, V sum, int negate) {
    V lo = _mm512_and_si512(a, _mm512_set1_epi32((int)((1u << k) - 1)));
    V u = _mm512_srli_epi32(a, (unsigned)k);
    V t = _mm512_slli_epi32(_mm512_sub_epi32(_mm512_slli_epi32(lo, 7), lo), (unsigned)(24 - k)); /* < p */
    return negate ? v_sub(v_add(sum, t), u) : v_add(v_sub(sum, t), u);
}

/* ---- Poseidon2 (width 16, x^3, 8 external + 13 internal rounds; constants as kb31_poseidon2.rs:35-50) on 16 states ----- */
static uint32_t RC_EXT[8][16], RC_INT[13]; /* Montgomery form */
static int rc_ready;
static void init_rc(void) {
    if (__atomic_load_n(&rc_ready, __ATOMIC_ACQUIRE)) return;
    for (int r = 0; r < 4; r++)
        for (int i = 0; i < 16; i++) {
            RC_EXT[r][i] = m_to(BF_RC_16_30[r][i]);
            RC_EXT[4 + r][i] = m_to(BF_RC_16_30[17 + r][i]);
        }
    for (int r = 0; r < 13; r++) RC_INT[r] = m_to(BF_RC_16_30[4 + r][0]);
    __atomic_store_n(&rc_ready, 1, __ATOMIC_RELEASE);
}
TGT static inline V v_sbox(V x) { return v_mul(v_mul(x, x), x); }
TGT static inline void v_mat4(V* x) {
    V a = x[0], b = x[1], c = x[2], d = x[3];
    V t01 = v_add(a, b), t23 = v_add(c, d), t = v_add(t01, t23);
    V t01123 = v_add(t, b), t01233 = v_add(t, d);
    x[3] = v_add(t01233, v_dbl(a));
    x[1] = v_add(t01123, v_dbl(c));
    x[0] = v_add(t01123, t01);
    x[2] = v_add(t01233, t23);
}
TGT static inline void v_external(V* s) {
    for (int k = 0; k < 4; k++) v_mat4(s + 4 * k);
    for (int i = 0; i < 4; i++) {
        V t = v_add(v_add(s[i], s[4 + i]), v_add(s[8 + i], s[12 + i]));
        for (int k = 0; k < 4; k++) s[4 * k + i] = v_add(s[4 * k + i], t);
    }
}
/* 1 + Diag(V), V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/8, 1/2^24, -1/2^8, -1/8, -1/16, -1/2^24] */
TGT static inline void v_internal(V* s) {
    V part = v_add(v_add(v_add(s[1], s[2]), v_add(s[3], s[4])), v_add(v_add(s[5], s[6]), v_add(s[7], s[8])));
    part = v_add(part, v_add(v_add(v_add(s[9], s[10]), v_add(s[11], s[12])), v_add(v_add(s[13], s[14]), s[15])));
    V sum = v_add(part, s[0]);
    s[0] = v_sub(part, s[0]);
    s[1] = v_add(s[1], sum);
    s[2] = v_add(v_dbl(s[2]), sum);
    s[3] = v_add(v_halve(s[3]), sum);
    s[4] = v_add(v_add(v_dbl(s[4]), s[4]), sum);
    s[5] = v_add(v_dbl(v_dbl(s[5])), sum);
    s[6] = v_sub(sum, v_halve(s[6]));
    s[7] = v_sub(sum, v_add(v_dbl(s[7]), s[7]));
    s[8] = v_sub(sum, v_dbl(v_dbl(s[8])));
    s[9] = v_div2k_addto(s[9], 8, sum, 0);
    s[10] = v_div2k_addto(s[10], 3, sum, 0);
    s[11] = v_div2k_addto(s[11], 24, sum, 0);
    s[12] = v_div2k_addto(s[12], 8, sum, 1);
    s[13] = v_div2k_addto(s[13], 3, sum, 1);
    s[14] = v_div2k_addto(s[14], 4, sum, 1);
    s[15] = v_div2k_addto(s[15], 24, sum, 1);
}
TGT static void v_permute(V* s) {
    v_external(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = v_sbox(v_add(s[i], _mm512_set1_epi32((int)RC_EXT[r][i])));
        v_external(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = v_sbox(v_add(s[0], _mm512_set1_epi32((int)RC_INT[r])));
        v_internal(s);
    }
    for (int r = 4; r < 8; r++) {
        for (int i = 0; i < 16; i++) s[i] = v_sbox(v_add(s[i], _mm512_set1_epi32((int)RC_EXT[r][i])));
        v_external(s);
    }
}

/* in-register transpose of a 16x16 block of 32-bit words: r[i] = row i  ->  r[c] = column c */
TGT static inline void v_transpose16(V* r) {
    V t[16], q[16];
    for (int i = 0; i < 16; i += 2) {
        t[i] = _mm512_unpacklo_epi32(r[i], r[i + 1]);
        t[i + 1] = _mm512_unpackhi_epi32(r[i], r[i + 1]);
    }
    /* q[4k + j], 128-bit lane L = (rows 4k..4k+3, column 4L + j) */
    for (int k = 0; k < 16; k += 4) {
        q[k] = _mm512_unpacklo_epi64(t[k], t[k + 2]);
        q[k + 1] = _mm512_unpackhi_epi64(t[k], t[k + 2]);
        q[k + 2] = _mm512_unpacklo_epi64(t[k + 1], t[k + 3]);
        q[k + 3] = _mm512_unpackhi_epi64(t[k + 1], t[k + 3]);
    }
    for (int j = 0; j < 4; j++) {
        V u0 = _mm512_shuffle_i32x4(q[j], q[4 + j], 0x88), u1 = _mm512_shuffle_i32x4(q[j], q[4 + j], 0xdd);
        V u2 = _mm512_shuffle_i32x4(q[8 + j], q[12 + j], 0x88), u3 = _mm512_shuffle_i32x4(q[8 + j], q[12 + j], 0xdd);
        r[j] = _mm512_shuffle_i32x4(u0, u2, 0x88);       /* lane 0 of the four row groups: column j */
        r[8 + j] = _mm512_shuffle_i32x4(u0, u2, 0xdd);   /* lane 2: column 8 + j */
        r[4 + j] = _mm512_shuffle_i32x4(u1, u3, 0x88);   /* lane 1: column 4 + j */
        r[12 + j] = _mm512_shuffle_i32x4(u1, u3, 0xdd);  /* lane 3: column 12 + j */
    }
}

/* ---- batch NTT passes ----------------------------------------------------------------------------------------------------- */
static inline uint64_t brev64(uint64_t x, unsigned bits) {
    if (!bits) return 0;
    x = __builtin_bswap64(x);
    x = ((x & 0x0f0f0f0f0f0f0f0full) << 4) | ((x >> 4) & 0x0f0f0f0f0f0f0f0full);
    x = ((x & 0x3333333333333333ull) << 2) | ((x >> 2) & 0x3333333333333333ull);
    x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
    return x >> (64 - bits);
}

typedef struct {
    unsigned log_n;  /* transform size */
    uint32_t* tw;    /* tw[e] = w^e (Montgomery), w primitive 2^log_n-th root (or its inverse), e < 2^(log_n-1) */
} twtab;
static void tw_build(twtab* t, unsigned log_n, int inverse) {
    t->log_n = log_n;
    uint64_t half = log_n ? (1ull << (log_n - 1)) : 1;
    t->tw = (uint32_t*)aligned_alloc(64, (half * 4 + 63) & ~63ull);
    uint32_t w = m_to(kb_two_adic_generator(log_n));
    if (inverse) w = m_pow(w, (uint64_t)FP - 2);
    /* chunked powers: every thread starts from w^(chunk start) */
#pragma omp parallel
    {
        int nt = omp_get_num_threads(), id = omp_get_thread_num();
        uint64_t a = half * (uint64_t)id / (uint64_t)nt, b = half * (uint64_t)(id + 1) / (uint64_t)nt;
        uint32_t x = m_pow(w, a);
        for (uint64_t e = a; e < b; e++) {
            t->tw[e] = x;
            x = m_mul(x, w);
        }
    }
}

/* One pass = the stages on index bits [b0, b0 + g) of a size-2^log_n transform over `ws`-strided rows, columns [c0, c0 + cb).
 * dif != 0: decimation in frequency, bits from high to low, (a, b) -> (a + b, (a - b) w); else decimation in time, low to high,
 * (a, b) -> (a + w b, a - w b).  The 2^g rows of a closed set are gathered into `loc` (cb words per row), transformed there and
 * written to dst.  Source variants: plain rows of `src` (stride ws_src); canonical input to convert (first inverse pass);
 * expansion `row r <- coef[r >> 1] * scale[r >> 1]` (first forward pass, the zero padding already folded in). */
typedef struct {
    const uint32_t* src;
    uint64_t ws_src;
    int src_canonical;
    const uint32_t* scale; /* non-null: expansion source (src = coefficient rows, half as many) */
    uint32_t* dst;
    uint64_t ws_dst;
    uint64_t cols; /* real columns (<= ws) */
} passio;

TGT static void ntt_pass_set(const twtab* T, const passio* io, int dif, unsigned b0, unsigned g, uint64_t high, uint64_t low, uint64_t c0, uint64_t cb,
                             uint32_t* loc) {
    const uint64_t nloc = 1ull << g;
    const V R2 = _mm512_set1_epi32((int)FR2);
    /* gather */
    for (uint64_t j = 0; j < nloc; j++) {
        uint64_t r = (high << (b0 + g)) | (j << b0) | low;
        uint32_t* d = loc + j * cb;
        if (io->scale) {
            const uint32_t* s = io->src + (r >> 1) * io->ws_src + c0;
            V sc = _mm512_set1_epi32((int)io->scale[r >> 1]);
            for (uint64_t c = 0; c < cb; c += 16) _mm512_store_si512((void*)(d + c), v_mul(_mm512_loadu_si512((const void*)(s + c)), sc));
        } else if (io->src_canonical) {
            const uint32_t* s = io->src + r * io->ws_src + c0;
            for (uint64_t c = 0; c < cb; c += 16) {
                uint64_t left = io->cols > c0 + c ? io->cols - (c0 + c) : 0;
                __mmask16 k = left >= 16 ? (__mmask16)0xFFFF : (__mmask16)((1u << left) - 1);
                _mm512_store_si512((void*)(d + c), v_mul(_mm512_maskz_loadu_epi32(k, (const void*)(s + c)), R2));
            }
        } else {
            const uint32_t* s = io->src + r * io->ws_src + c0;
            for (uint64_t c = 0; c < cb; c += 16) _mm512_store_si512((void*)(d + c), _mm512_loadu_si512((const void*)(s + c)));
        }
    }
    /* stages */
    for (unsigned st = 0; st < g; st++) {
        const unsigned lb = dif ? g - 1 - st : st, b = b0 + lb;
        const uint64_t lh = 1ull << lb;
        for (uint64_t blk = 0; blk < nloc; blk += 2 * lh)
            for (uint64_t k = 0; k < lh; k++) {
                const uint64_t t = (k << b0) | low;                   /* position inside the butterfly group of size 2^(b+1) */
                const uint64_t e = t << (T->log_n - b - 1);           /* exponent of the size-2^log_n root */
                uint32_t* pa = loc + (blk + k) * cb;
                uint32_t* pb = pa + lh * cb;
                if (e == 0) {
                    for (uint64_t c = 0; c < cb; c += 16) {
                        V a = _mm512_load_si512((const void*)(pa + c)), bb = _mm512_load_si512((const void*)(pb + c));
                        _mm512_store_si512((void*)(pa + c), v_add(a, bb));
                        _mm512_store_si512((void*)(pb + c), v_sub(a, bb));
                    }
                } else {
                    const V w = _mm512_set1_epi32((int)T->tw[e]);
                    if (dif)
                        for (uint64_t c = 0; c < cb; c += 16) {
                            V a = _mm512_load_si512((const void*)(pa + c)), bb = _mm512_load_si512((const void*)(pb + c));
                            _mm512_store_si512((void*)(pa + c), v_add(a, bb));
                            _mm512_store_si512((void*)(pb + c), v_mul(v_sub(a, bb), w));
                        }
                    else
                        for (uint64_t c = 0; c < cb; c += 16) {
                            V a = _mm512_load_si512((const void*)(pa + c)), bb = v_mul(_mm512_load_si512((const void*)(pb + c)), w);
                            _mm512_store_si512((void*)(pa + c), v_add(a, bb));
                            _mm512_store_si512((void*)(pb + c), v_sub(a, bb));
                        }
                }
            }
    }
    /* scatter */
    for (uint64_t j = 0; j < nloc; j++) {
        uint64_t r = (high << (b0 + g)) | (j << b0) | low;
        uint32_t* d = io->dst + r * io->ws_dst + c0;
        const uint32_t* s = loc + j * cb;
        for (uint64_t c = 0; c < cb; c += 16) _mm512_storeu_si512((void*)(d + c), _mm512_load_si512((const void*)(s + c)));
    }
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Work buffers are kept between calls (first-touch page faults of fresh multi-GiB allocations cost as much as the transforms; the
   GPU side caches its blocks the same way) and released by bfo_fast_release(). */
#include <sys/mman.h>
static struct { void* p; size_t bytes; } g_ws[4];
static void* ws_get(int slot, size_t bytes) {
    bytes = (bytes + 63) & ~(size_t)63;
    if (g_ws[slot].bytes < bytes) {
        free(g_ws[slot].p);
        size_t al = bytes >= (4u << 20) ? (2u << 20) : 64;
        g_ws[slot].p = aligned_alloc(al, (bytes + al - 1) / al * al);
        g_ws[slot].bytes = g_ws[slot].p ? bytes : 0;
#ifdef MADV_HUGEPAGE
        if (g_ws[slot].p && al > 64) madvise(g_ws[slot].p, (bytes + al - 1) / al * al, MADV_HUGEPAGE);
#endif
    }
    return g_ws[slot].p;
}
void bfo_fast_release(void) {
    for (int i = 0; i < 4; i++) {
        free(g_ws[i].p);
        g_ws[i].p = NULL;
        g_ws[i].bytes = 0;
    }
}
#define MAX_G 11
#define LOC_BYTES (1u << 20)

/* all stages on index bits [first_bit, log_n) (forward after the trivial stage: first_bit = 1; inverse: 0) */
static void ntt_run(const twtab* T, passio io, int dif, unsigned first_bit, uint64_t ws) {
    const unsigned nbits = T->log_n - first_bit;
    const unsigned npass = nbits ? (nbits + MAX_G - 1) / MAX_G : 1;
    unsigned gs[8], b0s[8];
    {
        unsigned base = nbits / npass, extra = nbits % npass, acc = first_bit;
        for (unsigned i = 0; i < npass; i++) {
            gs[i] = base + (i < extra ? 1 : 0);
            b0s[i] = acc;
            acc += gs[i];
        }
    }
    for (unsigned s = 0; s < npass; s++) {
        double tp0 = now_s();
        const unsigned i = dif ? npass - 1 - s : s; /* DIF: high bits first; DIT: low bits first */
        const unsigned g = gs[i], b0 = b0s[i];
        passio cur = io;
        if (s > 0) { /* later passes work in place on the destination */
            cur.src = io.dst;
            cur.ws_src = io.ws_dst;
            cur.src_canonical = 0;
            cur.scale = NULL;
        }
        uint64_t cb = LOC_BYTES / 4 >> g;
        cb = cb < 16 ? 16 : cb / 16 * 16;
        if (cb > ws) cb = ws;
        const uint64_t ncb = (ws + cb - 1) / cb;
        const uint64_t nlow = 1ull << b0, nhigh = 1ull << (T->log_n - b0 - g);
        const uint64_t nsets = nlow * nhigh;
#pragma omp parallel
        {
            uint32_t* loc = (uint32_t*)aligned_alloc(64, ((size_t)cb << g) * 4);
#pragma omp for schedule(dynamic, 1) collapse(2)
            for (uint64_t set = 0; set < nsets; set++)
                for (uint64_t k = 0; k < ncb; k++) {
                    uint64_t c0 = k * cb, w = c0 + cb <= ws ? cb : ws - c0;
                    ntt_pass_set(T, &cur, dif, b0, g, set / nlow, set % nlow, c0, w, loc);
                }
            free(loc);
        }
        if (getenv("BFO_FAST_DEBUG")) fprintf(stderr, "  pass dif=%d bits [%u,%u) cb=%llu sets=%llu: %.3f s\n", dif, b0, b0 + g, (unsigned long long)cb, (unsigned long long)nsets, now_s() - tp0);
    }
}

/* ---- hashing ---------------------------------------------------------------------------------------------------------------- */
/* leaf digests: leaf i = PaddingFreeSponge over LDE row bitrev(i) (rows held in natural order), 16 leaves per call */
TGT static void leaf_group(const uint32_t* lde, uint64_t ws, uint64_t cols, unsigned log_h, uint64_t i0, uint32_t* digests) {
    const uint32_t* rows[16];
    for (int l = 0; l < 16; l++) rows[l] = lde + brev64(i0 + (uint64_t)l, log_h) * ws;
    V s[16];
    for (int k = 0; k < 16; k++) s[k] = _mm512_setzero_si512();
    for (uint64_t c = 0; c < cols; c += 16) {
        V t[16];
        for (int l = 0; l < 16; l++) t[l] = _mm512_loadu_si512((const void*)(rows[l] + c)); /* ws is a multiple of 16: in bounds */
        v_transpose16(t);
        uint64_t nc = cols - c < 16 ? cols - c : 16;
        for (uint64_t k = 0; k < (nc < 8 ? nc : 8); k++) s[k] = t[k];
        v_permute(s);
        if (nc > 8) {
            for (uint64_t k = 8; k < nc; k++) s[k - 8] = t[k];
            v_permute(s);
        }
    }
    uint32_t tmp[8][16] __attribute__((aligned(64)));
    for (int k = 0; k < 8; k++) _mm512_store_si512((void*)tmp[k], s[k]);
    for (int l = 0; l < 16; l++)
        for (int k = 0; k < 8; k++) digests[(i0 + (uint64_t)l) * 8 + (uint64_t)k] = tmp[k][l];
}
/* next[k] = compress(prev[2k], prev[2k+1]) for k in [k0, k0 + cnt), cnt <= 16 */
TGT static void compress_group(const uint32_t* prev, uint32_t* next, uint64_t k0, uint64_t cnt) {
    V s[16];
    for (uint64_t l = 0; l < 16; l++) s[l] = l < cnt ? _mm512_loadu_si512((const void*)(prev + (k0 + l) * 16)) : _mm512_setzero_si512();
    v_transpose16(s);
    v_permute(s);
    uint32_t tmp[8][16] __attribute__((aligned(64)));
    for (int k = 0; k < 8; k++) _mm512_store_si512((void*)tmp[k], s[k]);
    for (uint64_t l = 0; l < cnt; l++)
        for (int k = 0; k < 8; k++) next[(k0 + l) * 8 + (uint64_t)k] = tmp[k][l];
}

/* Pcs::commit of ONE rows x cols matrix of canonical words on the natural domain (LDE shift GENERATOR, log_blowup 1).
 * root: canonical.  lde_out (optional, 2*rows x cols, canonical, rows bit-reversed as TwoAdicFriPcs stores them): filled when
 * non-null (checker use; not part of the timed baseline).  phase_sec: {lde, leaf hashing, compression layers}.
 * Returns 0, or -1 when the shape is unsupported here (rows < 16 or not a power of two, or no AVX-512). */
int bfo_fast_pcs_commit(const uint32_t* in, uint64_t rows, uint64_t cols, uint32_t root[8], uint32_t* lde_out, double phase_sec[3]) {
    if (!bfo_fast_available() || rows < 16 || (rows & (rows - 1)) || cols == 0) return -1;
    init_rc();
    const unsigned m = bfo_log2(rows), M = m + 1;
    if (M > KB_TWO_ADICITY) return -1;
    const uint64_t n = rows, N = 2 * rows, ws = (cols + 15) / 16 * 16;
    double t0 = now_s();
    twtab Ti, Tf;
    tw_build(&Ti, m, 1);
    tw_build(&Tf, M, 0);
    uint32_t* coef = (uint32_t*)ws_get(0, n * ws * 4);
    uint32_t* lde = (uint32_t*)ws_get(1, N * ws * 4);
    uint32_t* scale = (uint32_t*)aligned_alloc(64, (n * 4 + 63) & ~63ull);
    uint32_t* layer = (uint32_t*)ws_get(2, N * 32);
    uint32_t* layer2 = (uint32_t*)ws_get(3, N * 16);
    if (!coef || !lde || !scale || !layer || !layer2) {
        bfo_fast_release();
        free(scale); free(Ti.tw); free(Tf.tw);
        return -1;
    }
    { /* scale[q] = shift^bitrev(q) / n: the coefficient at bit-reversed position q, moved onto the coset */
        const uint32_t sh = m_to(KB_GENERATOR), ninv = m_pow(m_to((uint32_t)(n % FP)), (uint64_t)FP - 2);
#pragma omp parallel
        {
            int nt = omp_get_num_threads(), id = omp_get_thread_num();
            uint64_t a = n * (uint64_t)id / (uint64_t)nt, b = n * (uint64_t)(id + 1) / (uint64_t)nt;
            uint32_t x = m_mul(m_pow(sh, a), ninv);
            for (uint64_t k = a; k < b; k++) {
                scale[brev64(k, m)] = x;
                x = m_mul(x, sh);
            }
        }
    }
    if (getenv("BFO_FAST_DEBUG")) fprintf(stderr, "  setup (tables, allocation): %.3f s\n", now_s() - t0);
    /* inverse transform (DIF, inverse root): natural-order evaluations -> coefficients in bit-reversed order */
    passio inv = {in, cols, 1, NULL, coef, ws, cols};
    ntt_run(&Ti, inv, 1, 0, ws);
    /* forward transform of the zero-padded, shifted coefficients (DIT from bit-reversed input): the first stage pairs
       (coefficient, 0) and only duplicates it, so it is folded into the gather of the first pass */
    passio fwd = {coef, ws, 0, scale, lde, ws, cols};
    if (M > 1) ntt_run(&Tf, fwd, 0, 1, ws);
    free(scale);
    free(Ti.tw);
    free(Tf.tw);
    double t1 = now_s();
    /* Merkle tree over the bit-reversed rows */
#pragma omp parallel for schedule(dynamic, 64)
    for (uint64_t g = 0; g < N / 16; g++) leaf_group(lde, ws, cols, M, g * 16, layer);
    double t2 = now_s();
    if (lde_out) {
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < N; i++) {
            const uint32_t* s = lde + brev64(i, M) * ws;
            for (uint64_t c = 0; c < cols; c++) lde_out[i * cols + c] = m_from(s[c]);
        }
    }
    double t2b = now_s();
    uint64_t len = N;
    uint32_t *cur = layer, *nxt = layer2; /* ping-pong: every level is half the previous one */
    while (len > 1) {
        uint64_t nl = len / 2;
#pragma omp parallel for schedule(static) if (nl >= 1024)
        for (uint64_t k0 = 0; k0 < nl; k0 += 16) compress_group(cur, nxt, k0, nl - k0 < 16 ? nl - k0 : 16);
        uint32_t* t = cur;
        cur = nxt;
        nxt = t;
        len = nl;
    }
    for (int k = 0; k < 8; k++) root[k] = m_from(cur[k]);
    double t3 = now_s();
    if (phase_sec) {
        phase_sec[0] = t1 - t0;
        phase_sec[1] = t2 - t1;
        phase_sec[2] = t3 - t2b;
    }
    return 0;
}

/* throughput helper: n permutations (n multiple of 16) over packed states, for the Poseidon2 hashes/s baseline */
TGT static void permute_many_chunk(uint32_t* states, uint64_t n16) {
    for (uint64_t g = 0; g < n16; g++) {
        V s[16];
        for (int l = 0; l < 16; l++) s[l] = _mm512_loadu_si512((const void*)(states + (g * 16 + (uint64_t)l) * 16));
        v_transpose16(s);
        v_permute(s);
        v_transpose16(s);
        for (int l = 0; l < 16; l++) _mm512_storeu_si512((void*)(states + (g * 16 + (uint64_t)l) * 16), s[l]);
    }
}
/* states: n x 16 canonical words, permuted in place (n a multiple of 16); returns -1 without AVX-512 */
TGT static void convert_block(uint32_t* p, uint64_t words, int to_mont) {
    const V k = _mm512_set1_epi32(to_mont ? (int)FR2 : 1);
    for (uint64_t i = 0; i < words; i += 16) _mm512_storeu_si512((void*)(p + i), v_mul(_mm512_loadu_si512((const void*)(p + i)), k));
}
int bfo_fast_permute_many(uint32_t* states, uint64_t n) {
    if (!bfo_fast_available() || n % 16) return -1;
    init_rc();
#pragma omp parallel for schedule(static)
    for (uint64_t g = 0; g < n / 16; g++) {
        convert_block(states + g * 256, 256, 1);
        permute_many_chunk(states + g * 256, 1);
        convert_block(states + g * 256, 256, 0);
    }
    return 0;
}

/* ---- FRI commit phase (CPU arm of `pcs.open`, reference crates/stark/src/prover.rs:460 -> p3-fri prover::commit_phase) -----------------
 * inputs[k]: 2^log_len[k] extension elements (4 canonical words each), heights strictly decreasing, inputs[0] the tallest (the
 * per-height reduced openings).  Per round: commit the folded vector as an (len/2 x 8) matrix (row i = elements 2i, 2i+1; leaf =
 * one permutation of (row | 0^8), then TruncatedPermutation compressions: the same packed 16-lane Poseidon2 as the commitments),
 * fold with the CALLER-SUPPLIED beta of that round (fold_matrix: out[i] = (1/2 + beta/2 g^-br(i)) lo + (1/2 - beta/2 g^-br(i)) hi),
 * add the input of the new length.  Stops at 2 elements (log_blowup 1).  roots: rounds x 8 canonical; final_poly: 4 canonical words.
 * sec: {hashing, folding}.  rollin_beta2 != 0: a rolled-in input is multiplied by beta^2 first (the later upstream rule).
 * Returns the number of rounds, or -1. */
static inline void e_mul_m(const uint32_t* a, const uint32_t* b, uint32_t* r) { /* F_p[X]/(X^4 - 3), Montgomery words */
    uint32_t t[7];
    for (int k = 0; k < 7; k++) t[k] = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            uint32_t s = t[i + j] + m_mul(a[i], b[j]);
            t[i + j] = s >= FP ? s - FP : s;
        }
    for (int k = 0; k < 3; k++) {
        uint32_t w = m_mul(t[4 + k], m_to(3));
        uint32_t s = t[k] + w;
        t[k] = s >= FP ? s - FP : s;
    }
    for (int k = 0; k < 4; k++) r[k] = t[k];
}
static inline uint32_t m_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s >= FP ? s - FP : s;
}
static inline uint32_t m_sub(uint32_t a, uint32_t b) { return a >= b ? a - b : a + FP - b; }
TGT static void leaf_group_rows8(const uint32_t* mat, uint64_t i0, uint64_t cnt, uint32_t* digests) { /* rows of 8 Montgomery words, in order */
    V t[16];
    for (uint64_t l = 0; l < 16; l++)
        t[l] = l < cnt ? _mm512_maskz_loadu_epi32(0x00ff, (const void*)(mat + (i0 + l) * 8)) : _mm512_setzero_si512();
    v_transpose16(t); /* t[k] = word k of the 16 rows; words 8..15 are the zero capacity */
    v_permute(t);
    uint32_t tmp[8][16] __attribute__((aligned(64)));
    for (int k = 0; k < 8; k++) _mm512_store_si512((void*)tmp[k], t[k]);
    for (uint64_t l = 0; l < cnt; l++)
        for (int k = 0; k < 8; k++) digests[(i0 + l) * 8 + (uint64_t)k] = tmp[k][l];
}
int bfo_fast_fri_commit_phase(const uint32_t* const* inputs, const uint32_t* log_len, int n_inputs, const uint32_t* betas, int rollin_beta2, uint32_t* roots,
                              uint32_t final_poly[4], double sec[2]) {
    if (!bfo_fast_available() || n_inputs < 1 || !inputs || !log_len || !betas || !roots) return -1;
    init_rc();
    const unsigned L = log_len[0];
    if (L < 2 || L > KB_TWO_ADICITY) return -1;
    const uint64_t len0 = 1ull << L;
    uint32_t* cur = (uint32_t*)aligned_alloc(64, (len0 * 16 + 64 + 63) & ~63ull);
    uint32_t* nxt = (uint32_t*)aligned_alloc(64, (len0 * 8 + 64 + 63) & ~63ull);
    uint32_t* lay = (uint32_t*)aligned_alloc(64, (len0 * 16 + 63) & ~63ull); /* digests of the current level */
    uint32_t* lay2 = (uint32_t*)aligned_alloc(64, (len0 * 8 + 63) & ~63ull);
    if (!cur || !nxt || !lay || !lay2) {
        free(cur); free(nxt); free(lay); free(lay2);
        return -1;
    }
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len0 * 4; i++) cur[i] = m_to(inputs[0][i]);
    const uint32_t half = m_pow(m_to(2), (uint64_t)FP - 2);
    double t_hash = 0, t_fold = 0;
    int next_in = 1, round = 0;
    uint64_t len = len0;
    while (len > 2) {
        const uint64_t h = len / 2; /* leaves = output length */
        unsigned log_h = 0;
        while ((1ull << log_h) < h) log_h++;
        double t0 = now_s();
#pragma omp parallel for schedule(static) if (h >= 1024)
        for (uint64_t g = 0; g < (h + 15) / 16; g++) leaf_group_rows8(cur, g * 16, h - g * 16 < 16 ? h - g * 16 : 16, lay);
        uint64_t ll = h;
        uint32_t *a = lay, *b = lay2;
        while (ll > 1) {
            uint64_t nl = ll / 2;
#pragma omp parallel for schedule(static) if (nl >= 1024)
            for (uint64_t k0 = 0; k0 < nl; k0 += 16) compress_group(a, b, k0, nl - k0 < 16 ? nl - k0 : 16);
            uint32_t* t = a;
            a = b;
            b = t;
            ll = nl;
        }
        for (int k = 0; k < 8; k++) roots[round * 8 + k] = m_from(a[k]);
        double t1 = now_s();
        uint32_t beta[4], hb[4], b2[4];
        for (int k = 0; k < 4; k++) {
            beta[k] = m_to(betas[round * 4 + k] % FP);
            hb[k] = m_mul(beta[k], half);
        }
        e_mul_m(beta, beta, b2);
        const uint32_t ginv = m_pow(m_pow(m_to(KB_GENERATOR), (uint64_t)(FP - 1) >> (log_h + 1)), (uint64_t)FP - 2); /* inverse generator of order 2^(log_h+1) */
        const uint32_t* add = (next_in < n_inputs && log_len[next_in] == log_h) ? inputs[next_in] : NULL;
#pragma omp parallel
        {
            int nt = omp_get_num_threads(), id = omp_get_thread_num();
            uint64_t i0 = h * (uint64_t)id / (uint64_t)nt, i1 = h * (uint64_t)(id + 1) / (uint64_t)nt;
            /* pw = ginv^bitrev(i), kept up to date while i counts up: i -> i + 1 clears the trailing ones of i (their mirrored
               bits leave the exponent: multiply by g^(2^k)) and sets the next bit (multiply by ginv^(2^k)): ~2 products per step */
            uint32_t GI[32], GG[32];
            GI[0] = ginv;
            GG[0] = m_pow(ginv, (uint64_t)FP - 2);
            for (unsigned k = 1; k < 32; k++) {
                GI[k] = m_mul(GI[k - 1], GI[k - 1]);
                GG[k] = m_mul(GG[k - 1], GG[k - 1]);
            }
            uint32_t pw = m_pow(ginv, brev64(i0, log_h));
            for (uint64_t i = i0; i < i1; i++) {
                if (i != i0) {
                    uint64_t prev = i - 1;
                    unsigned b = 0;
                    while (prev & 1) { /* trailing ones of i - 1: bit b was 1, now 0 */
                        pw = m_mul(pw, GG[log_h - 1 - b]);
                        prev >>= 1;
                        b++;
                    }
                    pw = m_mul(pw, GI[log_h - 1 - b]);
                }
                uint32_t pa[4], pb[4], x[4], y[4];
                for (int k = 0; k < 4; k++) {
                    uint32_t t = m_mul(hb[k], pw);
                    pa[k] = t;
                    pb[k] = t ? FP - t : 0;
                }
                pa[0] = m_add(pa[0], half);
                pb[0] = m_add(pb[0], half);
                e_mul_m(pa, cur + 8 * i, x);
                e_mul_m(pb, cur + 8 * i + 4, y);
                for (int k = 0; k < 4; k++) x[k] = m_add(x[k], y[k]);
                if (add) {
                    uint32_t r[4];
                    for (int k = 0; k < 4; k++) r[k] = m_to(add[4 * i + (uint64_t)k]);
                    if (rollin_beta2) {
                        uint32_t r2[4];
                        e_mul_m(b2, r, r2);
                        for (int k = 0; k < 4; k++) r[k] = r2[k];
                    }
                    for (int k = 0; k < 4; k++) x[k] = m_add(x[k], r[k]);
                }
                for (int k = 0; k < 4; k++) nxt[4 * i + (uint64_t)k] = x[k];
            }
        }
        if (add) next_in++;
        double t2 = now_s();
        t_hash += t1 - t0;
        t_fold += t2 - t1;
        uint32_t* t = cur;
        cur = nxt;
        nxt = t;
        len = h;
        round++;
    }
    for (int k = 0; k < 4; k++) final_poly[k] = m_from(cur[k]);
    if (sec) {
        sec[0] = t_hash;
        sec[1] = t_fold;
    }
    free(cur); free(nxt); free(lay); free(lay2);
    (void)m_sub;
    return round;
}
