/* oracle/dft.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Two-adic DFT / coset low-degree extension over KoalaBear, batched over the columns of a
 * row-major matrix: the `TwoAdicSubgroupDft` surface the reference instantiates as
 * `Dft = Radix2DitParallel<Val>` (crates/stark/src/kb31_poseidon2.rs:30) and reaches through
 * `pcs.commit` (crates/stark/src/prover.rs:227,334,411; crates/stark/src/machine.rs:196).
 * Definitions restated from Plonky3 p3-dft v0.1.0 @93967fce (un-vendored), trait
 * `TwoAdicSubgroupDft`: dft_batch, idft_batch, coset_dft_batch, coset_lde_batch
 * (SURVEY.md Appendix B.6).  Results are mathematically unique, so butterfly order is free.
 *
 * PARITY UNPINNED (no reference golden vectors); pinned instead against the O(n^2)
 * definition below (tests/test_oracle_dft.py).
 */
#include "bf_oracle.h"
#include "kb31.h"
#include <stdlib.h>
#include <string.h>

/* out[j][c] = P_c(shift * w_N^j), N = rows << added_bits, where P_c interpolates in[i][c] at w_rows^i */
void bfo_coset_lde_naive(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out) {
    unsigned log_n = bfo_log2(rows);
    uint64_t N = rows << added_bits;
    uint32_t w = kb_two_adic_generator(log_n);
    uint32_t winv = kb_inv(w);
    uint32_t ninv = kb_inv((uint32_t)(rows % KB_P));
    /* coefficients: c_k = 1/n sum_i in[i] w^{-ik} */
    uint32_t* coef = (uint32_t*)malloc(rows * cols * 4);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)rows; k++) {
        uint32_t step = kb_pow(winv, (uint64_t)k), x = 1;
        for (uint64_t c = 0; c < cols; c++) coef[k * cols + c] = 0;
        for (uint64_t i = 0; i < rows; i++) {
            for (uint64_t c = 0; c < cols; c++) coef[k * cols + c] = kb_add(coef[k * cols + c], kb_mul(in[i * cols + c], x));
            x = kb_mul(x, step);
        }
        for (uint64_t c = 0; c < cols; c++) coef[k * cols + c] = kb_mul(coef[k * cols + c], ninv);
    }
    uint32_t W = kb_two_adic_generator(log_n + added_bits);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)N; j++) {
        uint32_t pt = kb_mul(shift, kb_pow(W, (uint64_t)j));
        for (uint64_t c = 0; c < cols; c++) { /* Horner */
            uint32_t acc = 0;
            for (int64_t k = (int64_t)rows - 1; k >= 0; k--) acc = kb_add(kb_mul(acc, pt), coef[k * cols + c]);
            out[j * cols + c] = acc;
        }
    }
    free(coef);
}

/* in-place radix-2 transform with root `w` of order `rows`; natural order in and out */
static void ntt_inplace(uint32_t* m, uint64_t rows, uint64_t cols, uint32_t w) {
    unsigned log_n = bfo_log2(rows);
    if (rows <= 1) return;
    /* bit-reverse rows */
    uint32_t* tmp = (uint32_t*)malloc(cols * 4);
    for (uint64_t i = 0; i < rows; i++) {
        uint64_t j = bfo_bitrev(i, log_n);
        if (i < j) {
            memcpy(tmp, m + i * cols, cols * 4);
            memcpy(m + i * cols, m + j * cols, cols * 4);
            memcpy(m + j * cols, tmp, cols * 4);
        }
    }
    free(tmp);
    uint32_t* tw = (uint32_t*)malloc((rows / 2) * 4);
    tw[0] = 1;
    for (uint64_t i = 1; i < rows / 2; i++) tw[i] = kb_mul(tw[i - 1], w);
    for (unsigned s = 0; s < log_n; s++) {
        uint64_t half = 1ull << s;
        uint64_t tstride = rows >> (s + 1);
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < (int64_t)(rows / 2); b++) {
            uint64_t grp = (uint64_t)b >> s, k = (uint64_t)b & (half - 1);
            uint32_t* lo = m + (grp * 2 * half + k) * cols;
            uint32_t* hi = lo + half * cols;
            uint32_t t = tw[k * tstride];
            for (uint64_t c = 0; c < cols; c++) {
                uint32_t u = lo[c], v = kb_mul(hi[c], t);
                lo[c] = kb_add(u, v);
                hi[c] = kb_sub(u, v);
            }
        }
    }
    free(tw);
}

void bfo_dft_batch(uint32_t* mat, uint64_t rows, uint64_t cols) { ntt_inplace(mat, rows, cols, kb_two_adic_generator(bfo_log2(rows))); }

void bfo_idft_batch(uint32_t* mat, uint64_t rows, uint64_t cols) {
    ntt_inplace(mat, rows, cols, kb_inv(kb_two_adic_generator(bfo_log2(rows))));
    uint32_t ninv = kb_inv((uint32_t)(rows % KB_P));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(rows * cols); i++) mat[i] = kb_mul(mat[i], ninv);
}

void bfo_coset_lde_batch(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out) {
    uint64_t N = rows << added_bits;
    memcpy(out, in, rows * cols * 4);
    memset(out + rows * cols, 0, (N - rows) * cols * 4);
    bfo_idft_batch(out, rows, cols);
    /* coset_dft_batch: coefficient i *= shift^i */
    uint32_t* pw = (uint32_t*)malloc(rows * 4);
    pw[0] = 1;
    for (uint64_t i = 1; i < rows; i++) pw[i] = kb_mul(pw[i - 1], shift);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)rows; i++)
        for (uint64_t c = 0; c < cols; c++) out[i * cols + c] = kb_mul(out[i * cols + c], pw[i]);
    free(pw);
    bfo_dft_batch(out, N, cols);
}

void bfo_coset_lde_batch_bitrev(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out) {
    uint64_t N = rows << added_bits;
    unsigned log_N = bfo_log2(N);
    uint32_t* nat = (uint32_t*)malloc(N * cols * 4);
    bfo_coset_lde_batch(in, rows, cols, added_bits, shift, nat);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)N; i++) memcpy(out + bfo_bitrev((uint64_t)i, log_N) * cols, nat + i * cols, cols * 4);
    free(nat);
}
