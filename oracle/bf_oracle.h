/* oracle/bf_oracle.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Public surface of the CPU restatement of the reference's proving hot path
 * (reference crates/stark/src/prover.rs:209-553 and the Plonky3 routines it calls).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or load this library.  All field elements are CANONICAL u32.
 *
 * PARITY UNPINNED at the Plonky3 boundary: no golden vectors exist in the reference
 * and the dependency source is not available offline.
 */
#ifndef BF_ORACLE_H
#define BF_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- Poseidon2 (width 16, x^3, 8 full + 13 partial rounds) ------------------------- */
void bfo_poseidon2_constants(uint32_t ext_initial[4][16], uint32_t internal[13], uint32_t ext_terminal[4][16]);
void bfo_poseidon2_permute(uint32_t state[16]);
/* PaddingFreeSponge<Perm,16,8,8>::hash_iter */
void bfo_sponge_hash(const uint32_t* in, uint64_t n, uint32_t out[8]);
/* TruncatedPermutation<Perm,2,8,16>::compress */
void bfo_compress(const uint32_t left[8], const uint32_t right[8], uint32_t out[8]);
/* throughput helper for bench.py: n independent permutations over a buffer of n*16 words */
void bfo_poseidon2_permute_many(uint32_t* states, uint64_t n);

/* ---- MerkleTreeMmcs ------------------------------------------------------------------ */
typedef struct {
    const uint32_t* data; /* row-major rows x cols */
    uint64_t rows;
    uint64_t cols;
} bfo_mat;
typedef struct bfo_tree bfo_tree;
/* commit to n matrices (heights powers of two); returns tree, writes root */
bfo_tree* bfo_mmcs_commit(const bfo_mat* mats, int n, uint32_t root[8]);
void bfo_tree_free(bfo_tree* t);
int bfo_tree_num_layers(const bfo_tree* t);                 /* log2(max_height)+1 */
uint64_t bfo_tree_layer_len(const bfo_tree* t, int layer);   /* digests in that layer */
const uint32_t* bfo_tree_layer(const bfo_tree* t, int layer);/* layer_len*8 words */
/* Mmcs::open_batch: rows are written back-to-back in INPUT matrix order
   (row `index >> (log_max_h - log_h)` of each matrix); siblings = log_max_h digests */
void bfo_mmcs_open_batch(const bfo_tree* t, const bfo_mat* mats, int n, uint64_t index, uint32_t* opened_rows, uint32_t* siblings);
/* Mmcs::verify_batch: dims given as (rows, cols) per matrix; returns 0 on success */
int bfo_mmcs_verify_batch(const uint32_t root[8], const uint64_t* rows, const uint64_t* cols, int n, uint64_t index,
                          const uint32_t* opened_rows, const uint32_t* siblings);

/* ---- two-adic DFT ------------------------------------------------------------------- */
/* naive O(n^2): out[j][c] = sum_i coeff... evaluates the interpolant of column c (evals over the
   order-`rows` subgroup, natural order) at shift*w^j for j < rows<<added_bits, natural order */
void bfo_coset_lde_naive(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out);
/* fast version of the same (TwoAdicSubgroupDft::coset_lde_batch), natural-order output */
void bfo_coset_lde_batch(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out);
/* same, rows of the output permuted by bit reversal (what TwoAdicFriPcs::commit stores) */
void bfo_coset_lde_batch_bitrev(const uint32_t* in, uint64_t rows, uint64_t cols, unsigned added_bits, uint32_t shift, uint32_t* out);
void bfo_dft_batch(uint32_t* mat, uint64_t rows, uint64_t cols);  /* in place, natural in/out */
void bfo_idft_batch(uint32_t* mat, uint64_t rows, uint64_t cols); /* in place, natural in/out */

/* ---- TwoAdicFriPcs::commit --------------------------------------------------------- */
typedef struct bfo_pcs_data bfo_pcs_data;
/* domain_shift[i] is the shift of matrix i's evaluation domain (1 for natural domains);
   LDE shift = GENERATOR / domain_shift, log_blowup = 1 */
bfo_pcs_data* bfo_pcs_commit(const bfo_mat* evals, const uint32_t* domain_shift, int n, unsigned log_blowup, uint32_t root[8]);
void bfo_pcs_data_free(bfo_pcs_data* d);
int bfo_pcs_num_mats(const bfo_pcs_data* d);
const uint32_t* bfo_pcs_lde(const bfo_pcs_data* d, int i, uint64_t* rows, uint64_t* cols); /* bit-reversed rows */
const bfo_tree* bfo_pcs_tree(const bfo_pcs_data* d);

/* ---- tuned CPU baseline (fast_commit.c: AVX-512 Montgomery, packed Poseidon2, cache-blocked row NTT) ------------------ */
int bfo_fast_available(void);
/* Pcs::commit of one rows x cols matrix (natural domain, log_blowup 1): same root as bfo_pcs_commit.  lde_out may be NULL.
   phase_sec = {LDE, leaf hashing, compression layers}.  Returns 0, -1 if unsupported (rows < 16, no AVX-512). */
int bfo_fast_pcs_commit(const uint32_t* in, uint64_t rows, uint64_t cols, uint32_t root[8], uint32_t* lde_out, double phase_sec[3]);
int bfo_fast_permute_many(uint32_t* states, uint64_t n);
/* CPU arm of the FRI commit phase with caller-supplied folding challenges (oracle/fast_commit.c) */
int bfo_fast_fri_commit_phase(const uint32_t* const* inputs, const uint32_t* log_len, int n_inputs, const uint32_t* betas, int rollin_beta2, uint32_t* roots,
                              uint32_t final_poly[4], double sec[2]);
void bfo_fast_release(void); /* drop the work buffers bfo_fast_pcs_commit keeps between calls */

void bfo_set_threads(int n);
int bfo_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
