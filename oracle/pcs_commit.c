/* oracle/pcs_commit.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * `TwoAdicFriPcs::commit` as the reference calls it (crates/stark/src/prover.rs:227,334,411,
 * crates/stark/src/machine.rs:196): for every (domain, evaluations) pair compute the coset LDE
 * with shift GENERATOR/domain.shift and blowup 2^log_blowup, store it with bit-reversed rows,
 * then MerkleTreeMmcs::commit over all LDE matrices.  Restated from Plonky3 p3-fri v0.1.0
 * @93967fce `TwoAdicFriPcs::commit` (un-vendored).
 *
 * PARITY UNPINNED (no reference golden vectors; see bf_oracle.h).
 */
#include "bf_oracle.h"
#include "kb31.h"
#include <stdlib.h>
#include <omp.h>

struct bfo_pcs_data {
    int n;
    bfo_mat* ldes; /* bit-reversed rows, owned */
    bfo_tree* tree;
};

bfo_pcs_data* bfo_pcs_commit(const bfo_mat* evals, const uint32_t* domain_shift, int n, unsigned log_blowup, uint32_t root[8]) {
    bfo_pcs_data* d = (bfo_pcs_data*)calloc(1, sizeof *d);
    d->n = n;
    d->ldes = (bfo_mat*)calloc((size_t)n, sizeof(bfo_mat));
    for (int i = 0; i < n; i++) {
        uint64_t N = evals[i].rows << log_blowup;
        uint32_t* out = (uint32_t*)malloc(N * evals[i].cols * 4 + 4);
        uint32_t shift = kb_mul(KB_GENERATOR, kb_inv(domain_shift ? domain_shift[i] : 1u));
        bfo_coset_lde_batch_bitrev(evals[i].data, evals[i].rows, evals[i].cols, log_blowup, shift, out);
        d->ldes[i].data = out;
        d->ldes[i].rows = N;
        d->ldes[i].cols = evals[i].cols;
    }
    d->tree = bfo_mmcs_commit(d->ldes, n, root);
    return d;
}

void bfo_pcs_data_free(bfo_pcs_data* d) {
    if (!d) return;
    for (int i = 0; i < d->n; i++) free((void*)d->ldes[i].data);
    free(d->ldes);
    bfo_tree_free(d->tree);
    free(d);
}
int bfo_pcs_num_mats(const bfo_pcs_data* d) { return d->n; }
const uint32_t* bfo_pcs_lde(const bfo_pcs_data* d, int i, uint64_t* rows, uint64_t* cols) {
    if (rows) *rows = d->ldes[i].rows;
    if (cols) *cols = d->ldes[i].cols;
    return d->ldes[i].data;
}
const bfo_tree* bfo_pcs_tree(const bfo_pcs_data* d) { return d->tree; }

void bfo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int bfo_get_threads(void) { return omp_get_max_threads(); }
