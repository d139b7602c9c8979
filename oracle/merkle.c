/* oracle/merkle.c — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * MerkleTreeMmcs<_, _, MyHash, MyCompress, 8> (reference alias
 * crates/stark/src/kb31_poseidon2.rs:27-28; used through Pcs::commit at
 * crates/stark/src/prover.rs:227,334,411 and crates/stark/src/machine.rs:196).
 * Algorithm restated from Plonky3 p3-merkle-tree v0.1.0 @93967fce (un-vendored):
 * `MerkleTree::new`, `first_digest_layer`, `compress_and_inject`, `MerkleTreeMmcs::open_batch`
 * and `verify_batch`.  Heights are powers of two on this path.
 *
 * PARITY UNPINNED (no reference golden vectors; see bf_oracle.h).
 */
#include "bf_oracle.h"
#include "kb31.h"
#include <stdlib.h>
#include <string.h>

struct bfo_tree {
    int num_layers;       /* log2(max_height) + 1 */
    uint64_t* layer_len;  /* [num_layers] */
    uint32_t** layers;    /* [num_layers][layer_len*8] */
    int n_mats;
    uint64_t* mat_rows;   /* input order */
    uint64_t max_height;
};

/* streaming overwrite-mode sponge (same result as bfo_sponge_hash over the concatenation) */
typedef struct { uint32_t st[16]; int fill; } sponge_t;
static inline void sponge_init(sponge_t* s) { memset(s, 0, sizeof *s); }
static inline void sponge_absorb(sponge_t* s, const uint32_t* in, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        s->st[s->fill++] = in[i];
        if (s->fill == 8) { bfo_poseidon2_permute(s->st); s->fill = 0; }
    }
}
static inline void sponge_finish(sponge_t* s, uint32_t out[8]) {
    if (s->fill) bfo_poseidon2_permute(s->st);
    memcpy(out, s->st, 32);
}

/* hash row `r` of every matrix in idx[0..k) (already in sorted = stable input order) */
static void hash_rows(const bfo_mat* mats, const int* idx, int k, uint64_t r, uint32_t out[8]) {
    sponge_t s;
    sponge_init(&s);
    for (int j = 0; j < k; j++) {
        const bfo_mat* m = &mats[idx[j]];
        sponge_absorb(&s, m->data + r * m->cols, m->cols);
    }
    sponge_finish(&s, out);
}

bfo_tree* bfo_mmcs_commit(const bfo_mat* mats, int n, uint32_t root[8]) {
    /* stable sort of matrix indices by height, tallest first */
    int* order = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) order[i] = i;
    for (int i = 1; i < n; i++) { /* insertion sort = stable */
        int v = order[i], j = i - 1;
        while (j >= 0 && mats[order[j]].rows < mats[v].rows) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    uint64_t max_h = mats[order[0]].rows;
    unsigned log_max = bfo_log2(max_h);
    bfo_tree* t = (bfo_tree*)calloc(1, sizeof *t);
    t->num_layers = (int)log_max + 1;
    t->layer_len = (uint64_t*)calloc((size_t)t->num_layers, sizeof(uint64_t));
    t->layers = (uint32_t**)calloc((size_t)t->num_layers, sizeof(uint32_t*));
    t->n_mats = n;
    t->mat_rows = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    for (int i = 0; i < n; i++) t->mat_rows[i] = mats[i].rows;
    t->max_height = max_h;

    int pos = 0; /* cursor into `order` */
    int k = 0;
    while (pos + k < n && mats[order[pos + k]].rows == max_h) k++;
    t->layer_len[0] = max_h;
    t->layers[0] = (uint32_t*)malloc(max_h * 32);
    {
        uint32_t* L = t->layers[0];
        const int* idx = order + pos;
#pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < (int64_t)max_h; r++) hash_rows(mats, idx, k, (uint64_t)r, L + 8 * r);
    }
    pos += k;
    for (int l = 1; l < t->num_layers; l++) {
        uint64_t len = t->layer_len[l - 1] / 2;
        t->layer_len[l] = len;
        t->layers[l] = (uint32_t*)malloc(len * 32);
        const uint32_t* prev = t->layers[l - 1];
        uint32_t* L = t->layers[l];
        k = 0;
        while (pos + k < n && mats[order[pos + k]].rows == len) k++;
        const int* idx = order + pos;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)len; i++) {
            uint32_t d[8];
            bfo_compress(prev + 16 * i, prev + 16 * i + 8, d);
            if (k > 0) { /* inject the rows of the matrices whose height equals this layer's length */
                uint32_t h[8];
                hash_rows(mats, idx, k, (uint64_t)i, h);
                bfo_compress(d, h, L + 8 * i);
            } else {
                memcpy(L + 8 * i, d, 32);
            }
        }
        pos += k;
    }
    memcpy(root, t->layers[t->num_layers - 1], 32);
    free(order);
    return t;
}

void bfo_tree_free(bfo_tree* t) {
    if (!t) return;
    for (int l = 0; l < t->num_layers; l++) free(t->layers[l]);
    free(t->layers);
    free(t->layer_len);
    free(t->mat_rows);
    free(t);
}
int bfo_tree_num_layers(const bfo_tree* t) { return t->num_layers; }
uint64_t bfo_tree_layer_len(const bfo_tree* t, int layer) { return t->layer_len[layer]; }
const uint32_t* bfo_tree_layer(const bfo_tree* t, int layer) { return t->layers[layer]; }

void bfo_mmcs_open_batch(const bfo_tree* t, const bfo_mat* mats, int n, uint64_t index, uint32_t* opened_rows, uint32_t* siblings) {
    unsigned log_max = (unsigned)t->num_layers - 1;
    uint32_t* o = opened_rows;
    for (int i = 0; i < n; i++) {
        unsigned lh = bfo_log2(mats[i].rows);
        uint64_t r = index >> (log_max - lh);
        memcpy(o, mats[i].data + r * mats[i].cols, mats[i].cols * 4);
        o += mats[i].cols;
    }
    for (unsigned l = 0; l < log_max; l++) memcpy(siblings + 8 * l, t->layers[l] + 8 * ((index >> l) ^ 1), 32);
}

int bfo_mmcs_verify_batch(const uint32_t root[8], const uint64_t* rows, const uint64_t* cols, int n, uint64_t index,
                          const uint32_t* opened_rows, const uint32_t* siblings) {
    /* group matrices by height, tallest first, stable */
    int* order = (int*)malloc(sizeof(int) * (size_t)n);
    uint64_t* off = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    uint64_t acc = 0;
    for (int i = 0; i < n; i++) { order[i] = i; off[i] = acc; acc += cols[i]; }
    for (int i = 1; i < n; i++) {
        int v = order[i], j = i - 1;
        while (j >= 0 && rows[order[j]] < rows[v]) { order[j + 1] = order[j]; j--; }
        order[j + 1] = v;
    }
    uint64_t cur_h = rows[order[0]];
    unsigned log_max = bfo_log2(cur_h);
    int pos = 0;
    uint32_t d[8];
    {
        sponge_t s;
        sponge_init(&s);
        while (pos < n && rows[order[pos]] == cur_h) { sponge_absorb(&s, opened_rows + off[order[pos]], cols[order[pos]]); pos++; }
        sponge_finish(&s, d);
    }
    for (unsigned l = 0; l < log_max; l++) {
        const uint32_t* sib = siblings + 8 * l;
        uint32_t nd[8];
        if ((index >> l) & 1) bfo_compress(sib, d, nd); else bfo_compress(d, sib, nd);
        memcpy(d, nd, 32);
        cur_h >>= 1;
        if (pos < n && rows[order[pos]] == cur_h) {
            sponge_t s;
            sponge_init(&s);
            while (pos < n && rows[order[pos]] == cur_h) { sponge_absorb(&s, opened_rows + off[order[pos]], cols[order[pos]]); pos++; }
            uint32_t h[8];
            sponge_finish(&s, h);
            bfo_compress(d, h, nd);
            memcpy(d, nd, 32);
        }
    }
    free(order);
    free(off);
    return memcmp(d, root, 32) == 0 ? 0 : -1;
}
