// dist_commit.cuh — ONE `Pcs::commit` spread over several GPUs, one process per GPU.
//
// The reference commits on one host (`TwoAdicFriPcs::commit`, call sites crates/stark/src/prover.rs:227,334,411):
// per matrix a coset LDE (column-wise independent) followed by `MerkleTreeMmcs::commit`, whose leaves hash whole
// ROWS.  So the path shards as: columns -> LDE -> [exchange] -> rows -> leaf hash + subtree -> cap -> top of tree.
//
//   rank s owns columns  [W*s/G, W*(s+1)/G)          of every matrix (col_range below)
//   rank g owns LDE rows [g*h/G, (g+1)*h/G)  (stored, i.e. bit-reversed, order)  of every matrix of LDE height h
//
// Stored rows are Merkle leaves in order, so rank g's rows are exactly subtree g of the global tree at depth
// log2(G): it builds that subtree with the single-GPU tree code (shorter matrices inject inside the subtree, which
// needs h_i >= G), the G cap digests (32 B each) are all-gathered by the host plumbing and the top log2(G) levels
// are hashed redundantly on every rank's host.
//
// The exchange is fused into the producer: as soon as a block of columns leaves the LDE, k_scatter_rows stores it
// with plain st.global into the receive matrices of the G peers (CUDA IPC mappings of peer HBM over
// NVLink/NVSwitch), on the copy stream, while the next block of columns is in the NTT.  Destinations are rotated by
// rank so the G senders never converge on one receiver.  The same kernel can instead pack into a local send buffer
// for a library all-to-all (`bfgpu_dist_commit_set_staging`), which is the comparison baseline in bench.py.
#pragma once
#include <array>
#include <cstring>
#include <memory>

namespace distk {

constexpr int MAX_WORLD = 16;
struct ScatterArgs {
    const uint32_t* src;      // local LDE columns, column-major, `rows` rows each
    uint64_t rows;            // LDE height h
    uint32_t log_rpg;         // log2(h / world)
    uint32_t world, rank;
    uint32_t dcol0;           // column index of local column 0 inside the destination matrices
    uint32_t* dst[MAX_WORLD];  // destination matrix of every rank: column-major, h/world rows
};
// thread -> VEC consecutive rows of one column;  blockIdx.y = local column
template <int VEC>
__global__ void __launch_bounds__(256) k_scatter_rows(ScatterArgs A) {
    const uint64_t per_dest = (1ull << A.log_rpg) / VEC;  // vectors per destination
    const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (v >= per_dest * A.world) return;
    const uint32_t sweep = (uint32_t)(v / per_dest);
    const uint32_t g = (sweep + A.rank + 1) % A.world;  // every sender starts at a different receiver
    const uint64_t lr = (v % per_dest) * VEC;           // row inside the destination's range
    const uint32_t c = blockIdx.y;
    const uint32_t* s = A.src + (uint64_t)c * A.rows + ((uint64_t)g << A.log_rpg) + lr;
    uint32_t* d = A.dst[g] + ((uint64_t)(A.dcol0 + c) << A.log_rpg) + lr;
    if (VEC == 4) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
    else *d = *s;
}

}  // namespace distk

struct bfgpu_dist_commit {
    bfgpu_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1;
    struct Mat {
        uint64_t rows = 0, lde_rows = 0, rpg = 0;  // trace height, LDE height, LDE rows per rank
        uint32_t total_cols = 0, col0 = 0, ncols = 0;
        uint64_t recv_off = 0;  // word offset of this matrix's row shard (column-major rpg x total_cols) in `recv`
        uint64_t send_off = 0;  // word offset inside one destination block of the staging buffer
    };
    std::vector<Mat> mats;
    uint32_t* recv = nullptr;
    uint64_t recv_words = 0, block_words = 0;
    std::vector<uint32_t*> peer_recv;  // P2P mode
    uint32_t* staging = nullptr;       // staged mode (caller-owned device buffer, world * block_words words)
    bool lde_done = false;
    bfgpu_tree* tree = nullptr;
    std::vector<std::vector<uint32_t>> top;  // top[0] = caps (world digests) ... top.back() = root; Montgomery
};

// Columns of matrix `mat_index` that rank r extends: the even split [W*q/G, W*(q+1)/G) taken at position q = (r + mat_index) mod G, so
// that commitments of many NARROW matrices (the sixteen 4-column quotient chunks over 8 ranks) spread over all ranks instead of
// landing on the same few.
static inline void dist_col_range(uint32_t total, uint32_t world, uint32_t r, uint32_t mat_index, uint32_t* c0, uint32_t* n) {
    const uint32_t q = (r + mat_index) % world;
    uint32_t a = (uint32_t)((uint64_t)total * q / world), b = (uint32_t)((uint64_t)total * (q + 1) / world);
    *c0 = a;
    *n = b - a;
}
// rank owning column c of matrix mat_index, and that rank's first column
static inline uint32_t dist_col_owner(uint32_t total, uint32_t world, uint32_t mat_index, uint32_t c, uint32_t* owner_col0) {
    for (uint32_t r = 0; r < world; r++) {
        uint32_t c0, n;
        dist_col_range(total, world, r, mat_index, &c0, &n);
        if (c >= c0 && c < c0 + n) {
            if (owner_col0) *owner_col0 = c0;
            return r;
        }
    }
    return 0;
}

extern "C" int32_t bfgpu_dist_commit_begin(bfgpu_ctx* ctx, uint32_t rank, uint32_t world, const uint64_t* rows, const uint32_t* total_cols, int32_t n,
                                           bfgpu_dist_commit** out) {
    if (!ctx || !rows || !total_cols || n <= 0 || !out) return fail(ctx, BFGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!is_pow2(world) || world > (uint32_t)distk::MAX_WORLD || rank >= world)
        return fail(ctx, BFGPU_ERR_INVALID, "world size %u must be a power of two <= %d and rank %u below it", world, distk::MAX_WORLD, rank);
    std::unique_ptr<bfgpu_dist_commit> dc(new bfgpu_dist_commit());
    dc->ctx = ctx;
    dc->rank = rank;
    dc->world = world;
    dc->mats.resize(n);
    for (int i = 0; i < n; i++) {
        auto& m = dc->mats[i];
        if (!is_pow2(rows[i]) || total_cols[i] == 0) return fail(ctx, BFGPU_ERR_INVALID, "matrix %d: height %llu must be a power of two and width non-zero", i, (unsigned long long)rows[i]);
        if (ilog2(rows[i]) + ctx->log_blowup > kb::TWO_ADICITY) return fail(ctx, BFGPU_ERR_INVALID, "LDE height exceeds the field's two-adicity");
        m.rows = rows[i];
        m.lde_rows = rows[i] << ctx->log_blowup;
        if (m.lde_rows < world) return fail(ctx, BFGPU_ERR_INVALID, "matrix %d: LDE height %llu is below the world size %u", i, (unsigned long long)m.lde_rows, world);
        m.rpg = m.lde_rows / world;
        m.total_cols = total_cols[i];
        dist_col_range(m.total_cols, world, rank, (uint32_t)i, &m.col0, &m.ncols);
        m.recv_off = dc->recv_words;
        dc->recv_words += m.rpg * m.total_cols;
        m.send_off = dc->block_words;
        dc->block_words += m.rpg * m.ncols;
    }
    // the receive matrix is IPC-exported to the peers: it lives in the context's exported pool, never in the trimmable block cache
    TRY(dalloc_export(ctx, (void**)&dc->recv, dc->recv_words * 4));
    *out = dc.release();
    return BFGPU_OK;
}

extern "C" uint32_t bfgpu_dist_commit_local_cols(const bfgpu_dist_commit* dc, int32_t i, uint32_t* col0) {
    if (!dc || i < 0 || (size_t)i >= dc->mats.size()) return 0;
    if (col0) *col0 = dc->mats[i].col0;
    return dc->mats[i].ncols;
}

// words rank `src` sends to every peer (= what this rank receives from `src`) in staged mode
extern "C" uint64_t bfgpu_dist_commit_block_words(const bfgpu_dist_commit* dc, uint32_t src) {
    if (!dc || src >= dc->world) return 0;
    uint64_t w = 0;
    for (size_t i = 0; i < dc->mats.size(); i++) {
        const auto& m = dc->mats[i];
        uint32_t c0, nc;
        dist_col_range(m.total_cols, dc->world, src, (uint32_t)i, &c0, &nc);
        w += m.rpg * nc;
    }
    return w;
}

extern "C" int32_t bfgpu_dist_commit_recv_handle(bfgpu_dist_commit* dc, uint8_t handle[64]) {
    if (!dc || !handle) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dc->recv));
    memcpy(handle, &h, 64);
    return BFGPU_OK;
}

// P2P mode: map the receive buffers of all ranks (handles[r] from rank r's bfgpu_dist_commit_recv_handle).
// Mappings are cached on the context: the block cache hands identical commits identical buffers.
extern "C" int32_t bfgpu_dist_commit_set_peers(bfgpu_dist_commit* dc, const uint8_t* handles) {
    if (!dc || !handles) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    if (dc->lde_done) return fail(ctx, BFGPU_ERR_STATE, "peers must be set before the LDE");
    dc->peer_recv.assign(dc->world, nullptr);
    for (uint32_t r = 0; r < dc->world; r++) {
        if (r == dc->rank) {
            dc->peer_recv[r] = dc->recv;
            continue;
        }
        std::array<uint8_t, 64> key;
        memcpy(key.data(), handles + 64 * r, 64);
        auto it = ctx->ipc_open.find(key);
        if (it == ctx->ipc_open.end()) {
            cudaIpcMemHandle_t h;
            memcpy(&h, key.data(), 64);
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            it = ctx->ipc_open.emplace(key, p).first;
        }
        dc->peer_recv[r] = (uint32_t*)it->second;
    }
    dc->staging = nullptr;
    return BFGPU_OK;
}

// staged mode: the scatter packs into a local device buffer laid out [destination][matrix][local column][row]
// (world * block_words(rank) words); the caller runs the all-to-all and hands the result to _unpack.
extern "C" int32_t bfgpu_dist_commit_set_staging(bfgpu_dist_commit* dc, uint32_t* dev_send) {
    if (!dc || !dev_send) return BFGPU_ERR_INVALID;
    if (dc->lde_done) return fail(dc->ctx, BFGPU_ERR_STATE, "the staging buffer must be set before the LDE");
    dc->staging = dev_send;
    dc->peer_recv.clear();
    return BFGPU_OK;
}

static int32_t dist_scatter(bfgpu_dist_commit* dc, const bfgpu_dist_commit::Mat& m, const uint32_t* src, uint32_t c_first, uint32_t nc, cudaStream_t st) {
    bfgpu_ctx* ctx = dc->ctx;
    distk::ScatterArgs A;
    memset(&A, 0, sizeof A);
    A.src = src;
    A.rows = m.lde_rows;
    A.log_rpg = ilog2(m.rpg);
    A.world = dc->world;
    A.rank = dc->rank;
    for (uint32_t g = 0; g < dc->world; g++) {
        if (dc->staging) A.dst[g] = dc->staging + (uint64_t)g * dc->block_words + m.send_off;
        else A.dst[g] = dc->peer_recv[g] + m.recv_off;
    }
    A.dcol0 = (dc->staging ? 0 : m.col0) + c_first;
    bool vec = m.rpg % 4 == 0 && ((uintptr_t)src & 15) == 0;
    for (uint32_t g = 0; g < dc->world; g++) vec = vec && ((uintptr_t)A.dst[g] & 15) == 0;
    if (vec) {
        uint64_t vecs = m.lde_rows / 4;
        dim3 grid((unsigned)((vecs + 255) / 256), nc);
        distk::k_scatter_rows<4><<<grid, 256, 0, st>>>(A);
    } else {
        dim3 grid((unsigned)((m.lde_rows + 255) / 256), nc);
        distk::k_scatter_rows<1><<<grid, 256, 0, st>>>(A);
    }
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    return BFGPU_OK;
}

constexpr uint32_t DIST_CHUNK_COLS = 64;
// LDE + exchange of this rank's columns of every matrix, from coefficient-side inputs already on the device in the prover's layout
// (column-major, Montgomery, bit-reversed rows; coefs[i] = the rank's ncols(i) columns, TRANSFORMED IN PLACE: pass a copy to keep them).
// shift_mont[i]: LDE coset shift of matrix i.  Asynchronous, same completion protocol as bfgpu_dist_commit_lde.
// first_pass_done[i]: the first inverse NTT pass of matrix i was executed by the fused ingest (ntt3::k_ingest_pass)
static int32_t dist_lde_coefs(bfgpu_dist_commit* dc, const std::vector<DMat>& coefs, const std::vector<uint32_t>& shift_mont,
                              const std::vector<char>* first_pass_done = nullptr) {
    bfgpu_ctx* ctx = dc->ctx;
    if (dc->lde_done) return fail(ctx, BFGPU_ERR_STATE, "LDE already done");
    if (!dc->staging && dc->peer_recv.empty()) return fail(ctx, BFGPU_ERR_STATE, "neither peers nor a staging buffer set");
    struct Pending { uint32_t* buf; cudaEvent_t done; };
    std::vector<Pending> pending;
    auto retire = [&](size_t keep) -> int32_t {  // compute stream waits for old scatters, then their blocks return to the cache
        while (pending.size() > keep) {
            CU(cudaStreamWaitEvent(ctx->stream, pending.front().done, 0));
            cudaEventDestroy(pending.front().done);
            dfree(ctx, pending.front().buf);
            pending.erase(pending.begin());
        }
        return BFGPU_OK;
    };
    int32_t rc = BFGPU_OK;
    for (size_t i = 0; i < dc->mats.size() && rc == BFGPU_OK; i++) {
        const auto& m = dc->mats[i];
        if (m.ncols == 0) continue;
        const DMat& coef = coefs[i];
        if (coef.rows != m.rows || coef.cols != m.ncols || !coef.d) {
            rc = fail(ctx, BFGPU_ERR_INVALID, "matrix %zu: expected this rank's %llu x %u column slice", i, (unsigned long long)m.rows, m.ncols);
            break;
        }
        // at least ~4 blocks per matrix so that only the last quarter of the exchange is exposed (16..64 columns each)
        const uint32_t chunk = std::min(DIST_CHUNK_COLS, std::max(ctx->dist_min_chunk, ((m.ncols + 3) / 4 + 7) / 8 * 8));
        for (uint32_t c = 0; c < m.ncols && rc == BFGPU_OK; c += chunk) {
            uint32_t nc = std::min(chunk, m.ncols - c);
            if ((rc = retire(1)) != BFGPU_OK) break;  // at most two LDE blocks alive: one in the NTT, one being scattered
            DMat slice = coef, lde;
            slice.d = coef.d + (uint64_t)c * coef.rows;
            slice.cols = nc;
            // P2P mode with row shards of at least one run of the last pass: that pass sends its results to their owners itself
            bool deferred = false;
            const bool try_fused = !dc->staging && ctx->log_blowup == 1;
            if ((rc = lde_from_bitrev(ctx, slice, ctx->log_blowup, shift_mont[i], &lde, /*consume=*/false, nullptr,
                                      first_pass_done && (*first_pass_done)[i], try_fused ? &deferred : nullptr)) != BFGPU_OK)
                break;
            Phase ph(ctx, BFGPU_PHASE_EXCHANGE);
            cudaEvent_t ready, done;
            cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
            cudaEventRecord(ready, ctx->stream);
            cudaStreamWaitEvent(ctx->copy_stream, ready, 0);
            cudaEventDestroy(ready);
            if (deferred) {
                ScatterTarget sc;
                sc.log_rpg = ilog2(m.rpg);
                sc.dcol0 = m.col0 + c;
                sc.world = dc->world;
                bool ok = sc.log_rpg >= 8;  // >= the longest run (2^8 rows) of a contiguous pass
                for (uint32_t g = 0; g < dc->world; g++) {
                    sc.dst[g] = dc->peer_recv[g] + m.recv_off;
                    ok = ok && ((uintptr_t)sc.dst[g] & 15) == 0;
                }
                if (ok) rc = lde_last_pass_scatter(ctx, lde, sc, ctx->copy_stream);
                else {  // finish the block in place, then the plain scatter
                    const std::vector<bfgpu_ctx::NttPass>* fplan = nullptr;
                    rc = get_plan(ctx, ilog2(lde.rows) - 1, false, &fplan);
                    if (rc == BFGPU_OK) rc = run_cfwd(ctx, lde.d, lde.rows / 2, 2 * lde.cols, ilog2(lde.rows) - 1, fplan->front(), nullptr, ctx->copy_stream);
                    if (rc == BFGPU_OK) rc = dist_scatter(dc, m, lde.d, c, nc, ctx->copy_stream);
                }
            } else {
                rc = dist_scatter(dc, m, lde.d, c, nc, ctx->copy_stream);
            }
            cudaEventRecord(done, ctx->copy_stream);
            pending.push_back({lde.d, done});
        }
    }
    int32_t rc2 = retire(0);
    if (rc != BFGPU_OK) return rc;
    TRY(rc2);
    dc->lde_done = true;
    return BFGPU_OK;
}

// Coset LDE of this rank's columns of every matrix (local[i]: rows[i] x local_cols(i), row-major, in the context's
// input space; data may be null where the rank owns no column) and the exchange.  Asynchronous: returns once
// everything is enqueued.  Before bfgpu_dist_commit_finish every rank must have synchronised its context and the
// ranks must have passed a barrier (P2P mode) or completed the all-to-all + _unpack (staged mode).
static int32_t dist_commit_lde_impl(bfgpu_dist_commit* dc, const bfgpu_mat* local, const uint32_t* domain_shifts);
extern "C" int32_t bfgpu_dist_commit_lde(bfgpu_dist_commit* dc, const bfgpu_mat* local, const uint32_t* domain_shifts) {
    AllocScope scope(dc ? dc->ctx : nullptr);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(dist_commit_lde_impl(dc, local, domain_shifts));
}
static int32_t dist_commit_lde_impl(bfgpu_dist_commit* dc, const bfgpu_mat* local, const uint32_t* domain_shifts) {
    if (!dc || !local) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    if (dc->lde_done) return fail(ctx, BFGPU_ERR_STATE, "LDE already done");
    if (!dc->staging && dc->peer_recv.empty()) return fail(ctx, BFGPU_ERR_STATE, "neither peers nor a staging buffer set");
    const uint32_t gen = kb::to_mont(kb::GEN);
    std::vector<DMat> coefs(dc->mats.size());
    std::vector<uint32_t> shifts(dc->mats.size(), gen);
    std::vector<char> first_done(dc->mats.size(), 0);
    int32_t rc = BFGPU_OK;
    for (size_t i = 0; i < dc->mats.size() && rc == BFGPU_OK; i++) {
        const auto& m = dc->mats[i];
        if (m.ncols == 0) continue;
        if (local[i].rows != m.rows || local[i].cols != m.ncols) {
            rc = fail(ctx, BFGPU_ERR_INVALID, "matrix %zu: expected this rank's %llu x %u column slice, got %llu x %llu", i, (unsigned long long)m.rows, m.ncols,
                      (unsigned long long)local[i].rows, (unsigned long long)local[i].cols);
            break;
        }
        if ((rc = check_mat(ctx, &local[i], true)) != BFGPU_OK) break;
        if (domain_shifts) {
            uint32_t ds = ctx->repr == BFGPU_REPR_CANONICAL ? kb::to_mont(domain_shifts[i] % kb::P) : domain_shifts[i];
            if (ds == 0) {
                rc = fail(ctx, BFGPU_ERR_INVALID, "zero domain shift");
                break;
            }
            shifts[i] = kb::mul(gen, kb::inv(ds));
        }
        bool fd = false;
        rc = ingest(ctx, local[i], /*bitrev=*/true, &coefs[i], &fd);  // transpose + first inverse pass in one kernel where the shape allows
        first_done[i] = fd;
    }
    if (rc == BFGPU_OK) rc = dist_lde_coefs(dc, coefs, shifts, &first_done);
    // the coefficient matrices were transformed in place slice by slice; all work on them is in compute-stream order
    for (DMat& c : coefs) dfree(ctx, c.d);
    return rc;
}

// staged mode: dev_recv = [source rank][matrix][source's columns][own rows]  ->  the receive matrices
static int32_t dist_commit_unpack_impl(bfgpu_dist_commit* dc, const uint32_t* dev_recv);
extern "C" int32_t bfgpu_dist_commit_unpack(bfgpu_dist_commit* dc, const uint32_t* dev_recv) {
    AllocScope scope(dc ? dc->ctx : nullptr);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(dist_commit_unpack_impl(dc, dev_recv));
}
static int32_t dist_commit_unpack_impl(bfgpu_dist_commit* dc, const uint32_t* dev_recv) {
    if (!dc || !dev_recv) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    Phase ph(ctx, BFGPU_PHASE_EXCHANGE);
    uint64_t off = 0;
    for (uint32_t s = 0; s < dc->world; s++)
        for (size_t i = 0; i < dc->mats.size(); i++) {
            const auto& m = dc->mats[i];
            uint32_t c0, nc;
            dist_col_range(m.total_cols, dc->world, s, (uint32_t)i, &c0, &nc);
            if (!nc) continue;
            CU(cudaMemcpyAsync(dc->recv + m.recv_off + (uint64_t)c0 * m.rpg, dev_recv + off, m.rpg * nc * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            off += m.rpg * nc;
        }
    return BFGPU_OK;
}

// leaf hashes + subtree over this rank's rows; cap = root of the subtree
static int32_t dist_commit_finish_impl(bfgpu_dist_commit* dc, uint32_t cap[8]);
extern "C" int32_t bfgpu_dist_commit_finish(bfgpu_dist_commit* dc, uint32_t cap[8]) {
    AllocScope scope(dc ? dc->ctx : nullptr);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(dist_commit_finish_impl(dc, cap));
}
static int32_t dist_commit_finish_impl(bfgpu_dist_commit* dc, uint32_t cap[8]) {
    if (!dc || !cap) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    if (!dc->lde_done || dc->tree) return fail(ctx, BFGPU_ERR_STATE, "finish needs a completed LDE/exchange and runs once");
    std::vector<DMat> shards(dc->mats.size());
    for (size_t i = 0; i < dc->mats.size(); i++) {
        shards[i].d = dc->recv + dc->mats[i].recv_off;
        shards[i].rows = dc->mats[i].rpg;
        shards[i].cols = dc->mats[i].total_cols;
    }
    int32_t rc = build_tree(ctx, std::move(shards), false, &dc->tree);
    if (rc == BFGPU_OK) rc = read_digest(ctx, dc->tree->layers.back(), cap);
    if (rc != BFGPU_OK) {  // no half-built tree stays behind in the handle
        tree_release(dc->tree);
        dc->tree = nullptr;
    }
    return rc;
}

// top of the tree from the all-gathered caps (caps[r] = rank r's cap, caller representation)
extern "C" int32_t bfgpu_dist_commit_root(bfgpu_dist_commit* dc, const uint32_t* caps, uint32_t root[8]) {
    if (!dc || !caps || !root) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    dc->top.clear();
    std::vector<uint32_t> layer(caps, caps + 8 * dc->world);
    if (ctx->repr == BFGPU_REPR_CANONICAL)
        for (auto& v : layer) v = kb::to_mont(v % kb::P);
    dc->top.push_back(layer);
    while (dc->top.back().size() > 8) {
        const auto& prev = dc->top.back();
        std::vector<uint32_t> next(prev.size() / 2);
        for (size_t k = 0; k < next.size() / 8; k++) {
            uint32_t s[16];
            memcpy(s, &prev[16 * k], 64);
            host_p2::permute(s);  // TruncatedPermutation: first 8 words of permute(left || right)
            memcpy(&next[8 * k], s, 32);
        }
        dc->top.push_back(std::move(next));
    }
    for (int k = 0; k < 8; k++) root[k] = ctx->repr == BFGPU_REPR_CANONICAL ? kb::from_mont(dc->top.back()[k]) : dc->top.back()[k];
    return BFGPU_OK;
}

// Mmcs::open_batch for a GLOBAL leaf index owned by this rank (index / (max LDE height / world) == rank):
// rows of every matrix + log2(max height) siblings, leaf level first.
static int32_t dist_commit_open_batch_impl(bfgpu_dist_commit* dc, uint64_t index, uint32_t* opened_rows, uint32_t* siblings);
extern "C" int32_t bfgpu_dist_commit_open_batch(bfgpu_dist_commit* dc, uint64_t index, uint32_t* opened_rows, uint32_t* siblings) {
    AllocScope scope(dc ? dc->ctx : nullptr);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(dist_commit_open_batch_impl(dc, index, opened_rows, siblings));
}
static int32_t dist_commit_open_batch_impl(bfgpu_dist_commit* dc, uint64_t index, uint32_t* opened_rows, uint32_t* siblings) {
    if (!dc) return BFGPU_ERR_INVALID;
    bfgpu_ctx* ctx = dc->ctx;
    if (!dc->tree || dc->top.empty()) return fail(ctx, BFGPU_ERR_STATE, "commit not finished");
    uint64_t per = 1ull << dc->tree->log_max;
    if (index / per != dc->rank) return fail(ctx, BFGPU_ERR_INVALID, "leaf %llu belongs to rank %llu", (unsigned long long)index, (unsigned long long)(index / per));
    TRY(bfgpu_mmcs_open_batch(dc->tree, index % per, opened_rows, siblings));
    uint64_t pos = dc->rank;
    for (size_t l = 0; l + 1 < dc->top.size(); l++, pos >>= 1)
        for (int k = 0; k < 8; k++) {
            uint32_t v = dc->top[l][8 * (pos ^ 1) + k];
            siblings[8 * (dc->tree->log_max + l) + k] = ctx->repr == BFGPU_REPR_CANONICAL ? kb::from_mont(v) : v;
        }
    return BFGPU_OK;
}

extern "C" uint64_t bfgpu_dist_commit_rows_per_rank(const bfgpu_dist_commit* dc) { return (dc && dc->tree) ? (1ull << dc->tree->log_max) : 0; }

extern "C" void bfgpu_dist_commit_free(bfgpu_dist_commit* dc) {
    if (!dc) return;
    tree_release(dc->tree);
    dfree_export(dc->ctx, dc->recv);
    delete dc;
}
