// dist_prove.cuh — ONE shard proof over several GPUs (BASELINE configs 3 and 5), one process per GPU.
//
// The reference proves a shard on one host: `CpuProver::commit` + `CpuProver::open` (crates/stark/src/prover.rs:209-236,242-553),
// with rayon parallelism per chip for the permutation traces (:280-296) and the quotient values (:356-388) and inside `pcs.open`
// (:460-470).  Here the same proof — word for word the one `bfgpu_machine_open` produces on one GPU — is computed by G ranks:
//
//   commitments (main, permutation, quotient)  column-sharded LDE -> rows stored into the peers' row shards over NVLink -> per-rank
//                                              Merkle subtree -> G caps -> top of the tree (dist_commit.cuh)
//   main traces / LogUp traces                 REPLICATED: every rank generates all of them (a few ms at 2^22 rows; the LogUp running
//                                              sum needs whole rows and all rows), then extends only its columns
//   quotient values                            row-sharded: a rank evaluates the constraints on ITS stored LDE rows, reads the "next"
//                                              row from the peer that holds it (P2P load) and stores each quotient word into the
//                                              coefficient buffer of the rank that owns that COLUMN of the quotient commitment
//   opened values                              row-sharded barycentric partial sums over the whole LDE coset + one small all-gather
//   reduced openings, FRI folds                row-local on the row shards; per round a subtree + cap exchange; at 2^20 elements the
//                                              vector is all-gathered by peer copies and every rank finishes the commit phase alone
//   queries                                    answered by the rank owning the leaf; one all-gather assembles the proof on every rank
//
// Only the data path touches peer HBM.  The control plane (64-byte IPC handles, 32-byte caps, partial sums, proof pieces, barriers)
// goes through the caller's communicator (bfgpu_comm: torch.distributed in the Python mirror, MPI/NCCL in a Rust shim): ~30 small
// host all-gathers and ~8 barriers per proof.  The Fiat-Shamir challenger is replicated: every rank observes the same roots and
// samples the same challenges.
#pragma once

// global length (log2) at which the FRI vector stops being sharded ($BFGPU_DIST_FRI_GATHER_LOG overrides, tests use 13): above it every round costs a subtree + a cap exchange through the
// host communicator, below it one device finishes alone; measured with the threshold at 2^13 the ~10 sharded rounds cost 8.5 ms at 8 ranks

namespace distp {

struct SMat {  // one committed matrix as THIS rank sees it: its row shard (or a window of a replicated matrix)
    const uint32_t* d;   // first local row, column-major
    uint64_t stride;     // words between columns
    uint64_t h;          // global LDE height
    uint64_t rpg;        // local rows = h / world
    uint32_t cols;
    std::vector<kb::Ext> pts;
    std::vector<std::vector<kb::Ext>> ys;  // [point][column]
};
struct SRound {
    std::vector<SMat> mats;
    const bfgpu_tree* tree = nullptr;   // replicated: the full tree; sharded: this rank's subtree
    bool replicated = false;
    const std::vector<std::vector<uint32_t>>* top = nullptr;  // sharded: caps ... root (Montgomery)
    unsigned log_max = 0;               // global log2 of the tallest LDE
};
struct SLayer {  // a FRI layer committed while the vector was still sharded
    uint32_t* vec;       // local slice of the folded input: len_local extension elements
    uint64_t len;        // GLOBAL length
    bfgpu_tree* tree;    // subtree over the local leaves
    std::vector<std::vector<uint32_t>> top;
};

// top of a Merkle tree over `world` caps (Montgomery words): layers caps ... root
static void cap_tree(const uint32_t* caps, uint32_t world, std::vector<std::vector<uint32_t>>* top) {
    top->clear();
    top->push_back(std::vector<uint32_t>(caps, caps + 8 * world));
    while (top->back().size() > 8) {
        const auto& prev = top->back();
        std::vector<uint32_t> next(prev.size() / 2);
        for (size_t k = 0; k < next.size() / 8; k++) {
            uint32_t s[16];
            memcpy(s, &prev[16 * k], 64);
            host_p2::permute(s);
            memcpy(&next[8 * k], s, 32);
        }
        top->push_back(std::move(next));
    }
}

}  // namespace distp

// map the exported buffers of all ranks from their 64-byte IPC handles (handles[r] at stride `stride`); mine = this rank's own pointer
static int32_t dist_map_peers(bfgpu_ctx* ctx, const uint8_t* handles, size_t stride, uint32_t rank, uint32_t world, void* mine, std::vector<uint32_t*>* out) {
    out->assign(world, nullptr);
    for (uint32_t r = 0; r < world; r++) {
        if (r == rank) {
            (*out)[r] = (uint32_t*)mine;
            continue;
        }
        std::array<uint8_t, 64> key;
        memcpy(key.data(), handles + stride * r, 64);
        auto it = ctx->ipc_open.find(key);
        if (it == ctx->ipc_open.end()) {
            cudaIpcMemHandle_t hh;
            memcpy(&hh, key.data(), 64);
            void* p = nullptr;
            CU(cudaIpcOpenMemHandle(&p, hh, cudaIpcMemLazyEnablePeerAccess));
            it = ctx->ipc_open.emplace(key, p).first;
        }
        (*out)[r] = (uint32_t*)it->second;
    }
    return BFGPU_OK;
}

#define COMM(call)                                                                                           \
    do {                                                                                                     \
        if ((call) != 0) return fail(ctx, BFGPU_ERR_STATE, "communicator callback failed: %s (%s:%d)", #call, __FILE__, __LINE__); \
    } while (0)

// every store this rank issued (both streams) has landed, and so has everybody else's
static int32_t dist_sync_barrier(bfgpu_ctx* ctx, const bfgpu_comm* comm) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    COMM(comm->barrier(comm->user));
    return BFGPU_OK;
}

// finish one sharded commitment: barrier, subtree, cap exchange, top tree; root_mont = the commitment (Montgomery)
static int32_t dist_finish_commit(bfgpu_ctx* ctx, const bfgpu_comm* comm, bfgpu_dist_commit* dc, uint32_t root_mont[8]) {
    TRY(dist_sync_barrier(ctx, comm));
    uint32_t cap[8];
    TRY(bfgpu_dist_commit_finish(dc, cap));  // caller representation
    std::vector<uint32_t> caps(8 * (size_t)dc->world);
    COMM(comm->all_gather(comm->user, cap, caps.data(), 32));
    uint32_t root[8];
    TRY(bfgpu_dist_commit_root(dc, caps.data(), root));
    memcpy(root_mont, dc->top.back().data(), 32);
    return BFGPU_OK;
}

// The proof of one shard by `world` ranks.  names / chips / traces: every included chip's main trace, sorted by (height desc, name),
// in the prover's layout, REPLICATED on every rank (consumed: released here).  ch: the replicated transcript, pk already observed.
static int32_t dist_prove_core(bfgpu_ctx* ctx, const bfgpu_comm* comm, uint32_t rank, uint32_t world, const bfgpu_pk* pk, const std::vector<std::string>& names,
                               const std::vector<int>& chips, std::vector<DMat>& traces, bfgpu_challenger& ch, int64_t fixed_pow_witness,
                               std::vector<uint32_t>* proof_out) {
    const size_t nchips = chips.size();
    const uint32_t gen = kb::to_mont(kb::GEN);
    const unsigned log_world = ilog2(world);
    struct TraceGuard {
        bfgpu_ctx* ctx;
        std::vector<DMat>& t;
        ~TraceGuard() {
            for (DMat& m : t) dfree(ctx, m.d);
            t.clear();
        }
    } trace_guard{ctx, traces};
    if (!is_pow2(world) || world > (uint32_t)distk::MAX_WORLD || rank >= world) return fail(ctx, BFGPU_ERR_INVALID, "bad rank / world size");
    if (ctx->log_blowup != 1) return fail(ctx, BFGPU_ERR_INVALID, "the machine prover requires log_blowup = 1 (kb31_poseidon2.rs:63)");
    std::vector<int> pk_idx(nchips, -1);
    for (size_t i = 0; i < nchips; i++)
        for (size_t k = 0; k < pk->names.size(); k++)
            if (pk->names[k] == names[i]) pk_idx[i] = (int)k;
    for (size_t i = 0; i < nchips; i++) {
        const air::ChipInfo& ci = air::CHIPS[chips[i]];
        if (ci.log_quotient_degree != 1) return fail(ctx, BFGPU_ERR_STATE, "chip %s: unsupported quotient degree", ci.name);
        if (ci.prep_w && (pk_idx[i] < 0 || pk->traces[pk_idx[i]].rows != traces[i].rows))
            return fail(ctx, BFGPU_ERR_INVALID, "chip %s: preprocessed trace missing or of a different height", ci.name);
        if (2 * traces[i].rows < world) return fail(ctx, BFGPU_ERR_INVALID, "chip %s: LDE height below the world size", ci.name);
    }
    for (const DMat& m : pk->data->ldes)
        if (m.rows < world) return fail(ctx, BFGPU_ERR_INVALID, "preprocessed LDE height below the world size");

    // ---- the three sharded commitments and the quotient coefficient buffer: exported blocks, handles exchanged ONCE ---------------
    struct DcGuard {
        bfgpu_dist_commit* dc[3] = {nullptr, nullptr, nullptr};
        bfgpu_ctx* ctx = nullptr;
        uint32_t *qcol = nullptr, *gbuf = nullptr;
        ~DcGuard() {
            for (auto* d : dc) bfgpu_dist_commit_free(d);
            if (qcol) dfree_export(ctx, qcol);
            if (gbuf) dfree_export(ctx, gbuf);
        }
    } G;
    G.ctx = ctx;
    std::vector<uint64_t> rows_main(nchips), rows_q(2 * nchips);
    std::vector<uint32_t> cols_main(nchips), cols_perm(nchips), cols_q(2 * nchips, 4);
    for (size_t i = 0; i < nchips; i++) {
        rows_main[i] = traces[i].rows;
        cols_main[i] = (uint32_t)air::CHIPS[chips[i]].main_w;
        cols_perm[i] = 4u * (uint32_t)air::CHIPS[chips[i]].perm_w;
        rows_q[2 * i] = rows_q[2 * i + 1] = traces[i].rows;
    }
    TRY(bfgpu_dist_commit_begin(ctx, rank, world, rows_main.data(), cols_main.data(), (int32_t)nchips, &G.dc[0]));
    TRY(bfgpu_dist_commit_begin(ctx, rank, world, rows_main.data(), cols_perm.data(), (int32_t)nchips, &G.dc[1]));
    TRY(bfgpu_dist_commit_begin(ctx, rank, world, rows_q.data(), cols_q.data(), (int32_t)(2 * nchips), &G.dc[2]));
    bfgpu_dist_commit *dcm = G.dc[0], *dcp = G.dc[1], *dcq = G.dc[2];
    // quotient coefficient columns owned by rank r: offsets inside r's buffer, computable by everybody
    auto qcol_offset = [&](uint32_t r, size_t mi) {
        uint64_t off = 0;
        for (size_t k = 0; k < mi; k++) {
            uint32_t c0, nc;
            dist_col_range(4, world, r, (uint32_t)k, &c0, &nc);
            off += (uint64_t)nc * rows_q[k];
        }
        return off;
    };
    const uint64_t qcol_words = qcol_offset(rank, 2 * nchips);
    TRY(dalloc_export(ctx, (void**)&G.qcol, std::max<uint64_t>(qcol_words, 1) * 4));
    // FRI gather buffer: below `gather_at` elements the folded vector and the reduced openings still to come are all-gathered by
    // plain peer copies into this buffer on every rank (less than 2 * gather_at extension elements in total)
    unsigned log_tallest = 0;
    for (size_t i = 0; i < nchips; i++) log_tallest = std::max(log_tallest, ilog2(rows_main[i]) + 1);
    for (const DMat& m : pk->data->ldes) log_tallest = std::max(log_tallest, ilog2(m.rows));
    const uint64_t gather_at = std::min<uint64_t>(1ull << log_tallest, std::max<uint64_t>(1ull << ctx->dist_fri_gather_log, 4ull * world));
    TRY(dalloc_export(ctx, (void**)&G.gbuf, 2 * gather_at * 16));
    std::vector<uint32_t*> qcol_peer, gbuf_peer;
    {
        uint8_t mine[5 * 64];
        for (int k = 0; k < 3; k++) TRY(bfgpu_dist_commit_recv_handle(G.dc[k], mine + 64 * k));
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, G.qcol));
        memcpy(mine + 192, &h, 64);
        CU(cudaIpcGetMemHandle(&h, G.gbuf));
        memcpy(mine + 256, &h, 64);
        std::vector<uint8_t> all((size_t)world * 320);
        COMM(comm->all_gather(comm->user, mine, all.data(), 320));
        std::vector<uint8_t> per(world * 64);
        for (int k = 0; k < 3; k++) {
            for (uint32_t r = 0; r < world; r++) memcpy(&per[64 * r], &all[320 * r + 64 * k], 64);
            TRY(bfgpu_dist_commit_set_peers(G.dc[k], per.data()));
        }
        TRY(dist_map_peers(ctx, all.data() + 192, 320, rank, world, G.qcol, &qcol_peer));
        TRY(dist_map_peers(ctx, all.data() + 256, 320, rank, world, G.gbuf, &gbuf_peer));
    }
    // nobody may still be reading these exported buffers for a previous proof (they are recycled through the exported pool)
    COMM(comm->barrier(comm->user));

    // ---- main commitment (prover.rs:209-236): this rank extends ITS columns of every main trace -----------------------------------
    uint32_t main_root[8], perm_root[8], quot_root[8];
    {
        std::vector<DMat> slices(nchips);
        std::vector<uint32_t> shifts(nchips, gen);
        Scratch copies(ctx);
        for (size_t i = 0; i < nchips; i++) {
            const auto& m = dcm->mats[i];
            if (!m.ncols) continue;
            slices[i].rows = m.rows;
            slices[i].cols = m.ncols;
            const size_t bytes = (size_t)m.rows * m.ncols * 4;
            TRY(copies.alloc((void**)&slices[i].d, bytes));  // the LogUp kernel reads the traces later: extend a copy
            CU(cudaMemcpyAsync(slices[i].d, traces[i].d + (uint64_t)m.col0 * m.rows, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        TRY(dist_lde_coefs(dcm, slices, shifts));
    }
    TRY(dist_finish_commit(ctx, comm, dcm, main_root));

    // ---- transcript: main commitment, LogUp challenges (prover.rs:266-272) ---------------------------------------------------------
    ch.observe_slice(main_root, 8);
    air::Challenges chal;
    chal.alpha = ch.sample_ext();
    kb::Ext beta = ch.sample_ext();
    chal.beta_pow[0] = kb::ext_one();
    for (int k = 1; k < 8; k++) chal.beta_pow[k] = kb::ext_mul(chal.beta_pow[k - 1], beta);
    chal.cumulative_sum = kb::ext_zero();

    // ---- LogUp traces (replicated) and their commitment ----------------------------------------------------------------------------
    std::vector<kb::Ext> csum(nchips);
    {
        std::vector<DMat> perm(nchips);
        struct PermGuard {
            bfgpu_ctx* ctx;
            std::vector<DMat>& p;
            ~PermGuard() {
                for (DMat& m : p) dfree(ctx, m.d);
            }
        } perm_guard{ctx, perm};
        Scratch cs(ctx);
        uint32_t* d_csums = nullptr;
        TRY(cs.alloc((void**)&d_csums, nchips * 16));
        TRY(perm_traces(ctx, pk, pk_idx, chips, traces, chal, &perm, d_csums));
        std::vector<DMat> slices(nchips);
        std::vector<uint32_t> shifts(nchips, gen);
        for (size_t i = 0; i < nchips; i++) {
            const auto& m = dcp->mats[i];
            if (!m.ncols) continue;
            slices[i].rows = m.rows;
            slices[i].cols = m.ncols;
            slices[i].d = perm[i].d + (uint64_t)m.col0 * m.rows;  // extended in place: the trace itself is not needed again
        }
        TRY(dist_lde_coefs(dcp, slices, shifts));
        CU(cudaMemcpyAsync(csum.data(), d_csums, nchips * 16, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(dist_finish_commit(ctx, comm, dcp, perm_root));  // synchronises: csum has arrived
    }
    // the main traces are not needed any more
    for (DMat& m : traces) dfree(ctx, m.d);
    traces.clear();
    ch.observe_slice(perm_root, 8);
    for (size_t i = 0; i < nchips; i++) ch.observe_ext(csum[i]);

    // ---- quotient values (prover.rs:343-388 -> quotient.rs:18-165) on this rank's rows ---------------------------------------------
    const kb::Ext alpha = ch.sample_ext();
    std::vector<uint32_t> qshifts(2 * nchips);
    {
        Phase ph(ctx, BFGPU_PHASE_QUOTIENT);
        std::vector<kb::Ext> apow(air::MAX_CONSTRAINTS);
        apow[0] = kb::ext_one();
        for (int k = 1; k < air::MAX_CONSTRAINTS; k++) apow[k] = kb::ext_mul(apow[k - 1], alpha);
        Scratch qs(ctx);
        kb::Ext* d_apow = nullptr;
        TRY(qs.alloc((void**)&d_apow, apow.size() * sizeof(kb::Ext)));
        TRY(upload_small(ctx, d_apow, apow.data(), apow.size() * sizeof(kb::Ext)));
        for (size_t i = 0; i < nchips; i++) {
            const unsigned log_n = ilog2(rows_main[i]);
            const uint64_t n = 1ull << log_n;
            air::QuotientShardArgs qa;
            memset(&qa, 0, sizeof qa);
            qa.chip = chips[i];
            for (uint32_t r = 0; r < world; r++) {
                qa.main.shard[r] = dcm->peer_recv[r] + dcm->mats[i].recv_off;
                qa.perm.shard[r] = dcp->peer_recv[r] + dcp->mats[i].recv_off;
            }
            qa.prep = pk_idx[i] >= 0 ? pk->data->ldes[pk_idx[i]].d : nullptr;
            qa.log_n = log_n;
            qa.lqd = 1;
            qa.log_rpg = log_n + 1 - log_world;
            qa.rank = rank;
            qa.shift = gen;
            qa.g_inv = kb::inv(kb::two_adic_generator(log_n));
            uint32_t sn = kb::pow(gen, n);
            qa.zh[0] = kb::sub(sn, kb::ONE);
            qa.zh[1] = kb::sub(kb::neg(sn), kb::ONE);
            qa.zh_inv[0] = kb::inv(qa.zh[0]);
            qa.zh_inv[1] = kb::inv(qa.zh[1]);
            qa.apow = d_apow;
            qa.tw = ctx->d_tw;
            const uint32_t w2n = kb::two_adic_generator(log_n + 1);
            for (int c = 0; c < 2; c++) {
                const size_t mi = 2 * i + (size_t)c;
                qshifts[mi] = c == 0 ? kb::ONE : kb::inv(w2n);  // split_domains (prover.rs:391-402): chunk c lives on 3 w_{2n}^c H
                for (uint32_t k = 0; k < 4; k++) {
                    uint32_t oc0 = 0;
                    const uint32_t owner = dist_col_owner(4, world, (uint32_t)mi, k, &oc0);
                    qa.out_col[c][k] = qcol_peer[owner] + qcol_offset(owner, mi) + (uint64_t)(k - oc0) * n;
                }
            }
            air::Challenges c2 = chal;
            c2.cumulative_sum = csum[i];
            const uint64_t rpg = 1ull << qa.log_rpg;
#define BF_QS_CASE(C) \
    case C: air::k_quotient_shard<C><<<(unsigned)((rpg + 127) / 128), 128, 0, ctx->stream>>>(qa, c2); break;
            switch (qa.chip) { BF_QS_CASE(0) BF_QS_CASE(1) BF_QS_CASE(2) BF_QS_CASE(3) BF_QS_CASE(4) BF_QS_CASE(5) BF_QS_CASE(6) BF_QS_CASE(7) }
#undef BF_QS_CASE
            LAUNCHED(ctx);
            CU(cudaGetLastError());
        }
        TRY(dist_sync_barrier(ctx, comm));  // every quotient word has reached the rank that owns its column
    }
    {
        std::vector<DMat> slices(2 * nchips);
        for (size_t mi = 0; mi < 2 * nchips; mi++) {
            const auto& m = dcq->mats[mi];
            if (!m.ncols) continue;
            slices[mi].rows = m.rows;
            slices[mi].cols = m.ncols;
            slices[mi].d = G.qcol + qcol_offset(rank, mi);
        }
        TRY(dist_lde_coefs(dcq, slices, qshifts));
    }
    TRY(dist_finish_commit(ctx, comm, dcq, quot_root));
    ch.observe_slice(quot_root, 8);

    // ---- opening points (prover.rs:415-458) -------------------------------------------------------------------------------------------
    const kb::Ext zeta = ch.sample_ext();
    auto next_point = [&](unsigned log_n) { return kb::ext_scale(zeta, kb::two_adic_generator(log_n)); };
    distp::SRound R[4];
    R[0].replicated = true;
    R[0].tree = pk->data->tree;
    for (size_t k = 0; k < pk->names.size(); k++) {
        const DMat& l = pk->data->ldes[k];
        distp::SMat m{l.d + (uint64_t)rank * (l.rows / world), l.rows, l.rows, l.rows / world, l.cols, {zeta}, {}};
        if (!air::CHIPS[pk->chip[k]].local_only) m.pts.push_back(next_point(ilog2(pk->traces[k].rows)));
        R[0].mats.push_back(m);
    }
    bfgpu_dist_commit* dcs[3] = {dcm, dcp, dcq};
    for (int r = 0; r < 3; r++) {
        R[r + 1].tree = dcs[r]->tree;
        R[r + 1].top = &dcs[r]->top;
        for (size_t k = 0; k < dcs[r]->mats.size(); k++) {
            const auto& dm = dcs[r]->mats[k];
            const size_t chip_pos = r == 2 ? k / 2 : k;
            distp::SMat m{dcs[r]->recv + dm.recv_off, dm.rpg, dm.lde_rows, dm.rpg, dm.total_cols, {zeta}, {}};
            const bool both = r == 1 || (r == 0 && !air::CHIPS[chips[chip_pos]].local_only);
            if (both) m.pts.push_back(next_point(ilog2(rows_main[chip_pos])));
            R[r + 1].mats.push_back(m);
        }
    }
    unsigned log_global_max = 0;
    for (auto& rd : R) {
        for (auto& m : rd.mats) rd.log_max = std::max(rd.log_max, ilog2(m.h));
        log_global_max = std::max(log_global_max, rd.log_max);
    }
    std::vector<uint32_t> opening;  // the bfgpu_opening layout

    // ---- (i) opened values: barycentric partial sums over this rank's rows of the WHOLE LDE coset (size h, shift GENERATOR) --------
    {
        Phase ph(ctx, BFGPU_PHASE_OPEN_EVAL);
        Scratch sc(ctx);
        std::map<ExtKey, uint32_t*> wcache;
        auto weights = [&](unsigned log_h, const kb::Ext& z, uint32_t** w) -> int32_t {
            ExtKey key{log_h, {z.c[0], z.c[1], z.c[2], z.c[3]}};
            auto it = wcache.find(key);
            if (it == wcache.end()) {
                uint32_t* d = nullptr;
                const uint32_t cnt = (1u << log_h) / world;
                TRY(sc.alloc((void**)&d, (size_t)16 * cnt));
                openk::k_bary_weights<<<(cnt + 255) / 256, 256, 0, ctx->stream>>>(d, log_h, gen, z, ctx->d_tw, rank * cnt, cnt);
                LAUNCHED(ctx);
                CU(cudaGetLastError());
                it = wcache.emplace(key, d).first;
            }
            *w = it->second;
            return BFGPU_OK;
        };
        struct Job {
            distp::SMat* m;
            size_t t0;
            uint32_t np;
            size_t off;
        };
        std::vector<Job> jobs;
        size_t total_words = 0;
        for (auto& rd : R)
            for (auto& m : rd.mats)
                for (size_t t0 = 0; t0 < m.pts.size(); t0 += 2) {
                    uint32_t np = (uint32_t)std::min<size_t>(2, m.pts.size() - t0);
                    jobs.push_back({&m, t0, np, total_words});
                    total_words += (size_t)m.cols * np * 4;
                }
        uint32_t* sums_all = nullptr;
        TRY(sc.alloc((void**)&sums_all, total_words * 4));
        std::vector<openk::BaryJob> batch[2];
        uint32_t batch_groups[2] = {0, 0};
        for (Job& jb : jobs) {
            distp::SMat& m = *jb.m;
            const uint32_t hl = (uint32_t)m.rpg;  // local rows
            const unsigned log_h = ilog2(m.h);
            const uint32_t nchunks = (hl + openk::BARY_ROWS - 1) / openk::BARY_ROWS;
            uint32_t *w0 = nullptr, *w1 = nullptr;
            TRY(weights(log_h, m.pts[jb.t0], &w0));
            if (jb.np == 2) TRY(weights(log_h, m.pts[jb.t0 + 1], &w1));
            if (nchunks == 1) {
                batch[jb.np - 1].push_back({m.d, m.stride, w0, w1, sums_all + jb.off, m.cols, hl});
                batch_groups[jb.np - 1] = std::max(batch_groups[jb.np - 1], (m.cols + openk::BARY_COLS - 1) / openk::BARY_COLS);
                continue;
            }
            uint32_t* partial = nullptr;
            const size_t nsum = (size_t)m.cols * jb.np * 4;
            TRY(sc.alloc((void**)&partial, nsum * nchunks * 4));
            dim3 grid((m.cols + openk::BARY_COLS - 1) / openk::BARY_COLS, nchunks);
            if (jb.np == 1) openk::k_bary_dot<1><<<grid, openk::BARY_THREADS, 0, ctx->stream>>>(m.d, m.stride, m.cols, hl, w0, w1, partial, nchunks);
            else openk::k_bary_dot<2><<<grid, openk::BARY_THREADS, 0, ctx->stream>>>(m.d, m.stride, m.cols, hl, w0, w1, partial, nchunks);
            LAUNCHED(ctx);
            openk::k_bary_finish<<<(unsigned)((nsum + 127) / 128), 128, 0, ctx->stream>>>(partial, sums_all + jb.off, m.cols, nchunks, jb.np);
            LAUNCHED(ctx);
            CU(cudaGetLastError());
        }
        for (int k = 0; k < 2; k++) {
            if (batch[k].empty()) continue;
            openk::BaryJob* d_jobs = nullptr;
            TRY(sc.alloc((void**)&d_jobs, batch[k].size() * sizeof(openk::BaryJob)));
            TRY(upload_small(ctx, d_jobs, batch[k].data(), batch[k].size() * sizeof(openk::BaryJob)));
            dim3 grid(batch_groups[k], (unsigned)batch[k].size());
            if (k == 0) openk::k_bary_dot_batch<1><<<grid, openk::BARY_THREADS, 0, ctx->stream>>>(d_jobs);
            else openk::k_bary_dot_batch<2><<<grid, openk::BARY_THREADS, 0, ctx->stream>>>(d_jobs);
            LAUNCHED(ctx);
            CU(cudaGetLastError());
        }
        std::vector<uint32_t> mine(total_words), all(total_words * (size_t)world);
        CU(cudaMemcpyAsync(mine.data(), sums_all, total_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        COMM(comm->all_gather(comm->user, mine.data(), all.data(), total_words * 4));
        for (size_t w = 0; w < total_words; w++) {
            uint32_t v = 0;
            for (uint32_t r = 0; r < world; r++) v = kb::add(v, all[(size_t)r * total_words + w]);
            mine[w] = v;
        }
        for (Job& jb : jobs) {
            distp::SMat& m = *jb.m;
            const uint64_t h = m.h;
            m.ys.resize(m.pts.size());
            for (uint32_t t = 0; t < jb.np; t++) {
                // p(z) = (z^h - s^h) / (h s^(h-1)) * sum   (interpolation over the coset s K, |K| = h; the value of a polynomial is unique)
                const kb::Ext& z = m.pts[jb.t0 + t];
                kb::Ext zer = ext_pow(z, h);
                zer.c[0] = kb::sub(zer.c[0], kb::pow(gen, h));
                uint32_t den = kb::mul(kb::pow(gen, h - 1), kb::to_mont((uint32_t)(h % kb::P)));
                kb::Ext scale = kb::ext_scale(zer, kb::inv(den));
                auto& ys = m.ys[jb.t0 + t];
                ys.resize(m.cols);
                for (uint32_t c = 0; c < m.cols; c++) {
                    const uint32_t* sp = &mine[jb.off + ((size_t)c * jb.np + t) * 4];
                    ys[c] = kb::ext_mul(scale, kb::Ext{{sp[0], sp[1], sp[2], sp[3]}});
                }
            }
        }
    }
    for (auto& rd : R)
        for (auto& m : rd.mats)
            for (auto& ys : m.ys)
                for (auto& y : ys) {
                    for (int k = 0; k < 4; k++) opening.push_back(out_word(ctx, y.c[k]));
                    if (ctx->opt[BFGPU_OPT_OBSERVE_OPENED_VALUES]) ch.observe_ext(y);
                }
    const kb::Ext fri_alpha = ch.sample_ext();

    // ---- (ii) reduced openings per LDE height, on this rank's rows -----------------------------------------------------------------------
    std::map<unsigned, uint32_t*, std::greater<unsigned>> reduced;  // log GLOBAL height -> local ext vector (h / world elements)
    std::vector<distp::SLayer> slayers;
    std::vector<FriLayer> layers;
    struct OpenGuard {
        bfgpu_ctx* ctx;
        std::map<unsigned, uint32_t*, std::greater<unsigned>>& reduced;
        std::vector<distp::SLayer>& sl;
        std::vector<FriLayer>& fl;
        ~OpenGuard() {
            for (auto& kv : reduced) dfree(ctx, kv.second);
            for (auto& L : sl) {
                tree_release(L.tree);
                dfree(ctx, L.vec);
            }
            for (auto& L : fl) {
                tree_release(L.tree);
                dfree(ctx, L.vec);
            }
        }
    } open_guard{ctx, reduced, slayers, layers};
    {
        Phase ph(ctx, BFGPU_PHASE_OPEN_REDUCE);
        uint32_t maxw = 1;
        for (auto& rd : R)
            for (auto& m : rd.mats) maxw = std::max(maxw, m.cols);
        std::vector<kb::Ext> apow(maxw);
        apow[0] = kb::ext_one();
        for (uint32_t k = 1; k < maxw; k++) apow[k] = kb::ext_mul(apow[k - 1], fri_alpha);
        Scratch sc(ctx);
        uint32_t* d_apow = nullptr;
        TRY(sc.alloc((void**)&d_apow, (size_t)maxw * 16));
        TRY(upload_small(ctx, d_apow, apow.data(), (size_t)maxw * 16));
        struct Group {
            std::vector<openk::RoMat> mats;
            std::vector<kb::Ext> pts;
            uint64_t num_reduced = 0;
        };
        std::map<unsigned, Group> groups;
        for (auto& rd : R)
            for (auto& m : rd.mats) {
                Group& g = groups[ilog2(m.h)];
                openk::RoMat rm;
                memset(&rm, 0, sizeof rm);
                rm.d = m.d;
                rm.stride = m.stride;
                rm.width = m.cols;
                rm.npoints = (uint32_t)m.pts.size();
                for (size_t t = 0; t < m.pts.size(); t++) {
                    size_t pi = 0;
                    for (; pi < g.pts.size(); pi++)
                        if (!memcmp(g.pts[pi].c, m.pts[t].c, 16)) break;
                    if (pi == g.pts.size()) g.pts.push_back(m.pts[t]);
                    if (g.pts.size() > 4) return fail(ctx, BFGPU_ERR_INVALID, "more than four distinct opening points per height are not supported");
                    rm.pt[t] = (uint32_t)pi;
                    kb::Ext yr = kb::ext_zero();
                    for (uint32_t k = 0; k < m.cols; k++) yr = kb::ext_add(yr, kb::ext_mul(apow[k], m.ys[t][k]));
                    kb::Ext ao = ext_pow(fri_alpha, g.num_reduced);
                    memcpy(rm.yred[t], yr.c, 16);
                    memcpy(rm.aoff[t], ao.c, 16);
                    g.num_reduced += m.cols;
                }
                g.mats.push_back(rm);
            }
        for (auto& kv : groups) {
            const unsigned lh = kv.first;
            Group& g = kv.second;
            const uint32_t cnt = (1u << lh) / world;
            openk::RoMat* d_m = nullptr;
            uint32_t *d_z = nullptr, *ro = nullptr;
            TRY(sc.alloc((void**)&d_m, g.mats.size() * sizeof(openk::RoMat)));
            TRY(sc.alloc((void**)&d_z, g.pts.size() * 16));
            TRY(dalloc(ctx, (void**)&ro, (size_t)16 * cnt));
            reduced[lh] = ro;
            TRY(upload_small(ctx, d_m, g.mats.data(), g.mats.size() * sizeof(openk::RoMat)));
            TRY(upload_small(ctx, d_z, g.pts.data(), g.pts.size() * 16));
            openk::k_reduce_openings<<<(cnt + 127) / 128, 128, 0, ctx->stream>>>(d_m, (uint32_t)g.mats.size(), d_z, (uint32_t)g.pts.size(), d_apow, lh, gen, ctx->d_tw,
                                                                              ro, rank * cnt, cnt);
            LAUNCHED(ctx);
            CU(cudaGetLastError());
        }
    }

    // ---- (iii) FRI commit phase: sharded rounds, then the gathered tail ----------------------------------------------------------------
    const unsigned log_max_height = reduced.begin()->first;
    std::vector<std::array<uint32_t, 8>> commits;
    uint32_t final_poly[4];
    {
        Phase ph(ctx, BFGPU_PHASE_FRI);
        auto it = reduced.begin();
        uint32_t* folded = it->second;  // local slice
        uint64_t len = 1ull << it->first;
        it->second = nullptr;
        ++it;
        struct FoldedGuard {
            bfgpu_ctx* ctx;
            uint32_t*& p;
            ~FoldedGuard() { dfree(ctx, p); }
        } folded_guard{ctx, folded};
        while (len > gather_at) {
            const uint64_t ll = len / world;  // local length
            DMat leaves;
            leaves.d = folded;
            leaves.rows = ll / 2;
            leaves.cols = 8;
            leaves.rs = 8;
            bfgpu_tree* t = nullptr;
            int32_t rc = build_tree(ctx, {leaves}, false, &t);
            slayers.push_back({folded, len, t, {}});
            folded = nullptr;  // owned by the layer now
            if (rc != BFGPU_OK) return rc;
            uint32_t cap[8];
            CU(cudaMemcpyAsync(cap, t->layers.back(), 32, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            std::vector<uint32_t> caps(8 * (size_t)world);
            COMM(comm->all_gather(comm->user, cap, caps.data(), 32));
            distp::cap_tree(caps.data(), world, &slayers.back().top);
            std::array<uint32_t, 8> root;
            memcpy(root.data(), slayers.back().top.back().data(), 32);
            ch.observe_slice(root.data(), 8);
            commits.push_back(root);
            const kb::Ext fbeta = ch.sample_ext();
            const uint64_t nlen = len / 2, nll = nlen / world;
            uint32_t* next = nullptr;
            TRY(dalloc(ctx, (void**)&next, nll * 16));
            const uint32_t* add = nullptr;
            if (it != reduced.end() && (1ull << it->first) == nlen) add = it->second;
            openk::k_fri_fold<<<(unsigned)((nll + 127) / 128), 128, 0, ctx->stream>>>(slayers.back().vec, next, add, ilog2(nlen), kb::ext_scale(fbeta, kb::halve(kb::ONE)),
                                                                                     ctx->d_tw, (int)ctx->opt[BFGPU_OPT_FRI_ROLLIN], (uint32_t)(rank * nll), (uint32_t)nll);
            LAUNCHED(ctx);
            CU(cudaGetLastError());
            if (add) {
                dfree(ctx, it->second);  // stream order keeps it alive for the fold
                it->second = nullptr;
                ++it;
            }
            folded = next;
            len = nlen;
        }
        // gather the folded vector and the reduced openings still to come: every rank stores its slice into the gather buffer of every
        // rank (peer copies over NVLink), one barrier, and every rank finishes the commit phase on the whole vector
        uint64_t goff = 0;  // words
        auto gather_vec = [&](uint32_t* local, uint64_t glen, uint32_t** full) -> int32_t {
            const uint64_t ll = glen / world;
            if (goff + glen * 4 > 2 * gather_at * 4) return fail(ctx, BFGPU_ERR_STATE, "internal: FRI gather buffer too small");
            for (uint32_t r = 0; r < world; r++) {
                const uint32_t dst = (rank + 1 + r) % world;  // senders start at different receivers
                CU(cudaMemcpyAsync(gbuf_peer[dst] + goff + (uint64_t)rank * ll * 4, local, ll * 16, cudaMemcpyDeviceToDevice, ctx->stream));
            }
            *full = G.gbuf + goff;
            goff += glen * 4;
            return BFGPU_OK;
        };
        uint32_t* full = nullptr;
        TRY(gather_vec(folded, len, &full));
        std::vector<std::pair<unsigned, uint32_t*>> gathered;
        for (auto jt = it; jt != reduced.end(); ++jt) {
            uint32_t* f = nullptr;
            TRY(gather_vec(jt->second, 1ull << jt->first, &f));
            gathered.push_back({jt->first, f});
        }
        TRY(dist_sync_barrier(ctx, comm));  // every slice has landed in this rank's buffer
        dfree(ctx, folded);
        folded = nullptr;
        for (auto& g : gathered) {
            dfree(ctx, reduced[g.first]);
            reduced[g.first] = g.second;  // lives in the exported gather buffer: dfree() ignores blocks it does not own
        }
        TRY(fri_commit_phase(ctx, ch, full, len, reduced, layers, commits, final_poly));
    }
    opening.push_back((uint32_t)commits.size());
    for (auto& c : commits)
        for (int k = 0; k < 8; k++) opening.push_back(out_word(ctx, c[k]));
    for (int k = 0; k < 4; k++) opening.push_back(out_word(ctx, final_poly[k]));

    // ---- (iv) proof of work (replicated: every rank finds the same witness) ----------------------------------------------------------
    uint32_t witness = 0;
    TRY(pow_grind(ctx, ch, fixed_pow_witness, &witness));
    opening.push_back(witness);

    // ---- (v) queries: the owner of a leaf gathers its words; ONE all-gather assembles the proof on every rank -------------------------
    {
        Phase ph(ctx, BFGPU_PHASE_QUERY);
        const uint32_t nq = ctx->num_queries;
        opening.push_back(nq);
        std::vector<uint32_t> indices(nq);
        for (uint32_t q = 0; q < nq; q++) indices[q] = ch.sample_bits(log_max_height);
        // A proof is a sequence of segments in a fixed order; a segment is either known to everybody (the index, siblings inside a
        // top tree) or comes from exactly one rank (device words: a pointer per word).
        struct Seg {
            int32_t owner;          // -1: host words (already in caller representation)
            uint32_t nwords;
            size_t host_off;        // owner < 0: offset into `hostw`
        };
        std::vector<Seg> segs;
        std::vector<uint32_t> hostw;
        std::vector<const uint32_t*> myptr;  // device words this rank contributes, in order
        std::vector<uint64_t> per_rank(world, 0);
        auto host_seg = [&](const uint32_t* w, uint32_t n) {
            segs.push_back({-1, n, hostw.size()});
            hostw.insert(hostw.end(), w, w + n);
        };
        auto top_siblings = [&](const std::vector<std::vector<uint32_t>>& top, uint64_t pos) {  // pos = owning rank
            for (size_t l = 0; l + 1 < top.size(); l++, pos >>= 1) {
                uint32_t w[8];
                for (int k = 0; k < 8; k++) w[k] = out_word(ctx, top[l][8 * (pos ^ 1) + k]);
                host_seg(w, 8);
            }
        };
        auto tree_path = [&](const bfgpu_tree* t, uint64_t leaf, unsigned levels, bool mine) {
            for (unsigned l = 0; l < levels; l++)
                for (uint32_t k = 0; k < 8; k++)
                    if (mine) myptr.push_back(t->layers[l] + 8 * ((leaf >> l) ^ 1) + k);
        };
        for (uint32_t q = 0; q < nq; q++) {
            const uint64_t index = indices[q];
            uint32_t iw = (uint32_t)index;
            host_seg(&iw, 1);
            for (auto& rd : R) {
                const uint64_t ridx = index >> (log_global_max - rd.log_max);
                const uint64_t per = (1ull << rd.log_max) / world;  // leaves per rank
                const uint32_t owner = (uint32_t)(ridx / per);
                const bool mine = owner == rank;
                uint32_t nwords = 0;
                for (auto& m : rd.mats) {
                    const uint64_t row = ridx >> (rd.log_max - ilog2(m.h));  // global stored row of this matrix
                    const uint64_t lrow = row - (uint64_t)owner * m.rpg;    // inside the owner's window
                    if (mine)
                        for (uint32_t c = 0; c < m.cols; c++) myptr.push_back(m.d + (uint64_t)c * m.stride + lrow);
                    nwords += m.cols;
                }
                if (rd.replicated) {  // the full tree is everywhere: the owner serves the whole path
                    tree_path(rd.tree, ridx, rd.log_max, mine);
                    nwords += 8 * rd.log_max;
                    segs.push_back({(int32_t)owner, nwords, 0});
                } else {
                    const unsigned local_levels = rd.log_max - log_world;
                    tree_path(rd.tree, ridx - (uint64_t)owner * per, local_levels, mine);
                    nwords += 8 * local_levels;
                    segs.push_back({(int32_t)owner, nwords, 0});
                    top_siblings(*rd.top, owner);
                }
                per_rank[owner] += nwords;
            }
            size_t li = 0;
            for (; li < slayers.size(); li++) {  // layers committed while sharded
                const distp::SLayer& L = slayers[li];
                const uint64_t idx = index >> li, pair = idx >> 1;
                const uint64_t ll = L.len / world, leaves_per = ll / 2;
                const uint32_t owner = (uint32_t)(pair / leaves_per);
                const bool mine = owner == rank;
                const unsigned local_levels = ilog2(leaves_per);
                if (mine)
                    for (uint32_t k = 0; k < 4; k++) myptr.push_back(L.vec + 4 * ((idx ^ 1) - (uint64_t)owner * ll) + k);
                tree_path(L.tree, pair - (uint64_t)owner * leaves_per, local_levels, mine);
                const uint32_t nwords = 4 + 8 * local_levels;
                segs.push_back({(int32_t)owner, nwords, 0});
                per_rank[owner] += nwords;
                top_siblings(L.top, owner);
            }
            for (size_t fi = 0; fi < layers.size(); fi++, li++) {  // replicated tail: served by the rank the index points at (balance)
                const FriLayer& L = layers[fi];
                const uint64_t idx = index >> li, pair = idx >> 1;
                const uint32_t owner = (uint32_t)(index % world);
                const bool mine = owner == rank;
                if (mine)
                    for (uint32_t k = 0; k < 4; k++) myptr.push_back(L.vec + 4 * (idx ^ 1) + k);
                tree_path(L.tree, pair, L.tree->log_max, mine);
                const uint32_t nwords = 4 + 8 * L.tree->log_max;
                segs.push_back({(int32_t)owner, nwords, 0});
                per_rank[owner] += nwords;
            }
        }
        if (myptr.size() != per_rank[rank]) return fail(ctx, BFGPU_ERR_STATE, "internal: query segment accounting");
        const uint64_t maxw = *std::max_element(per_rank.begin(), per_rank.end());
        std::vector<uint32_t> mine(std::max<uint64_t>(maxw, 1), 0), all(std::max<uint64_t>(maxw, 1) * world);
        if (!myptr.empty()) {
            Scratch sc(ctx);
            const uint32_t** d_ptr = nullptr;
            uint32_t* d_out = nullptr;
            TRY(sc.alloc((void**)&d_ptr, myptr.size() * sizeof(void*)));
            TRY(sc.alloc((void**)&d_out, myptr.size() * 4));
            CU(cudaMemcpyAsync((void*)d_ptr, myptr.data(), myptr.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
            openk::k_gather_words<<<(unsigned)((myptr.size() + 255) / 256), 256, 0, ctx->stream>>>(d_ptr, d_out, myptr.size(), ctx->repr == BFGPU_REPR_CANONICAL);
            LAUNCHED(ctx);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(mine.data(), d_out, myptr.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        }
        COMM(comm->all_gather(comm->user, mine.data(), all.data(), mine.size() * 4));
        std::vector<uint64_t> cursor(world, 0);
        for (const Seg& s : segs) {
            if (s.owner < 0) {
                opening.insert(opening.end(), hostw.begin() + s.host_off, hostw.begin() + s.host_off + s.nwords);
            } else {
                const uint32_t* src = all.data() + (size_t)s.owner * mine.size() + cursor[s.owner];
                opening.insert(opening.end(), src, src + s.nwords);
                cursor[s.owner] += s.nwords;
            }
        }
    }
    // nobody may recycle its exported buffers (next proof) while a peer still reads them
    TRY(dist_sync_barrier(ctx, comm));

    // ---- ShardProof (types.rs:32-73), same serialisation as bfgpu_machine_open --------------------------------------------------------
    std::vector<uint32_t>& flat = *proof_out;
    flat.clear();
    for (int k = 0; k < 8; k++) flat.push_back(out_word(ctx, main_root[k]));
    for (int k = 0; k < 8; k++) flat.push_back(out_word(ctx, perm_root[k]));
    for (int k = 0; k < 8; k++) flat.push_back(out_word(ctx, quot_root[k]));
    flat.push_back((uint32_t)nchips);
    for (size_t i = 0; i < nchips; i++) {
        flat.push_back((uint32_t)chips[i]);
        flat.push_back(ilog2(rows_main[i]));
        for (int k = 0; k < 4; k++) flat.push_back(out_word(ctx, csum[i].c[k]));
    }
    flat.insert(flat.end(), opening.begin(), opening.end());
    return BFGPU_OK;
}

// MachineProver::prove over `world` ranks from the execution record (every rank passes the SAME record and proving key; device-side
// trace generation is replicated) — the proof is returned on every rank.
extern "C" int32_t bfgpu_dist_prove_record(bfgpu_ctx* ctx, const bfgpu_comm* comm, uint32_t rank, uint32_t world, const bfgpu_pk* pk, const bfgpu_record* rec,
                                           bfgpu_challenger* chh, int64_t fixed_pow_witness, bfgpu_shard_proof** out) {
    if (!ctx || !comm || !comm->all_gather || !comm->barrier || !pk || !rec || !chh || !out) return fail(ctx, BFGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    AllocScope scope(ctx);
    std::vector<std::string> names;
    std::vector<int> chips;
    std::vector<DMat> traces;
    // the 16-byte cycle records (67 MB at 4.2 M cycles): every rank uploads 1/world of them and stores its slice into all peers
    // (8 ranks pulling the whole record over PCIe at once took 2.9 ms)
    struct RecGuard {
        bfgpu_ctx* ctx;
        uint32_t* buf = nullptr;
        ~RecGuard() { if (buf) dfree_export(ctx, buf); }
    } recg{ctx};
    const uint64_t nrec = rec->n_cycles + 1;
    if (world > 1 && nrec >= 64 * (uint64_t)world && is_pow2(world) && rank < world) {
        Phase ph(ctx, BFGPU_PHASE_H2D);
        TRY(dalloc_export(ctx, (void**)&recg.buf, nrec * 16));
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, recg.buf));
        std::vector<uint8_t> all((size_t)world * 64);
        COMM(comm->all_gather(comm->user, &h, all.data(), 64));
        std::vector<uint32_t*> peer;
        TRY(dist_map_peers(ctx, all.data(), 64, rank, world, recg.buf, &peer));
        const uint64_t lo = nrec * rank / world, hi = nrec * (rank + 1) / world;
        CU(cudaMemcpyAsync(recg.buf + lo * 4, (const uint32_t*)rec->cycles + lo * 4, (hi - lo) * 16, cudaMemcpyHostToDevice, ctx->stream));
        for (uint32_t r = 1; r < world; r++) {
            const uint32_t dst = (rank + r) % world;
            CU(cudaMemcpyAsync(peer[dst] + lo * 4, recg.buf + lo * 4, (hi - lo) * 16, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        TRY(dist_sync_barrier(ctx, comm));
    }
    int32_t rc = record_traces(ctx, rec, &names, &chips, &traces, (const uint4*)recg.buf);
    if (rc != BFGPU_OK) return rc;
    auto* proof = new bfgpu_shard_proof();
    rc = dist_prove_core(ctx, comm, rank, world, pk, names, chips, traces, box(chh)->ch, fixed_pow_witness, &proof->flat);
    if (rc != BFGPU_OK) {
        delete proof;
        return rc;
    }
    *out = proof;
    return scope.ok();
}

// Same from host (or device) traces: every rank passes ALL named main traces (as bfgpu_machine_commit takes them).
extern "C" int32_t bfgpu_dist_prove(bfgpu_ctx* ctx, const bfgpu_comm* comm, uint32_t rank, uint32_t world, const bfgpu_pk* pk, const char* const* names_in,
                                    const bfgpu_mat* mats, int32_t n, bfgpu_challenger* chh, int64_t fixed_pow_witness, bfgpu_shard_proof** out) {
    if (!ctx || !comm || !comm->all_gather || !comm->barrier || !pk || !names_in || !mats || n <= 0 || !chh || !out) return fail(ctx, BFGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    AllocScope scope(ctx);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        TRY(check_mat(ctx, &mats[i], true));
        int ci = chip_index(names_in[i]);
        if (ci < 0) return fail(ctx, BFGPU_ERR_INVALID, "unknown chip '%s'", names_in[i]);
        if (mats[i].cols != (uint64_t)air::CHIPS[ci].main_w) return fail(ctx, BFGPU_ERR_INVALID, "chip %s: wrong trace width", names_in[i]);
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (mats[a].rows != mats[b].rows) return mats[a].rows > mats[b].rows;
        return strcmp(names_in[a], names_in[b]) < 0;
    });
    std::vector<std::string> names;
    std::vector<int> chips;
    std::vector<DMat> traces;
    std::vector<bfgpu_mat> in_order(n);
    for (int k = 0; k < n; k++) in_order[k] = mats[order[k]];
    int32_t rc = prestage_all(ctx, in_order.data(), n);
    for (int k = 0; k < n && rc == BFGPU_OK; k++) {
        DMat t;
        rc = ingest(ctx, in_order[k], /*bitrev=*/true, &t);
        if (rc != BFGPU_OK) break;
        names.push_back(names_in[order[k]]);
        chips.push_back(chip_index(names_in[order[k]]));
        traces.push_back(t);
    }
    prestage_clear(ctx);
    if (rc != BFGPU_OK) {
        for (DMat& t : traces) dfree(ctx, t.d);
        return rc;
    }
    auto* proof = new bfgpu_shard_proof();
    rc = dist_prove_core(ctx, comm, rank, world, pk, names, chips, traces, box(chh)->ch, fixed_pow_witness, &proof->flat);
    if (rc != BFGPU_OK) {
        delete proof;
        return rc;
    }
    *out = proof;
    return scope.ok();
}
