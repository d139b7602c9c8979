// kernels_ntt3.cuh — strided NTT passes whose tiles travel by TMA in both directions (32-lane tiles, in-place shared tiles).
//
// Same role and arithmetic as the strided instantiations of ntt2::k_pass (reference: `Radix2DitParallel::coset_lde_batch`
// behind `TwoAdicFriPcs::commit`, crates/stark/src/prover.rs:227,334,411); what changes is how a tile moves:
//   * a tile is 2^g digits x 32 lanes (128-byte segments = whole L2 lines; ntt2 moves 64-byte segments), so every warp-wide
//     shared access of either register phase is ONE row of 32 consecutive words: dense tile, no padding, no swizzle;
//   * the input tile of a later column is fetched by ONE `cp.async.bulk.tensor.4d` (TMA box 32 x 2^g of the 4-d view
//     [column][hi][digit][lo] of the matrix) issued by thread 0 into a ring of slots, completion on an mbarrier;
//   * the column is transformed IN PLACE in its slot: each register phase reads and writes the words the thread owns in that
//     phase's layout, so the only hazards are the layout changes (one barrier each);
//   * results leave by TMA too: in the forward layout a warp owns 16 consecutive digits, so every warp stores its own 2 KB box
//     (`cp.async.bulk.tensor` shared -> global) after a warp-level sync — no per-element address arithmetic, no STG.
// The SM therefore issues no global load/store instructions at all for the data of a pass (ncu of the register-prefetching
// passes: ~20 % of the issued instructions were address arithmetic and LSU global traffic in an issue-bound kernel).
//
// Slot reuse: a slot is refilled by thread 0 right after the first barrier of an iteration; by then every warp has waited
// (`cp.async.bulk.wait_group.read`) for its own store out of that slot — the wait sits before that barrier in program order.
//
// Modes: FWD / INV = one forward / inverse pass in place; TURN = LAST inverse pass + coset scaling + FIRST forward pass of a
// blow-up-2 LDE in one kernel: both act on the same tile (top g index bits) and the inverse pass ends in exactly the register
// layout the forward pass starts from, so the coefficients never travel to HBM between the two.
#pragma once
#include <cuda.h>

#include "kernels_ntt2.cuh"

namespace ntt3 {

using ntt2::G2;
using ntt2::TWT;
using ntt2::phase;

constexpr int LANES = 32;
enum Mode { FWD = 0, INV = 1, TURN = 2 };

struct PassArgs {
    uint32_t ncols;
    uint32_t cols_per_cta;
    uint32_t p;          // low bit of the pass (>= 5)
    uint32_t log_n;
    const TWT* twA;      // tables of the pass (TURN: the inverse ones)
    const TWT* twB;
    const TWT* twA2;     // TURN: forward tables of the same (p, g)
    const TWT* twB2;
    const uint32_t* pw;  // TURN: coset scale vectors [2][n]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)tm), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ constexpr int nslots(int mode) { return mode == TURN ? 2 : 4; }
__host__ __device__ constexpr int ntiles(int mode) { return mode == TURN ? 6 : 4; }  // TURN: 2 slots + 2 exchange tiles + 2 tiles of coset scale factors
constexpr size_t smem_bytes(int mode, int G1) {
    return (size_t)ntiles(mode) * ((size_t)4 << (G1 + 4 + 5)) + 2 * 15 * LANES * sizeof(TWT) + 4 * sizeof(uint64_t);
}

extern __shared__ __align__(1024) uint32_t smem3[];

// Threads per CTA = 2^(g+1) (16 elements per thread).  Register budget: TURN 128 per thread (512 threads resident per SM), the
// plain passes 80 (768 threads: three CTAs of 256 at g = 7; their four slots are 64 KB per CTA).
#ifndef NTT3_RESIDENT
#define NTT3_RESIDENT 768
#endif
template <int MODE, int G1>
__global__ void __launch_bounds__(1 << (G1 + 5), G1 >= 4 ? 1 : ((MODE == TURN ? 512 : NTT3_RESIDENT) >> (G1 + 5)))
    k_pass3(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, PassArgs A) {
    static_assert(G1 >= 1 && G1 <= 4, "two register phases");
    constexpr int g = G1 + G2, NT = 1 << (g + 1);
    constexpr int RA = 1 << G1, NGA = 16 >> G1;
    constexpr int TILE = (1 << g) * LANES;  // words
    constexpr int NSLOT = nslots(MODE);
    uint32_t* const s_slot = smem3;                               // [NSLOT][TILE]
    uint32_t* const s_ex = smem3 + NSLOT * TILE;                   // TURN: [2][TILE]
    uint32_t* const s_pw = smem3 + (NSLOT + 2) * TILE;             // TURN: [2][TILE]
    TWT(*const s_twb)[LANES] = reinterpret_cast<TWT(*)[LANES]>(smem3 + ntiles(MODE) * TILE);  // [15][LANES]
    TWT(*const s_twb2)[LANES] = s_twb + 15;                                                      // [15][LANES] (TURN)
    uint64_t* const full = reinterpret_cast<uint64_t*>(s_twb2 + 15);                             // [NSLOT]

    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t p = A.p;
    const uint32_t lo_blocks_log = p - 5;
    const uint32_t lo_base = (blockIdx.x & ((1u << lo_blocks_log) - 1)) << 5;
    const uint32_t hi = blockIdx.x >> lo_blocks_log;
    const uint32_t c_begin = blockIdx.y * A.cols_per_cta;
    const uint32_t c_end = min(A.ncols, c_begin + A.cols_per_cta);
    if (c_begin >= c_end) return;
    const uint32_t ncol = c_end - c_begin;

    // ---- ring of tiles -------------------------------------------------------------------------------------------------
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    auto issue = [&](uint32_t i) {  // thread 0: tile of the i-th column of this CTA into slot i % NSLOT
        const uint32_t s = i % NSLOT;
        mbar_expect_tx(&full[s], TILE * 4);
        tma_load_4d(s_slot + s * TILE, &tm_in, &full[s], (int)lo_base, 0, (int)hi, (int)(c_begin + i));
    };
    constexpr uint32_t AHEAD = MODE == TURN ? 1 : 2;  // columns in flight beyond the current one
    if (t == 0) {
#pragma unroll
        for (uint32_t s = 0; s < AHEAD; s++)
            if (s < ncol) issue(s);
    }

    // ---- per-thread coordinates: phase A combos (r1, lane) x NGA, phase B combo (aB = warp, lane) ----------------------
    uint32_t offA[NGA];  // word offset of (digit r1, lane) in a tile; element a of the group sits (a << 9) words further
#pragma unroll
    for (int j = 0; j < NGA; j++) {
        const uint32_t cidx = t + j * NT;
        offA[j] = ((cidx >> 5) << 5) + (cidx & 31);
    }
    const uint32_t offB = (warp << (G2 + 5)) + lane;  // (digit warp*16, lane); element b sits (b << 5) words further

    // twiddles of this tile position: phase A in registers, phase B (lane-only) in shared memory
    TWT twa[15];
    TWT twa2[MODE == TURN ? 15 : 1];
    {
        const uint32_t M = 1u << (p + 4);
#pragma unroll
        for (int j = 0; j < NGA; j++) {
            const uint32_t cidx = t + j * NT;
            const uint32_t m = ((cidx >> 5) << p) | (lo_base + (cidx & 31));
#pragma unroll
            for (int q = 1; q < RA; q++) {
                twa[j * (RA - 1) + q - 1] = A.twA[(uint64_t)(q - 1) * M + m];
                if (MODE == TURN) twa2[MODE == TURN ? j * (RA - 1) + q - 1 : 0] = A.twA2[(uint64_t)(q - 1) * M + m];
            }
        }
        for (uint32_t i = t; i < 15 * LANES; i += NT) {
            s_twb[i / LANES][i % LANES] = A.twB[(uint64_t)(i / LANES) << p | (lo_base + i % LANES)];
            if (MODE == TURN) s_twb2[i / LANES][i % LANES] = A.twB2[(uint64_t)(i / LANES) << p | (lo_base + i % LANES)];
        }
        if (MODE == TURN) {  // coset scale factors of this tile position, same for every column
            const uint64_t n = 1ull << A.log_n;
            const uint64_t base = ((uint64_t)hi << (p + g)) | lo_base;
            for (uint32_t i = t; i < 2 * TILE; i += NT) {
                const uint32_t h = i / TILE, w = i % TILE;
                s_pw[i] = __ldg(A.pw + (uint64_t)h * n + base + ((uint64_t)(w >> 5) << p) + (w & 31));
            }
        }
        __syncthreads();
    }
    auto twA = [&](int i) { return twa[i]; };
    auto twB = [&](int i) { return s_twb[i][lane]; };
    auto twA2 = [&](int i) { return twa2[MODE == TURN ? i : 0]; };
    auto twB2 = [&](int i) { return s_twb2[i][lane]; };

    // ---- register <-> shared moves in the two phase layouts ----------------------------------------------------------------
    auto lds_A = [&](const uint32_t* s, uint32_t* r) {
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) r[j * RA + a] = s[offA[j] + (a << (G2 + 5))];
    };
    auto sts_A = [&](uint32_t* s, const uint32_t* r) {
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) s[offA[j] + (a << (G2 + 5))] = r[j * RA + a];
    };
    auto lds_B = [&](const uint32_t* s, uint32_t* r) {
#pragma unroll
        for (int b = 0; b < 16; b++) r[b] = s[offB + (b << 5)];
    };
    auto sts_B = [&](uint32_t* s, const uint32_t* r) {
#pragma unroll
        for (int b = 0; b < 16; b++) s[offB + (b << 5)] = r[b];
    };
    // the warp's 16 digits x 32 lanes of tile s (forward layout) -> column `col` of the output view
    auto store_warp_rows = [&](const uint32_t* s, uint32_t col) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_4d(&tm_out, s + (warp << (G2 + 5)), (int)lo_base, (int)(warp << G2), (int)hi, (int)col);
            bulk_commit();
        }
    };

    uint32_t v[16];
    for (uint32_t i = 0; i < ncol; i++) {
        const uint32_t c = c_begin + i;
        uint32_t* const S = s_slot + (i % NSLOT) * TILE;
        if (MODE == TURN && lane == 0) bulk_wait_read<1>();  // own stores of column i-1 (slot) and i-2 (exchange tile) have left shared memory
        while (!mbar_try_wait(&full[i % NSLOT], (i / NSLOT) & 1)) {
        }
        if constexpr (MODE == FWD) {
            lds_A(S, v);
            phase<false, G1, true>(v, twA);
            sts_A(S, v);
            __syncthreads();
            if (t == 0 && i + AHEAD < ncol) issue(i + AHEAD);  // slot of column i-2: every warp waited for its store before the barrier
            lds_B(S, v);
            phase<false, G2, true>(v, twB);
            sts_B(S, v);
            store_warp_rows(S, c);
            if (lane == 0) bulk_wait_read<1>();  // store of column i-1 has left its slot
        } else if constexpr (MODE == INV) {
            lds_B(S, v);
            phase<true, G2, true>(v, twB);
            sts_B(S, v);
            __syncthreads();
            if (t == 0 && i + AHEAD < ncol) issue(i + AHEAD);
            lds_A(S, v);
            phase<true, G1, true>(v, twA);
            sts_A(S, v);
            fence_async_smem();
            __syncthreads();
            if (t == 0) {  // the inverse layout scatters a warp's digits over the tile: one store of the whole tile
                tma_store_4d(&tm_out, S, (int)lo_base, 0, (int)hi, (int)c);
                bulk_commit();
                bulk_wait_read<1>();
            }
        } else {
            uint32_t* const E = s_ex + (i & 1) * TILE;
            lds_B(S, v);
            phase<true, G2, true>(v, twB);
            sts_B(S, v);
            __syncthreads();
            if (t == 0 && i + AHEAD < ncol) issue(i + AHEAD);  // other slot: its store was waited for at the top of this iteration
            lds_A(S, v);
            phase<true, G1, true>(v, twA);
            uint32_t x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = v[k];
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t* pw = s_pw + h * TILE;
                uint32_t* const X = h ? E : S;  // h = 0 continues in place; h = 1 needs a tile of its own (S is being stored)
#pragma unroll
                for (int j = 0; j < NGA; j++)
#pragma unroll
                    for (int a = 0; a < RA; a++) v[j * RA + a] = kb::mul(x[j * RA + a], pw[offA[j] + (a << (G2 + 5))]);
                phase<false, G1, true>(v, twA2);
                sts_A(X, v);
                __syncthreads();
                lds_B(X, v);
                phase<false, G2, true>(v, twB2);
                sts_B(X, v);
                store_warp_rows(X, 2 * c + h);
            }
        }
    }
    if (lane == 0) bulk_wait_all();  // shared memory must outlive the stores
}


// ---- ingest fused with the first inverse pass ---------------------------------------------------------------------------------
// `Pcs::commit` receives row-major matrices (RowMajorMatrix<Val>, reference crates/stark/src/prover.rs:227); the device layout is
// column-major with bit-reversed rows.  The stand-alone transpose (nttk::k_ingest_wide: 8 B of HBM traffic per element) and the
// first inverse pass (contiguous, stage bits [0, g): another 8 B) both stream the whole matrix; here one kernel does both:
//   * a tile is 2^g stored positions x 32 COLUMNS.  Stored position base + d holds natural row brev(base + d), so the tile's rows
//     are the 2^g natural rows  k * 2^(log_n - g) + brev(base >> g),  k < 2^g: a TMA box (32 columns, 1, 2^g) of the 3-d view
//     [k][row low bits][column] of the caller's matrix, landing as a dense [k][32] tile with d = brev_g(k);
//   * both register phases run in place on that tile exactly like the contiguous pass of ntt2::k_pass (lane = column);
//   * the result goes through a padded [column][2^g + 1] staging tile so that each warp stores whole 128-byte runs of a column.
// CANON: caller words are canonical residues (converted to Montgomery form on load) instead of Montgomery words.
struct IngestArgs {
    uint32_t* dst;        // column-major coefficients-in-progress: column c at dst + c * n
    uint32_t ncols;       // columns of the matrix
    uint32_t tiles_per_cta;   // tiles (of 2^g stored positions) a CTA walks through for its 32-column block
    uint32_t log_n;
    const TWT* twA;       // phase-A table of the contiguous inverse pass: [(2^G1 - 1)][16]
};
constexpr int INGEST_SLOTS = 2;  // 2 x 32 KB + 33 KB of staging at g = 8: two CTAs of 512 threads per SM
constexpr size_t ingest_smem_bytes(int G1) {
    return (size_t)INGEST_SLOTS * ((size_t)4 << (G1 + 4 + 5)) + (size_t)32 * ((1 << (G1 + 4)) + 1) * 4 + INGEST_SLOTS * sizeof(uint64_t) + 16;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ constexpr uint32_t brev_c(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <int G1, bool CANON>
__global__ void __launch_bounds__(1 << (G1 + 5), 1024 >> (G1 + 5)) k_ingest_pass(const __grid_constant__ CUtensorMap tm_src, IngestArgs A) {
    static_assert(G1 >= 1 && G1 <= 4, "two register phases");
    constexpr int g = G1 + G2, NT = 1 << (g + 1), ND = 1 << g;
    constexpr int RA = 1 << G1, NGA = 16 >> G1;
    constexpr int TILE = ND * LANES, PITCH = ND + 1;
    uint32_t* const s_slot = smem3;                        // [INGEST_SLOTS][TILE]
    uint32_t* const s_out = smem3 + INGEST_SLOTS * TILE;   // [32][PITCH]
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem3 + INGEST_SLOTS * TILE + ((32 * PITCH + 3) & ~3));

    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint64_t n = 1ull << A.log_n;
    // A CTA keeps ONE 32-column block (blockIdx.x) and walks through tiles_per_cta tiles.  Tile index = low bits of the natural rows,
    // so the CTAs dispatched together (consecutive blockIdx.x) read the neighbouring 128-byte segments of the SAME rows of the
    // row-major source at about the same time (one DRAM page per row instead of one per segment), and consecutive iterations read
    // consecutive rows; the tile's stored position is brev(tile), which costs nothing: a tile is written as whole 1 KB runs.
    const uint32_t ntiles_all = 1u << (A.log_n - g);
    const uint32_t t_begin = blockIdx.y * A.tiles_per_cta;
    const uint32_t t_end = min(ntiles_all, t_begin + A.tiles_per_cta);
    if (t_begin >= t_end) return;
    const uint32_t nblk = t_end - t_begin;  // iterations
    const uint32_t col0 = blockIdx.x << 5;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < INGEST_SLOTS; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();
    auto issue = [&](uint32_t i) {
        const uint32_t s = i % INGEST_SLOTS;
        mbar_expect_tx(&full[s], TILE * 4);
        tma_load_3d(s_slot + s * TILE, &tm_src, &full[s], (int)col0, (int)(t_begin + i), 0);
    };
    if (t == 0) {
#pragma unroll
        for (uint32_t s = 0; s < (uint32_t)INGEST_SLOTS; s++)
            if (s < nblk) issue(s);
    }
    // phase B: thread (aB = warp, lane) owns digits (aB << 4) | b  = tile rows  brev4(b) << G1 | brev_G1(aB)
    const uint32_t offB = ((__brev(warp) >> (32 - G1)) << 5) + lane;
    // phase A: combo j: r1 = (t + j NT) >> 5 owns digits (a << 4) | r1 = tile rows  brev4(r1) << G1 | brev_G1(a)
    uint32_t offA[NGA], r1A[NGA];
#pragma unroll
    for (int j = 0; j < NGA; j++) {
        r1A[j] = (t + j * NT) >> 5;
        offA[j] = (((__brev(r1A[j]) >> 28) << G1) << 5) + lane;
    }
    TWT twa[15];
#pragma unroll
    for (int j = 0; j < NGA; j++)
#pragma unroll
        for (int q = 1; q < RA; q++) twa[j * (RA - 1) + q - 1] = A.twA[(q - 1) * 16 + r1A[j]];
    auto twA = [&](int i) { return twa[i]; };
    auto twB = [&](int) { return TWT(); };

    uint32_t v[16];
    for (uint32_t i = 0; i < nblk; i++) {
        uint32_t* const S = s_slot + (i % INGEST_SLOTS) * TILE;
        while (!mbar_try_wait(&full[i % INGEST_SLOTS], (i / INGEST_SLOTS) & 1)) {
        }
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const uint32_t x = S[offB + ((brev_c(b, 4) << G1) << 5)];
            v[b] = CANON ? kb::to_mont(x) : x;
        }
        phase<true, G2, false>(v, twB);
#pragma unroll
        for (int b = 0; b < 16; b++) S[offB + ((brev_c(b, 4) << G1) << 5)] = v[b];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) v[j * RA + a] = S[offA[j] + (brev_c(a, G1) << 5)];
        phase<true, G1, true>(v, twA);
        // staging: [column = lane][digit], pitch 2^g + 1: conflict free here and in the row reads below
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) s_out[lane * PITCH + ((uint32_t)a << G2) + r1A[j]] = v[j * RA + a];
        __syncthreads();
        if (t == 0 && i + INGEST_SLOTS < nblk) issue(i + INGEST_SLOTS);  // the slot was last read before the barrier above
        // warp w stores columns w, w + NT/32, ... of the block: runs of 32 consecutive words
        const uint32_t tile = t_begin + i;
        const uint64_t base = (uint64_t)(A.log_n > (unsigned)g ? __brev(tile) >> (32 - (A.log_n - g)) : 0u) << g;  // first stored position of the tile
#pragma unroll
        for (int cc = 0; cc < 32 / (NT / 32); cc++) {
            const uint32_t c = warp + cc * (NT / 32);
            if (col0 + c < A.ncols) {
                uint32_t* o = A.dst + (uint64_t)(col0 + c) * n + base + lane;
#pragma unroll
                for (int k = 0; k < ND / 32; k++) o[k * 32] = s_out[c * PITCH + k * 32 + lane];
            }
        }
    }
}

// ---- contiguous forward pass (stage bits [0, g): the LAST pass of a forward transform) fed by bulk copies ----------------------------
// A tile is 32 consecutive runs of 2^g words = one contiguous chunk of a column.  Each run travels by its own 1-d bulk copy
// (`cp.async.bulk`, 2^g * 4 bytes, issued by lane `run` of warp 0) into a slot whose runs are 2^g + 4 words apart: with that pitch
//   * phase A (thread = 4 low digits x 8 runs per warp; element a of the thread 16 words further) and
//   * phase B (thread = 16 consecutive words of one run, 32 runs per warp, four 128-bit accesses)
// are both free of bank conflicts, and both phases work in place.  The result leaves by 32 bulk copies of whole runs.
// Ring of three slots: current, prefetched, being stored; the slot being stored is refilled one iteration later by the same lanes
// that stored from it (each lane waits for its own store only).
struct CfwdArgs {
    uint32_t* data;       // first column, transformed in place
    uint64_t col_stride;  // words between columns
    uint32_t ncols;
    uint32_t cols_per_cta;
    const TWT* twA;       // phase-A table of the contiguous forward pass: [(2^G1 - 1)][16]
    // Sharded commitment (dist_commit.cuh): the columns are the 2W half-columns [W][2][n] of an LDE block and the results do not go
    // back in place but straight to the rank that owns their ROWS: stored row R = (hc & 1) * n + position of LDE column hc >> 1 goes to
    // dst[R >> log_rpg] + ((dcol0 + (hc >> 1)) << log_rpg) + (R & (2^log_rpg - 1)) — peer HBM mapped over NVLink, written by the same
    // asynchronous bulk stores (a run of 2^g rows never straddles two ranks: log_rpg >= g).  scatter = 0: in place.
    uint32_t scatter, log_n, log_rpg, dcol0;
    uint32_t* dst[16];
};
constexpr int CFWD_SLOTS = 3;
constexpr size_t cfwd_smem_bytes(int G1) { return (size_t)CFWD_SLOTS * 32 * ((1 << (G1 + 4)) + 4) * 4 + 15 * 16 * sizeof(TWT) + CFWD_SLOTS * sizeof(uint64_t); }
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int G1>
__global__ void __launch_bounds__(1 << (G1 + 5), 1024 >> (G1 + 5)) k_cfwd(CfwdArgs A) {
    static_assert(G1 >= 1 && G1 <= 4, "two register phases");
    constexpr int g = G1 + G2, NT = 1 << (g + 1), ND = 1 << g;
    constexpr int RA = 1 << G1, NGA = 16 >> G1;
    constexpr int PITCH = ND + 4, SLOT = 32 * PITCH;
    uint32_t* const s_slot = smem3;                                                      // [CFWD_SLOTS][32][PITCH]
    TWT(*const s_tw)[16] = reinterpret_cast<TWT(*)[16]>(smem3 + CFWD_SLOTS * SLOT);      // [15][16]
    uint64_t* const full = reinterpret_cast<uint64_t*>(s_tw + 15);                        // [CFWD_SLOTS]

    const uint32_t t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint64_t base = (uint64_t)blockIdx.x << (g + 5);  // first word of the chunk inside a column
    const uint32_t c_begin = blockIdx.y * A.cols_per_cta;
    const uint32_t c_end = min(A.ncols, c_begin + A.cols_per_cta);
    if (c_begin >= c_end) return;
    const uint32_t ncol = c_end - c_begin;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < CFWD_SLOTS; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    for (uint32_t i = t; i < (RA - 1) * 16; i += NT) s_tw[i >> 4][i & 15] = A.twA[i];
    __syncthreads();
    auto issue = [&](uint32_t i) {  // warp 0, all lanes: run `lane` of the i-th column of this CTA into slot i % CFWD_SLOTS
        const uint32_t s = i % CFWD_SLOTS;
        if (lane == 0) mbar_expect_tx(&full[s], 32 * ND * 4);
        __syncwarp();
        bulk_load_1d(s_slot + s * SLOT + lane * PITCH, A.data + (uint64_t)(c_begin + i) * A.col_stride + base + (lane << g), ND * 4, &full[s]);
    };
    if (warp == 0) {
        if (0 < ncol) issue(0);
        if (1 < ncol) issue(1);
    }
    // phase A combos: cidx = t + j NT -> digit low part r1 = (cidx >> 5 & 3) * 4 + (cidx & 3), run = (cidx >> 7) * 8 + (cidx >> 2 & 7)
    uint32_t offA[NGA], r1A[NGA];
#pragma unroll
    for (int j = 0; j < NGA; j++) {
        const uint32_t cidx = t + j * NT;
        r1A[j] = (((cidx >> 5) & 3) << 2) | (cidx & 3);
        offA[j] = (((cidx >> 7) << 3) | ((cidx >> 2) & 7)) * PITCH + r1A[j];
    }
    const uint32_t offB = lane * PITCH + (warp << 4);  // phase B: run = lane, digits warp * 16 + b
    uint32_t v[16];
    for (uint32_t i = 0; i < ncol; i++) {
        uint32_t* const S = s_slot + (i % CFWD_SLOTS) * SLOT;
        while (!mbar_try_wait(&full[i % CFWD_SLOTS], (i / CFWD_SLOTS) & 1)) {
        }
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) v[j * RA + a] = S[offA[j] + (a << G2)];
        // flat twiddle index of ntt2::phase = group * (RA - 1) + (q - 1); the table is indexed by the group's low digit part
        phase<false, G1, true>(v, [&](int k) { return s_tw[k % (RA - 1)][r1A[k / (RA - 1)]]; });
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) S[offA[j] + (a << G2)] = v[j * RA + a];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 x = *reinterpret_cast<const uint4*>(S + offB + 4 * k);
            v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
        phase<false, G2, false>(v, [&](int) { return TWT(); });
#pragma unroll
        for (int k = 0; k < 4; k++) *reinterpret_cast<uint4*>(S + offB + 4 * k) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            uint32_t* o = A.data + (uint64_t)(c_begin + i) * A.col_stride + base + (lane << g);
            if (A.scatter) {
                const uint32_t hc = c_begin + i;
                const uint64_t R = ((uint64_t)(hc & 1) << A.log_n) + base + (lane << g);
                o = A.dst[R >> A.log_rpg] + ((uint64_t)(A.dcol0 + (hc >> 1)) << A.log_rpg) + (R & ((1ull << A.log_rpg) - 1));
            }
            bulk_store_1d(o, S + lane * PITCH, ND * 4);
            bulk_commit();
            bulk_wait_read<1>();  // this lane's store of column i-1 has left its slot: refill it with column i+2
            if (i + 2 < ncol) issue(i + 2);
        }
    }
    if (warp == 0) bulk_wait_all();
}

}  // namespace ntt3
