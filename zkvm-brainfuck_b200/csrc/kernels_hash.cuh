// kernels_hash.cuh — Poseidon2 Merkle kernels (K2/K3 of SURVEY.md §2): one thread per leaf / node,
// the 16-word state never leaves registers.
//
// Replaces MerkleTreeMmcs::commit as reached from Pcs::commit (reference
// crates/stark/src/prover.rs:227,334,411; crates/stark/src/machine.rs:196): first digest layer =
// PaddingFreeSponge over the concatenated same-height rows, upper layers = TruncatedPermutation
// 2-to-1 compression with injection of shorter matrices' row digests.
//
// Device layout: matrices are COLUMN-major (column c of a matrix with `rows` rows starts at
// base + c*rows), so thread r reading column c at row r is perfectly coalesced; digest layers are
// arrays of 8-word digests (32 B, loaded/stored as 2 x uint4).
#pragma once
#include "poseidon2.cuh"

namespace hashk {

constexpr int HASH_THREADS = 128;  // (256: leaf hash 46.59 instead of 46.80 ms at 2^23 x 32 permutations: not worth a second instantiation)
static_assert(HASH_THREADS >= 128, "the four-lane layer kernels stage the 128 external round constants with one thread each");

// absorb the r-th row of the column list (overwrite-mode sponge, rate 8) into state s
// (r is the ELEMENT offset of the row inside each column: row index * row stride)
__device__ __forceinline__ void sponge_rows(uint32_t (&s)[16], const uint32_t* const* __restrict__ colptr, uint32_t ncols,
                                            uint64_t r) {
    // single permute() call site (the permutation body must stay small enough for the I-cache);
    // the last block may be partial: it overwrites only its prefix of the state
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < ncols; c0 += 8) {
        if (c0 + 8 <= ncols) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = __ldg(colptr[c0 + k] + r);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (c0 + k < ncols) s[k] = __ldg(colptr[c0 + k] + r);
        }
        p2::permute(s);
    }
}

__device__ __forceinline__ void store_digest(uint32_t* out, const uint32_t (&s)[16]) {
    uint4* o = reinterpret_cast<uint4*>(out);
    o[0] = make_uint4(s[0], s[1], s[2], s[3]);
    o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// first digest layer: digests[r] = sponge(row r of every column in colptr)
// row_stride: words between consecutive rows inside a column (1 for column-major matrices, the row
// length for row-major ones such as the FRI layer matrices)
__global__ void __launch_bounds__(HASH_THREADS) k_leaf_hash(const uint32_t* const* __restrict__ colptr, uint32_t ncols, uint64_t rows,
                                                            uint32_t* __restrict__ digests, uint32_t row_stride) {
    uint64_t r = blockIdx.x * (uint64_t)HASH_THREADS + threadIdx.x;
    if (r >= rows) return;
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = 0;
    sponge_rows(s, colptr, ncols, r * row_stride);
    store_digest(digests + 8 * r, s);
}

// Incremental form of k_leaf_hash for the pipelined commit: absorbs `ncols` more columns (a multiple of 8 unless
// `last`) of a column-major block into the per-row sponge states.  state: 16 word-planes of `rows` words (plane k
// holds word k of every row, so thread r's accesses are coalesced); first: start from the zero state; last: write
// the digest instead of the state.
__global__ void __launch_bounds__(HASH_THREADS) k_leaf_absorb(const uint32_t* __restrict__ cols, uint64_t col_stride, uint32_t ncols, uint64_t rows,
                                                              uint32_t* __restrict__ state, int first, int last, uint32_t* __restrict__ digests) {
    uint64_t r = blockIdx.x * (uint64_t)HASH_THREADS + threadIdx.x;
    if (r >= rows) return;
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = first ? 0u : state[(uint64_t)k * rows + r];
    const uint32_t* p = cols + r;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < ncols; c0 += 8, p += 8 * col_stride) {
        if (c0 + 8 <= ncols) {
#pragma unroll
            for (int k = 0; k < 8; k++) s[k] = __ldg(p + k * col_stride);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (c0 + k < ncols) s[k] = __ldg(p + k * col_stride);
        }
        p2::permute(s);
    }
    if (last) store_digest(digests + 8 * r, s);
    else {
#pragma unroll
        for (int k = 0; k < 16; k++) state[(uint64_t)k * rows + r] = s[k];
    }
}

// next layer: out[i] = compress(prev[2i], prev[2i+1]); if ncols > 0 additionally
// out[i] = compress(out[i], sponge(row i of the injected columns))
__global__ void __launch_bounds__(HASH_THREADS) k_compress_layer(const uint32_t* __restrict__ prev, uint32_t* __restrict__ out, uint64_t len,
                                                                 const uint32_t* const* __restrict__ colptr, uint32_t ncols, uint32_t row_stride) {
    uint64_t i = blockIdx.x * (uint64_t)HASH_THREADS + threadIdx.x;
    if (i >= len) return;
    uint32_t s[16];
    const uint4* in = reinterpret_cast<const uint4*>(prev + 16 * i);
    uint4 a = in[0], b = in[1], c = in[2], d = in[3];
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    s[8] = c.x; s[9] = c.y; s[10] = c.z; s[11] = c.w;
    s[12] = d.x; s[13] = d.y; s[14] = d.z; s[15] = d.w;
    p2::permute(s);
    if (ncols) {
        uint32_t h[16];
#pragma unroll
        for (int k = 0; k < 16; k++) h[k] = 0;
        sponge_rows(h, colptr, ncols, i * row_stride);
#pragma unroll
        for (int k = 0; k < 8; k++) s[8 + k] = h[k];
        p2::permute(s);
    }
    store_digest(out + 8 * i, s);
}

// Same layer step with FOUR lanes per node (p2::permute_x4): for layers too narrow to fill the machine (a 2^14-node layer
// occupies 5 % of the thread slots) the launch costs one permutation LATENCY, and the four-lane form has a third of the
// dependent chain (ncu: every one-thread layer launch between 2^10 and 2^16 nodes takes ~10.5 us whatever its width).
__global__ void __launch_bounds__(HASH_THREADS) k_compress_layer_x4(const uint32_t* __restrict__ prev, uint32_t* __restrict__ out, uint64_t len,
                                                                    const uint32_t* const* __restrict__ colptr, uint32_t ncols, uint32_t row_stride) {
    __shared__ uint32_t s_ext[8 * 16], s_int[16];
    const uint32_t t = threadIdx.x;
    if (t < 128) s_ext[t] = p2::c_p2.ext_s[t >> 4][t & 15];
    if (t < 16) s_int[t] = p2::c_p2.internal_s[t];
    __syncthreads();
    const int q = t & 3;
    const p2::X4 xc = p2::x4_setup(s_ext, s_int, q);
    const uint64_t i = (blockIdx.x * (uint64_t)HASH_THREADS + t) >> 2;
    const bool on = i < len;
    if (((blockIdx.x * (uint64_t)HASH_THREADS + (t & ~31u)) >> 2) >= len) return;  // whole warp past the end
    uint32_t w[4] = {0, 0, 0, 0};
    if (on) {
        uint4 x = *reinterpret_cast<const uint4*>(prev + 16 * i + 4 * q);
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
    }
    p2::permute_x4(w, xc, q);
    if (ncols) {
        uint32_t h[4] = {0, 0, 0, 0};
        for (uint32_t c0 = 0; c0 < ncols; c0 += 8) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t c = c0 + 4 * q + j;
                if (on && q < 2 && c < ncols) h[j] = __ldg(colptr[c] + i * row_stride);
            }
            p2::permute_x4(h, xc, q);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t from_low = __shfl_xor_sync(0xffffffffu, h[j], 2);
            if (q >= 2) w[j] = from_low;
        }
        p2::permute_x4(w, xc, q);
    }
    if (on && q < 2) *reinterpret_cast<uint4*>(out + 8 * i + 4 * q) = make_uint4(w[0], w[1], w[2], w[3]);
}

// Top of the tree in ONE launch: starting from a layer of `len0` <= TOP_MAX digests, computes every remaining layer
// down to the root inside a single CTA (the layer being consumed sits in shared memory), writing each layer to its
// global array.  Removes ~10 launches per tree, which dominate small proofs and the FRI commit phase.
constexpr int TOP_MAX = 1024, TOP_THREADS = 1024, TOP_LEVELS = 10;
struct TopArgs {
    const uint32_t* in;                  // len0 digests
    uint32_t* out[TOP_LEVELS];            // out[k]: len0 >> (k+1) digests
    const uint32_t* const* colptr[TOP_LEVELS];  // injected columns of layer k (or null)
    uint32_t ncols[TOP_LEVELS];
    uint32_t row_stride[TOP_LEVELS];
    uint32_t len0;
    uint32_t nlevels;
    uint32_t x1_max;  // cluster kernel: levels with at most this many nodes per CTA use one thread per node (0: always four lanes)
};
// Levels with at most TOP_THREADS / 4 nodes run the four-lane permutation (p2::permute_x4): a level costs one
// permutation LATENCY whatever its width, and the four-lane form has a third of the dependent chain.
__global__ void __launch_bounds__(TOP_THREADS) k_compress_top(TopArgs A) {
    __shared__ __align__(16) uint32_t cur[TOP_MAX * 8];
    __shared__ uint32_t s_ext[8 * 16], s_int[16];
    const uint32_t t = threadIdx.x;
    for (uint32_t i = t; i < A.len0 * 2; i += TOP_THREADS) reinterpret_cast<uint4*>(cur)[i] = reinterpret_cast<const uint4*>(A.in)[i];
    if (t < 128) s_ext[t] = p2::c_p2.ext_s[t >> 4][t & 15];
    if (t < 16) s_int[t] = p2::c_p2.internal_s[t];
    __syncthreads();
    const int q = t & 3;
    const p2::X4 xc = p2::x4_setup(s_ext, s_int, q);
    uint32_t len = A.len0;
    for (uint32_t k = 0; k < A.nlevels; k++) {
        len >>= 1;
        if (4 * len <= TOP_THREADS) {
            // ---- four lanes per node ----
            const uint32_t i = t >> 2;
            const bool warp_on = ((t & ~31u) >> 2) < len, on = i < len;  // whole warps without a node skip the level
            uint32_t w[4] = {0, 0, 0, 0};
            if (on) {
                uint4 x = *reinterpret_cast<const uint4*>(cur + 16 * i + 4 * q);
                w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
            }
            __syncthreads();  // everyone has read its pair before the layer is overwritten
            if (warp_on) {
                p2::permute_x4(w, xc, q);
                if (A.ncols[k]) {
                    // digest of the injected row i: overwrite-mode sponge, rate 8 = the words of lanes q = 0, 1
                    uint32_t h[4] = {0, 0, 0, 0};
                    const uint32_t nc = A.ncols[k];
                    for (uint32_t c0 = 0; c0 < nc; c0 += 8) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            uint32_t c = c0 + 4 * q + j;
                            if (on && q < 2 && c < nc) h[j] = __ldg(A.colptr[k][c] + (uint64_t)i * A.row_stride[k]);
                        }
                        p2::permute_x4(h, xc, q);
                    }
                    // state = (compressed children, row digest): lanes 2, 3 take the digest words of lanes 0, 1
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t from_low = __shfl_xor_sync(0xffffffffu, h[j], 2);
                        if (q >= 2) w[j] = from_low;
                    }
                    p2::permute_x4(w, xc, q);
                }
                if (on && q < 2) {
                    uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(A.out[k] + 8 * i + 4 * q) = o;
                    *reinterpret_cast<uint4*>(cur + 8 * i + 4 * q) = o;
                }
            }
            __syncthreads();
            continue;
        }
        // ---- one thread per node (first level of a full-size top: 512 nodes) ----
        uint32_t s[16];
        const uint32_t i = t;
        if (i < len) {
#pragma unroll
            for (int w = 0; w < 16; w++) s[w] = cur[16 * i + w];
        }
        __syncthreads();  // everyone has read its pair before the layer is overwritten
        if (i < len) {
            p2::permute(s);
            if (A.ncols[k]) {
                uint32_t h[16];
#pragma unroll
                for (int w = 0; w < 16; w++) h[w] = 0;
                sponge_rows(h, A.colptr[k], A.ncols[k], (uint64_t)i * A.row_stride[k]);
#pragma unroll
                for (int w = 0; w < 8; w++) s[8 + w] = h[w];
                p2::permute(s);
            }
            store_digest(A.out[k] + 8 * i, s);
#pragma unroll
            for (int w = 0; w < 8; w++) cur[8 * i + w] = s[w];
        }
        __syncthreads();
    }
}

// ---- the same top of the tree on a CLUSTER of 8 CTAs (8 SMs) -------------------------------------------------------------------------
// k_compress_top runs on ONE SM, and its wide levels are bound by that SM's issue rate, not by latency: 512 + 256 + 128 permutations
// are ~26 us of the ~40 us a 1024-digest top takes (ncu launch list of a small proof: 14 launches x 85 us = 31 % of the proof).
// Here CTA r of an 8-CTA thread-block cluster reduces digests [r * len0/8, (r+1) * len0/8) to one digest with the four-lane
// permutation (every level of its subtree fits 256 threads), the cluster synchronises once (barrier.cluster, release/acquire at cluster
// scope, so the digests written to global memory are visible), and CTA 0 finishes the last three levels.  Each level is written to
// its global array exactly as k_compress_top does.
constexpr int TOPC_CLUSTER = 8, TOPC_THREADS = 256, TOPC_MIN_LEN0 = 256;
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// one level with four lanes per node: nodes [0, len) of this CTA are global nodes node0 + i of level k; cur holds the 2 len children
__device__ __forceinline__ void top_level_x4(const TopArgs& A, uint32_t k, uint32_t* cur, uint32_t len, uint32_t node0, const p2::X4& xc, int q, uint32_t t) {
    const uint32_t i = t >> 2;
    const bool warp_on = ((t & ~31u) >> 2) < len, on = i < len;  // whole warps without a node skip the level
    uint32_t w[4] = {0, 0, 0, 0};
    if (on) {
        uint4 x = *reinterpret_cast<const uint4*>(cur + 16 * i + 4 * q);
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
    }
    __syncthreads();  // everyone has read its pair before the layer is overwritten
    if (warp_on) {
        p2::permute_x4(w, xc, q);
        if (A.ncols[k]) {
            uint32_t h[4] = {0, 0, 0, 0};
            const uint32_t nc = A.ncols[k];
            for (uint32_t c0 = 0; c0 < nc; c0 += 8) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t c = c0 + 4 * q + j;
                    if (on && q < 2 && c < nc) h[j] = __ldg(A.colptr[k][c] + (uint64_t)(node0 + i) * A.row_stride[k]);
                }
                p2::permute_x4(h, xc, q);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t from_low = __shfl_xor_sync(0xffffffffu, h[j], 2);
                if (q >= 2) w[j] = from_low;
            }
            p2::permute_x4(w, xc, q);
        }
        if (on && q < 2) {
            uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(A.out[k] + 8 * (uint64_t)(node0 + i) + 4 * q) = o;
            *reinterpret_cast<uint4*>(cur + 8 * i + 4 * q) = o;
        }
    }
    __syncthreads();
}
// one level with ONE thread per node (len <= TOPC_THREADS): no shuffles; with at most one warp per scheduler the dependent chain of
// the plain permutation is shorter than the four-lane form's (whose every round pays two shuffle latencies)
__device__ __forceinline__ void top_level_x1(const TopArgs& A, uint32_t k, uint32_t* cur, uint32_t len, uint32_t node0, uint32_t t) {
    uint32_t s[16];
    const bool on = t < len;
    if (on) {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint4 x = *reinterpret_cast<const uint4*>(cur + 16 * t + 4 * w);
            s[4 * w] = x.x; s[4 * w + 1] = x.y; s[4 * w + 2] = x.z; s[4 * w + 3] = x.w;
        }
    }
    __syncthreads();  // everyone has read its pair before the layer is overwritten
    if (on) {
        p2::permute(s);
        if (A.ncols[k]) {
            uint32_t h[16];
#pragma unroll
            for (int w = 0; w < 16; w++) h[w] = 0;
            sponge_rows(h, A.colptr[k], A.ncols[k], (uint64_t)(node0 + t) * A.row_stride[k]);
#pragma unroll
            for (int w = 0; w < 8; w++) s[8 + w] = h[w];
            p2::permute(s);
        }
        store_digest(A.out[k] + 8 * (uint64_t)(node0 + t), s);
        *reinterpret_cast<uint4*>(cur + 8 * t) = make_uint4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<uint4*>(cur + 8 * t + 4) = make_uint4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
}
__global__ void __cluster_dims__(TOPC_CLUSTER, 1, 1) __launch_bounds__(TOPC_THREADS) k_compress_top_cluster(TopArgs A) {
    __shared__ __align__(16) uint32_t cur[(TOP_MAX / TOPC_CLUSTER) * 8];
    __shared__ uint32_t s_ext[8 * 16], s_int[16];
    const uint32_t t = threadIdx.x, rank = cluster_cta_rank();
    const uint32_t seg = A.len0 / TOPC_CLUSTER;  // >= 32 digests per CTA, 4 * seg / 2 <= TOPC_THREADS
    for (uint32_t i = t; i < seg * 2; i += TOPC_THREADS) reinterpret_cast<uint4*>(cur)[i] = reinterpret_cast<const uint4*>(A.in + (uint64_t)rank * seg * 8)[i];
    if (t < 128) s_ext[t] = p2::c_p2.ext_s[t >> 4][t & 15];
    if (t < 16) s_int[t] = p2::c_p2.internal_s[t];
    __syncthreads();
    const int q = t & 3;
    const p2::X4 xc = p2::x4_setup(s_ext, s_int, q);
    uint32_t k = 0, len = seg, node0 = rank * seg;
    while (len > 1) {  // this CTA's subtree
        len >>= 1;
        node0 >>= 1;
        if (len <= A.x1_max) top_level_x1(A, k, cur, len, node0, t);
        else top_level_x4(A, k, cur, len, node0, xc, q, t);
        k++;
    }
    __threadfence();
    cluster_sync_all();
    if (rank != 0) return;
    // CTA 0: the TOPC_CLUSTER subtree roots (level k - 1) -> root.  They were written by other SMs: read around L1.
    if (t < TOPC_CLUSTER * 2) reinterpret_cast<uint4*>(cur)[t] = __ldcg(reinterpret_cast<const uint4*>(A.out[k - 1]) + t);
    __syncthreads();
    len = TOPC_CLUSTER;
    while (k < A.nlevels) {
        len >>= 1;
        if (len <= A.x1_max) top_level_x1(A, k, cur, len, 0, t);
        else top_level_x4(A, k, cur, len, 0, xc, q, t);
        k++;
    }
}

// n independent permutations, states row-major n x 16 (Montgomery form)
__global__ void __launch_bounds__(HASH_THREADS) k_permute_many(uint32_t* __restrict__ st, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)HASH_THREADS + threadIdx.x;
    if (i >= n) return;
    uint4* p = reinterpret_cast<uint4*>(st + 16 * i);
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    uint32_t s[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    p2::permute(s);
    p[0] = make_uint4(s[0], s[1], s[2], s[3]);
    p[1] = make_uint4(s[4], s[5], s[6], s[7]);
    p[2] = make_uint4(s[8], s[9], s[10], s[11]);
    p[3] = make_uint4(s[12], s[13], s[14], s[15]);
}

// compress n pairs: out[i] = permute(left[i] || right[i])[0..8]
__global__ void __launch_bounds__(HASH_THREADS) k_compress_pairs(const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                                                 uint32_t* __restrict__ out, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)HASH_THREADS + threadIdx.x;
    if (i >= n) return;
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        s[k] = left[8 * i + k];
        s[8 + k] = right[8 * i + k];
    }
    p2::permute(s);
    store_digest(out + 8 * i, s);
}

// elementwise representation change (canonical <-> Montgomery) for small host-facing buffers
__global__ void k_convert(uint32_t* __restrict__ data, uint64_t n, int to_mont) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    data[i] = to_mont ? kb::to_mont(data[i]) : kb::from_mont(data[i]);
}

// gather one row (index `row`) from a list of columns into a contiguous buffer
__global__ void k_gather_row(const uint32_t* const* __restrict__ colptr, const uint64_t* __restrict__ rowidx, uint32_t ncols,
                             uint32_t* __restrict__ out, int to_canonical) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    uint32_t v = colptr[c][rowidx[c]];
    out[c] = to_canonical ? kb::from_mont(v) : v;
}

}  // namespace hashk
