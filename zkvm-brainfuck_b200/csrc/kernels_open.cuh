// kernels_open.cuh — kernels behind `TwoAdicFriPcs::open` (K6-K9 of SURVEY.md §2).
//
// Replaces the CPU work of `pcs.open(...)` as called from the reference
// (crates/stark/src/prover.rs:460-470): barycentric evaluation of every committed column at the
// opening points, the per-height reduced openings, the FRI commit phase folds, the proof-of-work
// search and the query gathers.  Algorithms restated from Plonky3 p3-fri v0.1.0.
// All matrices are the column-major Montgomery LDEs with bit-reversed rows left on the device by
// `bfgpu_pcs_commit`; extension elements are 4 consecutive words.
#pragma once
#include "kb31.cuh"
#include "poseidon2.cuh"

namespace openk {

using kb::Ext;

__device__ __forceinline__ Ext ld_ext(const uint32_t* p) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    return Ext{{v.x, v.y, v.z, v.w}};
}
__device__ __forceinline__ void st_ext(uint32_t* p, Ext e) { *reinterpret_cast<uint4*>(p) = make_uint4(e.c[0], e.c[1], e.c[2], e.c[3]); }

// w_{2^log_n}^j for j < 2^log_n from the table tw[e] = w_{2^24}^e, e < 2^23
__device__ __forceinline__ uint32_t root_pow(const uint32_t* __restrict__ tw, unsigned log_n, uint32_t j) {
    if (log_n == 0) return kb::ONE;
    uint32_t half = 1u << (log_n - 1);
    uint32_t jj = j & (half - 1);
    uint32_t v = __ldg(tw + ((uint64_t)jj << (kb::TWO_ADICITY - log_n)));
    return (j & half) ? kb::neg(v) : v;
}

// ---- barycentric weights: w[r] = g^{br(r)} / (z - s g^{br(r)}),  r < h = 2^log_h ---------------------
// (rows of the low coset of a stored LDE are in bit-reversed order; s = coset shift, Montgomery)
// rows [r0, r0 + count) of the coset only (w[0] = weight of row r0): a rank of the sharded prover evaluates over its own row shard
__global__ void k_bary_weights(uint32_t* __restrict__ w, unsigned log_h, uint32_t shift, Ext z, const uint32_t* __restrict__ tw, uint32_t r0,
                               uint32_t count) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint32_t r = r0 + k;
    uint32_t g = root_pow(tw, log_h, kb::bitrev(r, log_h));
    Ext d = z;
    d.c[0] = kb::sub(d.c[0], kb::mul(shift, g));
    st_ext(w + 4 * (uint64_t)k, kb::ext_scale(kb::ext_inv(d), g));
}

// partial[(c * nchunks + chunk) * NP + t] = sum over the chunk's rows of col_c[r] * w_t[r]
// Products are accumulated unreduced in 64 bits (kb::mac) and reduced once per thread.
#ifndef BFGPU_BARY_COLS
#define BFGPU_BARY_COLS 4
#endif
#ifndef BFGPU_BARY_UNROLL
#define BFGPU_BARY_UNROLL 2
#endif
constexpr int BARY_THREADS = 256, BARY_ROWS = 4096, BARY_COLS = BFGPU_BARY_COLS, BARY_UNROLL = BFGPU_BARY_UNROLL;  // BARY_COLS columns share each pair of 16-byte weights
template <int NP>
__device__ __forceinline__ void bary_dot_body(const uint32_t* __restrict__ mat, uint64_t col_stride, uint32_t ncols, uint32_t h,
                                              const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1, uint32_t* __restrict__ partial,
                                              uint32_t nchunks, const uint32_t chunk, const uint32_t c0) {
    uint64_t acc[BARY_COLS][NP][4];
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
        for (int t = 0; t < NP; t++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[c][t][k] = 0;
    uint32_t r_end = min(h, (chunk + 1) * BARY_ROWS);
    const uint32_t* colp[BARY_COLS];
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++) colp[c] = mat + (uint64_t)min(c0 + c, ncols - 1) * col_stride;  // tail columns recompute the last one
    // two rows per iteration: two rows of loads in flight (the kernel was latency bound: ncu long_scoreboard ~10 cycles / issue) and
    // their products share one conditional subtraction (kb::mac2)
    uint32_t r = chunk * BARY_ROWS + threadIdx.x;
    for (; r + BARY_THREADS < r_end; r += 2 * BARY_THREADS) {
        const uint32_t r2 = r + BARY_THREADS;
        Ext wa[NP], wb[NP];
        wa[0] = ld_ext(w0 + 4 * (uint64_t)r);
        wb[0] = ld_ext(w0 + 4 * (uint64_t)r2);
        if (NP > 1) {
            wa[NP - 1] = ld_ext(w1 + 4 * (uint64_t)r);
            wb[NP - 1] = ld_ext(w1 + 4 * (uint64_t)r2);
        }
        uint32_t va[BARY_COLS], vb[BARY_COLS];
#pragma unroll
        for (int c = 0; c < BARY_COLS; c++) {
            va[c] = colp[c][r];
            vb[c] = colp[c][r2];
        }
#pragma unroll
        for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
            for (int t = 0; t < NP; t++)
#pragma unroll
                for (int k = 0; k < 4; k++) kb::mac2(acc[c][t][k], wa[t].c[k], va[c], wb[t].c[k], vb[c]);
    }
    if (r < r_end) {
        Ext w[NP];
        w[0] = ld_ext(w0 + 4 * (uint64_t)r);
        if (NP > 1) w[NP - 1] = ld_ext(w1 + 4 * (uint64_t)r);
        uint32_t v[BARY_COLS];
#pragma unroll
        for (int c = 0; c < BARY_COLS; c++) v[c] = colp[c][r];
#pragma unroll
        for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
            for (int t = 0; t < NP; t++)
#pragma unroll
                for (int k = 0; k < 4; k++) kb::mac(acc[c][t][k], w[t].c[k], v[c]);
    }
    __shared__ uint32_t red[BARY_THREADS / 32][BARY_COLS * NP * 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < BARY_COLS; c++)
#pragma unroll
        for (int t = 0; t < NP; t++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t v = kb::mont_reduce(acc[c][t][k]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = kb::add(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) red[warp][(c * NP + t) * 4 + k] = v;
            }
    __syncthreads();
    if (threadIdx.x < BARY_COLS * NP * 4) {
        uint32_t v = 0;
        for (int wv = 0; wv < BARY_THREADS / 32; wv++) v = kb::add(v, red[wv][threadIdx.x]);
        int c = threadIdx.x / (NP * 4), rem = threadIdx.x % (NP * 4);
        if (c0 + c < ncols) partial[(((uint64_t)(c0 + c) * nchunks + chunk) * NP) * 4 + rem] = v;
    }
}

#ifndef BFGPU_BARY_MINBLOCKS
#define BFGPU_BARY_MINBLOCKS 1  // forcing 3 or 4 CTAs/SM spills the 64-bit accumulators: open_eval 2.69 -> 3.26 / 5.0 ms at 2^22 rows
#endif
template <int NP>
__global__ void __launch_bounds__(BARY_THREADS, BFGPU_BARY_MINBLOCKS) k_bary_dot(const uint32_t* __restrict__ mat, uint64_t col_stride, uint32_t ncols, uint32_t h,
                                                           const uint32_t* __restrict__ w0, const uint32_t* __restrict__ w1,
                                                           uint32_t* __restrict__ partial, uint32_t nchunks) {
    // column group is the FAST grid index: the blocks in flight share a row chunk, so its weights (2 x 16 B per row, two thirds of
    // the kernel's traffic when every column group re-read them from DRAM) are served by L2
    bary_dot_body<NP>(mat, col_stride, ncols, h, w0, w1, partial, nchunks, blockIdx.y, blockIdx.x * BARY_COLS);
}
// every matrix whose low coset fits one chunk (h <= BARY_ROWS), in one launch: blockIdx.y = job, blockIdx.x = column group;
// with a single chunk the partial sums ARE the sums, written straight to their final place
struct BaryJob {
    const uint32_t* mat;
    uint64_t col_stride;
    const uint32_t *w0, *w1;
    uint32_t* out;
    uint32_t ncols, h;
};
template <int NP>
__global__ void __launch_bounds__(BARY_THREADS) k_bary_dot_batch(const BaryJob* __restrict__ jobs) {
    const BaryJob j = jobs[blockIdx.y];
    const uint32_t c0 = blockIdx.x * BARY_COLS;
    if (c0 >= j.ncols) return;
    bary_dot_body<NP>(j.mat, j.col_stride, j.ncols, j.h, j.w0, j.w1, j.out, 1, 0, c0);
}

// out[i] = sum over chunks of partial[i][chunk]   (i over ncols*NP*4 words; layout as above)
__global__ void k_bary_finish(const uint32_t* __restrict__ partial, uint32_t* __restrict__ out, uint32_t ncols, uint32_t nchunks, uint32_t np) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncols * np * 4) return;
    uint32_t c = i / (np * 4), rem = i % (np * 4);
    uint32_t v = 0;
    for (uint32_t ch = 0; ch < nchunks; ch++) v = kb::add(v, partial[(((uint64_t)c * nchunks + ch) * np) * 4 + rem]);
    out[i] = v;
}

// ---- reduced openings of one height group ---------------------------------------------------------------
#ifndef BFGPU_RO_UNROLL
#define BFGPU_RO_UNROLL 4
#endif
constexpr int RO_UNROLL = BFGPU_RO_UNROLL;  // column loads in flight per thread in k_reduce_openings
struct RoMat {
    const uint32_t* d;  // column-major LDE: first row handled by the launch (a row shard points at its first local row)
    uint64_t stride;    // words between columns (the height of the buffer `d` points into)
    uint32_t width;
    uint32_t npoints;    // 1 or 2
    uint32_t pt[2];      // index into the group's point list
    uint32_t yred[2][4];  // sum_k alpha^k y_k   for each point
    uint32_t aoff[2][4];  // alpha^(columns already reduced at this height)
};
// ro[r] = sum_{mat, point} aoff * (yred - sum_k alpha^k M[r][k]) / (z_point - x_r),  x_r = shift * w_H^{br(r)}
__global__ void __launch_bounds__(128) k_reduce_openings(const RoMat* __restrict__ mats, uint32_t nmats, const uint32_t* __restrict__ zs /* npts x 4 */,
                                                         uint32_t npts, const uint32_t* __restrict__ apow /* ext alpha^k */, unsigned log_h,
                                                         uint32_t shift, const uint32_t* __restrict__ tw, uint32_t* __restrict__ ro, uint32_t r0,
                                                         uint32_t count) {
    // rows [r0, r0 + count) of the height-2^log_h group; matrices, ro and the thread index are relative to r0 (row shard of a rank)
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    uint32_t x = kb::mul(shift, root_pow(tw, log_h, kb::bitrev(r0 + r, log_h)));
    // 1 / (z_t - x) for the (at most four) opening points of this height: Montgomery's trick, ONE extension inversion per row
    // (an inversion is ~1 200 instructions, an extension product ~140; the two inversions were 60 % of this kernel)
    Ext inv[4], pre[4];
    const uint32_t np = npts < 4 ? npts : 4;
    for (uint32_t t = 0; t < np; t++) {
        Ext d = ld_ext(zs + 4 * t);
        d.c[0] = kb::sub(d.c[0], x);
        inv[t] = d;
        pre[t] = t ? kb::ext_mul(pre[t - 1], d) : d;
    }
    if (np) {
        Ext acc_inv = kb::ext_inv(pre[np - 1]);
        for (uint32_t t = np; t-- > 1;) {
            Ext d = inv[t];
            inv[t] = kb::ext_mul(acc_inv, pre[t - 1]);
            acc_inv = kb::ext_mul(acc_inv, d);
        }
        inv[0] = acc_inv;
    }
    Ext acc = kb::ext_zero();
    for (uint32_t m = 0; m < nmats; m++) {
        const RoMat& M = mats[m];
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;  // unreduced sum_k alpha^k M[r][k]
        const uint32_t* col = M.d + r;
        // (MEASURED slower here at 2^22 rows, 2.56 ms as written: pairing two columns per conditional subtraction with kb::mac2 3.27 ms; the
        // powers of alpha staged in shared memory instead of the uniform global load 2.95 ms)
#pragma unroll RO_UNROLL
        for (uint32_t k = 0; k < M.width; k++) {
            uint32_t v = col[(uint64_t)k * M.stride];
            Ext a = ld_ext(apow + 4 * k);
            kb::mac(a0, a.c[0], v);
            kb::mac(a1, a.c[1], v);
            kb::mac(a2, a.c[2], v);
            kb::mac(a3, a.c[3], v);
        }
        Ext rr = Ext{{kb::mont_reduce(a0), kb::mont_reduce(a1), kb::mont_reduce(a2), kb::mont_reduce(a3)}};
        for (uint32_t t = 0; t < M.npoints; t++) {
            Ext y = Ext{{M.yred[t][0], M.yred[t][1], M.yred[t][2], M.yred[t][3]}};
            Ext a = Ext{{M.aoff[t][0], M.aoff[t][1], M.aoff[t][2], M.aoff[t][3]}};
            Ext q = kb::ext_mul(kb::ext_sub(y, rr), inv[M.pt[t]]);
            acc = kb::ext_add(acc, kb::ext_mul(a, q));
        }
    }
    st_ext(ro + 4 * (uint64_t)r, acc);
}

// ---- FRI fold: out[i] = (1/2 + b g^-br(i)) in[2i] + (1/2 - b g^-br(i)) in[2i+1] (+ add[i]),  b = beta/2 -----
// rollin: how the reduced opening of the new height enters (BFGPU_OPT_FRI_ROLLIN): 0 = plain sum (Plonky3 of the pinned API era),
// 1 = beta^2 * add[i] (the later upstream rule), beta being this round's folding challenge
// i0: global index of element 0 of in/out/add (a rank of the sharded prover folds its contiguous slice of the vector)
__device__ __forceinline__ void fri_fold_one(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ add, unsigned log_h,
                                             Ext half_beta, const uint32_t* __restrict__ tw, uint32_t i, int rollin = 0, uint32_t i0 = 0) {
    // g = generator of order 2^(log_h+1); g^-j = w^(2^(log_h+1) - j)
    uint32_t j = kb::bitrev(i0 + i, log_h);
    uint32_t ginv = j ? root_pow(tw, log_h + 1, (1u << (log_h + 1)) - j) : kb::ONE;
    Ext lo = ld_ext(in + 8 * (uint64_t)i), hi = ld_ext(in + 8 * (uint64_t)i + 4);
    Ext pw = kb::ext_scale(half_beta, ginv);
    const uint32_t half = kb::halve(kb::ONE);
    Ext a = pw, b = kb::ext_neg(pw);
    a.c[0] = kb::add(a.c[0], half);
    b.c[0] = kb::add(b.c[0], half);
    Ext o = kb::ext_add(kb::ext_mul(a, lo), kb::ext_mul(b, hi));
    if (add) {
        Ext r = ld_ext(add + 4 * (uint64_t)i);
        if (rollin == 1) r = kb::ext_mul(kb::ext_sqr(kb::ext_add(half_beta, half_beta)), r);
        o = kb::ext_add(o, r);
    }
    st_ext(out + 4 * (uint64_t)i, o);
}
// host-supplied beta, elements [i0, i0 + count) of the output (sharded prover: the challenger runs on the host between the rounds)
__global__ void k_fri_fold(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ add, unsigned log_h /* out length */,
                           Ext half_beta, const uint32_t* __restrict__ tw, int rollin, uint32_t i0, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fri_fold_one(in, out, add, log_h, half_beta, tw, i, rollin, i0);
}

// ---- Fiat-Shamir on the device for the rounds that are too big for the tail kernel --------------------------------------------
// One warp: observe the 8-word root, sample beta (DuplexChallenger<_, _, 16, 8>, same buffer logic as challenger.h).  The
// challenger lives in device memory (ch[0..16) state, ch[16..24) input buffer, ch[24] its fill), so a commit-phase round is
// leaf hash -> tree -> this kernel -> fold with NO host round trip; the host replays all roots once at the end.
__global__ void __launch_bounds__(32) k_challenger_round(uint32_t* __restrict__ ch, const uint32_t* __restrict__ root, uint32_t* __restrict__ beta_out,
                                                         uint32_t* __restrict__ root_log) {
    __shared__ uint32_t s_ext[8 * 16], s_int[16], st[16], in[8];
    const int lane = threadIdx.x, q = lane & 3;
    for (int k = lane; k < 128; k += 32) s_ext[k] = p2::c_p2.ext_s[k >> 4][k & 15];
    if (lane < 16) {
        s_int[lane] = p2::c_p2.internal_s[lane];
        st[lane] = ch[lane];
    }
    if (lane < 8) in[lane] = ch[16 + lane];
    __syncwarp();
    const p2::X4 xc = p2::x4_setup(s_ext, s_int, q);
    uint32_t n_in = ch[24], n_out = 0;
    auto duplex = [&]() {
        if ((uint32_t)lane < n_in) st[lane] = in[lane];
        __syncwarp();
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = st[4 * q + j];
        p2::permute_x4(w, xc, q);
        __syncwarp();
        if (lane < 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) st[4 * lane + j] = w[j];
        }
        __syncwarp();
        n_in = 0;
        n_out = 8;
    };
    for (int k = 0; k < 8; k++) {
        n_out = 0;
        if (lane == 0) in[n_in] = root[k];
        n_in++;
        __syncwarp();
        if (n_in == 8) duplex();
    }
    uint32_t beta[4];
    for (int k = 0; k < 4; k++) {
        if (n_in != 0 || n_out == 0) duplex();
        beta[k] = st[--n_out];
    }
    __syncwarp();
    // after the samples the output buffer still holds n_out words, but the next transcript operation is an observe
    // (next root or the final polynomial), which discards it: only state, input buffer and its fill are carried
    if (lane < 16) ch[lane] = st[lane];
    if (lane < 8) {
        ch[16 + lane] = in[lane];
        root_log[lane] = root[lane];
    }
    if (lane == 0) ch[24] = n_in;
    if (lane < 4) beta_out[lane] = beta[lane];
}
// fold with beta read from device memory (written by k_challenger_round)
__global__ void k_fri_fold_dev(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ add, unsigned log_h,
                               const uint32_t* __restrict__ beta, const uint32_t* __restrict__ tw, int rollin, uint32_t i0, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Ext half_beta = kb::ext_scale(ld_ext(beta), kb::halve(kb::ONE));
    fri_fold_one(in, out, add, log_h, half_beta, tw, i, rollin, i0);
}

// ---- FRI tail: every commit-phase round whose input has at most 2^TAIL_MAX_LOG elements, in ONE single-CTA launch -----
// Per round the host path costs three launches (fold, leaf hash, tree top), a 32-byte device->host copy, a stream
// synchronisation and the host sponge — ~60 us of latency for microseconds of work, ten times per proof.  Here one CTA
// runs the rounds back to back: leaf digests and tree levels as in hashk::k_compress_top (four-lane permutation below
// 256 nodes), then warp 0 plays the DuplexChallenger (observe the root, sample beta: same buffer logic as
// challenger.h, state handed in by the host), then the fold.  All layers, folded vectors and roots go to global memory
// in the layout the query phase reads; the host replays the roots through its own challenger to advance it.
constexpr int TAIL_MAX_LOG = 11, TAIL_THREADS = 1024, TAIL_MAX_ROUNDS = 11;
struct FriTailArgs {
    uint32_t* vec[TAIL_MAX_ROUNDS + 1];              // vec[r]: input of round r, 2^(log_len - r) elements; vec[nrounds]: final vector
    const uint32_t* add[TAIL_MAX_ROUNDS];            // reduced opening added to the output of round r, or null
    int rollin;                                      // BFGPU_OPT_FRI_ROLLIN
    uint32_t* layer[TAIL_MAX_ROUNDS][TAIL_MAX_LOG];  // layer[r][l]: 2^(log_len - r - 1 - l) digests, l = 0 .. log_len - r - 1
    uint32_t* roots;                                 // nrounds x 8
    uint32_t log_len, nrounds;
    const uint32_t* ch;                              // device challenger at entry: state[16], input[8], n_in (DevChallenger layout)
    const uint32_t* tw;
};
__global__ void __launch_bounds__(TAIL_THREADS) k_fri_tail(FriTailArgs A) {
    __shared__ __align__(16) uint32_t cur[(1 << (TAIL_MAX_LOG - 1)) * 8];
    __shared__ uint32_t s_ext[8 * 16], s_int[16], ch_st[16], ch_in[8];
    __shared__ __align__(16) uint32_t s_beta[4];
    const uint32_t t = threadIdx.x;
    const int q = t & 3, lane = t & 31;
    if (t < 128) s_ext[t] = p2::c_p2.ext_s[t >> 4][t & 15];
    if (t < 16) {
        s_int[t] = p2::c_p2.internal_s[t];
        ch_st[t] = A.ch[t];
    }
    if (t < 8) ch_in[t] = A.ch[16 + t];
    __syncthreads();
    const p2::X4 xc = p2::x4_setup(s_ext, s_int, q);
    uint32_t n_in = A.ch[24], n_out = 0;  // replicated in every lane of warp 0 (the output buffer is dead: an observe comes next)
    auto duplex = [&]() {                 // warp 0 only, all 32 lanes
        if ((uint32_t)lane < n_in) ch_st[lane] = ch_in[lane];
        __syncwarp();
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = ch_st[4 * q + j];
        p2::permute_x4(w, xc, q);
        __syncwarp();
        if (lane < 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) ch_st[4 * lane + j] = w[j];
        }
        __syncwarp();
        n_in = 0;
        n_out = 8;
    };
    for (uint32_t r = 0; r < A.nrounds; r++) {
        const unsigned log_n = A.log_len - r - 1;  // leaves of this round = output length of its fold
        const uint32_t n = 1u << log_n;
        const uint32_t* vin = A.vec[r];
        // ---- leaf digests: sponge of the 8-word row (in[2i], in[2i+1]) = one permutation of (row | 0^8) -------------------
        if (4 * n <= (uint32_t)TAIL_THREADS) {
            const uint32_t i = t >> 2;
            const bool warp_on = ((t & ~31u) >> 2) < n, on = i < n;
            if (warp_on) {
                uint32_t w[4] = {0, 0, 0, 0};
                if (on && q < 2) {
                    uint4 x = *reinterpret_cast<const uint4*>(vin + 8 * i + 4 * q);
                    w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
                }
                p2::permute_x4(w, xc, q);
                if (on && q < 2) {
                    uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(A.layer[r][0] + 8 * i + 4 * q) = o;
                    *reinterpret_cast<uint4*>(cur + 8 * i + 4 * q) = o;
                }
            }
        } else {
            for (uint32_t i = t; i < n; i += TAIL_THREADS) {  // at most one iteration (n <= 1024)
                uint32_t s[16];
                uint4 x0 = *reinterpret_cast<const uint4*>(vin + 8 * i), x1 = *reinterpret_cast<const uint4*>(vin + 8 * i + 4);
                s[0] = x0.x; s[1] = x0.y; s[2] = x0.z; s[3] = x0.w; s[4] = x1.x; s[5] = x1.y; s[6] = x1.z; s[7] = x1.w;
#pragma unroll
                for (int k = 8; k < 16; k++) s[k] = 0;
                p2::permute(s);
                hashk::store_digest(A.layer[r][0] + 8 * i, s);
#pragma unroll
                for (int k = 0; k < 8; k++) cur[8 * i + k] = s[k];
            }
        }
        __syncthreads();
        // ---- tree levels ------------------------------------------------------------------------------------------------
        for (unsigned l = 1; l <= log_n; l++) {
            const uint32_t m = n >> l;
            if (4 * m <= (uint32_t)TAIL_THREADS) {
                const uint32_t i = t >> 2;
                const bool warp_on = ((t & ~31u) >> 2) < m, on = i < m;
                uint32_t w[4] = {0, 0, 0, 0};
                if (on) {
                    uint4 x = *reinterpret_cast<const uint4*>(cur + 16 * i + 4 * q);
                    w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
                }
                __syncthreads();
                if (warp_on) {
                    p2::permute_x4(w, xc, q);
                    if (on && q < 2) {
                        uint4 o = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(A.layer[r][l] + 8 * i + 4 * q) = o;
                        *reinterpret_cast<uint4*>(cur + 8 * i + 4 * q) = o;
                    }
                }
            } else {
                uint32_t s[16];
                const uint32_t i = t;  // m <= 512 here
                if (i < m) {
#pragma unroll
                    for (int k = 0; k < 16; k++) s[k] = cur[16 * i + k];
                }
                __syncthreads();
                if (i < m) {
                    p2::permute(s);
                    hashk::store_digest(A.layer[r][l] + 8 * i, s);
#pragma unroll
                    for (int k = 0; k < 8; k++) cur[8 * i + k] = s[k];
                }
            }
            __syncthreads();
        }
        // ---- Fiat-Shamir: observe the root, sample beta (DuplexChallenger<_, _, 16, 8>) -----------------------------------------
        if (t < 32) {
            for (int k = 0; k < 8; k++) {  // observe: clears the output buffer, duplexes when the rate is full
                n_out = 0;
                if (lane == 0) ch_in[n_in] = cur[k];
                n_in++;
                __syncwarp();
                if (n_in == 8) duplex();
            }
            uint32_t beta[4];
            for (int k = 0; k < 4; k++) {  // sample: duplex first if anything was observed since, pop from the back
                if (n_in != 0 || n_out == 0) duplex();
                beta[k] = ch_st[--n_out];
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; k++) s_beta[k] = beta[k];
#pragma unroll
                for (int k = 0; k < 8; k++) A.roots[8 * r + k] = cur[k];
            }
        }
        __syncthreads();
        // ---- fold -------------------------------------------------------------------------------------------------------------
        const Ext half_beta = kb::ext_scale(Ext{{s_beta[0], s_beta[1], s_beta[2], s_beta[3]}}, kb::halve(kb::ONE));
        for (uint32_t i = t; i < n; i += TAIL_THREADS) fri_fold_one(vin, A.vec[r + 1], A.add[r], log_n, half_beta, A.tw, i, A.rollin);
        __syncthreads();  // the next round reads vec[r + 1] (same CTA: block-level visibility suffices)
    }
}

// ---- proof of work: smallest (descending: largest) w in [start, start+count) with (permute(state | w at pos)[7] & mask) == 0 ------
__global__ void __launch_bounds__(128) k_pow_grind(const uint32_t* __restrict__ state16, uint32_t pos, uint32_t mask, uint32_t start, uint32_t count,
                                                   unsigned int* __restrict__ best, int descending) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t w = start + i;  // canonical candidate
    uint32_t s[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s[k] = state16[k];
    uint32_t wm = kb::to_mont(w);
#pragma unroll
    for (int k = 0; k < 8; k++)
        if ((uint32_t)k == pos) s[k] = wm;
    p2::permute(s);
    if ((kb::from_mont(s[7]) & mask) == 0) {
        if (descending) atomicMax(best, w + 1);  // 0 = nothing found
        else atomicMin(best, w);
    }
}

// ---- query gathers: out[i] = *src[i] (optionally converted to canonical) -----------------------------------
__global__ void k_gather_words(const uint32_t* const* __restrict__ src, uint32_t* __restrict__ out, uint64_t n, int to_canonical) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = *src[i];
    out[i] = to_canonical ? kb::from_mont(v) : v;
}

// Query answers from a per-proof TEMPLATE: every query opens the same list of words, only the index differs, and every
// word's address is base + stride * ((index >> shift) ^ flip):
//   matrix element  : base = column start,   stride = row stride, shift = log(max height) - log(rows),          flip 0
//   Merkle sibling  : base = layer + k,       stride = 8,          shift = (log max - log tree) + level,         flip 1
//   FRI sibling val : base = folded vec + k,  stride = 4,          shift = round,                                 flip 1
// so the host ships ~3 000 descriptors once instead of 84 x 3 000 pointers (2.3 MB of pageable host memory per proof).
// A null base yields the query index itself (the serialisation stores it in front of each query).
struct QueryWord {
    const uint32_t* base;
    uint32_t stride, shift, flip, pad;
};
__global__ void k_answer_queries(const QueryWord* __restrict__ tmpl, uint32_t per_query, const uint32_t* __restrict__ indices, uint32_t nq,
                                 uint32_t* __restrict__ out, int to_canonical) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (uint64_t)per_query * nq) return;
    const uint32_t q = (uint32_t)(i / per_query), e = (uint32_t)(i % per_query);
    const QueryWord w = tmpl[e];
    const uint32_t index = indices[q];
    if (!w.base) {
        out[i] = index;
        return;
    }
    uint32_t v = w.base[(uint64_t)w.stride * ((index >> w.shift) ^ w.flip)];
    out[i] = to_canonical ? kb::from_mont(v) : v;
}

}  // namespace openk
