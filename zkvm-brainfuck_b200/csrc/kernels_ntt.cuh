// kernels_ntt.cuh — two-adic NTT / coset LDE kernels (K1 of SURVEY.md §2).
//
// Replaces `Radix2DitParallel::coset_lde_batch` as reached from `TwoAdicFriPcs::commit`
// (reference crates/stark/src/kb31_poseidon2.rs:30; call sites crates/stark/src/prover.rs:227,334,411
// and crates/stark/src/machine.rs:196).
//
// Device layout: COLUMN-major (each trace column is one contiguous vector), Montgomery words.
// The host's row-major matrix is transposed on ingest; the inverse transform's bit-reversal is
// fused into that ingest (rows are gathered in bit-reversed order), the forward transform is
// decimation-in-frequency so the LDE comes out with bit-reversed rows, which is exactly the order
// TwoAdicFriPcs stores and MerkleTreeMmcs hashes.  No standalone bit-reversal pass exists.
//
//   ingest(bitrev) -> inverse DIT passes -> per-coset scale (shift^k / n) -> forward DIF passes
//
// A "pass" executes g consecutive radix-2 stages (index bits [p, p+g)) on tiles staged in shared
// memory, so a 2^22-point column needs 3 passes per transform instead of 22 global sweeps.
#pragma once
#include "kb31.cuh"

namespace nttk {

constexpr int TW_LOG = kb::TWO_ADICITY;  // twiddle table: tw[e] = w^e, w of order 2^24, e < 2^23
constexpr int NTT_THREADS = 256;
constexpr int LANES_LOG = 5;  // 32 lanes (128 B) per strided tile row
constexpr int GMAX = 8;       // stages per pass: tile = 2^8 * 32 words = 32 KiB of shared memory

// forward twiddle w_{2^k}^j, j < 2^(k-1)
__device__ __forceinline__ uint32_t tw_fwd(const uint32_t* __restrict__ tw, unsigned k, uint32_t j) {
    return __ldg(tw + ((uint64_t)j << (TW_LOG - k)));
}
// inverse twiddle w_{2^k}^{-j}, j < 2^(k-1):  w^{-j} = -w^{2^(k-1) - j}
__device__ __forceinline__ uint32_t tw_inv(const uint32_t* __restrict__ tw, unsigned k, uint32_t j) {
    if (j == 0) return kb::ONE;
    return kb::P - __ldg(tw + ((uint64_t)((1u << (k - 1)) - j) << (TW_LOG - k)));
}

__global__ void k_build_twiddles(uint32_t* __restrict__ tw, uint32_t w /* Montgomery generator of order 2^TW_LOG */) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;  // grid covers 2^(TW_LOG-1)
    tw[e] = kb::pow(w, e);
}

// powers[k] = scale * base^k for k < n (used for the coset shift s^k / n)
__global__ void k_powers(uint32_t* __restrict__ out, uint32_t base, uint32_t scale, uint64_t n) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = kb::mul(kb::pow(base, k), scale);
}

// ---- ingest / egress: row-major host layout <-> column-major Montgomery ---------------------
// dst[c][j] = conv(src[perm(j)][c]),  perm = bit reversal on log_rows bits when bitrev != 0
// (src_pitch = words between consecutive source rows: cols for a whole matrix, more for a column block of a wider one)
__global__ void k_ingest(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t rows, uint32_t cols, unsigned log_rows,
                         int bitrev, int to_mont, uint64_t src_pitch) {
    __shared__ uint32_t tile[32][33];
    uint64_t r0 = (uint64_t)blockIdx.x * 32;
    uint32_t c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint64_t j = r0 + i;
        uint32_t c = c0 + threadIdx.x;
        if (j < rows && c < cols) {
            uint64_t sr = bitrev ? kb::bitrev((uint32_t)j, log_rows) : j;
            uint32_t v = src[sr * src_pitch + c];
            tile[i][threadIdx.x] = to_mont ? kb::to_mont(v) : v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint32_t c = c0 + i;
        uint64_t j = r0 + threadIdx.x;
        if (j < rows && c < cols) dst[(uint64_t)c * rows + j] = tile[threadIdx.x][i];
    }
}

// Wide-tile variant for rows >= 128: a CTA moves 128 rows x 32 columns with 16-byte accesses on both sides (8 threads x 16 B
// per source row, 4 consecutive destination rows per store) and four independent loads in flight per thread; the 32 x 32
// kernel above ran at half the HBM rate (2.7 ms for 4 GiB in + 4 GiB out) on latency.  cols % 4 == 0 and 16-byte aligned
// source rows required (checked by the host; otherwise the kernel above is used).
__global__ void __launch_bounds__(256) k_ingest_wide(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t rows, uint32_t cols,
                                                     unsigned log_rows, int bitrev, int to_mont, uint64_t src_pitch) {
    __shared__ uint32_t tile[32][132];  // [column][row], row stride 132 words: 16-byte aligned rows, conflict-free transposed stores
    const uint32_t t = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * 128;
    const uint32_t c0 = blockIdx.y * 32;
    const uint32_t cg = (t & 7) * 4;  // column group inside the tile
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t jl = (t >> 3) + 32 * k;
        const uint64_t j = r0 + jl;
        if (j < rows && c0 + cg < cols) {
            const uint64_t sr = bitrev ? kb::bitrev((uint32_t)j, log_rows) : j;
            uint4 v = *reinterpret_cast<const uint4*>(src + sr * src_pitch + c0 + cg);
            if (to_mont) {
                v.x = kb::to_mont(v.x); v.y = kb::to_mont(v.y); v.z = kb::to_mont(v.z); v.w = kb::to_mont(v.w);
            }
            tile[cg + 0][jl] = v.x;
            tile[cg + 1][jl] = v.y;
            tile[cg + 2][jl] = v.z;
            tile[cg + 3][jl] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t idx = t + 256 * k;  // 1024 vectors: 32 columns x 32 groups of 4 rows
        const uint32_t c = idx >> 5, jl = (idx & 31) * 4;
        if (c0 + c < cols && r0 + jl < rows)
            *reinterpret_cast<uint4*>(dst + (uint64_t)(c0 + c) * rows + r0 + jl) = *reinterpret_cast<const uint4*>(&tile[c][jl]);
    }
}

// dst[perm(j)][c] = conv(src[c][j])
__global__ void k_egress(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint64_t rows, uint32_t cols, unsigned log_rows,
                         int bitrev, int from_mont) {
    __shared__ uint32_t tile[32][33];
    uint64_t r0 = (uint64_t)blockIdx.x * 32;
    uint32_t c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint32_t c = c0 + i;
        uint64_t j = r0 + threadIdx.x;
        if (j < rows && c < cols) {
            uint32_t v = src[(uint64_t)c * rows + j];
            tile[i][threadIdx.x] = from_mont ? kb::from_mont(v) : v;
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        uint64_t j = r0 + i;
        uint32_t c = c0 + threadIdx.x;
        if (j < rows && c < cols) {
            uint64_t dr = bitrev ? kb::bitrev((uint32_t)j, log_rows) : j;
            dst[dr * cols + c] = tile[threadIdx.x][i];
        }
    }
}

// ---- whole coset LDE of SMALL columns (<= 2^SMALL_LDE_MAX_LOG points), many matrices per launch -----------------------
// A proof commits to a handful of short matrices (Program, Memory, IO, AddSub, ... : 16 to a few thousand rows) three
// times; through the general path each costs 4-5 launches of a few microseconds.  Here one CTA owns one column of one
// matrix and does the whole chain in shared memory — inverse DIT (bit-reversed in, natural out), per-coset scaling,
// forward DIF (natural in, bit-reversed out) — and one launch covers every small matrix of the commitment.
constexpr int SMALL_LDE_MAX_LOG = 11, SMALL_LDE_MAX_MATS = 24;
struct SmallLdeArgs {
    uint32_t* coef[SMALL_LDE_MAX_MATS];       // column-major, bit-reversed rows, 2^log_n rows
    uint32_t* out[SMALL_LDE_MAX_MATS];        // column-major, ncosets * 2^log_n rows
    const uint32_t* pw[SMALL_LDE_MAX_MATS];   // pw[h * n + k] = shift_h^k / n
    uint32_t log_n[SMALL_LDE_MAX_MATS];
    uint32_t first_cta[SMALL_LDE_MAX_MATS + 1];  // prefix sums of the column counts
    uint32_t nmats, ncosets;
    const uint32_t* tw;                       // w_{2^24}^e table
};
__device__ __forceinline__ uint32_t small_root(const uint32_t* __restrict__ tw, unsigned log_m, uint32_t j) {  // w_{2^log_m}^j, j < 2^log_m
    if (log_m == 0) return kb::ONE;
    uint32_t half = 1u << (log_m - 1);
    uint32_t v = __ldg(tw + ((uint64_t)(j & (half - 1)) << (TW_LOG - log_m)));
    return (j & half) ? kb::neg(v) : v;
}
__global__ void __launch_bounds__(256) k_lde_small(SmallLdeArgs A) {
    __shared__ uint32_t x[1 << SMALL_LDE_MAX_LOG], y[1 << SMALL_LDE_MAX_LOG];
    uint32_t m = 0;
    while (m + 1 < A.nmats && blockIdx.x >= A.first_cta[m + 1]) m++;
    const uint32_t c = blockIdx.x - A.first_cta[m];
    const unsigned log_n = A.log_n[m];
    const uint32_t n = 1u << log_n;
    const uint32_t* src = A.coef[m] + (uint64_t)c * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) x[i] = src[i];
    __syncthreads();
    for (unsigned s = 0; s < log_n; s++) {  // inverse, decimation in time
        const uint32_t half = 1u << s;
        for (uint32_t q = threadIdx.x; q < n / 2; q += blockDim.x) {
            const uint32_t j = q & (half - 1), i0 = ((q >> s) << (s + 1)) | j, i1 = i0 + half;
            const uint32_t w = j ? small_root(A.tw, s + 1, (2 * half) - j) : kb::ONE;  // w_{2^(s+1)}^{-j}
            const uint32_t a = x[i0], b = kb::mul(x[i1], w);
            x[i0] = kb::add(a, b);
            x[i1] = kb::sub(a, b);
        }
        __syncthreads();
    }
    for (uint32_t h = 0; h < A.ncosets; h++) {
        const uint32_t* pw = A.pw[m] + (uint64_t)h * n;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = kb::mul(x[i], __ldg(pw + i));
        __syncthreads();
        for (int s = (int)log_n - 1; s >= 0; s--) {  // forward, decimation in frequency
            const uint32_t half = 1u << s;
            for (uint32_t q = threadIdx.x; q < n / 2; q += blockDim.x) {
                const uint32_t j = q & (half - 1), i0 = ((q >> s) << (s + 1)) | j, i1 = i0 + half;
                const uint32_t a = y[i0], b = y[i1];
                y[i0] = kb::add(a, b);
                const uint32_t d = kb::sub(a, b);
                y[i1] = j ? kb::mul(d, small_root(A.tw, s + 1, j)) : d;
            }
            __syncthreads();
        }
        uint32_t* dst = A.out[m] + ((uint64_t)c * A.ncosets + h) * n;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = y[i];
        __syncthreads();
    }
}

// ---- coset scaling: coefficients -> 2^added_bits scaled copies ---------------------------------
// out[c][h*n + k] = in[c][k] * pw[h*n + k]   (pw[h*n+k] = (shift_h)^k / n)
__global__ void k_scale_cosets(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ pw, uint64_t n,
                               uint32_t ncosets, uint32_t cols) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t c = blockIdx.y;
    if (k >= n) return;
    uint32_t v = in[(uint64_t)c * n + k];
    for (uint32_t h = 0; h < ncosets; h++) out[((uint64_t)c * ncosets + h) * n + k] = kb::mul(v, __ldg(pw + h * n + k));
}

__global__ void k_scale(uint32_t* __restrict__ data, uint64_t n, uint32_t s) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) data[i] = kb::mul(data[i], s);
}

// ---- one pass = g radix-2 stages on index bits [p, p+g) ------------------------------------------
// Every column is a vector of 2^log_n words at data + col*col_stride.
//   strided tile  (p >= LANES_LOG): idx(d, c) = hi<<(p+g) | d<<p | (lo_block*32 + c),  smem[d*32 + c]
//   contiguous    (p == 0)        : a run of 2^g * lanes words,                       smem[flat]
// FORWARD (decimation in frequency, natural in -> bit-reversed out): stages t = g-1 .. 0,
//   (x, y) -> (x + y, (x - y) * w^m),  w of order 2^(p+t+1), m = idx mod 2^(p+t)
// INVERSE (decimation in time, bit-reversed in -> natural out, unnormalised): stages t = 0 .. g-1,
//   (x, y) -> (x + y w^-m, x - y w^-m)
template <bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass(uint32_t* __restrict__ data, uint64_t col_stride, unsigned log_n, unsigned p,
                                                          unsigned g, unsigned lanes_log, const uint32_t* __restrict__ tw) {
    extern __shared__ uint32_t sm[];
    const bool strided = p != 0;
    uint32_t* col = data + (uint64_t)blockIdx.y * col_stride;
    const uint32_t tile = blockIdx.x;
    const uint32_t nwords = 1u << (g + lanes_log);
    uint64_t base;  // strided: index of (d=0,c=0); contiguous: first word of the run
    if (strided) {
        uint32_t lo_blocks_log = p - lanes_log;
        uint32_t lo_block = tile & ((1u << lo_blocks_log) - 1);
        uint64_t hi = tile >> lo_blocks_log;
        base = (hi << (p + g)) | ((uint64_t)lo_block << lanes_log);
    } else {
        base = (uint64_t)tile << (g + lanes_log);
    }
    // load
    for (uint32_t f = threadIdx.x; f < nwords; f += NTT_THREADS) {
        uint64_t idx = strided ? base + ((uint64_t)(f >> lanes_log) << p) + (f & ((1u << lanes_log) - 1)) : base + f;
        sm[f] = col[idx];
    }
    __syncthreads();
    const uint32_t npairs = nwords >> 1;
    for (unsigned s = 0; s < g; s++) {
        const unsigned t = INVERSE ? s : g - 1 - s;
        for (uint32_t q = threadIdx.x; q < npairs; q += NTT_THREADS) {
            uint32_t i0, i1, m;
            if (strided) {
                uint32_t c = q & ((1u << lanes_log) - 1), k = q >> lanes_log;
                uint32_t dl = ((k >> t) << (t + 1)) | (k & ((1u << t) - 1));
                i0 = (dl << lanes_log) | c;
                i1 = i0 + (1u << (t + lanes_log));
                uint32_t lo = (uint32_t)(base & ((1ull << p) - 1)) + c;
                m = ((dl & ((1u << t) - 1)) << p) | lo;
            } else {
                uint32_t k = q & ((1u << (g - 1)) - 1), grp = q >> (g - 1);
                uint32_t dl = ((k >> t) << (t + 1)) | (k & ((1u << t) - 1));
                i0 = (grp << g) | dl;
                i1 = i0 + (1u << t);
                m = dl & ((1u << t) - 1);
            }
            uint32_t x = sm[i0], y = sm[i1];
            if (INVERSE) {
                y = kb::mul(y, tw_inv(tw, p + t + 1, m));
                sm[i0] = kb::add(x, y);
                sm[i1] = kb::sub(x, y);
            } else {
                sm[i0] = kb::add(x, y);
                sm[i1] = kb::mul(kb::sub(x, y), tw_fwd(tw, p + t + 1, m));
            }
        }
        __syncthreads();
    }
    for (uint32_t f = threadIdx.x; f < nwords; f += NTT_THREADS) {
        uint64_t idx = strided ? base + ((uint64_t)(f >> lanes_log) << p) + (f & ((1u << lanes_log) - 1)) : base + f;
        col[idx] = sm[f];
    }
}

}  // namespace nttk
