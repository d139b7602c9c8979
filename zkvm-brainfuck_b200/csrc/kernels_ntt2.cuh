// kernels_ntt2.cuh — production NTT passes: radix-16 butterflies in registers, Shoup twiddles.
//
// Same role as kernels_ntt.cuh (reference: `Radix2DitParallel::coset_lde_batch` behind
// `TwoAdicFriPcs::commit`, crates/stark/src/prover.rs:227,334,411) for columns of >= 2^12 points.
//
// One pass executes g <= 8 radix-2 stages (index bits [p, p+g)) as two register phases:
//   phase A: radix 2^G1 on the top G1 = g-4 bits of the pass digit,
//   phase B: radix 2^G2 on the low G2 = 4 bits,
// with one shared-memory exchange in between.  A CTA owns one tile position (2^g digits x 16 lanes)
// and loops over a group of columns: every column at that position uses the same twiddles, so the
// 15 + 15 inter-phase twiddles per thread are loaded once (coalesced, from a per-pass table) and kept
// in registers for the whole loop.  Inside a radix-16 the twiddles are the eight 16th roots of unity,
// held in constant memory.
//
// Multiplications by twiddles use Shoup's precomputed-quotient form (w, w' = floor(w 2^32 / p)):
//   q = mulhi(a, w');  r = a*w - q*p  in [0, 2p)   for ANY a < 2^32
// i.e. IMAD.HI + 2 IMAD and one min-correction, one FMA-pipe slot and one ALU op fewer than a
// Montgomery product, and the multiplicand may be an unreduced difference a - b + p.
// Data stay in Montgomery form (twiddles are plain residues, so x~ * w = (x w)~).
//
// Shared-memory tile: idx(d, lane) = d*17 + lane  (row stride 17 => conflict-free for the global
// staging in both orientations and for both phase access patterns).
#pragma once
#include "kb31.cuh"

namespace ntt2 {

constexpr int LANES = 16;
constexpr int ROW = 17;  // padded row stride of the smem tile
constexpr int G2 = 4;    // phase B radix bits

struct Tw {
    uint32_t w, wp;  // w (canonical residue) and floor(w * 2^32 / p)
};

// 16th roots of unity: c_w16[dir][j] = w16^j (dir 0) or w16^-j (dir 1), j < 8
__constant__ Tw c_w16[2][8];

__device__ __forceinline__ uint32_t shoup_lazy(uint32_t a, Tw t) {  // -> [0, 2p)
    uint32_t q = __umulhi(a, t.wp);
    return a * t.w - q * kb::P;
}
__device__ __forceinline__ uint32_t red(uint32_t x) { return kb::umin_(x, x - kb::P); }  // [0,2p) -> [0,p)

// ---- radix-2^K butterflies on K-digit register arrays (constant twiddles) -----------------------
// forward DIF: natural in -> outputs at bit-reversed positions; all values canonical on exit unless
// LAZY_LAST (then the last stage leaves values in [0, 2p) for a following Shoup multiply)
template <int K, bool LAZY_LAST>
__device__ __forceinline__ void radix_dif(uint32_t* v) {
#pragma unroll
    for (int s = K - 1; s >= 0; s--) {
        const int half = 1 << s;
#pragma unroll
        for (int gb = 0; gb < (1 << K); gb += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; k++) {
                uint32_t a = v[gb + k], b = v[gb + k + half];
                const bool lazy = LAZY_LAST && s == 0;
                uint32_t x = a + b;
                uint32_t y = a - b + kb::P;  // (0, 2p)
                const int tj = k * (8 >> s);  // exponent of w16
                if (tj != 0) y = shoup_lazy(y, c_w16[0][tj]);
                v[gb + k] = lazy ? x : red(x);
                v[gb + k + half] = lazy ? y : red(y);
            }
        }
    }
}
// inverse DIT: inputs at bit-reversed positions -> natural out, canonical
template <int K>
__device__ __forceinline__ void radix_dit(uint32_t* v) {
#pragma unroll
    for (int s = 0; s < K; s++) {
        const int half = 1 << s;
#pragma unroll
        for (int gb = 0; gb < (1 << K); gb += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; k++) {
                uint32_t a = v[gb + k], b = v[gb + k + half];
                const int tj = k * (8 >> s);
                if (tj != 0) b = red(shoup_lazy(b, c_w16[1][tj]));
                v[gb + k] = red(a + b);
                v[gb + k + half] = red(a - b + kb::P);
            }
        }
    }
}

__device__ __forceinline__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// One phase over a K-bit digit held as v[j*2^K + digit] for NG = 16 >> K independent groups.
// tw[j*(2^K-1) + q-1] multiplies output/input q of group j.
template <bool INV, int K, bool HAS_TW>
__device__ __forceinline__ void phase(uint32_t* v, const Tw* tw) {
    constexpr int R = 1 << K, NG = 16 >> K;
#pragma unroll
    for (int j = 0; j < NG; j++) {
        uint32_t* x = v + j * R;
        if (!INV) {
            radix_dif<K, HAS_TW>(x);
            if (HAS_TW) {
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const int q = brev(i, K);  // position i holds output q
                    x[i] = q ? red(shoup_lazy(x[i], tw[j * (R - 1) + q - 1])) : red(x[i]);
                }
            }
        } else {
            if (HAS_TW) {
#pragma unroll
                for (int i = 1; i < R; i++) {
                    const int q = brev(i, K);
                    x[i] = red(shoup_lazy(x[i], tw[j * (R - 1) + q - 1]));
                }
            }
            radix_dit<K>(x);
        }
    }
}

struct PassArgs {
    uint32_t* data;       // first column
    uint64_t col_stride;  // words between columns
    uint32_t ncols;       // total columns
    uint32_t cols_per_cta;
    uint32_t p;           // low bit of the pass
    uint32_t strided;     // p != 0
    const Tw* twA;        // [(2^G1 - 1)][2^(p+4)]  exponent q*m, transform size 2^(p+g)
    const Tw* twB;        // [15][2^p]              exponent q*lo, transform size 2^(p+4); null when p == 0
    // optional fused epilogue of the LAST inverse pass (coset scaling + blow-up): when pw != null the
    // result is not written back in place but as out[c][h*n + k] = x[k] * pw[h*n + k] for h < ncosets
    // (pw = (shift_h)^k / n in Montgomery form), out columns being ncosets*n words long.
    const uint32_t* pw;
    uint32_t* out;
    uint32_t ncosets;
    uint32_t log_n;
};

// G1 in 0..4 (g = G1 + 4).  Threads per CTA = 2^g.
template <bool INV, int G1>
__global__ void __launch_bounds__(256, 2) k_pass(PassArgs A) {
    constexpr int g = G1 + G2, NT = 1 << g;
    constexpr int RA = 1 << G1, NGA = 16 >> G1;  // phase A: radix, groups per thread
    __shared__ uint32_t sm[2][NT * ROW];
    const uint32_t t = threadIdx.x;
    const uint32_t p = A.p;
    // tile position
    uint64_t base;
    uint32_t lo_base = 0;
    if (A.strided) {
        uint32_t lo_blocks_log = p - 4;
        uint32_t lo_block = blockIdx.x & ((1u << lo_blocks_log) - 1);
        uint64_t hi = blockIdx.x >> lo_blocks_log;
        lo_base = lo_block << 4;
        base = (hi << (p + g)) | lo_base;
    } else {
        base = (uint64_t)blockIdx.x << (g + 4);
    }
    // ---- twiddles for this tile position (registers, reused for every column) -------------------
    Tw twa[15], twb[15];
    // phase A combos: cidx = t + j*NT, cidx = (r1 << 4) | lane ; m = (r1 << p) | lo
    if (G1 > 0) {
#pragma unroll
        for (int j = 0; j < NGA; j++) {
            uint32_t cidx = t + j * NT;
            uint32_t lane = cidx & 15, r1 = cidx >> 4;
            uint32_t m = A.strided ? ((r1 << p) | (lo_base + lane)) : r1;
            uint32_t M = A.strided ? (1u << (p + 4)) : 16u;
#pragma unroll
            for (int q = 1; q < RA; q++) twa[j * (RA - 1) + q - 1] = A.twA[(uint64_t)(q - 1) * M + m];
        }
    }
    const bool has_b = A.strided;
    if (has_b) {
        uint32_t lane = t & 15;
#pragma unroll
        for (int q = 1; q < 16; q++) twb[q - 1] = A.twB[(uint64_t)(q - 1) << p | (lo_base + lane)];
    }
    const uint32_t c_begin = blockIdx.y * A.cols_per_cta;
    const uint32_t c_end = min(A.ncols, c_begin + A.cols_per_cta);
    uint32_t v[16];
    int buf = 0;
    // Software pipeline: the tile of column c+1 is loaded into registers (nx) while column c is being
    // transformed, so HBM latency overlaps the butterflies inside one CTA.  (cp.async 4-byte LDGSTS
    // was measured slower: 16 LDGSTS per thread per column saturate the LSU issue rate.)
    uint32_t nx[16];
    auto load_tile = [&](uint32_t c) {
        const uint32_t* col = A.data + (uint64_t)c * A.col_stride + base;
        if (A.strided) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t f = t + i * NT;  // d = f >> 4, lane = f & 15
                nx[i] = col[((uint64_t)(f >> 4) << p) + (f & 15)];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) nx[i] = col[t + i * NT];  // lane = f >> g, d = f & (NT-1)
        }
    };
    if (c_begin < c_end) load_tile(c_begin);
    for (uint32_t c = c_begin; c < c_end; c++, buf ^= 1) {
        uint32_t* s = sm[buf];
        if (A.strided) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t f = t + i * NT;
                s[(f >> 4) * ROW + (f & 15)] = nx[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t f = t + i * NT;
                s[(f & (NT - 1)) * ROW + (f >> g)] = nx[i];
            }
        }
        if (c + 1 < c_end) load_tile(c + 1);
        __syncthreads();  // tile c visible (double buffering: iteration c-1 read the other buffer)
        // forward: phase A (high bits) then B; inverse: B then A
#pragma unroll
        for (int ph = 0; ph < 2; ph++) {
            const bool doA = INV ? (ph == 1) : (ph == 0);
            if (doA) {
                if (G1 > 0) {
#pragma unroll
                    for (int j = 0; j < NGA; j++) {
                        uint32_t cidx = t + j * NT;
                        uint32_t lane = cidx & 15, r1 = cidx >> 4;
#pragma unroll
                        for (int a = 0; a < RA; a++) v[j * RA + a] = s[((a << G2) | r1) * ROW + lane];
                    }
                    phase<INV, G1, true>(v, twa);
#pragma unroll
                    for (int j = 0; j < NGA; j++) {
                        uint32_t cidx = t + j * NT;
                        uint32_t lane = cidx & 15, r1 = cidx >> 4;
#pragma unroll
                        for (int a = 0; a < RA; a++) s[((a << G2) | r1) * ROW + lane] = v[j * RA + a];
                    }
                }
            } else {
                // phase B combo: (a, lane) = (t >> 4, t & 15)
                uint32_t lane = t & 15, a = t >> 4;
#pragma unroll
                for (int b = 0; b < 16; b++) v[b] = s[((a << G2) | b) * ROW + lane];
                if (has_b) phase<INV, G2, true>(v, twb);
                else phase<INV, G2, false>(v, twb);
#pragma unroll
                for (int b = 0; b < 16; b++) s[((a << G2) | b) * ROW + lane] = v[b];
            }
            if (G1 > 0 || ph == 1 || !INV) __syncthreads();
        }
        // ---- stage out ----------------------------------------------------------------------------
        if (INV && A.pw != nullptr) {
            const uint64_t n = 1ull << A.log_n;
            uint32_t* ocol = A.out + (uint64_t)c * (n * A.ncosets);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint32_t f = t + i * NT;
                uint64_t idx = A.strided ? base + ((uint64_t)(f >> 4) << p) + (f & 15) : base + f;
                uint32_t x = A.strided ? s[(f >> 4) * ROW + (f & 15)] : s[(f & (NT - 1)) * ROW + (f >> g)];
                for (uint32_t h = 0; h < A.ncosets; h++) ocol[h * n + idx] = kb::mul(x, __ldg(A.pw + h * n + idx));
            }
        } else {
            uint32_t* col = A.data + (uint64_t)c * A.col_stride + base;
            if (A.strided) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t f = t + i * NT;
                    col[((uint64_t)(f >> 4) << p) + (f & 15)] = s[(f >> 4) * ROW + (f & 15)];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t f = t + i * NT;
                    col[f] = s[(f & (NT - 1)) * ROW + (f >> g)];
                }
            }
        }
    }
}

// table entry i of a phase table: exponent e = q * m of a root of order 2^order_log (inverse: -e)
__global__ void k_build_tw(Tw* out, uint32_t nq, uint32_t M, uint32_t order_log, int inverse, uint32_t w_max /* Montgomery, order 2^24 */) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq * M) return;
    uint32_t q = (uint32_t)(i / M) + 1, m = (uint32_t)(i % M);
    uint64_t e = ((uint64_t)q * m) << (kb::TWO_ADICITY - order_log);  // exponent of the order-2^24 root
    if (inverse) e = ((1ull << kb::TWO_ADICITY) - e) & ((1ull << kb::TWO_ADICITY) - 1);
    uint32_t w = kb::from_mont(kb::pow(w_max, e));
    Tw tw;
    tw.w = w;
    tw.wp = (uint32_t)(((uint64_t)w << 32) / kb::P);
    out[i] = tw;
}

}  // namespace ntt2
