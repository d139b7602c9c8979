// kernels_ntt2.cuh — production NTT passes: radix-16 butterflies in registers, Shoup twiddles.
//
// Same role as kernels_ntt.cuh (reference: `Radix2DitParallel::coset_lde_batch` behind
// `TwoAdicFriPcs::commit`, crates/stark/src/prover.rs:227,334,411) for columns of >= 2^12 points.
//
// One pass executes g <= 8 radix-2 stages (index bits [p, p+g)) as two register phases:
//   phase A: radix 2^G1 on the top G1 = g-4 bits of the pass digit,
//   phase B: radix 2^G2 on the low G2 = 4 bits,
// with ONE shared-memory exchange in between.  A CTA owns one tile position (2^g digits x 16 lanes)
// and loops over a group of columns: every column at that position uses the same twiddles, so the
// 15 + 15 inter-phase twiddles per thread are loaded once (coalesced, from a per-pass table) and kept
// in registers for the whole loop.  Inside a radix-16 the twiddles are the eight 16th roots of unity,
// held in constant memory.
//
// Memory traffic per element and pass: one global load and one global store issued DIRECTLY in the
// register layout of the first / last phase (64-byte segments in strided passes, 64-byte runs or 16-byte
// vectors in the contiguous pass), one shared store + one shared load for the exchange: 4 LSU
// operations (the first version staged global<->shared and paid 8; ncu showed the LSU pipe 80% busy,
// profiles/r1_ntt_pass_v2.md).  The next column's tile is prefetched into registers while the current
// one is transformed, and the shared tile is double buffered, so one barrier per column suffices.
//
// Multiplications by twiddles use Shoup's precomputed-quotient form (w, w' = floor(w 2^32 / p)):
//   q = mulhi(a, w');  r = a*w - q*p  in [0, 2p)   for ANY a < 2^32
// i.e. IMAD.HI + 2 IMAD and one min-correction, one FMA-pipe slot and one ALU op fewer than a
// Montgomery product, and the multiplicand may be an unreduced difference a - b + p.
// Data stay in Montgomery form (twiddles are plain residues, so x~ * w = (x w)~).
#pragma once
#include "kb31.cuh"

namespace ntt2 {

constexpr int LANES = 16;
constexpr int ROW = 17;  // padded row stride of the strided-mode tile  [digit][lane]
constexpr int G2 = 4;    // phase B radix bits

struct Tw {
    uint32_t w, wp;  // w (canonical residue) and floor(w * 2^32 / p)
};

// Inter-phase twiddle representation (compile-time switch): Shoup pairs cost two registers
// per twiddle (60 of the 128 registers of a strided pass), single Montgomery words one register and one extra
// instruction per multiply.
#ifndef NTT2_MONT_TW
#define NTT2_MONT_TW 0
#endif
#ifndef NTT2_MINBLOCKS
#define NTT2_MINBLOCKS 2
#endif
#ifndef NTT2_CONTIG_MINBLOCKS
#define NTT2_CONTIG_MINBLOCKS 3  // contiguous passes at 85 registers, 3 CTAs/SM: forward pass 3.61 -> 3.39 ms at 2^22 x 512 half-columns (the strided
                                  // register-prefetching passes lose at 3: 67.4 -> 68.7 ms per commit with $BFGPU_NTT_TMA=0)
#endif
#if NTT2_MONT_TW
using TWT = uint32_t;
#else
using TWT = Tw;
#endif

// 16th roots of unity: c_w16[dir][j] = w16^j (dir 0) or w16^-j (dir 1), j < 8
__constant__ Tw c_w16[2][8];

__device__ __forceinline__ uint32_t shoup_lazy(uint32_t a, Tw t) {  // -> [0, 2p)
    uint32_t q = __umulhi(a, t.wp);
    return a * t.w - q * kb::P;
}
__device__ __forceinline__ uint32_t red(uint32_t x) { return kb::umin_(x, x - kb::P); }  // [0,2p) -> [0,p)

// ---- radix-2^K butterflies on K-digit register arrays (constant twiddles) -----------------------
// Every stage is ONE flat loop with a compile-time trip count, so that nvcc unrolls it completely (the nested
// gb/k loops of the first version were left partially rolled: ptxas then emulated the register-array indexing
// with ~150 predicated moves per column).
// forward DIF stage S (butterfly span 2^S); LAZY: leave the results in [0, 2p) for a following Shoup multiply
template <int K, int S, bool LAZY>
__device__ __forceinline__ void dif_stage(uint32_t* v) {
    constexpr int half = 1 << S;
#pragma unroll
    for (int i = 0; i < (1 << (K - 1)); i++) {
        const int lo = ((i >> S) << (S + 1)) | (i & (half - 1)), hi = lo + half;
        const int tj = (i & (half - 1)) * (8 >> S);  // exponent of w16
        uint32_t a = v[lo], b = v[hi];
        uint32_t x = a + b;
        uint32_t y = a - b + kb::P;  // (0, 2p)
        if (tj != 0) y = shoup_lazy(y, c_w16[0][tj]);
        v[lo] = LAZY ? x : red(x);
        v[hi] = LAZY ? y : red(y);
    }
}
// forward DIF: natural in -> outputs at bit-reversed positions; canonical on exit unless LAZY_LAST
template <int K, bool LAZY_LAST>
__device__ __forceinline__ void radix_dif(uint32_t* v) {
    if constexpr (K >= 4) dif_stage<K, 3, false>(v);
    if constexpr (K >= 3) dif_stage<K, 2, false>(v);
    if constexpr (K >= 2) dif_stage<K, 1, false>(v);
    if constexpr (K >= 1) dif_stage<K, 0, LAZY_LAST>(v);
}
template <int K, int S>
__device__ __forceinline__ void dit_stage(uint32_t* v) {
    constexpr int half = 1 << S;
#pragma unroll
    for (int i = 0; i < (1 << (K - 1)); i++) {
        const int lo = ((i >> S) << (S + 1)) | (i & (half - 1)), hi = lo + half;
        const int tj = (i & (half - 1)) * (8 >> S);
        uint32_t a = v[lo], b = v[hi];
        if (tj != 0) b = red(shoup_lazy(b, c_w16[1][tj]));
        v[lo] = red(a + b);
        v[hi] = red(a - b + kb::P);
    }
}
// inverse DIT: inputs at bit-reversed positions -> natural out, canonical
template <int K>
__device__ __forceinline__ void radix_dit(uint32_t* v) {
    if constexpr (K >= 1) dit_stage<K, 0>(v);
    if constexpr (K >= 2) dit_stage<K, 1>(v);
    if constexpr (K >= 3) dit_stage<K, 2>(v);
    if constexpr (K >= 4) dit_stage<K, 3>(v);
}

__device__ __forceinline__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// inter-phase twiddle multiply: Shoup pair (any a < 2^32 -> canonical) or a single Montgomery word (twiddle stored
// as w * 2^32 mod p; one register instead of two, one instruction more)
__device__ __forceinline__ uint32_t twmul(uint32_t a, Tw t) { return red(shoup_lazy(a, t)); }
__device__ __forceinline__ uint32_t twmul(uint32_t a, uint32_t t) { return kb::mul(a, t); }

// One phase over a K-bit digit held as v[j*2^K + digit] for NG = 16 >> K independent groups.
// tw[j*(2^K-1) + q-1] multiplies output/input q of group j.
template <bool INV, int K, bool HAS_TW, class TWF>
__device__ __forceinline__ void phase(uint32_t* v, TWF tw) {
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
                    x[i] = q ? twmul(x[i], tw(j * (R - 1) + q - 1)) : red(x[i]);
                }
            }
        } else {
            if (HAS_TW) {
#pragma unroll
                for (int i = 1; i < R; i++) {
                    const int q = brev(i, K);
                    x[i] = twmul(x[i], tw(j * (R - 1) + q - 1));
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
    uint32_t p;           // low bit of the pass (0 in the contiguous pass)
    const TWT* twA;       // [(2^G1 - 1)][2^(p+4)]  exponent q*m, transform size 2^(p+g)
    const TWT* twB;       // [15][2^p]              exponent q*lo, transform size 2^(p+4); null when p == 0
    // optional fused epilogue of the LAST inverse pass (coset scaling + blow-up): when pw != null the
    // result is not written back in place but as out[c][h*n + k] = x[k] * pw[h*n + k] for h < ncosets
    // (pw = (shift_h)^k / n in Montgomery form), out columns being ncosets*n words long.
    const uint32_t* pw;
    uint32_t* out;
    uint32_t ncosets;
    uint32_t log_n;
    // TURN (last inverse pass fused with the first forward pass): forward tables of the same (p, g)
    const TWT* twA2;
    const TWT* twB2;
};

// Shared tile index of element (digit d, lane l).
//   strided pass   : d*17 + l  (phase A warps read 16 lanes x 2 digits, phase B 16 lanes x 2 high digits: both conflict free)
//   contiguous pass: l*(2^g + 16) + a*16 + swizzled b, where d = a*16 + b and the four 16-byte chunks of a thread's
//                    64-byte row are XOR-ed with (a >> 1) & 3, so that phase B can move its 16 consecutive words with
//                    four conflict-free 128-bit accesses while phase A (16 consecutive b, fixed a) stays conflict free.
template <int g, bool CONTIG>
__device__ __forceinline__ uint32_t tile_idx(uint32_t d, uint32_t l) {
    if (!CONTIG) return d * ROW + l;
    uint32_t a = d >> 4, b = d & 15;
    return l * ((1u << g) + 16) + a * 16 + ((((b >> 2) ^ (a >> 1)) & 3) << 2) + (b & 3);
}

// G1 in 0..4 (g = G1 + 4).  Threads per CTA = 2^g.  CONTIG: p == 0 (lanes = 16 consecutive runs of 2^g words).
// EPI: the fused coset epilogue (separate instantiation: keeps its registers out of the plain passes)
// DUAL (forward, strided, G1 > 0): the FIRST forward pass of a blow-up-2 LDE reads each coefficient tile ONCE and produces both
// cosets from it: v = x * pw_h on load (pw_h[k] = shift_h^k / n), transform, store into half-column h of the output — the coset
// scaling needs no pass of its own and no second read of the coefficients (round 2: replaces the EPI epilogue of the last inverse
// pass, which wrote 2n words per column that this pass then read back; -9 GB of 90 at 2^22 x 256).
// TURN (inverse, strided, G1 > 0): the LAST inverse pass and the DUAL first forward pass in one kernel.  Both act on the same tile
// (top g index bits) and the inverse pass ends in exactly the register layout the forward pass starts from (phase A), so the
// coefficients never travel to HBM and back between the two: 4 B read + 8 B written per element instead of 8 + 12, one set of global
// loads / stores instead of two.  Three shared exchanges per column alternate between the two tiles.
#ifndef NTT2_TURN_MINBLOCKS
#define NTT2_TURN_MINBLOCKS 3  // CTAs of 2^g <= 128 threads per SM the register budget is set for (g = 8: one)
#endif
template <bool INV, int G1, bool CONTIG, bool EPI = false, bool DUAL = false, bool TURN = false>
__global__ void __launch_bounds__(TURN ? (1 << (G1 + 4)) : 256, TURN ? (G1 == 4 ? 1 : NTT2_TURN_MINBLOCKS) : EPI ? 1 : CONTIG ? NTT2_CONTIG_MINBLOCKS : NTT2_MINBLOCKS) k_pass(PassArgs A) {
    static_assert(!DUAL || (!INV && !CONTIG && !EPI && G1 > 0), "DUAL is the strided forward pass with two register phases");
    static_assert(!TURN || (INV && !CONTIG && !EPI && !DUAL && G1 > 0), "TURN is the strided last inverse pass with two register phases");
    constexpr int g = G1 + G2, NT = 1 << g;
    constexpr int RA = 1 << G1, NGA = 16 >> G1;  // phase A: radix, groups per thread
    constexpr int TILE_WORDS = CONTIG ? 16 * (NT + 16) : NT * ROW;
    __shared__ __align__(16) uint32_t sm[2][TILE_WORDS];
    const uint32_t t = threadIdx.x;
    const uint32_t p = A.p;
    // tile position
    uint64_t base;
    uint32_t lo_base = 0;
    if (!CONTIG) {
        uint32_t lo_blocks_log = p - 4;
        uint32_t lo_block = blockIdx.x & ((1u << lo_blocks_log) - 1);
        uint64_t hi = blockIdx.x >> lo_blocks_log;
        lo_base = lo_block << 4;
        base = (hi << (p + g)) | lo_base;
    } else {
        base = (uint64_t)blockIdx.x << (g + 4);
    }
    // ---- per-thread coordinates ------------------------------------------------------------------------
    // phase A combos j < NGA: (r1, lane); strided: lane fastest across the warp, contiguous: r1 fastest.
    // phase B: one combo (a, lane); strided: lane fastest; contiguous: lane = t >> G1, a = t & (RA-1).
    uint32_t laneA[NGA], r1A[NGA];
#pragma unroll
    for (int j = 0; j < NGA; j++) {
        uint32_t cidx = t + j * NT;
        laneA[j] = CONTIG ? cidx >> 4 : cidx & 15;
        r1A[j] = CONTIG ? cidx & 15 : cidx >> 4;
    }
    const uint32_t laneB = CONTIG ? t >> G1 : t & 15;
    const uint32_t aB = CONTIG ? t & (RA - 1) : t >> 4;
    // word offset (inside the tile's address range, < 2^32) of element (d, lane): 32-bit arithmetic on purpose, the
    // 64-bit column base is CTA-uniform (64-bit per-element address math was 5 instructions per access)
    auto goff = [&](uint32_t d, uint32_t lane) -> uint32_t { return CONTIG ? (lane << g) + d : (d << p) + lane; };

    // ---- twiddles for this tile position (registers, reused for every column) -------------------
    // phase A twiddles (per thread) live in registers; phase B twiddles depend on the lane only and are read from
    // shared memory (15 broadcast LDS.64 per column instead of 30 more registers, which spilled)
    TWT twa[15];
    __shared__ TWT s_twb[15][LANES];
    if (G1 > 0) {
#pragma unroll
        for (int j = 0; j < NGA; j++) {
            uint32_t m = CONTIG ? r1A[j] : ((r1A[j] << p) | (lo_base + laneA[j]));
            uint32_t M = CONTIG ? 16u : (1u << (p + 4));
#pragma unroll
            for (int q = 1; q < RA; q++) twa[j * (RA - 1) + q - 1] = A.twA[(uint64_t)(q - 1) * M + m];
        }
    }
    if (!CONTIG) {
        for (uint32_t i = t; i < 15 * LANES; i += NT) s_twb[i / LANES][i % LANES] = A.twB[(uint64_t)(i / LANES) << p | (lo_base + i % LANES)];
        __syncthreads();
    }
    // TURN: forward twiddles of the same tile position; phase A in registers, phase B in shared memory like the inverse ones
    TWT twa2[TURN ? 15 : 1];
    __shared__ TWT s_twb2[TURN ? 15 : 1][LANES];
    if (TURN) {
#pragma unroll
        for (int j = 0; j < NGA; j++) {
            uint32_t m = (r1A[j] << p) | (lo_base + laneA[j]);
            uint32_t M = 1u << (p + 4);
#pragma unroll
            for (int q = 1; q < RA; q++) twa2[TURN ? j * (RA - 1) + q - 1 : 0] = A.twA2[(uint64_t)(q - 1) * M + m];
        }
        for (uint32_t i = t; i < 15 * LANES; i += NT) s_twb2[TURN ? i / LANES : 0][i % LANES] = A.twB2[(uint64_t)(i / LANES) << p | (lo_base + i % LANES)];
        __syncthreads();
    }
    auto twA = [&](int i) { return twa[i]; };
    auto twB = [&](int i) { return s_twb[i][laneB]; };
    auto twA2 = [&](int i) { return twa2[TURN ? i : 0]; };
    auto twB2 = [&](int i) { return s_twb2[TURN ? i : 0][laneB]; };
    const uint32_t c_begin = blockIdx.y * A.cols_per_cta;
    const uint32_t c_end = min(A.ncols, c_begin + A.cols_per_cta);

    // ---- global <-> register moves in the phase layouts --------------------------------------------------
    uint32_t nx[16];
    auto load_A = [&](const uint32_t* col, uint32_t* r) {
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) r[j * RA + a] = col[goff(((uint32_t)a << G2) | r1A[j], laneA[j])];
    };
    auto load_B = [&](const uint32_t* col, uint32_t* r) {
        if (CONTIG) {
            const uint4* q = reinterpret_cast<const uint4*>(col + goff(aB << G2, laneB));
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint4 x = q[k];
                r[4 * k] = x.x; r[4 * k + 1] = x.y; r[4 * k + 2] = x.z; r[4 * k + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) r[b] = col[goff((aB << G2) | b, laneB)];
        }
    };
    auto store_B = [&](uint32_t* col, const uint32_t* r) {
        if (CONTIG) {
            uint4* q = reinterpret_cast<uint4*>(col + goff(aB << G2, laneB));
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) col[goff((aB << G2) | b, laneB)] = r[b];
        }
    };
    // ---- shared exchange -------------------------------------------------------------------------------------
    auto sts_A = [&](uint32_t* s, const uint32_t* r) {
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) s[tile_idx<g, CONTIG>(((uint32_t)a << G2) | r1A[j], laneA[j])] = r[j * RA + a];
    };
    auto lds_A = [&](const uint32_t* s, uint32_t* r) {
#pragma unroll
        for (int j = 0; j < NGA; j++)
#pragma unroll
            for (int a = 0; a < RA; a++) r[j * RA + a] = s[tile_idx<g, CONTIG>(((uint32_t)a << G2) | r1A[j], laneA[j])];
    };
    auto sts_B = [&](uint32_t* s, const uint32_t* r) {
        if (CONTIG) {
            uint32_t rowbase = laneB * (NT + 16) + aB * 16;
#pragma unroll
            for (int k = 0; k < 4; k++)
                *reinterpret_cast<uint4*>(s + rowbase + (((k ^ (aB >> 1)) & 3) << 2)) = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) s[tile_idx<g, CONTIG>((aB << G2) | b, laneB)] = r[b];
        }
    };
    auto lds_B = [&](const uint32_t* s, uint32_t* r) {
        if (CONTIG) {
            uint32_t rowbase = laneB * (NT + 16) + aB * 16;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint4 x = *reinterpret_cast<const uint4*>(s + rowbase + (((k ^ (aB >> 1)) & 3) << 2));
                r[4 * k] = x.x; r[4 * k + 1] = x.y; r[4 * k + 2] = x.z; r[4 * k + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int b = 0; b < 16; b++) r[b] = s[tile_idx<g, CONTIG>((aB << G2) | b, laneB)];
        }
    };

    auto colptr = [&](uint32_t c) { return A.data + (uint64_t)c * A.col_stride + base; };
    // first phase of the pass: forward = A (when G1 > 0), inverse = B
    auto load_first = [&](uint32_t c) {
        if (!INV && G1 > 0) load_A(colptr(c), nx);
        else load_B(colptr(c), nx);
    };
    if (c_begin < c_end) load_first(c_begin);
    uint32_t v[16];
    int buf = 0;
    for (uint32_t c = c_begin; c < c_end; c++, buf ^= DUAL ? 0 : 1) {
        uint32_t* s = sm[buf];
        if (TURN) {
            // inverse pass (phase B, exchange, phase A), then per coset: scale, forward phase A, exchange, phase B, store
            const uint64_t n = 1ull << A.log_n;
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = nx[i];
            if (c + 1 < c_end) load_first(c + 1);
            phase<true, G2, true>(v, twB);
            sts_B(s, v);
            __syncthreads();
            lds_A(s, v);
            phase<true, G1, true>(v, twA);
            uint32_t x[16];
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = v[i];
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t* pw = A.pw + (uint64_t)h * n + base;
#pragma unroll
                for (int j = 0; j < NGA; j++)
#pragma unroll
                    for (int a = 0; a < RA; a++) {
                        const uint32_t off = goff(((uint32_t)a << G2) | r1A[j], laneA[j]);
                        v[j * RA + a] = kb::mul(x[j * RA + a], __ldg(pw + off));
                    }
                uint32_t* sh = sm[buf ^ 1 ^ h];  // exchanges alternate tiles: inverse sm[buf], h = 0 sm[buf^1], h = 1 sm[buf]; next column starts at sm[buf^1]
                phase<false, G1, true>(v, twA2);
                sts_A(sh, v);
                __syncthreads();
                lds_B(sh, v);
                phase<false, G2, true>(v, twB2);
                store_B(A.out + (uint64_t)c * (2 * n) + (uint64_t)h * n + base, v);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = nx[i];
        if (c + 1 < c_end) load_first(c + 1);  // software pipeline: next column's tile in flight during the butterflies
        if (DUAL) {
            // two exchanges per column: they alternate between the two shared tiles (h = 0: sm[buf], h = 1: sm[buf ^ 1]), so one barrier
            // per exchange still separates every reuse of a tile from the reads of its previous contents
            const uint64_t n = 1ull << A.log_n;
            uint32_t x[16];
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = v[i];
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t* pw = A.pw + (uint64_t)h * n + base;
#pragma unroll
                for (int j = 0; j < NGA; j++)
#pragma unroll
                    for (int a = 0; a < RA; a++) {
                        const uint32_t off = goff(((uint32_t)a << G2) | r1A[j], laneA[j]);
                        v[j * RA + a] = kb::mul(x[j * RA + a], __ldg(pw + off));
                    }
                uint32_t* sh = sm[buf ^ h];
                phase<false, G1, true>(v, twA);
                sts_A(sh, v);
                __syncthreads();
                lds_B(sh, v);
                phase<false, G2, true>(v, twB);
                store_B(A.out + (uint64_t)c * (2 * n) + (uint64_t)h * n + base, v);
            }
            continue;
        }
        if (!INV) {
            if (G1 > 0) {
                phase<false, G1, true>(v, twA);
                sts_A(s, v);
                __syncthreads();  // double-buffered tile: one barrier per column
                lds_B(s, v);
            }
            if (!CONTIG) phase<false, G2, true>(v, twB);
            else phase<false, G2, false>(v, twB);
            store_B(colptr(c), v);
        } else {
            if (!CONTIG) phase<true, G2, true>(v, twB);
            else phase<true, G2, false>(v, twB);
            if (G1 > 0) {
                sts_B(s, v);
                __syncthreads();
                lds_A(s, v);
                phase<true, G1, true>(v, twA);
            }
            // ---- store in the phase-A layout (phase-B layout when G1 == 0) ---------------------------------
            if (EPI) {
                // fused coset epilogue for blow-up 2 (the host only requests it when ncosets == 2): CTA-uniform 64-bit
                // bases, 32-bit per-element offsets
                const uint64_t n = 1ull << A.log_n;
                const uint32_t* pw0 = A.pw + base;
                const uint32_t* pw1 = pw0 + n;
                uint32_t* o0 = A.out + (uint64_t)c * (2 * n) + base;
                uint32_t* o1 = o0 + n;
                if (G1 > 0) {
#pragma unroll
                    for (int j = 0; j < NGA; j++)
#pragma unroll
                        for (int a = 0; a < RA; a++) {
                            uint32_t off = goff(((uint32_t)a << G2) | r1A[j], laneA[j]);
                            uint32_t x = v[j * RA + a];
                            o0[off] = kb::mul(x, __ldg(pw0 + off));
                            o1[off] = kb::mul(x, __ldg(pw1 + off));
                        }
                } else {
#pragma unroll
                    for (int b = 0; b < 16; b++) {
                        uint32_t off = goff((aB << G2) | b, laneB);
                        o0[off] = kb::mul(v[b], __ldg(pw0 + off));
                        o1[off] = kb::mul(v[b], __ldg(pw1 + off));
                    }
                }
            } else if (G1 > 0) {
                uint32_t* col = colptr(c);
#pragma unroll
                for (int j = 0; j < NGA; j++)
#pragma unroll
                    for (int a = 0; a < RA; a++) col[goff(((uint32_t)a << G2) | r1A[j], laneA[j])] = v[j * RA + a];
            } else {
                store_B(colptr(c), v);
            }
        }
    }
}

// table entry i of a phase table: exponent e = q * m of a root of order 2^order_log (inverse: -e)
__global__ void k_build_tw(TWT* out, uint32_t nq, uint32_t M, uint32_t order_log, int inverse, uint32_t w_max /* Montgomery, order 2^24 */) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq * M) return;
    uint32_t q = (uint32_t)(i / M) + 1, m = (uint32_t)(i % M);
    uint64_t e = ((uint64_t)q * m) << (kb::TWO_ADICITY - order_log);  // exponent of the order-2^24 root
    if (inverse) e = ((1ull << kb::TWO_ADICITY) - e) & ((1ull << kb::TWO_ADICITY) - 1);
#if NTT2_MONT_TW
    out[i] = kb::pow(w_max, e);
#else
    uint32_t w = kb::from_mont(kb::pow(w_max, e));
    Tw tw;
    tw.w = w;
    tw.wp = (uint32_t)(((uint64_t)w << 32) / kb::P);
    out[i] = tw;
#endif
}

}  // namespace ntt2
