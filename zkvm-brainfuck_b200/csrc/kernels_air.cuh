// kernels_air.cuh — LogUp permutation trace (K4) and quotient (K5) kernels, driven by the generated
// per-chip programs in gen_air.cuh.
//
// Replaces `generate_permutation_trace` (reference crates/stark/src/permutation.rs:75-148) and
// `quotient_values` (crates/stark/src/quotient.rs:18-165) with their `selectors_on_coset` inputs
// (SURVEY.md Appendix B.7).  Everything is indexed by STORED row: traces and LDEs live column-major with
// bit-reversed rows, so thread t reads row t coalesced and its "next" row (natural index + 2^lqd) is a
// constant stored-index offset for almost all t.
#pragma once
#include "kb31.cuh"

namespace air {

struct Selectors {
    uint32_t is_first, is_last, is_trans;
};
struct Challenges {
    kb::Ext alpha;         // LogUp alpha
    kb::Ext beta_pow[8];   // LogUp beta^0..beta^7
    kb::Ext cumulative_sum;
};

// Extension products / inversions of the generated programs.  Inlined, the Cpu constraint program is ~100 KB of straight-line
// SASS (48 products of ~140 instructions each) and the warps of an SM, all at different places in it, stall on instruction fetch
// (ncu r2: stall_no_instruction 3.8 / 5.2 warps per issue slot in k_quotient<Cpu> / k_perm_rows<Cpu>, the top stall reason);
// as calls to ONE copy of the product the program is a few KB and the product's body stays in the instruction cache.
#ifndef BFGPU_AIR_NOINLINE
#define BFGPU_AIR_NOINLINE 1
#endif
#if BFGPU_AIR_NOINLINE
__device__ __noinline__ kb::Ext ext_mul_call(kb::Ext a, kb::Ext b) { return kb::ext_mul(a, b); }
__device__ __noinline__ kb::Ext ext_inv_call(kb::Ext a) { return kb::ext_inv(a); }
#define AIR_EXT_MUL air::ext_mul_call
#define AIR_EXT_INV air::ext_inv_call
#else
#define AIR_EXT_MUL kb::ext_mul
#define AIR_EXT_INV kb::ext_inv
#endif

}  // namespace air
#include "gen_air.cuh"

namespace air {

// ---- loaders -----------------------------------------------------------------------------------------------
// trace rows (perm trace generation): local row only
struct TraceLoader {
    const uint32_t* main;
    const uint32_t* prep;
    uint64_t rows;
    uint64_t t;
    __device__ __forceinline__ uint32_t main0(int c) const { return main[(uint64_t)c * rows + t]; }
    __device__ __forceinline__ uint32_t prep0(int c) const { return prep[(uint64_t)c * rows + t]; }
    __device__ __forceinline__ uint32_t main1(int) const { return 0; }
    __device__ __forceinline__ uint32_t prep1(int) const { return 0; }
};
// LDE rows (quotient): local = stored row t, next = stored row tn
struct LdeLoader {
    const uint32_t* main;
    const uint32_t* prep;
    const uint32_t* perm;  // 4 base columns per ext column
    uint64_t rows;
    uint64_t t, tn;
    __device__ __forceinline__ uint32_t main0(int c) const { return main[(uint64_t)c * rows + t]; }
    __device__ __forceinline__ uint32_t main1(int c) const { return main[(uint64_t)c * rows + tn]; }
    __device__ __forceinline__ uint32_t prep0(int c) const { return prep[(uint64_t)c * rows + t]; }
    __device__ __forceinline__ uint32_t prep1(int c) const { return prep[(uint64_t)c * rows + tn]; }
    __device__ __forceinline__ kb::Ext permx(int j, uint64_t r) const {
        const uint32_t* p = perm + (uint64_t)(4 * j) * rows + r;
        return kb::Ext{{p[0], p[rows], p[2 * rows], p[3 * rows]}};
    }
    __device__ __forceinline__ kb::Ext perm0(int j) const { return permx(j, t); }
    __device__ __forceinline__ kb::Ext perm1(int j) const { return permx(j, tn); }
};

// ---- permutation trace ----------------------------------------------------------------------------------------
// Thread t handles stored (bit-reversed) trace row t = natural row br(t).  Writes the batch columns
// (perm_w - 1 ext columns = 4 base columns each) at stored position t and the row sum at NATURAL position
// br(t) of `rowsum` (ext, AoS) for the prefix scan.
template <int CHIP>  // one instantiation per chip: register allocation follows the chip's own program
__global__ void __launch_bounds__(128) k_perm_rows(int chip, const uint32_t* __restrict__ main, const uint32_t* __restrict__ prep, unsigned log_n,
                                                   Challenges ch, int perm_w, uint32_t* __restrict__ perm_out, uint32_t* __restrict__ rowsum) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t n = 1ull << log_n;
    if (t >= n) return;
    TraceLoader ld{main, prep, n, t};
    kb::Ext out[MAX_PERM_W];
    (void)chip;
    air_perm_row(CHIP, ld, ch, out);
    kb::Ext s = kb::ext_zero();
    for (int j = 0; j < perm_w - 1; j++) {
        s = kb::ext_add(s, out[j]);
#pragma unroll
        for (int e = 0; e < 4; e++) perm_out[(uint64_t)(4 * j + e) * n + t] = out[j].c[e];
    }
    uint64_t nat = kb::bitrev((uint32_t)t, log_n);
    *reinterpret_cast<uint4*>(rowsum + 4 * nat) = make_uint4(s.c[0], s.c[1], s.c[2], s.c[3]);
}

// inclusive prefix sums of n ext elements (AoS), three kernels: block scans, scan of block totals, fix-up.
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;
__device__ __forceinline__ uint4 ext4_add(uint4 a, uint4 b) {
    return make_uint4(kb::add(a.x, b.x), kb::add(a.y, b.y), kb::add(a.z, b.z), kb::add(a.w, b.w));
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocks(uint4* __restrict__ data, uint64_t n, uint4* __restrict__ totals) {
    __shared__ uint4 sm[SCAN_THREADS];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint4 v[SCAN_ITEMS];
    uint4 run = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = base + i < n ? data[base + i] : make_uint4(0, 0, 0, 0);
        run = ext4_add(run, v[i]);
        v[i] = run;
    }
    sm[threadIdx.x] = run;
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {  // Hillis-Steele over the per-thread totals
        uint4 add = threadIdx.x >= off ? sm[threadIdx.x - off] : make_uint4(0, 0, 0, 0);
        __syncthreads();
        sm[threadIdx.x] = ext4_add(sm[threadIdx.x], add);
        __syncthreads();
    }
    uint4 prefix = threadIdx.x ? sm[threadIdx.x - 1] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) data[base + i] = ext4_add(v[i], prefix);
    if (threadIdx.x == SCAN_THREADS - 1 && totals) totals[blockIdx.x] = sm[threadIdx.x];
}
// data[i] += totals_scanned[block - 1]  (inner recursion levels)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint4* __restrict__ data, uint64_t n, const uint4* __restrict__ totals_scanned) {
    uint64_t i = blockIdx.x * (uint64_t)SCAN_THREADS + threadIdx.x;
    if (i >= n) return;
    uint64_t blk = i / SCAN_BLOCK;
    if (blk) data[i] = ext4_add(data[i], totals_scanned[blk - 1]);
}
// data[i] += exclusive prefix of the block totals; also scatter the final value into the last perm column at the
// bit-reversed position (4 base columns starting at `col`)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_fixup(const uint4* __restrict__ data, uint64_t n, const uint4* __restrict__ totals_scanned,
                                                             unsigned log_n, uint32_t* __restrict__ col) {
    uint64_t i = blockIdx.x * (uint64_t)SCAN_THREADS + threadIdx.x;
    if (i >= n) return;
    uint64_t blk = i / SCAN_BLOCK;
    uint4 v = data[i];
    if (blk && totals_scanned) v = ext4_add(v, totals_scanned[blk - 1]);
    uint64_t pos = kb::bitrev((uint32_t)i, log_n);
    col[pos] = v.x;
    col[n + pos] = v.y;
    col[2 * n + pos] = v.z;
    col[3 * n + pos] = v.w;
}

// ---- quotient ---------------------------------------------------------------------------------------------------------
struct QuotientArgs {
    int chip;
    const uint32_t* main;  // LDEs, column-major, bit-reversed rows, 2^(log_n + lqd) rows
    const uint32_t* prep;
    const uint32_t* perm;
    unsigned log_n;         // trace height
    unsigned lqd;           // log quotient degree (1 for every chip of this machine)
    uint32_t shift;         // coset shift (GENERATOR), Montgomery
    uint32_t g_inv;         // inverse of the trace-domain generator
    uint32_t zh[2];         // Z_H on the coset takes 2^lqd values (lqd = 1: even / odd natural index)
    uint32_t zh_inv[2];
    const kb::Ext* apow;    // alpha^0 .. alpha^(n_constraints-1)
    const uint32_t* tw;     // w_{2^24}^e table
    uint32_t* out;          // 2^lqd chunk matrices, each 4 base columns x 2^log_n rows, chunk-major, rows bit-reversed
};
__device__ __forceinline__ uint32_t root_pow_(const uint32_t* __restrict__ tw, unsigned log_n, uint32_t j) {
    if (log_n == 0) return kb::ONE;
    uint32_t half = 1u << (log_n - 1);
    uint32_t v = __ldg(tw + ((uint64_t)(j & (half - 1)) << (kb::TWO_ADICITY - log_n)));
    return (j & half) ? kb::neg(v) : v;
}
#ifndef BFGPU_QUOT_MINBLOCKS
#define BFGPU_QUOT_MINBLOCKS 6  // 85 registers (a few spills) instead of 102: 6 CTAs/SM; quotient phase 5.6 -> 5.0 ms at 2^22 rows (8: 5.04)
#endif
template <int CHIP>
__global__ void __launch_bounds__(128, BFGPU_QUOT_MINBLOCKS) k_quotient(QuotientArgs A, Challenges ch) {
    const unsigned L = A.log_n + A.lqd;
    const uint64_t N = 1ull << L, n = 1ull << A.log_n;
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= N) return;
    const uint32_t i = kb::bitrev((uint32_t)t, L);  // natural index on the quotient domain
    const uint32_t inext = (i + (1u << A.lqd)) & (uint32_t)(N - 1);
    LdeLoader ld{A.main, A.prep, A.perm, N, t, kb::bitrev(inext, L)};
    // selectors_on_coset: x = shift * w^i;  Z_H(x) = shift^n (w^n)^i - 1 depends on i mod 2^lqd only
    const uint32_t x = kb::mul(A.shift, root_pow_(A.tw, L, i));
    const uint32_t zh = A.zh[i & ((1u << A.lqd) - 1)], zh_inv = A.zh_inv[i & ((1u << A.lqd) - 1)];
    Selectors sel;
    // 1/(x - 1) and 1/(x - g^-1) with one inversion (x runs over a coset disjoint from the trace domain: neither is zero)
    const uint32_t d_first = kb::sub(x, kb::ONE), d_last = kb::sub(x, A.g_inv);
    const uint32_t ip = kb::mul(zh, kb::inv(kb::mul(d_first, d_last)));
    sel.is_first = kb::mul(ip, d_last);
    sel.is_last = kb::mul(ip, d_first);
    sel.is_trans = d_last;
    kb::Ext acc = kb::ext_zero();
    air_constraints(CHIP, ld, sel, ch, A.apow, acc);
    acc = kb::ext_scale(acc, zh_inv);
    // chunk c = i mod 2^lqd holds natural rows i >> lqd; its bit-reversed position is t mod n, and c = t >> log_n
    uint32_t c = (uint32_t)(t >> A.log_n);
    uint64_t pos = t & (n - 1);
    uint32_t* o = A.out + (uint64_t)c * 4 * n + pos;
    o[0] = acc.c[0];
    o[n] = acc.c[1];
    o[2 * n] = acc.c[2];
    o[3 * n] = acc.c[3];
}


// ---- quotient over ROW SHARDS (one proof over several GPUs, dist_prove.cuh) ----------------------------------------------------
// The committed LDEs of a sharded commitment live as row shards: rank r holds stored rows [r * rpg, (r + 1) * rpg) of every matrix,
// column-major with rpg words per column, and every rank has all shards mapped (CUDA IPC over NVLink).  A thread handles one LOCAL
// stored row; its "next" row (natural index + 2) sits at a fixed stored offset that usually belongs to another rank and is read
// straight out of that peer's HBM.  The preprocessed LDE is replicated (full height on every rank).  Each quotient word goes to the
// rank that owns that COLUMN of the quotient-chunk matrix (column-sharded LDE of the next commitment): out_col[chunk][k] points into
// the owner's coefficient buffer, so the row -> column exchange of the quotient commit is the kernel's own store.
constexpr int AIR_MAX_WORLD = 16;
struct ShardedLde {
    const uint32_t* shard[AIR_MAX_WORLD];  // shard r of this matrix (null when the chip has no such trace)
};
struct QuotientShardArgs {
    int chip;
    ShardedLde main, perm;
    const uint32_t* prep;  // full preprocessed LDE (2^(log_n + lqd) rows) or null
    unsigned log_n, lqd, log_rpg;
    uint32_t rank;
    uint32_t shift, g_inv, zh[2], zh_inv[2];
    const kb::Ext* apow;
    const uint32_t* tw;
    uint32_t* out_col[2][4];  // [chunk][extension coefficient]: column of 2^log_n words on the owning rank
};
struct ShardLoader {
    const uint32_t *m0, *m1, *p0, *p1, *q0, *q1;  // row pointers (column 0) of the local / next row in main, prep, perm
    uint64_t ms, ps;                               // words between columns: shard height (main, perm), full height (prep)
    __device__ __forceinline__ uint32_t main0(int c) const { return m0[(uint64_t)c * ms]; }
    __device__ __forceinline__ uint32_t main1(int c) const { return m1[(uint64_t)c * ms]; }
    __device__ __forceinline__ uint32_t prep0(int c) const { return p0[(uint64_t)c * ps]; }
    __device__ __forceinline__ uint32_t prep1(int c) const { return p1[(uint64_t)c * ps]; }
    __device__ __forceinline__ kb::Ext perm0(int j) const {
        const uint32_t* p = q0 + (uint64_t)(4 * j) * ms;
        return kb::Ext{{p[0], p[ms], p[2 * ms], p[3 * ms]}};
    }
    __device__ __forceinline__ kb::Ext perm1(int j) const {
        const uint32_t* p = q1 + (uint64_t)(4 * j) * ms;
        return kb::Ext{{p[0], p[ms], p[2 * ms], p[3 * ms]}};
    }
};
template <int CHIP>
__global__ void __launch_bounds__(128, BFGPU_QUOT_MINBLOCKS) k_quotient_shard(QuotientShardArgs A, Challenges ch) {
    const unsigned L = A.log_n + A.lqd;
    const uint64_t N = 1ull << L, n = 1ull << A.log_n, rpg = 1ull << A.log_rpg;
    const uint64_t tl = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (tl >= rpg) return;
    const uint64_t t = ((uint64_t)A.rank << A.log_rpg) + tl;  // global stored row
    const uint32_t i = kb::bitrev((uint32_t)t, L);
    const uint32_t inext = (i + (1u << A.lqd)) & (uint32_t)(N - 1);
    const uint64_t tn = kb::bitrev(inext, L);
    const uint32_t rn = (uint32_t)(tn >> A.log_rpg);
    const uint64_t tnl = tn & (rpg - 1);
    ShardLoader ld;
    ld.ms = rpg;
    ld.ps = N;
    ld.m0 = A.main.shard[A.rank] + tl;
    ld.m1 = A.main.shard[rn] + tnl;
    ld.q0 = A.perm.shard[A.rank] + tl;
    ld.q1 = A.perm.shard[rn] + tnl;
    ld.p0 = A.prep ? A.prep + t : nullptr;
    ld.p1 = A.prep ? A.prep + tn : nullptr;
    const uint32_t x = kb::mul(A.shift, root_pow_(A.tw, L, i));
    const uint32_t zh = A.zh[i & ((1u << A.lqd) - 1)], zh_inv = A.zh_inv[i & ((1u << A.lqd) - 1)];
    Selectors sel;
    const uint32_t d_first = kb::sub(x, kb::ONE), d_last = kb::sub(x, A.g_inv);
    const uint32_t ip = kb::mul(zh, kb::inv(kb::mul(d_first, d_last)));
    sel.is_first = kb::mul(ip, d_last);
    sel.is_last = kb::mul(ip, d_first);
    sel.is_trans = d_last;
    kb::Ext acc = kb::ext_zero();
    air_constraints(CHIP, ld, sel, ch, A.apow, acc);
    acc = kb::ext_scale(acc, zh_inv);
    const uint32_t c = (uint32_t)(t >> A.log_n);
    const uint64_t pos = t & (n - 1);
#pragma unroll
    for (int k = 0; k < 4; k++) A.out_col[c][k][pos] = acc.c[k];
}

}  // namespace air
