// oracle/fast_air_packed.cpp — TEST / BENCH INFRASTRUCTURE (CPU arm), not product code.
//
// The AIR-driven prover phases of oracle/fast_air.cpp, SIXTEEN rows at a time: the generated per-chip programs (csrc/gen_air.cuh) are
// compiled here against the packed AVX-512 field of oracle/packed_kb.h (`kb` -> `pkb`, `uint32_t` -> `pkb::V`), which is how the
// reference evaluates them (Plonky3 packs rows into `PackedMontyField31AVX512`; crates/stark/src/quotient.rs:64-70,
// permutation.rs:75-148).  Same interfaces and bit-identical results as bfo_air_perm_trace / bfo_air_quotient (checked in
// tests/test_oracle_fast_air.py); callers fall back to the scalar arm when the host has no AVX-512.
#include <immintrin.h>
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __restrict__ __restrict
#include "../zkvm-brainfuck_b200/csrc/kb31.cuh"  // scalar field (set-up values), real `kb`
#include "packed_kb.h"

// ---- the generated programs over the packed field -------------------------------------------------------------------------------
#define kb pkb
#define uint32_t pkb::V
namespace air {
struct Selectors {
    uint32_t is_first, is_last, is_trans;
};
struct Challenges {
    kb::Ext alpha;
    kb::Ext beta_pow[8];
    kb::Ext cumulative_sum;
};
}  // namespace air
#define AIR_EXT_MUL kb::ext_mul
#define AIR_EXT_INV kb::ext_inv
#pragma GCC push_options
#pragma GCC target("avx512f,avx512dq,avx512bw,avx512vl")
#include "../zkvm-brainfuck_b200/csrc/gen_air.cuh"
#undef uint32_t
#undef kb

namespace {

using pkb::V;
inline pkb::Ext bcast(const kb::Ext& e) { return pkb::Ext{{V(e.c[0]), V(e.c[1]), V(e.c[2]), V(e.c[3])}}; }
inline kb::Ext ext_to_mont(const uint32_t c[4]) { return kb::Ext{{kb::to_mont(c[0] % kb::P), kb::to_mont(c[1] % kb::P), kb::to_mont(c[2] % kb::P), kb::to_mont(c[3] % kb::P)}}; }
air::Challenges make_challenges(const uint32_t alpha[4], const uint32_t beta[4], const uint32_t* csum) {
    air::Challenges ch;
    ch.alpha = bcast(ext_to_mont(alpha));
    const kb::Ext b = ext_to_mont(beta);
    kb::Ext bp = kb::ext_one();
    for (int k = 0; k < 8; k++) {
        ch.beta_pow[k] = bcast(bp);
        bp = kb::ext_mul(bp, b);
    }
    ch.cumulative_sum = bcast(csum ? ext_to_mont(csum) : kb::ext_zero());
    return ch;
}
struct PTraceRows {
    V m[64], p[8];
    V main0(int c) const { return m[c]; }
    V prep0(int c) const { return p[c]; }
    V main1(int) const { return V(0u); }
    V prep1(int) const { return V(0u); }
};
struct PLdeRows {
    V m0[64], m1[64], p0[8], p1[8], q0[40], q1[40];
    V main0(int c) const { return m0[c]; }
    V main1(int c) const { return m1[c]; }
    V prep0(int c) const { return p0[c]; }
    V prep1(int c) const { return p1[c]; }
    pkb::Ext perm0(int j) const { return pkb::Ext{{q0[4 * j], q0[4 * j + 1], q0[4 * j + 2], q0[4 * j + 3]}}; }
    pkb::Ext perm1(int j) const { return pkb::Ext{{q1[4 * j], q1[4 * j + 1], q1[4 * j + 2], q1[4 * j + 3]}}; }
};
// dst[c] = Montgomery form of column c of the 16 rows whose first words sit at base + idx[k]
inline void gather_mont(V* dst, const uint32_t* base, __m512i idx, int w) {
    for (int c = 0; c < w; c++) dst[c] = pkb::to_mont(V(_mm512_i32gather_epi32(idx, base + c, 4)));
}
inline uint64_t brev(uint64_t x, unsigned bits) {
    uint64_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1ull) << (bits - 1 - i);
    return r;
}
inline unsigned ilog2(uint64_t x) {
    unsigned l = 0;
    while ((1ull << l) < x) l++;
    return l;
}
inline void store16(uint32_t* tmp, V v) { _mm512_storeu_si512((void*)tmp, v.v); }

}  // namespace

extern "C" {

int bfo_air_packed_available(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl");
}

// same contract as bfo_air_perm_trace (fast_air.cpp); rows must be a multiple of 16 and rows * main_w below 2^31
int bfo_air_perm_trace_packed(int chip, const uint32_t* main, const uint32_t* prep, uint64_t rows, const uint32_t alpha[4], const uint32_t beta[4],
                              uint32_t* perm_out, uint32_t csum[4]) {
    if (!bfo_air_packed_available() || chip < 0 || chip >= air::NUM_CHIPS || !main || !perm_out || rows < 16 || rows % 16) return -1;
    const int mw = air::CHIPS[chip].main_w, pw = air::CHIPS[chip].prep_w, ew = air::CHIPS[chip].perm_w;
    if (mw > 64 || pw > 8 || ew > air::MAX_PERM_W || (pw && !prep) || rows * (uint64_t)std::max(mw, 4 * ew) >= (1ull << 31)) return -1;
    const air::Challenges ch = make_challenges(alpha, beta, nullptr);
    std::vector<kb::Ext> rowsum(rows);
    alignas(64) int lane[16];
    for (int k = 0; k < 16; k++) lane[k] = k;
    const __m512i lanes = _mm512_load_si512((const void*)lane);
    const __m512i idx_m = _mm512_mullo_epi32(lanes, _mm512_set1_epi32(mw)), idx_p = _mm512_mullo_epi32(lanes, _mm512_set1_epi32(pw ? pw : 1));
#pragma omp parallel for schedule(static)
    for (uint64_t r0 = 0; r0 < rows; r0 += 16) {
        PTraceRows ld;
        gather_mont(ld.m, main + r0 * (uint64_t)mw, idx_m, mw);
        if (pw) gather_mont(ld.p, prep + r0 * (uint64_t)pw, idx_p, pw);
        pkb::Ext out[air::MAX_PERM_W];
        air::air_perm_row(chip, ld, ch, out);
        pkb::Ext s = pkb::ext_zero();
        alignas(64) uint32_t tmp[16];
        for (int j = 0; j < ew - 1; j++) {
            s = pkb::ext_add(s, out[j]);
            for (int e = 0; e < 4; e++) {
                store16(tmp, pkb::from_mont(out[j].c[e]));
                for (int k = 0; k < 16; k++) perm_out[(r0 + (uint64_t)k) * (uint64_t)(4 * ew) + 4 * j + e] = tmp[k];
            }
        }
        for (int e = 0; e < 4; e++) {
            store16(tmp, s.c[e]);
            for (int k = 0; k < 16; k++) rowsum[r0 + (uint64_t)k].c[e] = tmp[k];
        }
    }
    kb::Ext run = kb::ext_zero();  // inclusive running sum (permutation.rs:131-146)
    for (uint64_t r = 0; r < rows; r++) {
        run = kb::ext_add(run, rowsum[r]);
        uint32_t* o = perm_out + r * (uint64_t)(4 * ew) + 4 * (ew - 1);
        for (int e = 0; e < 4; e++) o[e] = kb::from_mont(run.c[e]);
    }
    for (int e = 0; e < 4; e++) csum[e] = kb::from_mont(run.c[e]);
    return 0;
}

// same contract as bfo_air_quotient (fast_air.cpp); n >= 8 (16 quotient-domain points per step)
int bfo_air_quotient_packed(int chip, const uint32_t* main_lde, const uint32_t* prep_lde, const uint32_t* perm_lde, uint64_t n, const uint32_t alpha_logup[4],
                            const uint32_t beta[4], const uint32_t csum[4], const uint32_t alpha[4], uint32_t* q_out) {
    if (!bfo_air_packed_available() || chip < 0 || chip >= air::NUM_CHIPS || !main_lde || !perm_lde || !q_out || n < 8 || (n & (n - 1))) return -1;
    const int mw = air::CHIPS[chip].main_w, pw = air::CHIPS[chip].prep_w, ew = air::CHIPS[chip].perm_w;
    const uint64_t N = 2 * n;
    if (mw > 64 || pw > 8 || 4 * ew > 40 || (pw && !prep_lde) || N * (uint64_t)std::max(mw, 4 * ew) >= (1ull << 31)) return -1;
    const unsigned log_n = ilog2(n), L = log_n + 1;
    const air::Challenges ch = make_challenges(alpha_logup, beta, csum);
    alignas(64) pkb::Ext apow[air::MAX_CONSTRAINTS];  // (a std::vector of this 64-byte-aligned type faulted: its storage came back 16-byte aligned)
    {
        kb::Ext a = ext_to_mont(alpha), p = kb::ext_one();
        for (int k = 0; k < air::MAX_CONSTRAINTS; k++) {
            apow[k] = bcast(p);
            p = kb::ext_mul(p, a);
        }
    }
    const uint32_t shift = kb::to_mont(kb::GEN), wN = kb::two_adic_generator(L), g_inv = kb::inv(kb::two_adic_generator(log_n));
    const uint32_t sn = kb::pow(shift, n);
    const uint32_t zh[2] = {kb::sub(sn, kb::ONE), kb::sub(kb::neg(sn), kb::ONE)};
    const uint32_t zh_inv[2] = {kb::inv(zh[0]), kb::inv(zh[1])};
    alignas(64) uint32_t w16[16], zhv[16], zhiv[16];
    for (int k = 0; k < 16; k++) {
        w16[k] = kb::pow(wN, (uint64_t)k);
        zhv[k] = zh[k & 1];
        zhiv[k] = zh_inv[k & 1];
    }
    const V W16(_mm512_load_si512((const void*)w16)), ZH(_mm512_load_si512((const void*)zhv)), ZHI(_mm512_load_si512((const void*)zhiv));
#pragma omp parallel for schedule(static)
    for (uint64_t i0 = 0; i0 < N; i0 += 16) {
        alignas(64) int it[16], itn[16];
        for (int k = 0; k < 16; k++) {
            it[k] = (int)brev(i0 + (uint64_t)k, L);
            itn[k] = (int)brev((i0 + (uint64_t)k + 2) & (N - 1), L);
        }
        const __m512i t = _mm512_load_si512((const void*)it), tn = _mm512_load_si512((const void*)itn);
        PLdeRows ld;
        gather_mont(ld.m0, main_lde, _mm512_mullo_epi32(t, _mm512_set1_epi32(mw)), mw);
        gather_mont(ld.m1, main_lde, _mm512_mullo_epi32(tn, _mm512_set1_epi32(mw)), mw);
        if (pw) {
            gather_mont(ld.p0, prep_lde, _mm512_mullo_epi32(t, _mm512_set1_epi32(pw)), pw);
            gather_mont(ld.p1, prep_lde, _mm512_mullo_epi32(tn, _mm512_set1_epi32(pw)), pw);
        }
        gather_mont(ld.q0, perm_lde, _mm512_mullo_epi32(t, _mm512_set1_epi32(4 * ew)), 4 * ew);
        gather_mont(ld.q1, perm_lde, _mm512_mullo_epi32(tn, _mm512_set1_epi32(4 * ew)), 4 * ew);
        const V x = pkb::mul(V(kb::mul(shift, kb::pow(wN, i0))), W16);  // g w^(i0 + k)
        air::Selectors sel;
        const V d_first = pkb::sub(x, V(kb::ONE)), d_last = pkb::sub(x, V(g_inv));
        const V ip = pkb::mul(ZH, pkb::inv(pkb::mul(d_first, d_last)));
        sel.is_first = pkb::mul(ip, d_last);
        sel.is_last = pkb::mul(ip, d_first);
        sel.is_trans = d_last;
        pkb::Ext acc = pkb::ext_zero();
        air::air_constraints(chip, ld, sel, ch, apow, acc);
        acc = pkb::ext_scale(acc, ZHI);
        alignas(64) uint32_t tmp[16];
        for (int e = 0; e < 4; e++) {
            store16(tmp, pkb::from_mont(acc.c[e]));
            for (int k = 0; k < 16; k++) {
                const uint64_t i = i0 + (uint64_t)k;
                q_out[((i & 1) * n + (i >> 1)) * 4 + (uint64_t)e] = tmp[k];
            }
        }
    }
    return 0;
}

}  // extern "C"
// ---- openings, sixteen rows per step (same contracts as bfo_open_eval / bfo_open_reduce_add in fast_air.cpp) ---------------------------
namespace {
// g^{bitrev(r0 + k, log_n)} for k < 16, r0 a multiple of 16, log_n >= 4:  g^{bitrev(r0 >> 4, log_n - 4)} * (g^(2^(log_n-4)))^{bitrev4(k)}
struct BrevPowers16 {
    uint32_t g;
    unsigned log_n;
    V t16;
    BrevPowers16(uint32_t gen, unsigned ln) : g(gen), log_n(ln) {
        alignas(64) uint32_t t[16];
        const uint32_t G = kb::pow(g, 1ull << (log_n - 4));
        for (uint64_t k = 0; k < 16; k++) t[k] = kb::pow(G, brev(k, 4));
        t16 = V(_mm512_load_si512((const void*)t));
    }
    V block(uint64_t r0) const { return pkb::mul(V(kb::pow(g, brev(r0 >> 4, log_n - 4))), t16); }
};
inline kb::Ext hsum(const pkb::Ext& e) {  // sum of the 16 lanes of every coefficient
    kb::Ext r = kb::ext_zero();
    alignas(64) uint32_t tmp[16];
    for (int c = 0; c < 4; c++) {
        store16(tmp, e.c[c]);
        uint32_t s = 0;
        for (int k = 0; k < 16; k++) s = kb::add(s, tmp[k]);
        r.c[c] = s;
    }
    return r;
}
}  // namespace

extern "C" {

int bfo_open_eval_packed(const uint32_t* lde, uint64_t n, uint32_t w, const uint32_t z[4], uint32_t* out) {
    if (!bfo_air_packed_available() || !lde || !out || n < 16 || (n & (n - 1)) || w == 0 || w > 256 || 2 * n * (uint64_t)w >= (1ull << 31)) return -1;
    const unsigned log_n = ilog2(n);
    const kb::Ext zs = ext_to_mont(z);
    const pkb::Ext zz = bcast(zs);
    const uint32_t shift = kb::to_mont(kb::GEN), g = kb::two_adic_generator(log_n);
    const BrevPowers16 bp(g, log_n);
    const int nt_max = omp_get_max_threads();
    std::vector<kb::Ext> partial((size_t)nt_max * w, kb::ext_zero());
    alignas(64) int lane[16];
    for (int k = 0; k < 16; k++) lane[k] = k * (int)w;
    const __m512i idx = _mm512_load_si512((const void*)lane);
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), id = omp_get_thread_num();
        pkb::Ext* acc = (pkb::Ext*)aligned_alloc(64, (size_t)w * sizeof(pkb::Ext));
        for (uint32_t c = 0; c < w; c++) acc[c] = pkb::ext_zero();
        const uint64_t nb = n / 16;
        for (uint64_t blk = nb * (uint64_t)id / (uint64_t)nt; blk < nb * (uint64_t)(id + 1) / (uint64_t)nt; blk++) {
            const uint64_t r0 = blk * 16;
            const V gp = bp.block(r0);
            pkb::Ext d = zz;
            d.c[0] = pkb::sub(d.c[0], pkb::mul(V(shift), gp));
            const pkb::Ext wgt = pkb::ext_scale(pkb::ext_inv(d), gp);
            const uint32_t* base = lde + r0 * (uint64_t)w;
            for (uint32_t c = 0; c < w; c++) {
                const V m = pkb::to_mont(V(_mm512_i32gather_epi32(idx, base + c, 4)));
                pkb::ext_mac(acc[c], wgt, m);
            }
        }
        for (uint32_t c = 0; c < w; c++) partial[(size_t)id * w + c] = hsum(acc[c]);
        free(acc);
    }
    kb::Ext zn = zs;
    for (unsigned k = 0; k < log_n; k++) zn = kb::ext_sqr(zn);
    zn.c[0] = kb::sub(zn.c[0], kb::pow(shift, n));
    const uint32_t denom = kb::mul(kb::pow(shift, n - 1), kb::to_mont((uint32_t)(n % kb::P)));
    const kb::Ext scale = kb::ext_scale(zn, kb::inv(denom));
    for (uint32_t c = 0; c < w; c++) {
        kb::Ext e = kb::ext_zero();
        for (int t = 0; t < nt_max; t++) e = kb::ext_add(e, partial[(size_t)t * w + c]);
        e = kb::ext_mul(e, scale);
        for (int k = 0; k < 4; k++) out[4 * c + k] = kb::from_mont(e.c[k]);
    }
    return 0;
}

int bfo_open_reduce_add_packed(const uint32_t* lde, uint64_t h, uint32_t w, uint32_t npts, const uint32_t* zs, const uint32_t* ys, const uint32_t alpha[4],
                               uint64_t reduced_before, uint32_t* ro) {
    if (!bfo_air_packed_available() || !lde || !ro || !zs || !ys || h < 16 || (h & (h - 1)) || w == 0 || npts == 0 || npts > 2 || h * (uint64_t)w >= (1ull << 31))
        return -1;
    const unsigned log_h = ilog2(h);
    const kb::Ext a = ext_to_mont(alpha);
    std::vector<kb::Ext> apow(w);
    apow[0] = kb::ext_one();
    for (uint32_t k = 1; k < w; k++) apow[k] = kb::ext_mul(apow[k - 1], a);
    const kb::Ext aw = kb::ext_mul(apow[w - 1], a);
    kb::Ext o = kb::ext_one();
    {
        kb::Ext base = a;
        for (uint64_t e = reduced_before; e; e >>= 1) {
            if (e & 1) o = kb::ext_mul(o, base);
            base = kb::ext_sqr(base);
        }
    }
    alignas(64) pkb::Ext off[2], yred[2], zp[2];
    for (uint32_t t = 0; t < npts; t++) {
        off[t] = bcast(o);
        o = kb::ext_mul(o, aw);
        zp[t] = bcast(ext_to_mont(zs + 4 * t));
        kb::Ext y = kb::ext_zero();
        for (uint32_t k = 0; k < w; k++) y = kb::ext_add(y, kb::ext_mul(apow[k], ext_to_mont(ys + ((uint64_t)t * w + k) * 4)));
        yred[t] = bcast(y);
    }
    pkb::Ext* pa = (pkb::Ext*)aligned_alloc(64, (size_t)w * sizeof(pkb::Ext));  // alpha^k in every lane
    for (uint32_t k = 0; k < w; k++) pa[k] = bcast(apow[k]);
    const uint32_t shift = kb::to_mont(kb::GEN), g = kb::two_adic_generator(log_h);
    const BrevPowers16 bp(g, log_h);
    alignas(64) int lane[16], lane4[16];
    for (int k = 0; k < 16; k++) {
        lane[k] = k * (int)w;
        lane4[k] = k * 4;
    }
    const __m512i idx = _mm512_load_si512((const void*)lane), idx4 = _mm512_load_si512((const void*)lane4);
#pragma omp parallel for schedule(static)
    for (uint64_t r0 = 0; r0 < h; r0 += 16) {
        const V x = pkb::mul(V(shift), bp.block(r0));
        pkb::Ext rr = pkb::ext_zero();
        const uint32_t* base = lde + r0 * (uint64_t)w;
        for (uint32_t k = 0; k < w; k++) pkb::ext_mac(rr, pa[k], pkb::to_mont(V(_mm512_i32gather_epi32(idx, base + k, 4))));
        pkb::Ext d[2], inv[2];
        for (uint32_t t = 0; t < npts; t++) {
            d[t] = zp[t];
            d[t].c[0] = pkb::sub(d[t].c[0], x);
        }
        if (npts == 2) {  // one inversion for both points
            const pkb::Ext ip = pkb::ext_inv(pkb::ext_mul(d[0], d[1]));
            inv[0] = pkb::ext_mul(ip, d[1]);
            inv[1] = pkb::ext_mul(ip, d[0]);
        } else {
            inv[0] = pkb::ext_inv(d[0]);
        }
        uint32_t* rp = ro + r0 * 4;
        pkb::Ext sum;
        for (int e = 0; e < 4; e++) sum.c[e] = pkb::to_mont(V(_mm512_i32gather_epi32(idx4, rp + e, 4)));
        for (uint32_t t = 0; t < npts; t++) sum = pkb::ext_add(sum, pkb::ext_mul(off[t], pkb::ext_mul(pkb::ext_sub(yred[t], rr), inv[t])));
        alignas(64) uint32_t tmp[16];
        for (int e = 0; e < 4; e++) {
            store16(tmp, pkb::from_mont(sum.c[e]));
            for (int k = 0; k < 16; k++) rp[4 * k + e] = tmp[k];
        }
    }
    free(pa);
    return 0;
}

}  // extern "C"
#pragma GCC pop_options
