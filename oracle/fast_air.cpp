// oracle/fast_air.cpp — TEST / BENCH INFRASTRUCTURE (CPU arm), not product code.
//
// CPU implementation of the two AIR-driven phases of the reference's shard prover, used by bench.py's per-phase CPU prove baseline
// and checked against the numpy oracle (tests/test_oracle_fast_air.py):
//   bfo_air_perm_trace : generate_permutation_trace          reference crates/stark/src/permutation.rs:75-148   (span prover.rs:281)
//   bfo_air_quotient   : quotient_values + ProverConstraintFolder   crates/stark/src/quotient.rs:18-165, folder.rs:68-89 (span prover.rs:355)
// One row at a time, rows spread over the host threads with OpenMP (what the reference's rayon `par_chunks` does), scalar
// Montgomery arithmetic (kb31.cuh compiled for the host).  The per-chip row programs are the ones air/codegen.py generates from
// air/chips.py (csrc/gen_air.cuh, compiled here for the host): this file is a TIMING arm, not an independent check of the AIR
// description — that is what tests/ref_air.py is for.
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __restrict__ __restrict
#include "../zkvm-brainfuck_b200/csrc/kb31.cuh"

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
#include "../zkvm-brainfuck_b200/csrc/gen_air.cuh"

namespace {

inline kb::Ext ext_to_mont(const uint32_t c[4]) { return kb::Ext{{kb::to_mont(c[0] % kb::P), kb::to_mont(c[1] % kb::P), kb::to_mont(c[2] % kb::P), kb::to_mont(c[3] % kb::P)}}; }
air::Challenges make_challenges(const uint32_t alpha[4], const uint32_t beta[4], const uint32_t* csum) {
    air::Challenges ch;
    ch.alpha = ext_to_mont(alpha);
    const kb::Ext b = ext_to_mont(beta);
    ch.beta_pow[0] = kb::ext_one();
    for (int k = 1; k < 8; k++) ch.beta_pow[k] = kb::ext_mul(ch.beta_pow[k - 1], b);
    ch.cumulative_sum = csum ? ext_to_mont(csum) : kb::ext_zero();
    return ch;
}
// one trace row, converted to Montgomery form on the fly (the reference's matrices hold Montgomery words already: this costs the port
// one multiplication per cell on top)
struct TraceRow {
    uint32_t m[64], p[8];
    uint32_t main0(int c) const { return m[c]; }
    uint32_t prep0(int c) const { return p[c]; }
    uint32_t main1(int) const { return 0; }
    uint32_t prep1(int) const { return 0; }
};
struct LdeRows {
    uint32_t m0[64], m1[64], p0[8], p1[8], q0[40], q1[40];
    uint32_t main0(int c) const { return m0[c]; }
    uint32_t main1(int c) const { return m1[c]; }
    uint32_t prep0(int c) const { return p0[c]; }
    uint32_t prep1(int c) const { return p1[c]; }
    kb::Ext perm0(int j) const { return kb::Ext{{q0[4 * j], q0[4 * j + 1], q0[4 * j + 2], q0[4 * j + 3]}}; }
    kb::Ext perm1(int j) const { return kb::Ext{{q1[4 * j], q1[4 * j + 1], q1[4 * j + 2], q1[4 * j + 3]}}; }
};
inline void load_mont(uint32_t* dst, const uint32_t* src, int w) {
    for (int c = 0; c < w; c++) dst[c] = kb::to_mont(src[c]);
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

// g^{bitrev(r, log_n)} for 64 consecutive r at a time: bitrev(r0 + k) = bitrev6(k) << (log_n - 6) | bitrev(r0 >> 6), so the block needs one
// power for its base and a fixed table of the 64th roots (instead of a ~1.5 log n product chain per row)
struct BrevPowers {
    uint32_t g, t64[64];
    unsigned log_n;
    BrevPowers(uint32_t gen, unsigned ln) : g(gen), log_n(ln) {
        if (log_n >= 6) {
            const uint32_t G = kb::pow(g, 1ull << (log_n - 6));
            for (uint64_t k = 0; k < 64; k++) t64[k] = kb::pow(G, brev(k, 6));
        }
    }
    // out[k] = g^{bitrev(r0 + k)}, k < cnt <= 64, r0 a multiple of 64 when log_n >= 6
    void block(uint64_t r0, uint64_t cnt, uint32_t* out) const {
        if (log_n < 6) {
            for (uint64_t k = 0; k < cnt; k++) out[k] = kb::pow(g, brev(r0 + k, log_n));
            return;
        }
        const uint32_t base = kb::pow(g, brev(r0 >> 6, log_n - 6));
        for (uint64_t k = 0; k < cnt; k++) out[k] = kb::mul(base, t64[k]);
    }
};

}  // namespace

extern "C" {

int bfo_air_chip_info(int chip, int info[4]) {
    if (chip < 0 || chip >= air::NUM_CHIPS) return -1;
    info[0] = air::CHIPS[chip].main_w;
    info[1] = air::CHIPS[chip].prep_w;
    info[2] = air::CHIPS[chip].perm_w;
    info[3] = air::CHIPS[chip].n_constraints;
    return 0;
}

// main: rows x main_w, prep: rows x prep_w (null when the chip has none), canonical row-major.  alpha, beta: canonical extension
// elements (the LogUp challenges).  perm_out: rows x 4*perm_w canonical (extension column j = base columns 4j .. 4j+3, last
// extension column = running sum); csum: its last entry.
int bfo_air_perm_trace(int chip, const uint32_t* main, const uint32_t* prep, uint64_t rows, const uint32_t alpha[4], const uint32_t beta[4], uint32_t* perm_out,
                       uint32_t csum[4]) {
    if (chip < 0 || chip >= air::NUM_CHIPS || !main || !perm_out || !rows) return -1;
    const int mw = air::CHIPS[chip].main_w, pw = air::CHIPS[chip].prep_w, ew = air::CHIPS[chip].perm_w;
    if (mw > 64 || pw > 8 || ew > air::MAX_PERM_W || (pw && !prep)) return -1;
    const air::Challenges ch = make_challenges(alpha, beta, nullptr);
    std::vector<kb::Ext> rowsum(rows);
#pragma omp parallel for schedule(static)
    for (uint64_t r = 0; r < rows; r++) {
        TraceRow ld;
        load_mont(ld.m, main + r * (uint64_t)mw, mw);
        if (pw) load_mont(ld.p, prep + r * (uint64_t)pw, pw);
        kb::Ext out[air::MAX_PERM_W];
        air::air_perm_row(chip, ld, ch, out);
        kb::Ext s = kb::ext_zero();
        uint32_t* o = perm_out + r * (uint64_t)(4 * ew);
        for (int j = 0; j < ew - 1; j++) {
            s = kb::ext_add(s, out[j]);
            for (int e = 0; e < 4; e++) o[4 * j + e] = kb::from_mont(out[j].c[e]);
        }
        rowsum[r] = s;
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

// LDEs on the quotient domain g * H_{2n} (log_quotient_degree 1): 2n x width canonical row-major with BIT-REVERSED rows (how
// TwoAdicFriPcs stores them and how bfo_fast_pcs_commit returns them).  q_out: 2 chunks x n x 4 canonical; chunk c holds the quotient
// values at the natural points 2k + c (split_evals), row k.
int bfo_air_quotient(int chip, const uint32_t* main_lde, const uint32_t* prep_lde, const uint32_t* perm_lde, uint64_t n, const uint32_t alpha_logup[4],
                     const uint32_t beta[4], const uint32_t csum[4], const uint32_t alpha[4], uint32_t* q_out) {
    if (chip < 0 || chip >= air::NUM_CHIPS || !main_lde || !perm_lde || !q_out || n < 2 || (n & (n - 1))) return -1;
    const int mw = air::CHIPS[chip].main_w, pw = air::CHIPS[chip].prep_w, ew = air::CHIPS[chip].perm_w, nc = air::CHIPS[chip].n_constraints;
    if (mw > 64 || pw > 8 || 4 * ew > 40 || (pw && !prep_lde)) return -1;
    const unsigned log_n = ilog2(n), L = log_n + 1;
    const uint64_t N = 2 * n;
    const air::Challenges ch = make_challenges(alpha_logup, beta, csum);
    std::vector<kb::Ext> apow(air::MAX_CONSTRAINTS);
    apow[0] = kb::ext_one();
    const kb::Ext a = ext_to_mont(alpha);
    for (int k = 1; k < air::MAX_CONSTRAINTS; k++) apow[k] = kb::ext_mul(apow[k - 1], a);
    (void)nc;
    const uint32_t shift = kb::to_mont(kb::GEN), wN = kb::two_adic_generator(L), g_inv = kb::inv(kb::two_adic_generator(log_n));
    const uint32_t sn = kb::pow(shift, n);
    const uint32_t zh[2] = {kb::sub(sn, kb::ONE), kb::sub(kb::neg(sn), kb::ONE)};  // Z_H(g w^i) = g^n (w^n)^i - 1, w^n = -1
    const uint32_t zh_inv[2] = {kb::inv(zh[0]), kb::inv(zh[1])};
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), id = omp_get_thread_num();
        const uint64_t i0 = N * (uint64_t)id / (uint64_t)nt, i1 = N * (uint64_t)(id + 1) / (uint64_t)nt;
        uint32_t x = kb::mul(shift, kb::pow(wN, i0));
        LdeRows ld;
        for (uint64_t i = i0; i < i1; i++, x = kb::mul(x, wN)) {
            const uint64_t t = brev(i, L), tn = brev((i + 2) & (N - 1), L);
            load_mont(ld.m0, main_lde + t * (uint64_t)mw, mw);
            load_mont(ld.m1, main_lde + tn * (uint64_t)mw, mw);
            if (pw) {
                load_mont(ld.p0, prep_lde + t * (uint64_t)pw, pw);
                load_mont(ld.p1, prep_lde + tn * (uint64_t)pw, pw);
            }
            load_mont(ld.q0, perm_lde + t * (uint64_t)(4 * ew), 4 * ew);
            load_mont(ld.q1, perm_lde + tn * (uint64_t)(4 * ew), 4 * ew);
            // selectors_on_coset: 1/(x - 1) and 1/(x - g^-1) with one inversion
            air::Selectors sel;
            const uint32_t d_first = kb::sub(x, kb::ONE), d_last = kb::sub(x, g_inv);
            const uint32_t ip = kb::mul(zh[i & 1], kb::inv(kb::mul(d_first, d_last)));
            sel.is_first = kb::mul(ip, d_last);
            sel.is_last = kb::mul(ip, d_first);
            sel.is_trans = d_last;
            kb::Ext acc = kb::ext_zero();
            air::air_constraints(chip, ld, sel, ch, apow.data(), acc);
            acc = kb::ext_scale(acc, zh_inv[i & 1]);
            uint32_t* o = q_out + ((i & 1) * n + (i >> 1)) * 4;
            for (int e = 0; e < 4; e++) o[e] = kb::from_mont(acc.c[e]);
        }
    }
    return 0;
}

// ---- openings (first half of `pcs.open`, reference prover.rs:460 -> TwoAdicFriPcs::open) -------------------------------------------------
// lde: 2n x w canonical, rows bit-reversed; its first n stored rows are the evaluations on the coset GEN * H_n (bit-reversed).
// out[c] = p_c(z) (w x 4 canonical) by the barycentric formula  p(z) = (z^n - s^n) / (n s^(n-1)) * sum_r p(x_r) g^r / (z - x_r).
int bfo_open_eval(const uint32_t* lde, uint64_t n, uint32_t w, const uint32_t z[4], uint32_t* out) {
    if (!lde || !out || n < 1 || (n & (n - 1)) || w == 0) return -1;
    const unsigned log_n = ilog2(n);
    const kb::Ext zz = ext_to_mont(z);
    const uint32_t shift = kb::to_mont(kb::GEN), g = kb::two_adic_generator(log_n);
    int nt_max = omp_get_max_threads();
    std::vector<uint64_t> partial((size_t)nt_max * w * 4, 0);
    constexpr int B = 64;  // rows per batch inversion
    const BrevPowers bp(g, log_n);
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), id = omp_get_thread_num();
        std::vector<kb::ExtAcc> acc(w, kb::ext_acc_zero());
        const uint64_t nb = (n + B - 1) / B;
        for (uint64_t blk = (uint64_t)id; blk < nb; blk += (uint64_t)nt) {
            const uint64_t r0 = blk * B, cnt = std::min<uint64_t>(B, n - r0);
            kb::Ext d[B], pre[B];
            uint32_t gp[B];
            bp.block(r0, cnt, gp);  // g^{br(r)}
            for (uint64_t k = 0; k < cnt; k++) {
                d[k] = zz;
                d[k].c[0] = kb::sub(d[k].c[0], kb::mul(shift, gp[k]));
                pre[k] = k ? kb::ext_mul(pre[k - 1], d[k]) : d[k];
            }
            kb::Ext run = kb::ext_inv(pre[cnt - 1]);
            for (uint64_t k = cnt; k-- > 0;) {
                const kb::Ext inv = k ? kb::ext_mul(run, pre[k - 1]) : run;
                run = kb::ext_mul(run, d[k]);
                const kb::Ext wgt = kb::ext_scale(inv, gp[k]);
                const uint32_t* row = lde + (r0 + k) * (uint64_t)w;
                for (uint32_t c = 0; c < w; c++) kb::ext_mac(acc[c], wgt, kb::to_mont(row[c]));
            }
        }
        for (uint32_t c = 0; c < w; c++) {
            const kb::Ext e = kb::ext_acc_reduce(acc[c]);
            for (int k = 0; k < 4; k++) partial[((size_t)id * w + c) * 4 + k] = e.c[k];
        }
    }
    // (z^n - s^n) / (n s^(n-1))
    kb::Ext zn = zz;
    for (unsigned k = 0; k < log_n; k++) zn = kb::ext_sqr(zn);
    zn.c[0] = kb::sub(zn.c[0], kb::pow(shift, n));
    const uint32_t denom = kb::mul(kb::pow(shift, n - 1), kb::to_mont((uint32_t)(n % kb::P)));
    const kb::Ext scale = kb::ext_scale(zn, kb::inv(denom));
    for (uint32_t c = 0; c < w; c++) {
        kb::Ext e = kb::ext_zero();
        for (int t = 0; t < nt_max; t++) {
            kb::Ext pt{{(uint32_t)partial[((size_t)t * w + c) * 4], (uint32_t)partial[((size_t)t * w + c) * 4 + 1], (uint32_t)partial[((size_t)t * w + c) * 4 + 2],
                        (uint32_t)partial[((size_t)t * w + c) * 4 + 3]}};
            e = kb::ext_add(e, pt);
        }
        e = kb::ext_mul(e, scale);
        for (int k = 0; k < 4; k++) out[4 * c + k] = kb::from_mont(e.c[k]);
    }
    return 0;
}

// Reduced openings of ONE matrix added into the vector of its height (TwoAdicFriPcs::open, "reduced openings by height"):
//   ro[r] += sum_t alpha^(off) * (y_red_t - sum_k alpha^k M[r][k]) / (z_t - x_r),   x_r = GEN w_h^{br(r)},  y_red_t = sum_k alpha^k ys_t[k],
// off = number of columns already reduced at this height (+ w per point).  lde: h x w canonical (rows bit-reversed); zs: npts x 4; ys: npts x w x 4;
// ro: h x 4 canonical, updated in place.
int bfo_open_reduce_add(const uint32_t* lde, uint64_t h, uint32_t w, uint32_t npts, const uint32_t* zs, const uint32_t* ys, const uint32_t alpha[4],
                        uint64_t reduced_before, uint32_t* ro) {
    if (!lde || !ro || !zs || !ys || h < 2 || (h & (h - 1)) || w == 0 || npts == 0 || npts > 2) return -1;
    const unsigned log_h = ilog2(h);
    const kb::Ext a = ext_to_mont(alpha);
    std::vector<kb::Ext> apow(w);
    apow[0] = kb::ext_one();
    for (uint32_t k = 1; k < w; k++) apow[k] = kb::ext_mul(apow[k - 1], a);
    kb::Ext aw = kb::ext_mul(apow[w - 1], a);  // alpha^w
    kb::Ext off[2], yred[2], zp[2];
    kb::Ext o = kb::ext_one();
    {  // alpha^reduced_before by square and multiply
        kb::Ext base = a;
        uint64_t e = reduced_before;
        while (e) {
            if (e & 1) o = kb::ext_mul(o, base);
            base = kb::ext_sqr(base);
            e >>= 1;
        }
    }
    for (uint32_t t = 0; t < npts; t++) {
        off[t] = o;
        o = kb::ext_mul(o, aw);
        zp[t] = ext_to_mont(zs + 4 * t);
        kb::Ext y = kb::ext_zero();
        for (uint32_t k = 0; k < w; k++) y = kb::ext_add(y, kb::ext_mul(apow[k], ext_to_mont(ys + ((uint64_t)t * w + k) * 4)));
        yred[t] = y;
    }
    const uint32_t shift = kb::to_mont(kb::GEN), g = kb::two_adic_generator(log_h);
    const BrevPowers bp(g, log_h);
#pragma omp parallel for schedule(static)
    for (uint64_t r0 = 0; r0 < h; r0 += 64) {
      uint32_t gp[64];
      const uint64_t cnt = std::min<uint64_t>(64, h - r0);
      bp.block(r0, cnt, gp);
      for (uint64_t r = r0; r < r0 + cnt; r++) {
        const uint32_t x = kb::mul(shift, gp[r - r0]);
        kb::ExtAcc acc = kb::ext_acc_zero();
        const uint32_t* row = lde + r * (uint64_t)w;
        for (uint32_t k = 0; k < w; k++) kb::ext_mac(acc, apow[k], kb::to_mont(row[k]));
        const kb::Ext rr = kb::ext_acc_reduce(acc);
        kb::Ext d[2], inv[2];
        for (uint32_t t = 0; t < npts; t++) {
            d[t] = zp[t];
            d[t].c[0] = kb::sub(d[t].c[0], x);
        }
        if (npts == 2) {  // one inversion for both points
            const kb::Ext ip = kb::ext_inv(kb::ext_mul(d[0], d[1]));
            inv[0] = kb::ext_mul(ip, d[1]);
            inv[1] = kb::ext_mul(ip, d[0]);
        } else {
            inv[0] = kb::ext_inv(d[0]);
        }
        kb::Ext sum{{kb::to_mont(ro[4 * r]), kb::to_mont(ro[4 * r + 1]), kb::to_mont(ro[4 * r + 2]), kb::to_mont(ro[4 * r + 3])}};
        for (uint32_t t = 0; t < npts; t++) sum = kb::ext_add(sum, kb::ext_mul(off[t], kb::ext_mul(kb::ext_sub(yred[t], rr), inv[t])));
        for (int k = 0; k < 4; k++) ro[4 * r + k] = kb::from_mont(sum.c[k]);
      }
    }
    return 0;
}

}  // extern "C"
