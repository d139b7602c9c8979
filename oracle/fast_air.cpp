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

}  // extern "C"
