// challenger.h — host-side Poseidon2 and DuplexChallenger<Val, Perm, 16, 8>.
//
// The Fiat–Shamir transcript is strictly sequential and tiny (a few hundred permutations per proof),
// so it stays on the host and serialises the GPU phases; only 32-byte roots and a few field elements
// cross PCIe.  Mirrors the reference's `Challenger = DuplexChallenger<Val, Perm, 16, 8>`
// (reference crates/stark/src/kb31_poseidon2.rs:31, created at :126-128; transcript order
// crates/stark/src/prover.rs:266-272,337-340,354,412-415,595-601).  Semantics restated from Plonky3
// p3-challenger v0.1.0.  All words are Montgomery residues.
#pragma once
#include <cstdint>
#include <vector>

#include "kb31.cuh"
#include "rc_16_30.h"

namespace host_p2 {

inline uint32_t sbox(uint32_t x) { return kb::mul(kb::mul(x, x), x); }

inline void mat4(uint32_t* x) {
    uint32_t a = x[0], b = x[1], c = x[2], d = x[3];
    uint32_t s = kb::add(kb::add(a, b), kb::add(c, d));
    x[0] = kb::add(kb::add(s, a), kb::dbl(b));
    x[1] = kb::add(kb::add(s, b), kb::dbl(c));
    x[2] = kb::add(kb::add(s, c), kb::dbl(d));
    x[3] = kb::add(kb::add(s, d), kb::dbl(a));
}
inline void external_linear(uint32_t* s) {
    for (int k = 0; k < 4; k++) mat4(s + 4 * k);
    for (int i = 0; i < 4; i++) {
        uint32_t t = kb::add(kb::add(s[i], s[4 + i]), kb::add(s[8 + i], s[12 + i]));
        for (int k = 0; k < 4; k++) s[4 * k + i] = kb::add(s[4 * k + i], t);
    }
}
struct Tables {
    uint32_t ext[8][16], internal[13], diag[16];
    Tables() {
        for (int r = 0; r < 4; r++)
            for (int i = 0; i < 16; i++) {
                ext[r][i] = kb::to_mont(BFGPU_RC_16_30[r][i]);
                ext[4 + r][i] = kb::to_mont(BFGPU_RC_16_30[17 + r][i]);
            }
        for (int r = 0; r < 13; r++) internal[r] = kb::to_mont(BFGPU_RC_16_30[4 + r][0]);
        auto frac = [](int sign, unsigned k) {
            uint32_t v = kb::ONE;
            for (unsigned i = 0; i < k; i++) v = kb::halve(v);
            return sign < 0 ? kb::neg(v) : v;
        };
        auto small = [](int v) { return v >= 0 ? kb::to_mont((uint32_t)v) : kb::neg(kb::to_mont((uint32_t)(-v))); };
        uint32_t d[16] = {small(-2), small(1), small(2), frac(1, 1), small(3), small(4), frac(-1, 1), small(-3),
                          small(-4), frac(1, 8), frac(1, 3), frac(1, 24), frac(-1, 8), frac(-1, 3), frac(-1, 4), frac(-1, 24)};
        for (int i = 0; i < 16; i++) diag[i] = d[i];
    }
};
inline const Tables& tables() {
    static const Tables t;
    return t;
}
inline void permute(uint32_t* s) {
    const Tables& T = tables();
    external_linear(s);
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(kb::add(s[i], T.ext[r][i]));
        external_linear(s);
    }
    for (int r = 0; r < 13; r++) {
        s[0] = sbox(kb::add(s[0], T.internal[r]));
        uint32_t sum = 0;
        for (int i = 0; i < 16; i++) sum = kb::add(sum, s[i]);
        for (int i = 0; i < 16; i++) s[i] = kb::add(kb::mul(s[i], T.diag[i]), sum);
    }
    for (int r = 4; r < 8; r++) {
        for (int i = 0; i < 16; i++) s[i] = sbox(kb::add(s[i], T.ext[r][i]));
        external_linear(s);
    }
}

}  // namespace host_p2

struct bfgpu_challenger {
    uint32_t state[16] = {0};
    std::vector<uint32_t> input, output;  // Montgomery words

    void duplex() {
        for (size_t i = 0; i < input.size(); i++) state[i] = input[i];
        input.clear();
        host_p2::permute(state);
        output.assign(state, state + 8);
    }
    void observe(uint32_t v) {
        output.clear();
        input.push_back(v);
        if (input.size() == 8) duplex();
    }
    void observe_slice(const uint32_t* v, size_t n) {
        for (size_t i = 0; i < n; i++) observe(v[i]);
    }
    void observe_ext(const kb::Ext& e) { observe_slice(e.c, 4); }
    uint32_t sample() {
        if (!input.empty() || output.empty()) duplex();
        uint32_t v = output.back();
        output.pop_back();
        return v;
    }
    kb::Ext sample_ext() {
        kb::Ext e;
        for (int i = 0; i < 4; i++) e.c[i] = sample();
        return e;
    }
    uint32_t sample_bits(unsigned bits) { return kb::from_mont(sample()) & ((1u << bits) - 1); }
    bool check_witness(unsigned bits, uint32_t w_mont) {
        observe(w_mont);
        return sample_bits(bits) == 0;
    }
};
