// oracle/packed_kb.h — TEST / BENCH INFRASTRUCTURE (CPU arm): 16 KoalaBear residues per AVX-512 register, with the subset of the
// `kb::` interface (csrc/kb31.cuh) that the generated AIR programs use, so that csrc/gen_air.cuh can be compiled a second time with
// `kb` -> `pkb` and `uint32_t` -> `pkb::V` and evaluate SIXTEEN rows per call — what Plonky3's `PackedMontyField31AVX512` does for
// the reference's `quotient_values` / `generate_permutation_trace` (crates/stark/src/quotient.rs:64-70 packs rows the same way).
// Words are Montgomery residues (R = 2^32) in [0, p), exactly as in kb31.cuh; every function is the lane-wise image of its scalar
// namesake, which tests/test_oracle_fast_air.py checks bit for bit.
#pragma once
#include <immintrin.h>

#include <cstdint>

#define PKB_TGT __attribute__((target("avx512f,avx512dq,avx512bw,avx512vl")))

namespace pkb {

constexpr uint32_t P = 0x7f000001u, PINV = 0x81000001u, ONE = 0x01fffffeu, R2 = 402124772u;

struct V {
    __m512i v;
    V() = default;
    PKB_TGT V(__m512i x) : v(x) {}
    PKB_TGT V(uint32_t c) : v(_mm512_set1_epi32((int)c)) {}  // a (Montgomery) constant in every lane
    PKB_TGT V(int c) : v(_mm512_set1_epi32(c)) {}
};

PKB_TGT inline V add(V a, V b) {
    __m512i t = _mm512_add_epi32(a.v, b.v);
    return _mm512_min_epu32(t, _mm512_sub_epi32(t, _mm512_set1_epi32((int)P)));
}
PKB_TGT inline V sub(V a, V b) {
    __m512i t = _mm512_sub_epi32(a.v, b.v);
    return _mm512_min_epu32(t, _mm512_add_epi32(t, _mm512_set1_epi32((int)P)));
}
PKB_TGT inline V neg(V a) { return sub(V(_mm512_setzero_si512()), a); }
PKB_TGT inline V dbl(V a) { return add(a, a); }
// Montgomery product on the even and odd lanes (six 32x32->64 multiplies), corrected into [0, p)
PKB_TGT inline V mul(V a, V b) {
    const __m512i p = _mm512_set1_epi32((int)P), mu = _mm512_set1_epi32((int)PINV);
    __m512i ao = _mm512_srli_epi64(a.v, 32), bo = _mm512_srli_epi64(b.v, 32);
    __m512i pe = _mm512_mul_epu32(a.v, b.v), po = _mm512_mul_epu32(ao, bo);
    __m512i qe = _mm512_mul_epu32(pe, mu), qo = _mm512_mul_epu32(po, mu);
    __m512i de = _mm512_sub_epi64(pe, _mm512_mul_epu32(qe, p)), dv = _mm512_sub_epi64(po, _mm512_mul_epu32(qo, p));
    __m512i r = _mm512_mask_blend_epi32(0xAAAA, _mm512_srli_epi64(de, 32), dv);  // high halves: values in (-p, p)
    return _mm512_min_epu32(r, _mm512_add_epi32(r, p));
}
PKB_TGT inline V sqr(V a) { return mul(a, a); }
PKB_TGT inline V to_mont(V canonical) { return mul(canonical, V(R2)); }
PKB_TGT inline V from_mont(V m) { return mul(m, V(1u)); }
PKB_TGT inline V mul3(V a) { return add(dbl(a), a); }
PKB_TGT inline V sqr_n(V a, int n) {
    for (int i = 0; i < n; i++) a = sqr(a);
    return a;
}
// a^(p-2): the same addition chain as kb::inv
PKB_TGT inline V inv(V a) {
    V x2 = mul(sqr(a), a);
    V x3 = mul(sqr(x2), a);
    V x6 = mul(sqr_n(x3, 3), x3);
    V x12 = mul(sqr_n(x6, 6), x6);
    V x24 = mul(sqr_n(x12, 12), x12);
    return mul(sqr_n(x6, 25), x24);
}

struct Ext {
    V c[4];
};
using ExtAcc = Ext;  // no lazy 64-bit accumulation on the packed side: every term is reduced
PKB_TGT inline Ext ext_zero() { return Ext{{V(0u), V(0u), V(0u), V(0u)}}; }
PKB_TGT inline Ext ext_one() { return Ext{{V(ONE), V(0u), V(0u), V(0u)}}; }
PKB_TGT inline Ext ext_from_base(V a) { return Ext{{a, V(0u), V(0u), V(0u)}}; }
PKB_TGT inline Ext ext_add(Ext a, Ext b) { return Ext{{add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2]), add(a.c[3], b.c[3])}}; }
PKB_TGT inline Ext ext_sub(Ext a, Ext b) { return Ext{{sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2]), sub(a.c[3], b.c[3])}}; }
PKB_TGT inline Ext ext_scale(Ext a, V s) { return Ext{{mul(a.c[0], s), mul(a.c[1], s), mul(a.c[2], s), mul(a.c[3], s)}}; }
PKB_TGT inline ExtAcc ext_acc_zero() { return ext_zero(); }
PKB_TGT inline ExtAcc ext_acc_from(Ext e) { return e; }
PKB_TGT inline Ext ext_acc_reduce(ExtAcc a) { return a; }
PKB_TGT inline void ext_mac(ExtAcc& a, Ext e, V s) {
    for (int k = 0; k < 4; k++) a.c[k] = add(a.c[k], mul(e.c[k], s));
}
PKB_TGT inline Ext ext_mul(Ext a, Ext b) {  // F_p[X]/(X^4 - 3)
    V a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    V b0 = b.c[0], b1 = b.c[1], b2 = b.c[2], b3 = b.c[3];
    V t4 = add(add(mul(a1, b3), mul(a2, b2)), mul(a3, b1));
    V t5 = add(mul(a2, b3), mul(a3, b2));
    V t6 = mul(a3, b3);
    Ext r;
    r.c[0] = add(mul(a0, b0), mul3(t4));
    r.c[1] = add(add(mul(a0, b1), mul(a1, b0)), mul3(t5));
    r.c[2] = add(add(add(mul(a0, b2), mul(a1, b1)), mul(a2, b0)), mul3(t6));
    r.c[3] = add(add(mul(a0, b3), mul(a1, b2)), add(mul(a2, b1), mul(a3, b0)));
    return r;
}
PKB_TGT inline Ext ext_inv(Ext a) {  // same tower formula as kb::ext_inv
    V a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    V A2_0 = add(sqr(a0), mul3(sqr(a2)));
    V A2_1 = dbl(mul(a0, a2));
    V B2_0 = add(sqr(a1), mul3(sqr(a3)));
    V B2_1 = dbl(mul(a1, a3));
    V n0 = sub(A2_0, mul3(B2_1));
    V n1 = sub(A2_1, B2_0);
    V d = sub(sqr(n0), mul3(sqr(n1)));
    V di = inv(d);
    V m0 = mul(n0, di), m1 = neg(mul(n1, di));
    Ext r;
    r.c[0] = add(mul(a0, m0), mul3(mul(a2, m1)));
    r.c[2] = add(mul(a0, m1), mul(a2, m0));
    r.c[1] = neg(add(mul(a1, m0), mul3(mul(a3, m1))));
    r.c[3] = neg(add(mul(a1, m1), mul(a3, m0)));
    return r;
}

}  // namespace pkb
