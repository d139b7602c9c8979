// kb31.cuh — KoalaBear (p = 2^31 - 2^24 + 1) arithmetic for sm_100a, Montgomery form R = 2^32.
//
// Device-side counterpart of the reference's `Val = KoalaBear` / `Challenge =
// BinomialExtensionField<Val, 4>` (reference crates/stark/src/kb31_poseidon2.rs:20-21).  All device
// buffers hold fully reduced Montgomery residues in [0, p); host buffers cross the C ABI either
// canonical or Montgomery (bfgpu_set_repr).  Everything runs on the INT32 pipes: a product is
// IMAD.WIDE + IMAD + IMAD.HI, the final correction an IADD + unsigned min.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#define KB_D __device__ __forceinline__
#else
#define KB_HD inline
#define KB_D inline
#endif

namespace kb {

constexpr uint32_t P = 0x7f000001u;        // 2130706433
constexpr uint32_t PINV = 0x81000001u;     // p^{-1} mod 2^32
constexpr uint32_t ONE = 0x01fffffeu;      // 2^32 mod p  (Montgomery form of 1)
constexpr uint32_t R2 = 402124772u;        // 2^64 mod p
constexpr uint32_t GEN = 3u;               // multiplicative generator (canonical)
constexpr int TWO_ADICITY = 24;

KB_HD uint32_t umin_(uint32_t a, uint32_t b) { return a < b ? a : b; }

// a + b mod p for a, b in [0, p)
KB_HD uint32_t add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return umin_(s, s - P);
}
// a - b mod p for a, b in [0, p)
KB_HD uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    return umin_(d, d + P);
}
KB_HD uint32_t neg(uint32_t a) { return a ? P - a : 0u; }
KB_HD uint32_t dbl(uint32_t a) { return add(a, a); }

// Montgomery reduction of t < 2^32 * p : returns t / 2^32 mod p in [0, p)
KB_HD uint32_t mont_reduce(uint64_t t) {
    uint32_t m = (uint32_t)t * PINV;
#if defined(__CUDA_ARCH__)
    uint32_t u = __umulhi(m, P);
#else
    uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
    uint32_t r = (uint32_t)(t >> 32) - u;
    return umin_(r, r + P);
}
// Lazy multiply-accumulate for long dot products: acc (< 2^32 p) += a * b with a, b in [0, p), kept below 2^32 p by
// a conditional subtraction of 2^32 p on the high word (one IMAD.WIDE + one add/min instead of a full Montgomery
// product and a modular add).  mont_reduce(acc) at the end is the Montgomery-form sum of the products.
KB_HD void mac(uint64_t& acc, uint32_t a, uint32_t b) {
    acc += (uint64_t)a * b;  // < 2^32 p + p^2 < 2^64
    uint32_t hi = (uint32_t)(acc >> 32);
    hi = umin_(hi, hi - P);
    acc = ((uint64_t)hi << 32) | (uint32_t)acc;
}
// Two products per conditional subtraction: acc < 2^32 p and 2 p^2 < 0.4923 * 2^64 keep the sum below 2^64, and one subtraction of
// 2^32 p brings it back under 2^32 p (0.9884 - 0.4961 < 0.4961): one add/min per TWO terms of a long dot product.
KB_HD void mac2(uint64_t& acc, uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1) {
    acc += (uint64_t)a0 * b0;
    acc += (uint64_t)a1 * b1;
    uint32_t hi = (uint32_t)(acc >> 32);
    hi = umin_(hi, hi - P);
    acc = ((uint64_t)hi << 32) | (uint32_t)acc;
}
// Montgomery product.  Exact for a in [0, 2^32), b in [0, p): result in [0, p).
KB_HD uint32_t mul(uint32_t a, uint32_t b) { return mont_reduce((uint64_t)a * b); }
KB_HD uint32_t sqr(uint32_t a) { return mul(a, a); }

KB_HD uint32_t to_mont(uint32_t canonical) { return mul(canonical, R2); }
KB_HD uint32_t from_mont(uint32_t m) { return mont_reduce((uint64_t)m); }

// a / 2 mod p (representation independent)
KB_HD uint32_t halve(uint32_t a) { return (a >> 1) + ((a & 1u) ? ((P + 1u) >> 1) : 0u); }

KB_HD uint32_t pow(uint32_t a_mont, uint64_t e) {
    uint32_t r = ONE;
    while (e) {
        if (e & 1) r = mul(r, a_mont);
        a_mont = sqr(a_mont);
        e >>= 1;
    }
    return r;
}
// a^(p-2) by a fixed addition chain: p - 2 = 2^31 - 2^24 - 1 = (2^6 - 1) * 2^25 + (2^24 - 1), so with x_k = a^(2^k - 1)
// (x2, x3, x6, x12, x24 by doubling) the inverse is x6^(2^25) * x24: 48 squarings + 6 products, no data-dependent branch
// (the generic square-and-multiply ladder takes 31 + 30 with a loop and a branch per bit).
KB_HD uint32_t sqr_n(uint32_t a, int n) {
    for (int i = 0; i < n; i++) a = sqr(a);
    return a;
}
KB_HD uint32_t inv(uint32_t a) {
    uint32_t x2 = mul(sqr(a), a);
    uint32_t x3 = mul(sqr(x2), a);
    uint32_t x6 = mul(sqr_n(x3, 3), x3);
    uint32_t x12 = mul(sqr_n(x6, 6), x6);
    uint32_t x24 = mul(sqr_n(x12, 12), x12);
    return mul(sqr_n(x6, 25), x24);
}
// generator of the order-2^bits subgroup, Montgomery form: 3^((p-1)/2^bits)
KB_HD uint32_t two_adic_generator(unsigned bits) { return pow(to_mont(GEN), (uint64_t)(P - 1) >> bits); }

KB_HD uint32_t bitrev(uint32_t x, unsigned bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

// ---- F_p^4 = F_p[X]/(X^4 - 3), Montgomery coefficients ------------------------------------------
struct Ext {
    uint32_t c[4];
};
constexpr uint32_t W_MONT = 0x05fffffau;  // 3 * 2^32 mod p

KB_HD Ext ext_zero() { return Ext{{0, 0, 0, 0}}; }
KB_HD Ext ext_one() { return Ext{{ONE, 0, 0, 0}}; }
KB_HD Ext ext_from_base(uint32_t a) { return Ext{{a, 0, 0, 0}}; }
KB_HD Ext ext_add(Ext a, Ext b) { return Ext{{add(a.c[0], b.c[0]), add(a.c[1], b.c[1]), add(a.c[2], b.c[2]), add(a.c[3], b.c[3])}}; }
KB_HD Ext ext_sub(Ext a, Ext b) { return Ext{{sub(a.c[0], b.c[0]), sub(a.c[1], b.c[1]), sub(a.c[2], b.c[2]), sub(a.c[3], b.c[3])}}; }
KB_HD Ext ext_neg(Ext a) { return Ext{{neg(a.c[0]), neg(a.c[1]), neg(a.c[2]), neg(a.c[3])}}; }
KB_HD Ext ext_scale(Ext a, uint32_t s) { return Ext{{mul(a.c[0], s), mul(a.c[1], s), mul(a.c[2], s), mul(a.c[3], s)}}; }
// unreduced accumulator of ext * base products (see mac): sum_k e_k * s_k with one reduction per coefficient at the end
struct ExtAcc {
    uint64_t c[4];
};
KB_HD ExtAcc ext_acc_zero() { return ExtAcc{{0, 0, 0, 0}}; }
// accumulator holding the reduced element e (mont_reduce(e * 2^32) = e)
KB_HD ExtAcc ext_acc_from(Ext e) { return ExtAcc{{(uint64_t)e.c[0] << 32, (uint64_t)e.c[1] << 32, (uint64_t)e.c[2] << 32, (uint64_t)e.c[3] << 32}}; }
KB_HD void ext_mac(ExtAcc& a, Ext e, uint32_t s) {
    mac(a.c[0], e.c[0], s);
    mac(a.c[1], e.c[1], s);
    mac(a.c[2], e.c[2], s);
    mac(a.c[3], e.c[3], s);
}
KB_HD Ext ext_acc_reduce(ExtAcc a) { return Ext{{mont_reduce(a.c[0]), mont_reduce(a.c[1]), mont_reduce(a.c[2]), mont_reduce(a.c[3])}}; }
KB_HD uint32_t mul3(uint32_t a) { return add(dbl(a), a); }
// Sum of four products of reduced words, then ONE Montgomery reduction: 4 p^2 < 2^64 and 4 p^2 < 2 * 2^32 p (2p < 2^32), so the
// products chain through one 64-bit accumulator (IMAD.WIDE with accumulate) and a single conditional subtraction of 2^32 p on the
// high word brings the sum below 2^32 p, where mont_reduce is exact.
KB_HD uint32_t dot4(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2, uint32_t a3, uint32_t b3) {
    uint64_t t = (uint64_t)a0 * b0 + (uint64_t)a1 * b1 + (uint64_t)a2 * b2 + (uint64_t)a3 * b3;
    uint32_t hi = (uint32_t)(t >> 32);
    hi = umin_(hi, hi - P);
    return mont_reduce(((uint64_t)hi << 32) | (uint32_t)t);
}
KB_HD Ext ext_mul(Ext a, Ext b) {
    // schoolbook with X^4 = 3: every coefficient is a four-term dot product against (b, 3b): 16 wide products + 4 reductions
    // (the product-by-product form paid 16 reductions and 9 modular additions: ~110 instructions against ~50)
    const uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    const uint32_t b0 = b.c[0], b1 = b.c[1], b2 = b.c[2], b3 = b.c[3];
    const uint32_t w1 = mul3(b1), w2 = mul3(b2), w3 = mul3(b3);
    Ext r;
    r.c[0] = dot4(a0, b0, a1, w3, a2, w2, a3, w1);
    r.c[1] = dot4(a0, b1, a1, b0, a2, w3, a3, w2);
    r.c[2] = dot4(a0, b2, a1, b1, a2, b0, a3, w3);
    r.c[3] = dot4(a0, b3, a1, b2, a2, b1, a3, b0);
    return r;
}
KB_HD Ext ext_sqr(Ext a) { return ext_mul(a, a); }
KB_HD Ext ext_inv(Ext a) {
    // a = A + B X over K = F_p[Y]/(Y^2 - 3), Y = X^2; a^{-1} = (A - B X) / (A^2 - Y B^2)
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    uint32_t A2_0 = add(sqr(a0), mul3(sqr(a2)));
    uint32_t A2_1 = dbl(mul(a0, a2));
    uint32_t B2_0 = add(sqr(a1), mul3(sqr(a3)));
    uint32_t B2_1 = dbl(mul(a1, a3));
    uint32_t n0 = sub(A2_0, mul3(B2_1));
    uint32_t n1 = sub(A2_1, B2_0);
    uint32_t d = sub(sqr(n0), mul3(sqr(n1)));
    uint32_t di = inv(d);
    uint32_t m0 = mul(n0, di), m1 = neg(mul(n1, di));
    Ext r;
    r.c[0] = add(mul(a0, m0), mul3(mul(a2, m1)));
    r.c[2] = add(mul(a0, m1), mul(a2, m0));
    r.c[1] = neg(add(mul(a1, m0), mul3(mul(a3, m1))));
    r.c[3] = neg(add(mul(a1, m1), mul(a3, m0)));
    return r;
}

}  // namespace kb
