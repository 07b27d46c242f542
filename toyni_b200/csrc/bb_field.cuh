// BabyBear arithmetic for sm_100a, 32-bit lanes.
//
// The reference computes on canonical u64 values with `(a*b as u128) % p` (src/babybear.rs:169-178)
// and, in its CUDA path, 64-bit Barrett (cuda/ntt_kernel.cu:49-67).  Field arithmetic is exact, so any
// reduction strategy gives identical canonical results; here everything is 32-bit:
//   * Shoup multiplication by a precomputed constant (twiddles): 3 IMAD, result in [0,2p)
//   * Montgomery multiplication (R = 2^32) for data-dependent products: canonical result
//   * reductions as min(v, v-p), which ptxas turns into one VIADDMNMX.U32
// 2p < 2^32 < 3p, so "lazy" values live in [0,2p) and nothing wider fits a lane.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BB_HD __host__ __device__ __forceinline__
#define BB_D __device__ __forceinline__
#else
#define BB_HD inline
#define BB_D inline
#endif

namespace bb {

constexpr uint32_t P = 2013265921u;          // 2^31 - 2^27 + 1, src/babybear.rs:8
constexpr uint32_t P_INV = 0x88000001u;      // p^-1 mod 2^32
constexpr uint32_t R_MOD_P = 268435454u;     // 2^32 mod p   (Montgomery one)
constexpr uint32_t R2_MOD_P = 1172168163u;   // 2^64 mod p
constexpr uint32_t HALF = 1006632961u;       // 2^-1 mod p
constexpr uint32_t GEN27 = 440564289u;       // generator of the 2^27 subgroup, src/babybear.rs:122
constexpr uint32_t EXT_W = 11u;              // X^4 = 11, src/ext.rs:20

// ---------------------------------------------------------------- host+device scalar helpers
BB_HD uint32_t add(uint32_t a, uint32_t b) {  // canonical in, canonical out
    uint32_t s = a + b;
    uint32_t t = s - P;
    return s < t ? s : t;  // min(s, s-p): s-p wraps to a huge value when s < p
}
BB_HD uint32_t sub(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    uint32_t t = d + P;
    return d < t ? d : t;  // when a<b, d wraps and d+p is the small (correct) one
}
BB_HD uint32_t reduce2p(uint32_t v) {  // [0,2p) -> [0,p)
    uint32_t t = v - P;
    return v < t ? v : t;
}
BB_HD uint32_t halve(uint32_t a) {  // a/2 mod p for canonical a
    return (a >> 1) + ((a & 1u) ? HALF : 0u);
}

// Montgomery product a*b*2^-32 mod p, canonical.  Requires a*b < 2^32 * p (e.g. a < 2^32, b < p).
BB_HD uint32_t monty_mul(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a * b;
    uint32_t m = (uint32_t)t * P_INV;
    uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
    uint32_t r = (uint32_t)(t >> 32) - u;  // in (-p, p)
    uint32_t c = r + P;
    return r < c ? r : c;
}
BB_HD uint32_t to_monty(uint32_t a) { return monty_mul(a, R2_MOD_P); }
BB_HD uint32_t from_monty(uint32_t a) { return monty_mul(a, 1u); }
// plain product of two plain canonical values (two Montgomery steps)
BB_HD uint32_t mul(uint32_t a, uint32_t b) { return monty_mul(monty_mul(a, b), R2_MOD_P); }

BB_HD uint32_t pow(uint32_t base, uint64_t e) {
    uint32_t b = to_monty(base), r = R_MOD_P;
    while (e) {
        if (e & 1) r = monty_mul(r, b);
        b = monty_mul(b, b);
        e >>= 1;
    }
    return from_monty(r);
}
BB_HD uint32_t inv(uint32_t a) { return pow(a, P - 2); }

// principal 2^log_n-th root of unity, src/babybear.rs:118-126
BB_HD uint32_t root_of_unity(uint32_t log_n) { return pow(GEN27, 1ull << (27 - log_n)); }

// Shoup companion of a constant w < p: floor(w * 2^32 / p)
BB_HD uint32_t shoup_companion(uint32_t w) { return (uint32_t)(((uint64_t)w << 32) / P); }

// The same companion for a canonical w without a 64-bit division (device-side twiddle generation):
// q^ = floor(w K / 2^31) with K = floor(2^63 / p) = 2^32 + KLO is q or q - 1, and the remainder
// w 2^32 - q^ p < 2p fits 32 bits, so one comparison settles it.
BB_HD uint32_t shoup_companion_fast(uint32_t w) {
    constexpr uint64_t K = (1ull << 63) / P;
    static_assert((K >> 32) == 1, "K = 2^32 + KLO");
    constexpr uint32_t KLO = (uint32_t)K;
    const uint64_t y = (uint64_t)w * KLO;
    uint32_t q = (w << 1) + (uint32_t)(y >> 31);
    const uint32_t rem = 0u - q * P;
    return rem >= P ? q + 1u : q;
}

// x*w mod p in [0,2p) for ANY 32-bit x, given wp = shoup_companion(w)
BB_HD uint32_t shoup_mul_lazy(uint32_t x, uint32_t w, uint32_t wp) {
#ifdef __CUDA_ARCH__
    uint32_t q = __umulhi(x, wp);
#else
    uint32_t q = (uint32_t)(((uint64_t)x * wp) >> 32);
#endif
    return x * w - q * P;
}

// ---------------------------------------------------------------- quartic extension (device layout: 4 x u32)
struct Ext {
    uint32_t c[4];
};

BB_HD Ext ext_add(const Ext& a, const Ext& b) {
    Ext r;
    for (int k = 0; k < 4; k++) r.c[k] = add(a.c[k], b.c[k]);
    return r;
}
BB_HD Ext ext_sub(const Ext& a, const Ext& b) {
    Ext r;
    for (int k = 0; k < 4; k++) r.c[k] = sub(a.c[k], b.c[k]);
    return r;
}
// Montgomery-form product: inputs a (plain) and bm (Montgomery form) -> plain a*b.  src/ext.rs:178-192.
// Each output limb is a sum of at most four 62-bit products plus W-scaled ones, accumulated in 64 bits
// after one Montgomery reduction per partial sum keeps everything exact mod p.
BB_HD Ext ext_mul_monty(const Ext& a, const Ext& bm) {
    // wb[k] = 11 * bm[k] mod p, still in Montgomery form (small-constant multiple via repeated adds is
    // slower than one monty_mul with 11 in Montgomery form)
    const uint32_t W_M = (uint32_t)(((uint64_t)EXT_W << 32) % P);
    uint32_t w1 = monty_mul(bm.c[1], W_M), w2 = monty_mul(bm.c[2], W_M), w3 = monty_mul(bm.c[3], W_M);
    // w*b in Montgomery form needs one more factor R: monty_mul(bm, W_M) = b*R*11*R/R = 11*b*R  (ok)
    Ext r;
    r.c[0] = add(add(monty_mul(a.c[0], bm.c[0]), monty_mul(a.c[1], w3)), add(monty_mul(a.c[2], w2), monty_mul(a.c[3], w1)));
    r.c[1] = add(add(monty_mul(a.c[0], bm.c[1]), monty_mul(a.c[1], bm.c[0])), add(monty_mul(a.c[2], w3), monty_mul(a.c[3], w2)));
    r.c[2] = add(add(monty_mul(a.c[0], bm.c[2]), monty_mul(a.c[1], bm.c[1])), add(monty_mul(a.c[2], bm.c[0]), monty_mul(a.c[3], w3)));
    r.c[3] = add(add(monty_mul(a.c[0], bm.c[3]), monty_mul(a.c[1], bm.c[2])), add(monty_mul(a.c[2], bm.c[1]), monty_mul(a.c[3], bm.c[0])));
    return r;
}

}  // namespace bb
