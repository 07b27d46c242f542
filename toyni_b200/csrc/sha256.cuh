// SHA-256 (FIPS 180-4) device primitives shared by the Merkle kernels and the fused fold + leaf-hash kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bb {

static __constant__ uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
// SHA-256 is ALU-pipe bound on sm_100 (rotates and LOP3 only run there: ~980 of the ~1330 instructions of a
// compression); its ~350 additions are steered to the otherwise idle FMA pipe by writing them as a*1+b with a
// multiplier the compiler cannot see through (a __constant__ word that happens to hold 1).
static __constant__ uint32_t c_sha_one = 1u;
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b) {
#ifdef BB_SHA_PLAIN_ADD
    return a + b;
#else
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(c_sha_one), "r"(b));
    return r;
#endif
}
__device__ __forceinline__ uint32_t bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

struct Sha {
    uint32_t h[8];
};

__device__ __forceinline__ void sha_init(Sha& s) {
    s.h[0] = 0x6a09e667; s.h[1] = 0xbb67ae85; s.h[2] = 0x3c6ef372; s.h[3] = 0xa54ff53a;
    s.h[4] = 0x510e527f; s.h[5] = 0x9b05688c; s.h[6] = 0x1f83d9ab; s.h[7] = 0x5be0cd19;
}

// FIPS 180-4 compression with a rolling 16-word schedule, fully unrolled so constant words fold away
__device__ __forceinline__ void sha_compress(Sha& s, uint32_t w[16]) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t wi;
        if (i < 16) {
            wi = w[i];
        } else {
            uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
            uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
            wi = fadd(fadd(w[i & 15], s0), fadd(w[(i - 7) & 15], s1));
            w[i & 15] = wi;
        }
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = fadd(fadd(fadd(h, S1), fadd(ch, K256[i])), wi);
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = fadd(S0, mj);
        h = g; g = f; f = e; e = fadd(d, t1); d = c; c = b; b = a; a = fadd(t1, t2);
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

// The second block of a node hash (0x01 || L || R is 65 bytes) holds ONE message byte, 0x80, zeros and the length:
// its whole 64-word message schedule is a function of that byte.  sha_pad_table_kernel tabulates K[i] + W_b[i] for the
// 256 values of b (64 KB, [round][byte] so that a round reads one 1 KB row), and sha_compress_tab runs the 64 rounds
// of that block straight from the table: no schedule arithmetic (48 x 8 ALU-pipe instructions) and no K additions.
constexpr int SHA_PAD_TAB_WORDS = 64 * 256;
static __global__ void sha_pad_table_kernel(uint32_t* __restrict__ tab, uint32_t bit_length) {
    const uint32_t b = threadIdx.x;  // 256 threads
    uint32_t w[64];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    w[0] = (b << 24) | 0x00800000u;
    w[15] = bit_length;
    for (int i = 16; i < 64; i++) {
        const uint32_t w15 = w[i - 15], w2 = w[i - 2];
        const uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
        const uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    for (int i = 0; i < 64; i++) tab[i * 256 + b] = w[i] + K256[i];
}

template <bool SMEM_TAB>
__device__ __forceinline__ void sha_compress_tab(Sha& s, const uint32_t* __restrict__ tab, uint32_t byte) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
    const uint32_t* t = tab + byte;
#pragma unroll
    for (int i = 0; i < 64; i++) {
        const uint32_t kw = SMEM_TAB ? t[i * 256] : __ldg(t + i * 256);
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = fadd(fadd(h, S1), fadd(ch, kw));
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = fadd(S0, mj);
        h = g; g = f; f = e; e = fadd(d, t1); d = c; c = b; b = a; a = fadd(t1, t2);
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

__device__ __forceinline__ void store_digest(uint8_t* dst, const Sha& s) {
    uint4* o = reinterpret_cast<uint4*>(dst);  // 32-byte aligned
    o[0] = make_uint4(bswap(s.h[0]), bswap(s.h[1]), bswap(s.h[2]), bswap(s.h[3]));
    o[1] = make_uint4(bswap(s.h[4]), bswap(s.h[5]), bswap(s.h[6]), bswap(s.h[7]));
}


// Digest of one prover leaf: SHA256(0x00 || salt[16] (optional) || LE-u64 of each limb), src/merkle.rs:109-114 with
// the leaf bytes of src/fibonacci.rs:340-363.  The byte stream after the tag is a sequence of little-endian words q[]:
// salt (4 words) then (value, 0) per limb; big-endian message word i is (bs[i-1] << 24) | (bs[i] >> 8) with
// bs = bswap(q) and bs[-1] = tag.  Every such leaf (9 / 25 / 33 / 49 bytes) is a single block.
template <int LIMBS, bool SALTED>
__device__ __forceinline__ void leaf_digest(const uint32_t (&vals)[LIMBS], uint4 salt, Sha& s) {
    constexpr int NQ = (SALTED ? 4 : 0) + 2 * LIMBS;  // stream words
    uint32_t q[NQ + 1];
    int k = 0;
    if (SALTED) {
        q[k++] = salt.x; q[k++] = salt.y; q[k++] = salt.z; q[k++] = salt.w;
    }
#pragma unroll
    for (int l = 0; l < LIMBS; l++) {
        q[k++] = vals[l];
        q[k++] = 0;
    }
    q[NQ] = 0x80u;  // padding byte right after the message
    uint32_t w[16];
    uint32_t prev = 0x00u;  // LEAF_TAG
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (j <= NQ) {
            uint32_t cur = bswap(q[j]);
            w[j] = (prev << 24) | (cur >> 8);
            prev = cur;
        } else if (j == NQ + 1) {
            w[j] = prev << 24;  // last byte of the 0x80 word (zero) spills over: always 0
        } else {
            w[j] = 0;
        }
    }
    w[15] = (uint32_t)((1 + 4 * NQ) * 8);  // message length in bits (NQ <= 12, so word 15 is free)
    sha_init(s);
    sha_compress(s, w);
}

}  // namespace bb
