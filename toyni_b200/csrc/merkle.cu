// SHA-256 Merkle commitment on the device.
//   leaf  = SHA256(0x00 || leaf bytes)            src/merkle.rs:105,109-114
//   node  = SHA256(0x01 || left || right)         src/merkle.rs:106,117-123
//   prover leaf bytes = salt[16] || LE-u64(value) (or the value alone)   src/fibonacci.rs:340-363
// One thread per hash.  Every prover leaf fits one SHA-256 block; a node is two blocks, the second of which is
// almost entirely padding, so its message schedule constant-folds.  Digests are stored as the standard 32 bytes.
// The reference builds Vec<Vec<u8>> per level and clones each level (src/merkle.rs:29-47); here the levels
// live back to back in one device array.
#include "merkle.cuh"

namespace bb {

__constant__ uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
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
            wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
            w[i & 15] = wi;
        }
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K256[i] + wi;
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

__device__ __forceinline__ void store_digest(uint8_t* dst, const Sha& s) {
    uint4* o = reinterpret_cast<uint4*>(dst);  // 32-byte aligned
    o[0] = make_uint4(bswap(s.h[0]), bswap(s.h[1]), bswap(s.h[2]), bswap(s.h[3]));
    o[1] = make_uint4(bswap(s.h[4]), bswap(s.h[5]), bswap(s.h[6]), bswap(s.h[7]));
}

// Leaf hash of one field value.  The byte stream after the tag is a sequence of little-endian words q[]:
// salt (4 words, optional) then (value, 0) per limb; big-endian message word i is (bs[i-1] << 24) | (bs[i] >> 8)
// with bs = bswap(q) and bs[-1] = tag.
template <int LIMBS, bool SALTED>
__global__ void __launch_bounds__(256) leaf_hash_kernel(const uint32_t* __restrict__ vals, const uint4* __restrict__ salts,
                                                        uint8_t* __restrict__ nodes, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int NQ = (SALTED ? 4 : 0) + 2 * LIMBS;  // stream words
    uint32_t q[NQ + 1];
    int k = 0;
    if (SALTED) {
        uint4 sv = salts[i];
        q[k++] = sv.x; q[k++] = sv.y; q[k++] = sv.z; q[k++] = sv.w;
    }
    if (LIMBS == 1) {
        q[k++] = vals[i];
        q[k++] = 0;
    } else {
        uint4 v = reinterpret_cast<const uint4*>(vals)[i];
        q[k++] = v.x; q[k++] = 0; q[k++] = v.y; q[k++] = 0; q[k++] = v.z; q[k++] = 0; q[k++] = v.w; q[k++] = 0;
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
    Sha s;
    sha_init(s);
    sha_compress(s, w);
    store_digest(nodes + 32 * i, s);
}

// Parent level: node j = SHA256(0x01 || child[2j] || child[2j+1]); an odd level pairs the last child with itself.
__global__ void __launch_bounds__(256) node_hash_kernel(const uint8_t* __restrict__ child, uint8_t* __restrict__ parent,
                                                        size_t n_child, size_t n_parent) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_parent) return;
    size_t li = 2 * j, ri = (2 * j + 1 < n_child) ? 2 * j + 1 : 2 * j;
    const uint4* L = reinterpret_cast<const uint4*>(child + 32 * li);
    const uint4* R = reinterpret_cast<const uint4*>(child + 32 * ri);
    uint4 l0 = L[0], l1 = L[1], r0 = R[0], r1 = R[1];
    uint32_t m[16] = {bswap(l0.x), bswap(l0.y), bswap(l0.z), bswap(l0.w), bswap(l1.x), bswap(l1.y), bswap(l1.z), bswap(l1.w),
                      bswap(r0.x), bswap(r0.y), bswap(r0.z), bswap(r0.w), bswap(r1.x), bswap(r1.y), bswap(r1.z), bswap(r1.w)};
    uint32_t w[16];
    uint32_t prev = 0x01u;  // NODE_TAG
#pragma unroll
    for (int k = 0; k < 16; k++) {
        w[k] = (prev << 24) | (m[k] >> 8);
        prev = m[k];
    }
    Sha s;
    sha_init(s);
    sha_compress(s, w);
    // second block: last message byte, 0x80, zeros, bit length 65*8
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = 0;
    w[0] = (prev << 24) | 0x00800000u;
    w[15] = 65 * 8;
    sha_compress(s, w);
    store_digest(parent + 32 * j, s);
}

// Generic leaves of arbitrary length (MerkleTree::new over raw byte strings)
__global__ void __launch_bounds__(128) leaf_hash_bytes_kernel(const uint8_t* __restrict__ leaves, size_t n, size_t leaf_len,
                                                              uint8_t* __restrict__ nodes) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* src = leaves + i * leaf_len;
    Sha s;
    sha_init(s);
    const size_t total = leaf_len + 1;  // with the tag
    size_t pos = 0;                     // message bytes consumed
    bool done = false, pad_written = false;
    while (!done) {
        uint32_t w[16];
#pragma unroll 1
        for (int k = 0; k < 16; k++) {
            uint32_t word = 0;
            for (int b = 0; b < 4; b++) {
                size_t idx = pos + 4 * k + b;
                uint32_t byte = 0;
                if (idx == 0)
                    byte = 0x00;
                else if (idx < total)
                    byte = src[idx - 1];
                else if (idx == total)
                    byte = 0x80;
                word = (word << 8) | byte;
            }
            w[k] = word;
        }
        if (pos + 64 > total) pad_written = true;                 // the 0x80 byte landed in this block
        if (pad_written && (pos + 56 >= total + 1)) {               // room for the 8 length bytes
            unsigned long long bits = (unsigned long long)total * 8ull;
            w[14] = (uint32_t)(bits >> 32);
            w[15] = (uint32_t)bits;
            done = true;
        }
        sha_compress(s, w);
        pos += 64;
    }
    store_digest(nodes + 32 * i, s);
}

__global__ void gather_path_kernel(const uint8_t* __restrict__ nodes, size_t nleaves, size_t index, uint8_t* __restrict__ path) {
    // one thread per byte of each sibling digest; walks the levels exactly as src/merkle.rs:59-77
    size_t level_off = 0, level_n = nleaves, cur = index;
    int depth = 0;
    while (level_n > 1) {
        size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
        size_t src = (sib >= level_n) ? cur : sib;
        if (threadIdx.x < 32) path[32 * depth + threadIdx.x] = nodes[32 * (level_off + src) + threadIdx.x];
        depth++;
        cur /= 2;
        level_off += level_n;
        level_n = (level_n + 1) / 2;
    }
}

size_t merkle_node_count(size_t nleaves) {
    size_t total = nleaves, cur = nleaves;
    while (cur > 1) {
        cur = (cur + 1) / 2;
        total += cur;
    }
    return total;
}

static int upper_levels(uint8_t* d_nodes, size_t n, cudaStream_t s) {
    uint8_t* cur = d_nodes;
    size_t cur_n = n;
    while (cur_n > 1) {
        uint8_t* next = cur + 32 * cur_n;
        size_t next_n = (cur_n + 1) / 2;
        node_hash_kernel<<<(unsigned)((next_n + 255) / 256), 256, 0, s>>>(cur, next, cur_n, next_n);
        cur = next;
        cur_n = next_n;
    }
    return (int)cudaGetLastError();
}

int merkle_commit(const uint32_t* d_vals, int limbs, size_t n, const uint8_t* d_salts, uint8_t* d_nodes, cudaStream_t s) {
    if (n == 0 || (limbs != 1 && limbs != 4)) return (int)cudaErrorInvalidValue;
    unsigned blocks = (unsigned)((n + 255) / 256);
    const uint4* salts = reinterpret_cast<const uint4*>(d_salts);
    if (limbs == 1 && d_salts)
        leaf_hash_kernel<1, true><<<blocks, 256, 0, s>>>(d_vals, salts, d_nodes, n);
    else if (limbs == 1)
        leaf_hash_kernel<1, false><<<blocks, 256, 0, s>>>(d_vals, salts, d_nodes, n);
    else if (d_salts)
        leaf_hash_kernel<4, true><<<blocks, 256, 0, s>>>(d_vals, salts, d_nodes, n);
    else
        leaf_hash_kernel<4, false><<<blocks, 256, 0, s>>>(d_vals, salts, d_nodes, n);
    int rc = (int)cudaGetLastError();
    if (rc) return rc;
    return upper_levels(d_nodes, n, s);
}

int merkle_build_bytes(const uint8_t* d_leaves, size_t n, size_t leaf_len, uint8_t* d_nodes, cudaStream_t s) {
    if (n == 0) return (int)cudaErrorInvalidValue;
    leaf_hash_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_leaves, n, leaf_len, d_nodes);
    int rc = (int)cudaGetLastError();
    if (rc) return rc;
    return upper_levels(d_nodes, n, s);
}

int merkle_open(const uint8_t* d_nodes, size_t nleaves, size_t index, uint8_t* d_path, uint8_t* h_pos, size_t* depth,
                cudaStream_t s) {
    if (index >= nleaves) return (int)cudaErrorInvalidValue;
    size_t level_n = nleaves, cur = index, d = 0;
    while (level_n > 1) {  // position flags need no device data (src/merkle.rs:67-73)
        size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
        h_pos[d++] = (sib >= level_n) ? 1 : (uint8_t)(cur % 2 == 1);
        cur /= 2;
        level_n = (level_n + 1) / 2;
    }
    *depth = d;
    if (d > 0) gather_path_kernel<<<1, 32, 0, s>>>(d_nodes, nleaves, index, d_path);
    return (int)cudaGetLastError();
}

}  // namespace bb
