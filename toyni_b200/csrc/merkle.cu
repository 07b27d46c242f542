// SHA-256 Merkle commitment on the device.
//   leaf  = SHA256(0x00 || leaf bytes)            src/merkle.rs:105,109-114
//   node  = SHA256(0x01 || left || right)         src/merkle.rs:106,117-123
//   prover leaf bytes = salt[16] || LE-u64(value) (or the value alone)   src/fibonacci.rs:340-363
// One thread per hash.  Every prover leaf fits one SHA-256 block; a node is two blocks, the second of which is
// almost entirely padding, so its message schedule constant-folds.  Digests are stored as the standard 32 bytes.
// The reference builds Vec<Vec<u8>> per level and clones each level (src/merkle.rs:29-47); here the levels
// live back to back in one device array.
#include "merkle.cuh"

#include "sha256.cuh"

#include <mutex>

namespace bb {

// Leaf hash of one field value (leaf_digest in sha256.cuh).
template <int LIMBS, bool SALTED>
__global__ void __launch_bounds__(256) leaf_hash_kernel(const uint32_t* __restrict__ vals, const uint4* __restrict__ salts,
                                                        uint8_t* __restrict__ nodes, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v[LIMBS];
    if (LIMBS == 1) {
        v[0] = vals[i];
    } else {
        uint4 t = reinterpret_cast<const uint4*>(vals)[i];
        v[0] = t.x; v[LIMBS > 1 ? 1 : 0] = t.y; v[LIMBS > 2 ? 2 : 0] = t.z; v[LIMBS > 3 ? 3 : 0] = t.w;
    }
    uint4 sv = make_uint4(0, 0, 0, 0);
    if (SALTED) sv = salts[i];
    Sha s;
    leaf_digest<LIMBS, SALTED>(v, sv, s);
    store_digest(nodes + 32 * i, s);
}

// Parent level: node j = SHA256(0x01 || child[2j] || child[2j+1]); an odd level pairs the last child with itself.
template <bool SMEM_TAB = false>
__device__ __forceinline__ void node_hash_one(const uint8_t* __restrict__ child, uint8_t* __restrict__ parent, size_t n_child, size_t j,
                                              const uint32_t* __restrict__ pad_tab) {
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
    // second block: last message byte, 0x80, zeros, bit length 65*8 - its schedule comes from the table
    sha_compress_tab<SMEM_TAB>(s, pad_tab, prev & 0xFFu);
    store_digest(parent + 32 * j, s);
}

__global__ void __launch_bounds__(256) node_hash_kernel(const uint8_t* __restrict__ child, uint8_t* __restrict__ parent,
                                                        size_t n_child, size_t n_parent, const uint32_t* __restrict__ pad_tab) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_parent) return;
    node_hash_one(child, parent, n_child, j, pad_tab);
}

// Large levels: persistent CTAs keep the 64 KB padding-block table in shared memory (random 4-byte reads cost ~3 bank
// cycles there instead of up to 8 L1 wavefronts), 512 threads each so that three CTAs still fill an SM.
__global__ void __launch_bounds__(512) node_hash_smem_kernel(const uint8_t* __restrict__ child, uint8_t* __restrict__ parent,
                                                             size_t n_child, size_t n_parent, const uint32_t* __restrict__ pad_tab) {
    extern __shared__ uint32_t s_tab[];
    for (int i = threadIdx.x; i < SHA_PAD_TAB_WORDS / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(s_tab)[i] = __ldg(reinterpret_cast<const uint4*>(pad_tab) + i);
    __syncthreads();
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_parent; j += (size_t)gridDim.x * blockDim.x)
        node_hash_one<true>(child, parent, n_child, j, s_tab);
}

// Below 2^18 nodes a level no longer fills the GPU and a tree is a chain of latencies: one hash (two compressions, ~2 us
// of dependent instructions) per level, plus a launch and a drain if every level is its own kernel.  Two kernels
// walk several levels per launch instead, with a block barrier between levels:
//   subtree kernel: CTA b owns SUB_NODES consecutive nodes of a level whose size is a multiple of SUB_NODES and reduces
//                   them to ONE node, eight levels up (every intermediate level has an exact power-of-two share per
//                   CTA, so no odd-node duplication can occur inside); the CTAs of a launch spread over the SMs;
//   tail kernel:    a single CTA finishes what is left (at most 2 * TAIL_THREADS nodes, any size, odd levels included).
constexpr int SUB_LEVELS = 8, SUB_NODES = 1 << SUB_LEVELS, SUB_THREADS = SUB_NODES / 2;
__global__ void __launch_bounds__(SUB_THREADS) node_hash_subtree_kernel(uint8_t* __restrict__ level, size_t n_child,
                                                                        const uint32_t* __restrict__ pad_tab) {
    size_t base = (size_t)blockIdx.x * SUB_NODES;  // this CTA's first node of the current level
    uint32_t mine = SUB_NODES;                     // and how many it owns there
#pragma unroll 1
    for (int l = 0; l < SUB_LEVELS; l++) {
        uint8_t* parent = level + 32 * n_child;
        if (threadIdx.x < mine / 2) node_hash_one(level, parent, n_child, base / 2 + threadIdx.x, pad_tab);
        __syncthreads();  // block-scope ordering of the global stores above with the loads of the next level
        level = parent;
        n_child /= 2;
        base /= 2;
        mine /= 2;
    }
}

constexpr int TAIL_THREADS = 128;
__global__ void __launch_bounds__(TAIL_THREADS) node_hash_tail_kernel(uint8_t* __restrict__ level, size_t n_child,
                                                                      const uint32_t* __restrict__ pad_tab) {
    while (n_child > 1) {
        uint8_t* parent = level + 32 * n_child;
        const size_t n_parent = (n_child + 1) / 2;
        if (threadIdx.x < n_parent) node_hash_one(level, parent, n_child, threadIdx.x, pad_tab);
        __syncthreads();  // block-scope ordering of the global stores above with the loads of the next level
        level = parent;
        n_child = n_parent;
    }
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

// K[i] + W_b[i] of the padding block of a node hash, built once per device (sha256.cuh)
static int pad_table_get(const uint32_t** out) {
    static std::mutex mu;
    static uint32_t* tabs[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lk(mu);
    uint32_t*& t = tabs[dev & 63];
    if (!t) {
        e = cudaMalloc(&t, SHA_PAD_TAB_WORDS * sizeof(uint32_t));
        if (e != cudaSuccess) return (int)e;
        sha_pad_table_kernel<<<1, 256>>>(t, 65 * 8);
        e = cudaDeviceSynchronize();  // built on the legacy stream: visible to every stream from here on
        if (e != cudaSuccess) {
            cudaFree(t);
            t = nullptr;
            return (int)e;
        }
    }
    *out = t;
    return 0;
}

int merkle_upper_launches(size_t n) {  // kernel launches of merkle_upper_levels(n): the same walk, counting
    int launches = 0;
    while (n > (size_t)2 * TAIL_THREADS) {
        if ((n + 1) / 2 < ((size_t)1 << 18) && n % SUB_NODES == 0)
            n >>= SUB_LEVELS;
        else
            n = (n + 1) / 2;
        launches++;
    }
    return launches + (n > 1 ? 1 : 0);
}

int merkle_upper_levels(uint8_t* d_nodes, size_t n, cudaStream_t s) {
    const uint32_t* pad_tab = nullptr;
    if (n > 1) {
        int rc = pad_table_get(&pad_tab);
        if (rc) return rc;
    }
    uint8_t* cur = d_nodes;
    size_t cur_n = n;
    static int smem_ok[64] = {};  // per device: 0 unknown, 1 configured, -1 unavailable
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    dev &= 63;
    constexpr int TAB_BYTES = SHA_PAD_TAB_WORDS * 4;
    while (cur_n > (size_t)2 * TAIL_THREADS) {
        uint8_t* next = cur + 32 * cur_n;
        size_t next_n = (cur_n + 1) / 2;
        if (next_n >= ((size_t)1 << 18)) {
            if (smem_ok[dev] == 0)
                smem_ok[dev] = cudaFuncSetAttribute(node_hash_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_BYTES) == cudaSuccess ? 1 : -1;
            if (smem_ok[dev] == 1) {
                cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
                node_hash_smem_kernel<<<(unsigned)(3 * n_sm), 512, TAB_BYTES, s>>>(cur, next, cur_n, next_n, pad_tab);
                cur = next;
                cur_n = next_n;
                continue;
            }
        }
        if (cur_n % SUB_NODES == 0) {  // eight levels in one launch
            node_hash_subtree_kernel<<<(unsigned)(cur_n / SUB_NODES), SUB_THREADS, 0, s>>>(cur, cur_n, pad_tab);
            for (int l = 0; l < SUB_LEVELS; l++) {
                cur += 32 * cur_n;
                cur_n /= 2;
            }
            continue;
        }
        node_hash_kernel<<<(unsigned)((next_n + 255) / 256), 256, 0, s>>>(cur, next, cur_n, next_n, pad_tab);
        cur = next;
        cur_n = next_n;
    }
    if (cur_n > 1) node_hash_tail_kernel<<<1, TAIL_THREADS, 0, s>>>(cur, cur_n, pad_tab);
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
    return merkle_upper_levels(d_nodes, n, s);
}

int merkle_build_bytes(const uint8_t* d_leaves, size_t n, size_t leaf_len, uint8_t* d_nodes, cudaStream_t s) {
    if (n == 0) return (int)cudaErrorInvalidValue;
    leaf_hash_bytes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_leaves, n, leaf_len, d_nodes);
    int rc = (int)cudaGetLastError();
    if (rc) return rc;
    return merkle_upper_levels(d_nodes, n, s);
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
