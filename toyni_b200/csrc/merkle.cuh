#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bb {
// number of 32-byte digests over all levels (odd levels pair their last node with itself, src/merkle.rs:36-43)
size_t merkle_node_count(size_t nleaves);
// Salted / unsalted commit of device-resident field values (src/fibonacci.rs:340-363 + src/merkle.rs:25-48).
//   d_vals : n * limbs canonical u32 values (limbs = 1 base field, 4 extension field)
//   d_salts: n * 16 bytes, or nullptr for the unsalted form (leaf = value bytes only)
//   d_nodes: merkle_node_count(n) * 32 bytes, every level, leaf level first
int merkle_commit(const uint32_t* d_vals, int limbs, size_t n, const uint8_t* d_salts, uint8_t* d_nodes, cudaStream_t s);
// Levels above the leaf digests (already in d_nodes[0 .. 32 n)).
int merkle_upper_levels(uint8_t* d_nodes, size_t n, cudaStream_t s);
int merkle_upper_launches(size_t n);  // how many kernels merkle_upper_levels(n) launches
// Generic leaves: n byte strings of leaf_len bytes (any length), back to back (MerkleTree::new, src/merkle.rs:16-23).
int merkle_build_bytes(const uint8_t* d_leaves, size_t n, size_t leaf_len, uint8_t* d_nodes, cudaStream_t s);
// Gather the authentication path of `index` (src/merkle.rs:50-80) into d_path (depth * 32 bytes); pos bits on host.
int merkle_open(const uint8_t* d_nodes, size_t nleaves, size_t index, uint8_t* d_path, uint8_t* h_pos, size_t* depth,
                cudaStream_t s);
}  // namespace bb
