// Element-wise stages of the Fibonacci prover and batched Merkle openings on the device (prover_ew.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bb {
// c[i] = (T[i+2 step] - T[i+step] - T[i]) (x_i - b1) (x_i - b2) over x_i = shift w_N^i, N = 2^log_n (src/fibonacci.rs:133-143)
int fib_constraint(const uint32_t* d_t, uint32_t* d_out, int log_n, uint32_t step, uint32_t shift, uint32_t b1, uint32_t b2,
                   cudaStream_t s);
// v[i] *= table[i mod period], period a power of two <= 64 (the 32 inverses of Z_H on the blowup-32 coset, :147-150)
int scale_periodic(uint32_t* d_v, size_t n, const uint32_t* h_table, uint32_t period, cudaStream_t s);
// DEEP composition (src/fibonacci.rs:186-198) with batched inversion of x_i - z
int fib_deep(const uint32_t* d_q, const uint32_t* d_t, uint32_t* d_out, int log_n, uint32_t step, uint32_t shift, uint32_t z,
             uint32_t q_z, uint32_t t_z, uint32_t t_gz, uint32_t t_ggz, cudaStream_t s);
// *d_acc = sum of the canonical terms of sum_k c[k] z^k (reduce mod p on the host)
int poly_eval(const uint32_t* d_c, size_t n, uint32_t z, unsigned long long* d_acc, cudaStream_t s);
// authentication paths of nq leaves in one launch; d_paths: nq * depth * 32 bytes
int merkle_gather_paths(const uint8_t* d_nodes, size_t nleaves, const unsigned long long* d_idx, size_t nq, uint32_t depth, uint8_t* d_paths,
                        cudaStream_t s);
int gather_elems(const void* d_src, uint32_t elem_bytes, const unsigned long long* d_idx, size_t nq, void* d_out, cudaStream_t s);
// one query of merkle_open_multi: tree, opened index, where its path goes (byte offset), leaf value and salt arrays
struct OpenQuery {
    const uint8_t* nodes;
    const uint8_t* vals;
    const uint8_t* salts;  // nullptr: unsalted tree
    unsigned long long nleaves, index, path_off;
};
// paths, values (val_bytes each) and salts (16 bytes, zero for unsalted trees) of nq queries over any number of trees
int merkle_open_multi(const OpenQuery* d_queries, size_t nq, uint32_t val_bytes, uint8_t* d_paths, uint8_t* d_vals, uint8_t* d_salts,
                      cudaStream_t s);
// dst[j*G + r] = src[r*chunk + j]: G runs of `chunk` elements (limbs words each) interleaved
int interleave(const uint32_t* d_src, uint32_t* d_dst, uint32_t groups, size_t chunk, uint32_t limbs, cudaStream_t s);
}  // namespace bb
