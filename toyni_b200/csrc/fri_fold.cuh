#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ntt_engine.cuh"

namespace bb {
// Fold a layer whose evaluation points are x_i = x0 * omega_m^i (m = 2^log_m_global).  `d_in` holds m_local
// values of `limbs` u32 each; local index t stands for global index t*idx_mul + idx_add (1, 0 on one GPU;
// G, rank for the cyclic multi-GPU layout).  Pairs are (t, t + m_local/2).
// hash_mode: 0 fold only; 1 / 2 also write the unsalted / salted leaf digest of every new value to d_leaf_nodes
// (the leaf level of the next layer's Merkle tree) from the same kernel.
int fri_fold_coset(const uint32_t* d_in, uint32_t* d_out, size_t m_local, int limbs, int log_m_global, uint32_t x0,
                   const uint32_t beta[4], uint32_t idx_mul, uint32_t idx_add, cudaStream_t s, int hash_mode = 0,
                   const uint8_t* d_salts = nullptr, uint8_t* d_leaf_nodes = nullptr);
// The last `nfolds` folds of a chain in ONE single-CTA launch (layers of at most 2^13 local values): layer f reads the output
// of layer f-1, x0 is squared from fold to fold, betas holds `limbs` values per fold, outputs land back to back in d_out.
bool fri_fold_tail_applies(size_t m_local, size_t nfolds);
int fri_fold_chain_tail(const uint32_t* d_in, uint32_t* d_out, size_t m_local, int limbs, int log_m_global, uint32_t x0,
                        const uint32_t* betas, size_t nfolds, uint32_t idx_mul, uint32_t idx_add, cudaStream_t s);
// Reference signature: arbitrary evaluation points xs[0..m/2) on the device.
int fri_fold_xs(const uint32_t* d_in, const uint32_t* d_xs, uint32_t* d_out, size_t m, int limbs, const uint32_t beta[4],
                cudaStream_t s);
// four-step NTT, step 2: d[k1][c] *= w_n^((col_offset + c) * k1) over a local block of 2^log_n1 rows x cols columns
int fourstep_twiddle(uint32_t* d, int log_n, int log_n1, size_t cols, size_t col_offset, bool inverse, cudaStream_t s);
// device-side rendezvous over peer-mapped flag words (d_peer_flags: device array of nranks pointers)
int peer_signal(uint32_t* const* d_peer_flags, uint32_t nranks, uint32_t rank, uint32_t epoch, cudaStream_t s);
int peer_wait(uint32_t* d_flags, uint32_t nranks, uint32_t epoch, uint32_t* d_err, cudaStream_t s);
// cached g^t tables (shared with the NTT engine)
int engine_pow_table(uint32_t g, int log_total, uint32_t scale, PowTable* out);
}  // namespace bb
