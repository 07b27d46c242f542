// Internal bridge between c_abi.cu (library state: sticky error, per-thread stream, launch counter) and mg.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ntt_engine.cuh"

namespace bb {
namespace abi {
int note_error(int rc);            // records the first error of this host thread, returns rc
cudaStream_t get_stream();         // stream of this host thread's calls (bb_set_stream)
void set_stream(cudaStream_t s);
void count_launches(unsigned n);
// stream-ordered NTT on the current device and stream (run_ntt of c_abi.cu without a coset shift)
int ntt(const uint32_t* in, uint32_t* out, uint32_t log_n, int log_inner, size_t n_in, size_t batch, int dir, const FourStepScatter* scatter);
}  // namespace abi
}  // namespace bb
