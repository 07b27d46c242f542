// Launcher of the TMA-staged 4096-row NTT pass (ntt_pass_v7.cuh).
#pragma once
#include "ntt_pass_v7.cuh"

namespace bb {
// One pass over `batch` row-major [4096][ncols] matrices at `in`: 4096-point transforms down the columns.
//   pass2 = false: no input twiddle, transposing store out[col * 4096 + e] (values left in [0, 2p))
//   pass2 = true : input (d, col) multiplied by w_n^(d * col) (tables in p), canonical row store
// p.tiles_x / p.total_tiles are filled in here.  Returns a cudaError_t value.
int launch_pass_v7(bool pass2, const uint32_t* in, size_t ncols, size_t batch, size_t in_batch_stride, V7Params p, bool pdl, cudaStream_t s);
}  // namespace bb
