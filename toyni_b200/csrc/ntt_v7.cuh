// Launcher of the TMA-staged 4096-row NTT pass (ntt_pass_v7.cuh).
#pragma once
#include "ntt_pass_v7.cuh"

namespace bb {
// One pass over `batch` row-major [4096][ncols] matrices at `in`: 4096-point transforms down the columns.
//   pass2 = false: no input twiddle, transposing store out[col * 4096 + e] (values left in [0, 2p))
//   pass2 = true : input (d, col) multiplied by w_n^(d * col) (tables in p), canonical row store
// p.tiles_x / p.total_tiles are filled in here.  Returns a cudaError_t value.
int launch_pass_v7(bool pass2, const uint32_t* in, size_t ncols, size_t batch, size_t in_batch_stride, V7Params p, bool pdl, cudaStream_t s);
struct LdeParams {
    const uint32_t* in;    // n_coeffs canonical coefficients
    uint32_t n_coeffs;     // <= 2^21
    uint32_t* out;         // Z[j][k1]: 4096 rows of 8192
    const uint2* tw;       // (w, w') of omega_8192^i, i < 4096
    PowTable shift;        // shift^i, i < 2^25 (read when has_shift)
    uint32_t has_shift;
    uint32_t total_tiles;  // 1024
};

struct LDE {
    static constexpr int NT = 512;
    static constexpr int LOG_ROWS = 13, LOG_COLS = 12;  // 8192 x 4096
    static constexpr uint32_t OFF_TILE = 0;                  // uint4 [16 a'][16 b][32 c]
    static constexpr uint32_t OFF_TB = 131072;               // round-B twiddles [15][16 a'][32 c]: lanes read consecutive entries
    static constexpr uint32_t OFF_TA = OFF_TB + 15 * 512 * 8;  // round-A twiddles [15][32 c]
    static constexpr uint32_t OFF_W32 = OFF_TA + 15 * 32 * 8;  // omega_32^c
    static constexpr uint32_t OFF_IN = OFF_W32 + 32 * 8;     // 2 x uint4 [512 rows]
    static constexpr uint32_t SMEM = OFF_IN + 2 * 512 * 16;
};

// First pass of the blowup-32 LDE 2^20 -> 2^25 (lde_expand.cuh, compiled in ntt_v7.cu): Z[j][k1] for the 4096 x 8192 intermediate array.
int launch_lde_expand(LdeParams p, bool pdl, cudaStream_t s);
}  // namespace bb
