// Explicit instantiations of the warp-private NTT pass kernel (R = 256).
#include "ntt_pass_v5.cuh"
namespace bb {
template int launch_pass_v5<V5_ROWS_CANON>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
template int launch_pass_v5<V5_ROWS_TWIDDLE>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
template int launch_pass_v5<V5_COLS_TWIDDLE>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
}  // namespace bb
