// Explicit instantiations of the wide-strip warp-private NTT pass kernel (R = 256, 16 columns).
#include "ntt_pass_v6.cuh"
namespace bb {
template int launch_pass_v6<V5_ROWS_CANON>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
template int launch_pass_v6<V5_ROWS_TWIDDLE>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
template int launch_pass_v6<V5_COLS_TWIDDLE>(const PassParams&, const uint2*, uint32_t, uint32_t, cudaStream_t);
}  // namespace bb
