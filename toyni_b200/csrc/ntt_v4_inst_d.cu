// Explicit instantiations of the vectorised NTT pass kernel (split so nvcc runs in parallel).
#include "ntt_pass_v4.cuh"
namespace bb {
template void launch_pass_v4<11, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<11, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<11, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<12, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<12, 3>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
