// Explicit instantiations of the vectorised NTT pass kernel (split so nvcc runs in parallel).
#include "ntt_pass_v4.cuh"
namespace bb {
template void launch_pass_v4<7, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<7, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<7, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<7, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<8, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<8, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<8, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<8, 5>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
