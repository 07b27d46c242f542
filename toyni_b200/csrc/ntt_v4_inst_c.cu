// Explicit instantiations of the vectorised NTT pass kernel (split so nvcc runs in parallel).
#include "ntt_pass_v4.cuh"
namespace bb {
template void launch_pass_v4<9, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<9, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<9, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<9, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<10, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<10, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<10, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<10, 5>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
