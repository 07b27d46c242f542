// Explicit instantiations of the vectorised NTT pass kernel (split so nvcc runs in parallel).
#include "ntt_pass_v4.cuh"
namespace bb {
template void launch_pass_v4<0, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<0, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<0, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<0, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<1, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<1, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<1, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<1, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<2, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<2, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<2, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<2, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<3, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<3, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<3, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<3, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<4, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<4, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<4, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<4, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<5, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<5, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<5, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<5, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<6, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<6, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<6, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass_v4<6, 5>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
