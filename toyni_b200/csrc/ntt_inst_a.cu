// Explicit instantiations of the NTT pass kernel (split across files so nvcc runs in parallel).
#include "ntt_pass.cuh"
namespace bb {
template void launch_pass<0, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<0, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<0, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<0, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<0, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<1, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<1, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<1, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<1, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<1, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<2, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<2, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<2, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<2, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<2, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<3, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<3, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<3, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<3, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<3, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<4, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<4, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<4, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<4, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<4, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<5, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<5, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<5, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<5, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<5, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<6, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<6, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<6, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<6, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<6, 5>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
