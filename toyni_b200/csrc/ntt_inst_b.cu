// Explicit instantiations of the NTT pass kernel (split across files so nvcc runs in parallel).
#include "ntt_pass.cuh"
namespace bb {
template void launch_pass<7, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<7, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<7, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<7, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<7, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<8, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<8, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<8, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<8, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<8, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<9, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<9, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<9, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<9, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<9, 5>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
