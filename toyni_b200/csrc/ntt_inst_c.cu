// Explicit instantiations of the NTT pass kernel (split across files so nvcc runs in parallel).
#include "ntt_pass.cuh"
namespace bb {
template void launch_pass<10, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<10, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<10, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<10, 4>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<10, 5>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<11, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<11, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<11, 3>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<11, 4>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
