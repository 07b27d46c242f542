// Explicit instantiations of the NTT pass kernel (split across files so nvcc runs in parallel).
#include "ntt_pass.cuh"
namespace bb {
template void launch_pass<12, 0>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<12, 2>(const PassParams&, dim3, cudaStream_t);
template void launch_pass<12, 3>(const PassParams&, dim3, cudaStream_t);
}  // namespace bb
