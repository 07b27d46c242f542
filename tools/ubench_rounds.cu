// Micro-benchmark of the in-shared-memory radix-16 rounds of the NTT pass kernel, isolated from global memory:
// every CTA keeps one tile resident and runs the rounds `iters` times.  Reports butterflies/clk/SM to compare with the
// 13.0 of tools/ubench_intpipe.cu.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I toyni_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "ntt_pass_v4.cuh"
using namespace bb;

template <int LR, int LC>
__global__ void __launch_bounds__(V4<LR, LC>::NT) rounds_kernel(PassParams p, int iters, uint32_t* sink) {
    using T = V4<LR, LC>;
    extern __shared__ uint4 smv[];
    uint2* stw = reinterpret_cast<uint2*>(smv + T::CHUNKS);
    for (uint32_t i = threadIdx.x; i < (uint32_t)(T::R / 2); i += T::NT) stw[i] = p.tw[i << (LOG_TW - LR)];
    for (int i = threadIdx.x; i < T::CHUNKS; i += T::NT) smv[i] = make_uint4(i * 7 + 1, i * 3 + 2, i + 5, i ^ 0x1234);
    __syncthreads();
    for (int it = 0; it < iters; it++) {
        if constexpr (T::G1 > 0) dit_round_v4<LR, LC, 0, T::G1>(smv, stw, p);
        if constexpr (T::G2 > 0) { __syncthreads(); dit_round_v4<LR, LC, 4, T::G2>(smv, stw, p); }
        if constexpr (T::G3 > 0) { __syncthreads(); dit_round_v4<LR, LC, 8, T::G3>(smv, stw, p); }
        __syncthreads();
    }
    if (threadIdx.x == 0) sink[blockIdx.x] = smv[blockIdx.x % T::CHUNKS].x;
}

template <int LR, int LC>
void run(const PassParams& p, int sms, double mhz, uint32_t* sink) {
    using T = V4<LR, LC>;
    size_t smem = T::SMEM + T::TW_BYTES;
    cudaFuncSetAttribute(rounds_kernel<LR, LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rounds_kernel<LR, LC>, T::NT, smem);
    for (int target = 1; target <= occ; target = (target < occ && target * 2 > occ) ? occ : target * 2) {
        int blocks = sms * target, iters = 200;
        rounds_kernel<LR, LC><<<blocks, T::NT, smem>>>(p, 10, sink);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        rounds_kernel<LR, LC><<<blocks, T::NT, smem>>>(p, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bfly = (double)blocks * iters * (double)T::R * T::C * LR / 2.0;
        printf("LR=%2d LC=%d VW=%d NT=%3d smem=%6zu  CTAs/SM=%d (%2d warps/SM): %7.3f ms  %6.2f bfly/clk/SM  (%4.1f%% of 13.0)\n", LR, LC, T::VW,
               T::NT, smem, target, target * T::NT / 32, ms, bfly / (ms * 1e-3) / sms / (mhz * 1e6), 100.0 * bfly / (ms * 1e-3) / sms / (mhz * 1e6) / 13.0);
        if (target == occ) break;
    }
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount, khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1000.0;
    uint2* tw; cudaMalloc(&tw, 2048 * sizeof(uint2));
    uint2 h[2048]; for (int i = 0; i < 2048; i++) { uint32_t w = pow(root_of_unity(12), i); h[i] = make_uint2(w, shoup_companion(w)); }
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    uint32_t* sink; cudaMalloc(&sink, 1 << 20);
    PassParams p{}; p.tw = tw;
    uint32_t w16 = root_of_unity(4), cur = 1;
    for (int i = 0; i < 8; i++) { p.tw16[i] = make_uint2(cur, shoup_companion(cur)); cur = mul(cur, w16); }
    run<8, 4>(p, sms, mhz, sink);
    run<8, 5>(p, sms, mhz, sink);
    run<8, 3>(p, sms, mhz, sink);
    run<12, 3>(p, sms, mhz, sink);
    run<4, 5>(p, sms, mhz, sink);
    return 0;
}
