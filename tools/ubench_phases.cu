// Where does the NTT pass kernel lose throughput relative to its isolated rounds?  Builds the tile pipeline up phase by
// phase on real-sized data (2^24 words in, 2^24 out) and reports microseconds per "pass" (4096 tiles of 256 x 16):
//   0: rounds only on a resident tile            3: loads (double buffered) + rounds
//   1: rounds + epilogue arithmetic (no store)   4: loads + rounds + row stores (EPI_NONE)
//   2: rounds + row stores (EPI_NONE)            5: the real kernel (EPI_NONE)
#include <cstdio>
#include <cuda_runtime.h>
#include "ntt_pass_v4.cuh"
using namespace bb;

template <int LR, int LC, int MODE>
__global__ void __launch_bounds__(V4<LR, LC>::NT) phase_kernel(PassParams p, uint32_t tiles_x, uint32_t total_tiles) {
    using T = V4<LR, LC>;
    constexpr int C = T::C;
    extern __shared__ uint4 smv_all[];
    uint2* stw = reinterpret_cast<uint2*>(smv_all + 2 * T::CHUNKS);
    for (uint32_t i = threadIdx.x; i < (uint32_t)(T::R / 2); i += T::NT) stw[i] = p.tw[i << (LOG_TW - LR)];
    for (int i = threadIdx.x; i < 2 * T::CHUNKS; i += T::NT) smv_all[i] = make_uint4(i * 7 + 1, i * 3 + 2, i + 5, i ^ 0x1234);
    __syncthreads();
    uint32_t tile = blockIdx.x;
    constexpr bool LOADS = (MODE == 3 || MODE == 4);
    constexpr bool STORES = (MODE == 2 || MODE == 4);
    if (LOADS) {
        load_tile_v4<LR, LC>(smv_all, p.in, p, (tile % tiles_x) * C);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    for (uint32_t iter = 0; tile < total_tiles; tile += gridDim.x, iter++) {
        uint4* smv = smv_all + ((iter & 1) ? T::CHUNKS : 0);
        const uint32_t col0 = (tile % tiles_x) * C;
        if (LOADS) {
            const uint32_t nt = tile + gridDim.x;
            if (nt < total_tiles) load_tile_v4<LR, LC>(smv_all + ((iter & 1) ? 0 : T::CHUNKS), p.in, p, (nt % tiles_x) * C);
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;\n" ::: "memory");
        }
        __syncthreads();
        dit_round_v4<LR, LC, 0, T::G1>(smv, stw, p);
        __syncthreads();
        dit_round_v4<LR, LC, 4, T::G2>(smv, stw, p);
        __syncthreads();
        if (MODE == 1) {  // epilogue arithmetic of a twiddle pass, results kept in shared memory
            PassParams q = p;
            for (int i = threadIdx.x; i < T::CHUNKS; i += T::NT) {
                uint4 v = smv[i];
                v = mul4(v, p.epi_const + i);
                smv[i] = v;
            }
        }
        if (STORES) store_rows_v4<LR, LC, EPI_NONE>(smv, p.out, p, col0);
        __syncthreads();
    }
}

template <int MODE>
float run(const PassParams& p, uint32_t tiles_x, uint32_t total, int ctas) {
    using T = V4<8, 4>;
    size_t smem = 2 * T::SMEM + T::TW_BYTES;
    cudaFuncSetAttribute(phase_kernel<8, 4, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int w = 0; w < 2; w++) phase_kernel<8, 4, MODE><<<ctas, T::NT, smem>>>(p, tiles_x, total);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; r++) phase_kernel<8, 4, MODE><<<ctas, T::NT, smem>>>(p, tiles_x, total);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / reps;
}

int main() {
    using T = V4<8, 4>;
    const uint32_t log_n = 24, n = 1u << log_n;
    uint32_t *in, *out; uint2* tw;
    cudaMalloc(&in, (size_t)n * 4); cudaMalloc(&out, (size_t)n * 4); cudaMalloc(&tw, 2048 * sizeof(uint2));
    cudaMemset(in, 1, (size_t)n * 4);
    uint2 h[2048]; for (int i = 0; i < 2048; i++) { uint32_t w = pow(root_of_unity(12), i); h[i] = make_uint2(w, shoup_companion(w)); }
    cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
    PassParams p{}; p.tw = tw; p.in = in; p.out = out;
    p.ncols = n >> 8; p.log_pfull = 16; p.n_in_limit = ~0ull; p.epi_const = 12345;  // a last-pass-like layout: rows of 64 B
    uint32_t w16 = root_of_unity(4), cur = 1;
    for (int i = 0; i < 8; i++) { p.tw16[i] = make_uint2(cur, shoup_companion(cur)); cur = mul(cur, w16); }
    const uint32_t tiles_x = p.ncols / T::C, total = tiles_x;
    int occ = 0; size_t smem = 2 * T::SMEM + T::TW_BYTES;
    cudaFuncSetAttribute(phase_kernel<8, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phase_kernel<8, 4, 4>, T::NT, smem);
    int ctas = occ * 148;
    printf("2^24 words, %u tiles of 256x16, %d CTAs (%d per SM), one pass each:\n", total, ctas, occ);
    printf("  rounds only                      %7.1f us\n", run<0>(p, tiles_x, total, ctas));
    printf("  rounds + epilogue multiplies     %7.1f us\n", run<1>(p, tiles_x, total, ctas));
    printf("  rounds + row stores              %7.1f us\n", run<2>(p, tiles_x, total, ctas));
    printf("  prefetched loads + rounds        %7.1f us\n", run<3>(p, tiles_x, total, ctas));
    printf("  loads + rounds + stores          %7.1f us\n", run<4>(p, tiles_x, total, ctas));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
