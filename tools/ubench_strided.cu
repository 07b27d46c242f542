// Micro-benchmark of the global-memory access shape of an NTT pass.  A 2^24-word array is seen as rows of 65536 words
// (row stride 256 KB).  Every warp moves 8 KB units of (8 KB / W) rows x W bytes through its own shared-memory buffer
// (cp.async in, STG.128 out), consecutive warps taking horizontally adjacent units, 16 warps per SM - the traffic of
// round 1's warp-private strip kernel (removed in round 2) without the arithmetic.  W = 32 B is the kernel's strip (256 rows); larger W shows what wider row
// segments would buy; W = 8 KB is a plain streaming copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/ubench_strided.cu -o tools/ubench_strided.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>  // 0 read, 1 write, 2 copy
__global__ void __launch_bounds__(512) strided_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t log_w16,
                                                       uint32_t row_stride16, uint32_t units, uint32_t* sink) {
    extern __shared__ uint4 sm[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp, nw = gridDim.x * (blockDim.x >> 5);
    uint4* buf = sm + warp * 512;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(buf);
    const uint32_t segs_per_row = row_stride16 >> log_w16;  // units side by side in one row block
    const uint32_t rows_per_unit = 512u >> log_w16;
    for (uint32_t u = gw; u < units; u += nw) {
        const uint32_t rb = u / segs_per_row, sg = u - rb * segs_per_row;
        const size_t base = (size_t)rb * rows_per_unit * row_stride16 + ((size_t)sg << log_w16);
        if (MODE != 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t c = lane + 32u * i, r = c >> log_w16, cc = c & ((1u << log_w16) - 1u);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa + c * 16u), "l"(in + base + (size_t)r * row_stride16 + cc) : "memory");
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
            __syncwarp();
        }
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t c = lane + 32u * i, r = c >> log_w16, cc = c & ((1u << log_w16) - 1u);
                out[base + (size_t)r * row_stride16 + cc] = (MODE == 2) ? buf[c] : make_uint4(u, c, lane, i);
            }
            __syncwarp();
        }
    }
    if (MODE == 0 && buf[lane].x == 0x12345678u) sink[0] = 1;
}

int main() {
    const size_t words = 1u << 24;
    uint4 *a, *b; uint32_t* sink;
    cudaMalloc(&a, words * 4 * 4); cudaMalloc(&b, words * 4 * 4); cudaMalloc(&sink, 64);  // 4 rotating 64 MiB buffers each
    cudaMemset(a, 1, words * 16); cudaMemset(b, 2, words * 16);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const uint32_t row_stride16 = 65536 / 4, units = (uint32_t)(words * 4 / 8192);
    const char* names[3] = {"read ", "write", "copy "};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int smem = 16 * 8192;
    cudaFuncSetAttribute(strided_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(strided_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(strided_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int resident = 0; resident < 2; resident++)
    for (uint32_t log_w16 = 1; log_w16 <= 9; log_w16 += (log_w16 >= 3 ? 3 : 1)) {
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e9f;
            for (int rep = 0; rep < 14; rep++) {
                // resident: the same 64 MiB buffer every time, updated in place (stays in the 126 MB L2)
                const uint4* in = a + (size_t)(resident ? 0 : (rep & 3)) * (words / 4);
                uint4* out = resident ? a : b + (size_t)(rep & 3) * (words / 4);
                cudaEventRecord(e0);
                if (mode == 0) strided_kernel<0><<<sms, 512, smem>>>(in, out, log_w16, row_stride16, units, sink);
                if (mode == 1) strided_kernel<1><<<sms, 512, smem>>>(in, out, log_w16, row_stride16, units, sink);
                if (mode == 2) strided_kernel<2><<<sms, 512, smem>>>(in, out, log_w16, row_stride16, units, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep >= 2 && ms < best) best = ms;
            }
            const double bytes = (double)words * 4 * (mode == 2 ? 2 : 1);
            printf("%s row segment %5u B x %3u rows  %s  %6.1f us  %7.1f GB/s\n", resident ? "L2-resident (in place)" : "HBM (rotating buffers) ", 16u << log_w16, 512u >> log_w16, names[mode], best * 1e3, bytes / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
