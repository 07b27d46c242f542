// Warp-private NTT pass, wide strips (R = 256, 16 columns): the same scheme as ntt_pass_v5.cuh, but a warp owns 256 rows
// x 64 bytes (16 KB) and works through it as two register blocks of 16 rows x 4 columns per round.
//
// Why wider rows: tools/ubench_strided.cu shows that a pass is bounded by its global-memory access shape as much as
// by its arithmetic - moving 64 MB in and 64 MB out as 32-byte row segments takes 37 us on a B200, as 64-byte
// segments 30 us, as a plain stream 27 us (the arithmetic of a pass takes about 27 us).  64-byte rows need 16 KB per
// warp, so there are 8 warps per SM (two per scheduler) instead of 16; the two warps of a "pair" take the two strips
// that share each 128-byte line.
//
// Layout of a strip in shared memory (16-byte chunk units): row r, chunk c at
//     ((r ^ ((r >> 4) & 1)) << 2) | (c ^ ((r >> 1) & 3))
// which makes every LDS.128 / STS.128 quarter-warp (2 row groups x 4 chunks, or 8 consecutive rows x 1 chunk in the
// transposing first pass) and every cp.async quarter-warp touch eight different 16-byte bank groups.
#pragma once
#include "ntt_pass_v5.cuh"

namespace bb {

// warps per SM: 12 (three per scheduler) where shared memory allows, 8 for the first pass, whose 32 KB table of
// column-dependent twiddles takes the room of two strips
__host__ __device__ constexpr int v6_warps(int mode) { return mode == V5_COLS_TWIDDLE ? 8 : 12; }
constexpr int V6_STRIP_U4 = V5_R * 4;  // 16 KB

__host__ __device__ constexpr size_t v6_smem_bytes(int mode) {
    return (size_t)V5_TW_ENTRIES * 8 + (mode == V5_COLS_TWIDDLE ? (size_t)V5_R * 16 * 8 : 0) +
           (size_t)v6_warps(mode) * (V6_STRIP_U4 * 16 + (mode != V5_ROWS_CANON ? V5_R * 8 : 0));
}

#ifdef __CUDACC__
BB_D uint32_t v6_phys(uint32_t r, uint32_t c) { return ((r ^ ((r >> 4) & 1u)) << 2) | (c ^ ((r >> 1) & 3u)); }

template <int MODE>
__global__ void __launch_bounds__(v6_warps(MODE) * 32, 1)
    ntt_pass_v6_kernel(const PassParams p, const uint2* __restrict__ btab, uint32_t strips_x, uint32_t total_strips) {
    constexpr int R = V5_R;
    constexpr int V6_WARPS = v6_warps(MODE), V6_THREADS = V6_WARPS * 32;
    extern __shared__ __align__(128) uint4 smem6[];
    uint2* const stw = reinterpret_cast<uint2*>(smem6);
    uint4* const sB = smem6 + V5_TW_ENTRIES / 2;
    constexpr int B_U4 = (MODE == V5_COLS_TWIDDLE) ? R * 8 : 0;  // 16 pairs = 8 chunks per row
    constexpr int WARP_U4 = V6_STRIP_U4 + (MODE != V5_ROWS_CANON ? R / 2 : 0);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* const strip = sB + B_U4 + warp * WARP_U4;
    uint2* const sA = reinterpret_cast<uint2*>(strip + V6_STRIP_U4);

    // Work assignment: pair q = warps 2q, 2q+1 (schedulers 0,1 for even q, 2,3 for odd q) walks a contiguous range
    // of line groups (two strips sharing each 128-byte line).  The CTA's groups are split evenly over the pairs, the
    // remainder going to pairs 0, 1, 2, ... so that the two halves of the schedulers carry the same load (+-1 group).
    const uint32_t pair = warp >> 1, pi = warp & 1u;
    const uint32_t lg_total = total_strips >> 1;
    const uint32_t cta_lo = (uint32_t)(((unsigned long long)blockIdx.x * lg_total) / gridDim.x);
    const uint32_t cta_hi = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * lg_total) / gridDim.x);
    constexpr uint32_t NP = V6_WARPS / 2;
    const uint32_t cnt = cta_hi - cta_lo, basep = cnt / NP, rem = cnt - basep * NP;
    const uint32_t lg_lo = cta_lo + pair * basep + (pair < rem ? pair : rem);
    const uint32_t lg_hi = lg_lo + basep + (pair < rem ? 1u : 0u);
    uint32_t s = 2u * lg_lo + pi;
    const uint32_t s_end = 2u * lg_hi;
    const bool active = s < s_end;

    const uint32_t strip_sa = (uint32_t)__cvta_generic_to_shared(strip);
    const uint32_t ncols = p.ncols;

    // load: lane = c + 4 dr; data row d = 128 (dr & 1) + (dr >> 1) + 4 it (it < 32) lands at logical row brev8(d)
    const uint32_t lc = lane & 3u, dr = lane >> 2;
    const uint32_t d_lane = 128u * (dr & 1u) + (dr >> 1);
    // rounds: lane = c + 4 g; round 1 row group blk = g + 8 hf, round 2 (row stores) b = g + 8 hf; transposing stores:
    // lane = b + 16 c', chunk c = c' + 2 hf
    const uint32_t rc = lane & 3u, rg = lane >> 2;

    uint32_t bz = s / strips_x, sx = s - bz * strips_x;
    uint32_t nbz = bz, nsx = sx;
    auto issue_load = [&]() {
        const uint32_t* src = p.in + (size_t)nbz * p.in_batch_stride + (size_t)d_lane * ncols + nsx * 16u + 4u * lc;
        const size_t step = (size_t)4 * ncols;
#pragma unroll
        for (int it = 0; it < 32; it++) {
            const uint32_t d = d_lane + 4u * (uint32_t)it;
            const uint32_t r = __brev(d) >> 24;
            cp_async16(strip_sa + (v6_phys(r, lc) << 4), src);
            src += step;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    if (active) issue_load();  // in flight while the tables are set up

    for (uint32_t i = threadIdx.x; i < (uint32_t)V5_TW_ENTRIES; i += V6_THREADS) {
        const uint32_t t = (i >= 112u) ? 3u : (i >= 48u) ? 2u : (i >= 16u) ? 1u : 0u;
        const uint32_t m = i - 16u * ((1u << t) - 1u);
        stw[i] = __ldg(&p.tw[(m << (3u - t)) << (LOG_TW - V5_LR)]);
    }
    if (MODE == V5_COLS_TWIDDLE) {
        // btab[e*16 + c] (Shoup pairs), two pairs per chunk: chunk (e, q) at e*8 + (q ^ (e & 7))
        for (uint32_t i = threadIdx.x; i < (uint32_t)(R * 8); i += V6_THREADS) {
            const uint32_t e = i >> 3, q = i & 7u;
            sB[(e << 3) + (q ^ (e & 7u))] = __ldg(reinterpret_cast<const uint4*>(btab) + i);
        }
    }
    __syncthreads();

    if (!active) return;
    uint32_t cur_j = 0xFFFFFFFFu;
    const uint32_t log_pfull = p.log_pfull;

    for (; s < s_end; s += 2) {
        const uint32_t col0 = sx * 16u;
        __syncwarp();
        if (MODE != V5_ROWS_CANON) {
            const uint32_t jj = (MODE == V5_COLS_TWIDDLE) ? col0 : (col0 >> log_pfull);
            if (MODE == V5_COLS_TWIDDLE || jj != cur_j) {
                cur_j = jj;
#pragma unroll
                for (int i = 0; i < R / 32; i++) {
                    const uint32_t e = lane + 32u * i;
                    const uint32_t w = pow_plain(p.epi, (jj * e) << p.epi_shift);
                    sA[e] = make_uint2(w, shoup_companion_fast(w));
                }
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();

        uint32_t x[16][4];
        // ---- round 1: rows blk*16 + k
#pragma unroll 1
        for (uint32_t hf = 0; hf < 2; hf++) {
            const uint32_t blk = rg + 8u * hf;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint4 v = strip[v6_phys(blk * 16u + (uint32_t)k, rc)];
                x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k & (1 << t)) continue;
                    const int kp = k & ((1 << t) - 1);
                    if (kp == 0) {
#pragma unroll
                        for (int c = 0; c < 4; c++) bfly_one(x[k][c], x[k + (1 << t)][c]);
                    } else {
                        const uint2 w = p.tw16[kp << (3 - t)];
#pragma unroll
                        for (int c = 0; c < 4; c++) bfly(x[k][c], x[k + (1 << t)][c], w);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++) strip[v6_phys(blk * 16u + (uint32_t)k, rc)] = make_uint4(x[k][0], x[k][1], x[k][2], x[k][3]);
        }
        __syncwarp();

        // ---- round 2: rows b + 16 k, then epilogue + store straight from registers
#pragma unroll 1
        for (uint32_t hf = 0; hf < 2; hf++) {
            const uint32_t b = (MODE == V5_COLS_TWIDDLE) ? (lane & 15u) : (rg + 8u * hf);
            const uint32_t c = (MODE == V5_COLS_TWIDDLE) ? ((lane >> 4) + 2u * hf) : rc;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint4 v = strip[v6_phys(b + 16u * (uint32_t)k, c)];
                x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint2* tws = stw + 16 * ((1 << t) - 1) + b;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    if (k & (1 << t)) continue;
                    const int kp = k & ((1 << t) - 1);
                    const uint2 w = tws[16 * kp];
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) bfly(x[k][cc], x[k + (1 << t)][cc], w);
                }
                if (t == 0 && hf == 1) {
                    // every lane has consumed its last loads: the strip buffer is free for the next strip
                    __syncwarp();
                    nsx += 2;
                    if (nsx >= strips_x) {
                        nsx -= strips_x;
                        nbz++;
                    }
                    if (s + 2 < s_end) issue_load();
                }
            }

            if (MODE == V5_COLS_TWIDDLE) {
                uint32_t* o = p.out + (size_t)bz * p.out_batch_stride + ((size_t)(col0 + 4u * c) << V5_LR) + b;
                const uint4* bq0 = sB + (b << 3) + ((2u * c) ^ (b & 7u));
                const uint4* bq1 = sB + (b << 3) + ((2u * c + 1u) ^ (b & 7u));
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint2 a = sA[b + 16 * k];
                    const uint4 b01 = bq0[128 * k], b23 = bq1[128 * k];  // row b + 16 k: (b + 16 k) & 7 == b & 7
                    o[16 * k] = shoup_mul_lazy(shoup_mul_lazy(x[k][0], b01.x, b01.y), a.x, a.y);
                    o[16 * k + R] = shoup_mul_lazy(shoup_mul_lazy(x[k][1], b01.z, b01.w), a.x, a.y);
                    o[16 * k + 2 * R] = shoup_mul_lazy(shoup_mul_lazy(x[k][2], b23.x, b23.y), a.x, a.y);
                    o[16 * k + 3 * R] = shoup_mul_lazy(shoup_mul_lazy(x[k][3], b23.z, b23.w), a.x, a.y);
                }
            } else {
                const uint32_t j = col0 >> log_pfull, low0 = col0 & ((1u << log_pfull) - 1u);
                uint32_t* o = p.out + (size_t)bz * p.out_batch_stride + ((((size_t)j << V5_LR) + b) << log_pfull) + low0 + 4u * c;
                const size_t step = (size_t)16 << log_pfull;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    uint4 v;
                    if (MODE == V5_ROWS_TWIDDLE) {
                        const uint2 a = sA[b + 16 * k];
                        v.x = shoup_mul_lazy(x[k][0], a.x, a.y);
                        v.y = shoup_mul_lazy(x[k][1], a.x, a.y);
                        v.z = shoup_mul_lazy(x[k][2], a.x, a.y);
                        v.w = shoup_mul_lazy(x[k][3], a.x, a.y);
                    } else {
                        v.x = min(x[k][0], x[k][0] - P);
                        v.y = min(x[k][1], x[k][1] - P);
                        v.z = min(x[k][2], x[k][2] - P);
                        v.w = min(x[k][3], x[k][3] - P);
                    }
                    *reinterpret_cast<uint4*>(o) = v;
                    o += step;
                }
            }
        }
        bz = nbz;
        sx = nsx;
    }
}

// strips_x = ncols / 16 strips per vector (even); `batch` vectors
template <int MODE>
int launch_pass_v6(const PassParams& p, const uint2* btab, uint32_t strips_x, uint32_t batch, cudaStream_t s) {
    static int n_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    constexpr size_t smem = v6_smem_bytes(MODE);
    if (n_sm[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_v6_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    const uint32_t total = strips_x * batch;
    uint32_t ctas = (uint32_t)n_sm[dev];
    if (ctas > total / 2u) ctas = total / 2u;
    ntt_pass_v6_kernel<MODE><<<ctas, v6_warps(MODE) * 32, smem, s>>>(p, btab, strips_x, total);
    return (int)cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace bb
