// Warp-private NTT pass (R = 256): the hot kernel of every large transform.
//
// Same mathematics and the same PassParams contract as ntt_pass.cuh / ntt_pass_v4.cuh (read in[d*ncols + col], R-point
// DFT over d, inter-pass twiddle, write out[((j*R + e) << log_pfull) + low]), restructured around ONE WARP per work
// unit so that nothing in the steady state is CTA-wide:
//   * a warp owns a "strip" of 256 rows x 8 columns (one 32-byte sector per row, 8 KB of shared memory); the four
//     warps of a quad take the four strips that share each 128-byte line at the same time, and the quad walks a
//     contiguous range of such line groups; warps never wait for each other (no __syncthreads after the table set-up), so the
//     load latency of one warp is covered by the arithmetic of the other 13 on the SM;
//   * every lane holds a 16 x 4 block (16 rows, one 16-byte chunk of four columns) = 64 values in registers per
//     radix-16 round: one LDS.128 / STS.128 per four values, 32 independent butterflies per stage per lane, and the
//     twiddles of a lane are shared by its four columns;
//   * the second round's results go from registers straight to global memory (STG.128 sectors, or for the
//     transposing first pass 16 lanes writing 64 contiguous bytes); the strip buffer is free as soon as the second
//     round has loaded it, so the next strip's cp.async traffic overlaps the second round, the epilogue and the stores
//     with a single buffer;
//   * inter-pass twiddles are Shoup multiplications (1 IMAD.HI + 2 IMAD) instead of Montgomery chains: a per-row
//     factor A[e] generated once per strip (or once per j) with its Shoup companion, and for the first pass, where
//     the eight columns of a strip carry different twiddles, a second factor B[e][c] = w^(c*e) from a fixed table;
//   * non-final passes store lazily reduced values in [0, 2p) (every consumer starts with a reduction).
// Shared-memory addresses are swizzled (row ^ f(row), chunk ^ g(row)) so that all LDS.128 / STS.128 / cp.async
// accesses of both rounds and both lane mappings are bank-conflict free, and so that every address is
// (lane constant) + (compile-time constant).
//
// 16 warps per SM (128 registers each): four quads, one warp of each quad per scheduler.
#pragma once
#include "ntt_pass.cuh"

namespace bb {

#ifndef V5_WARPS_N
#define V5_WARPS_N 16
#endif
constexpr int V5_WARPS = V5_WARPS_N;
constexpr int V5_THREADS = V5_WARPS * 32;
constexpr int V5_LR = 8, V5_R = 256;
constexpr int V5_STRIP_BYTES = V5_R * 32;
constexpr int V5_TW_ENTRIES = 240;  // per-stage compact twiddle tables of the second round: 16 + 32 + 64 + 128
enum : int { V5_ROWS_CANON = 0, V5_ROWS_TWIDDLE = 1, V5_COLS_TWIDDLE = 2 };

__host__ __device__ constexpr size_t v5_smem_bytes(int mode) {
    return (size_t)V5_TW_ENTRIES * 8 + (mode == V5_COLS_TWIDDLE ? (size_t)V5_R * 8 * 8 : 0) +
           (size_t)V5_WARPS * (V5_STRIP_BYTES + (mode != V5_ROWS_CANON ? V5_R * 8 : 0));
}

__host__ __device__ constexpr uint32_t brev4c(uint32_t v) { return ((v & 1u) << 3) | ((v & 2u) << 1) | ((v & 4u) >> 1) | ((v & 8u) >> 3); }

#ifdef __CUDACC__
// canonical g^e (times the table's constant factor)
BB_D uint32_t pow_plain(const PowTable& t, uint32_t e) {
    const uint2 lo = __ldg(&t.lo[e & ((1u << t.lo_bits) - 1u)]);
    const uint2 hi = __ldg(&t.hi[e >> t.lo_bits]);
    const uint32_t w = shoup_mul_lazy(lo.x, hi.x, hi.y);
    return min(w, w - P);
}

BB_D void cp_async16(uint32_t dst, const void* src) {
#if defined(V5_CP_CA)
    asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
#elif defined(V5_CP_PLAIN)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
#endif
}

template <int MODE>
__global__ void __launch_bounds__(V5_THREADS, 1)
    ntt_pass_v5_kernel(const PassParams p, const uint2* __restrict__ btab, uint32_t strips_x, uint32_t total_strips,
                       uint32_t unused) {
    constexpr int R = V5_R;
    extern __shared__ __align__(128) uint4 smem5[];
    uint2* const stw = reinterpret_cast<uint2*>(smem5);
    uint4* const sB = smem5 + V5_TW_ENTRIES / 2;
    constexpr int B_U4 = (MODE == V5_COLS_TWIDDLE) ? R * 4 : 0;
    constexpr int WARP_U4 = V5_STRIP_BYTES / 16 + (MODE != V5_ROWS_CANON ? R / 2 : 0);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* const strip = sB + B_U4 + warp * WARP_U4;
    uint2* const sA = reinterpret_cast<uint2*>(strip + V5_STRIP_BYTES / 16);

    // Work assignment.  Four consecutive strips share every 128-byte line they read (and, in the row-store modes,
    // write), so they go to the four warps of a "quad" at the same time: lines are fetched and completed once,
    // whole, instead of sector by sector at distant times.  The four warps of a quad sit on the four schedulers of
    // the SM and every quad works on whole line groups, so the schedulers carry identical loads; line groups are
    // split evenly (+-1) over the CTAs and, inside a CTA, evenly (+-1) and contiguously over its quads.
    const uint32_t quad = warp >> 2, qi = warp & 3u;
    constexpr uint32_t NQ = V5_WARPS / 4;
    const uint32_t lg_total = total_strips >> 2;
    const uint32_t cta_lo = (uint32_t)(((unsigned long long)blockIdx.x * lg_total) / gridDim.x);
    const uint32_t cta_hi = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * lg_total) / gridDim.x);
    const uint32_t lg_lo = cta_lo + (quad * (cta_hi - cta_lo)) / NQ, lg_hi = cta_lo + ((quad + 1) * (cta_hi - cta_lo)) / NQ;
    uint32_t s = 4u * lg_lo + qi;
    const uint32_t s_end = 4u * lg_hi;
    const bool active = s < s_end;

    const uint32_t strip_sa = (uint32_t)__cvta_generic_to_shared(strip);
    const uint32_t ncols = p.ncols;

    // ---- lane constants
    // load: lane = lh + 2 dl; data row d = 16 it + dl lands at logical row brev8(d) = brev4(dl) << 4 | brev4(it)
    const uint32_t lh = lane & 1u, dl = lane >> 1, rt = __brev(dl) >> 28;
    const uint32_t ld_base = strip_sa + rt * 512u + lh * 16u;
    const uint32_t ld_s = rt & 3u;
    // round 1: lane = h1 + 2 blk, rows blk*16 + k
    const uint32_t h1 = lane & 1u, blk = lane >> 1;
    const uint32_t r1_base = blk * 32u + h1;  // 16-byte units
    const uint32_t r1_s = blk & 3u;
    // round 2: rows b + 16 k; lanes walk the columns first (row stores) or the rows first (transposing stores)
    const uint32_t b = (MODE == V5_COLS_TWIDDLE) ? (lane & 15u) : (lane >> 1);
    const uint32_t h2 = (MODE == V5_COLS_TWIDDLE) ? (lane >> 4) : (lane & 1u);
    const uint32_t r2_base = h2 ^ ((b >> 2) & 1u);  // 16-byte units

    // (vector, strip within the vector) of the strip being computed and of the one being prefetched; a quad's strips
    // advance by 4 within a vector (strips_x is a multiple of 4), so no division is needed after the first
    uint32_t bz = s / strips_x, sx = s - bz * strips_x;
    uint32_t nbz = bz, nsx = sx;
    auto issue_load = [&]() {
        const uint32_t* src = p.in + (size_t)nbz * p.in_batch_stride + (size_t)dl * ncols + nsx * 8u + 4u * lh;
        const size_t step = (size_t)16 * ncols;
#pragma unroll
        for (int it = 0; it < 16; it++) {
            const uint32_t rl = brev4c((uint32_t)it);
            const uint32_t off = (((rl & 12u) | ((rl & 3u) ^ ld_s)) << 5);
            uint32_t dst = ld_base + off;
            if (rl & 4u) dst ^= 16u;  // chunk ^ tau(row), tau = bit 2 of the logical row
#ifndef V5_NO_LOAD
            cp_async16(dst, src);
#else
            if (p.ncols == 12345u) cp_async16(dst, src);
#endif
            src += step;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    if (active) issue_load();  // in flight while the tables are set up

    // ---- CTA-wide tables (the only CTA-wide step)
    for (uint32_t i = threadIdx.x; i < (uint32_t)V5_TW_ENTRIES; i += V5_THREADS) {
        // stage t occupies [16 (2^t - 1), 16 (2^(t+1) - 1)); entry m = b + 16 kp holds omega_R^(m << (3 - t))
        const uint32_t t = (i >= 112u) ? 3u : (i >= 48u) ? 2u : (i >= 16u) ? 1u : 0u;
        const uint32_t m = i - 16u * ((1u << t) - 1u);
        stw[i] = __ldg(&p.tw[(m << (3u - t)) << (LOG_TW - V5_LR)]);
    }
    if (MODE == V5_COLS_TWIDDLE) {
        // btab[e*8 + c] -> 16-byte slot (e, si = c/2) at e*4 + (si ^ ((e >> 1) & 3)): eight consecutive rows hit the
        // eight different slots of a 128-byte line
        for (uint32_t i = threadIdx.x; i < (uint32_t)(R * 4); i += V5_THREADS) {
            const uint32_t e = i >> 2, si = i & 3u;
            sB[(e << 2) + (si ^ ((e >> 1) & 3u))] = __ldg(reinterpret_cast<const uint4*>(btab) + i);
        }
    }
    __syncthreads();

    if (!active) return;
    uint32_t cur_j = 0xFFFFFFFFu;
    const uint32_t log_pfull = p.log_pfull;

    for (; s < s_end; s += 4) {
        const uint32_t col0 = sx * 8u;
        __syncwarp();
        // ---- per-row inter-pass twiddles A[e] = w^((jj * e) << shift) as Shoup pairs (overlaps the load in flight)
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
        // ---- round 1: rows blk*16 + k, twiddles are powers of omega_16 (kernel parameters)
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t idx = r1_base + ((((uint32_t)k & 12u) | (((uint32_t)k & 3u) ^ r1_s)) << 1);
            const uint4 v = strip[(k & 4) ? (idx ^ 1u) : idx];
            x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
        }
#ifndef V5_NO_COMPUTE
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
#endif
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint32_t idx = r1_base + ((((uint32_t)k & 12u) | (((uint32_t)k & 3u) ^ r1_s)) << 1);
            strip[(k & 4) ? (idx ^ 1u) : idx] = make_uint4(x[k][0], x[k][1], x[k][2], x[k][3]);
        }
        __syncwarp();

        // ---- round 2: rows b + 16 k
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint4 v = strip[r2_base + (uint32_t)k * 32u + ((b ^ ((uint32_t)k & 3u)) << 1)];
            x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
        }
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const uint2* tws = stw + 16 * ((1 << t) - 1) + b;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << t)) continue;
                const int kp = k & ((1 << t) - 1);
#ifndef V5_NO_COMPUTE
                const uint2 w = tws[16 * kp];
#pragma unroll
                for (int c = 0; c < 4; c++) bfly(x[k][c], x[k + (1 << t)][c], w);
#endif
            }
            if (t == 0) {
                // every lane has consumed its 16 loads: the strip buffer is free for the next strip
                __syncwarp();
                nsx += 4;
                if (nsx >= strips_x) {
                    nsx -= strips_x;
                    nbz++;
                }
                if (s + 4 < s_end) issue_load();
            }
        }

        // ---- epilogue + store straight from registers; x[k] is output row e = b + 16 k
        if (MODE == V5_COLS_TWIDDLE) {
            uint32_t* o = p.out + (size_t)bz * p.out_batch_stride + ((size_t)(col0 + 4u * h2) << V5_LR) + b;
            const uint32_t xb = (b >> 1) & 3u;
            const uint4* bq0 = sB + (b << 2) + ((2u * h2) ^ xb);
            const uint4* bq1 = sB + (b << 2) + ((2u * h2 + 1u) ^ xb);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint2 a = sA[b + 16 * k];
                const uint4 b01 = bq0[64 * k], b23 = bq1[64 * k];
                uint32_t v0 = shoup_mul_lazy(shoup_mul_lazy(x[k][0], b01.x, b01.y), a.x, a.y);
                uint32_t v1 = shoup_mul_lazy(shoup_mul_lazy(x[k][1], b01.z, b01.w), a.x, a.y);
                uint32_t v2 = shoup_mul_lazy(shoup_mul_lazy(x[k][2], b23.x, b23.y), a.x, a.y);
                uint32_t v3 = shoup_mul_lazy(shoup_mul_lazy(x[k][3], b23.z, b23.w), a.x, a.y);
#ifdef V5_NO_STORE
                if (v0 == 0xFFFFFFFFu && v1 == 0xFFFFFFFEu)
#endif
                {
                    o[16 * k] = v0;
                    o[16 * k + R] = v1;
                    o[16 * k + 2 * R] = v2;
                    o[16 * k + 3 * R] = v3;
                }
            }
        } else {
            const uint32_t j = col0 >> log_pfull, low0 = col0 & ((1u << log_pfull) - 1u);
            uint32_t* o = p.out + (size_t)bz * p.out_batch_stride + ((((size_t)j << V5_LR) + b) << log_pfull) + low0 + 4u * h2;
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
#ifndef V5_NO_STORE
                *reinterpret_cast<uint4*>(o) = v;
#else
                if (v.x == 0xFFFFFFFFu && v.y == 0xFFFFFFFEu) *reinterpret_cast<uint4*>(o) = v;
#endif
                o += step;
            }
        }
        bz = nbz;
        sx = nsx;
    }
}

// strips_x = ncols / 8 strips per vector; `batch` vectors
template <int MODE>
int launch_pass_v5(const PassParams& p, const uint2* btab, uint32_t strips_x, uint32_t batch, cudaStream_t s) {
    static int n_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    constexpr size_t smem = v5_smem_bytes(MODE);
    if (n_sm[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_v5_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    }
    const uint32_t total = strips_x * batch;  // a multiple of 4 (checked by the caller)
    uint32_t ctas = (uint32_t)n_sm[dev];
    if (ctas > total / 4u) ctas = total / 4u;
    ntt_pass_v5_kernel<MODE><<<ctas, V5_THREADS, smem, s>>>(p, btab, strips_x, total, 0u);
    return (int)cudaGetLastError();
}
#endif  // __CUDACC__

}  // namespace bb
