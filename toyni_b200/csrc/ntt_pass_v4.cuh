// Vectorised NTT pass: the hot kernel.  Same mathematics and parameters as ntt_pass.cuh (which stays as the
// scalar path for batches of contiguous short vectors), but the tile is kept row-major in shared memory as
// 16-byte chunks of four adjacent columns, so that
//   * global -> shared is one 16-byte cp.async per chunk (no register staging, the whole tile in flight at once),
//   * every radix-16 round moves data with LDS.128 / STS.128 and the four columns of a chunk share their twiddles,
//   * shared -> global is STG.128 (row-granular passes) or four coalesced STG.32 (the transposing first pass).
// Bank conflicts are removed by an XOR swizzle on the chunk index that is GF(2)-linear in (row, chunk), so the
// address of row (row0 | k*S) is base ^ constant_k.
#pragma once
#include "ntt_pass.cuh"

namespace bb {

template <int LR, int LC>
struct V4 {
    static_assert(LC >= 2 && LC <= 5, "4..32 columns");
    static constexpr int R = 1 << LR, C = 1 << LC, LCV = LC - 2, CV = 1 << LCV;
    static constexpr int Q = 3 - LCV;  // log2(rows per 128-byte line)
    static constexpr int G1 = LR < 4 ? LR : 4;
    static constexpr int G2 = (LR - G1) < 4 ? (LR - G1) : 4;
    static constexpr int G3 = LR - G1 - G2;
    static constexpr int ITEMS = (R >> G1) * CV;  // work items of the first round
    static constexpr int NT = ITEMS < 32 ? 32 : (ITEMS > 512 ? 512 : ITEMS);
    static constexpr size_t SMEM = (size_t)R * C * 4;

    // physical 16-byte chunk index of (row, cv)
    __host__ __device__ static constexpr uint32_t chunk(uint32_t row, uint32_t cv) {
        constexpr uint32_t qm = (1u << Q) - 1u;
        uint32_t rin = (row & qm) ^ ((row >> 4) & qm);
        uint32_t cvv = cv ^ ((row >> Q) & (CV - 1u));
        return ((row >> Q) << 3) | (rin << LCV) | cvv;
    }
};

BB_D void bfly4(uint4& u, uint4& x, uint2 w) {
    bfly(u.x, x.x, w);
    bfly(u.y, x.y, w);
    bfly(u.z, x.z, w);
    bfly(u.w, x.w, w);
}
BB_D void bfly4_one(uint4& u, uint4& x) {
    bfly_one(u.x, x.x);
    bfly_one(u.y, x.y);
    bfly_one(u.z, x.z);
    bfly_one(u.w, x.w);
}

template <int LR, int LC, int S_LOG, int G_LOG>
BB_D void dit_round_v4(uint4* __restrict__ sm, const PassParams& p) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, CV = T::CV, NT = T::NT, S = 1 << S_LOG, G = 1 << G_LOG;
    constexpr int ITEMS = (R / G) * CV;
    const uint2* __restrict__ tw = p.tw;
    const uint32_t log_tw = p.log_tw;
#pragma unroll 1
    for (int it = threadIdx.x; it < ITEMS; it += NT) {
        const uint32_t cv = it & (CV - 1);
        const uint32_t rg = it >> T::LCV;
        const uint32_t b = rg & (S - 1), blk = rg >> S_LOG;
        const uint32_t row0 = blk * (G * S) + b;
        const uint32_t base = T::chunk(row0, cv);
        uint4 x[G];
#pragma unroll
        for (int k = 0; k < G; k++) x[k] = sm[base ^ T::chunk((uint32_t)k << S_LOG, 0)];
#pragma unroll
        for (int t = 0; t < G_LOG; t++) {
#pragma unroll
            for (int k = 0; k < G; k++) {
                if (k & (1 << t)) continue;
                const int kp = k & ((1 << t) - 1);
                if (S_LOG == 0) {
                    if (kp == 0)
                        bfly4_one(x[k], x[k + (1 << t)]);
                    else
                        bfly4(x[k], x[k + (1 << t)], p.tw16[kp << (3 - t)]);
                } else {
                    const uint32_t idx = (b + (uint32_t)kp * S) << (log_tw - (S_LOG + t + 1));
                    bfly4(x[k], x[k + (1 << t)], __ldg(&tw[idx]));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < G; k++) sm[base ^ T::chunk((uint32_t)k << S_LOG, 0)] = x[k];
    }
}

BB_D uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

BB_D uint4 mul4(uint4 v, uint32_t w_m) {
    return make_uint4(monty_mul(v.x, w_m), monty_mul(v.y, w_m), monty_mul(v.z, w_m), monty_mul(v.w, w_m));
}

template <int LR, int LC>
__global__ void __launch_bounds__(V4<LR, LC>::NT) ntt_pass_v4_kernel(const PassParams p) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, C = T::C, CV = T::CV, LCV = T::LCV, NT = T::NT;
    extern __shared__ uint4 smv[];

    const int tid = threadIdx.x;
    const uint32_t col0 = blockIdx.x * C;
    const uint32_t* __restrict__ in = p.in + (size_t)blockIdx.y * p.in_batch_stride;
    uint32_t* __restrict__ out = p.out + (size_t)blockIdx.y * p.out_batch_stride;
    const uint32_t ncols = p.ncols;

    // ---- load: one 16-byte cp.async per chunk, row d of the tile lands at row bitrev(d); bytes past the
    //      zero-padding limit are zero-filled by the copy engine
#pragma unroll 4
    for (int i = tid; i < R * CV; i += NT) {
        const uint32_t cv = i & (CV - 1), r = i >> LCV;
        const uint32_t d = (LR == 0) ? 0u : (__brev(r) >> (32 - LR));
        const unsigned long long lidx = (unsigned long long)d * ncols + col0 + 4u * cv;
        const unsigned long long remw = p.n_in_limit > lidx ? p.n_in_limit - lidx : 0ull;  // valid words from here
        const uint32_t bytes = remw >= 4ull ? 16u : (uint32_t)remw * 4u;
        const uint32_t* src = in + (bytes ? lidx : 0ull);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(&smv[T::chunk(r, cv)])), "l"(src), "r"(bytes)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    if (p.pro_mode == PRO_INIDX) {  // coset shift of the (few) real input rows: x[i] *= s^i
        const uint32_t rows_in = (uint32_t)((p.n_in_limit + ncols - 1) / ncols);  // rows that hold any input
#pragma unroll 1
        for (uint32_t i = tid; i < rows_in * CV && i < (uint32_t)(R * CV); i += NT) {
            const uint32_t cv = i & (CV - 1), d = i >> LCV;
            const uint32_t r = (LR == 0) ? 0u : (__brev(d) >> (32 - LR));
            const uint32_t lidx = d * ncols + col0 + 4u * cv;
            uint4 v = smv[T::chunk(r, cv)];
            if (p.log_inner >= 2) {
                v = mul4(v, pow_lookup(p.pro, lidx >> p.log_inner));
            } else {
                v.x = monty_mul(v.x, pow_lookup(p.pro, (lidx + 0) >> p.log_inner));
                v.y = monty_mul(v.y, pow_lookup(p.pro, (lidx + 1) >> p.log_inner));
                v.z = monty_mul(v.z, pow_lookup(p.pro, (lidx + 2) >> p.log_inner));
                v.w = monty_mul(v.w, pow_lookup(p.pro, (lidx + 3) >> p.log_inner));
            }
            smv[T::chunk(r, cv)] = v;
        }
        __syncthreads();
    }

    // ---- radix-16 DIT rounds, strides 1, 16, 256
    if constexpr (T::G1 > 0) dit_round_v4<LR, LC, 0, T::G1>(smv, p);
    if constexpr (T::G2 > 0) {
        __syncthreads();
        dit_round_v4<LR, LC, 4, T::G2>(smv, p);
    }
    if constexpr (T::G3 > 0) {
        __syncthreads();
        dit_round_v4<LR, LC, 8, T::G3>(smv, p);
    }
    __syncthreads();

    // ---- epilogue + store
    const uint32_t log_pfull = p.log_pfull;
    const uint32_t pfull_mask = (1u << log_pfull) - 1u;
    if (log_pfull >= 2) {
        // chunks stay whole in the output: the four values of a chunk share j (and the inter-pass twiddle)
        const uint32_t log_clv = (log_pfull < (uint32_t)LC ? log_pfull : (uint32_t)LC) - 2;  // chunks per contiguous run
        const uint32_t j0 = col0 >> log_pfull, low0 = col0 & pfull_mask;
#pragma unroll 2
        for (int i = tid; i < R * CV; i += NT) {
            const uint32_t lv = i & ((1u << log_clv) - 1u);
            const uint32_t e = (i >> log_clv) & (R - 1);
            const uint32_t jj = i >> (log_clv + LR);
            const uint32_t cv = (jj << log_clv) + lv;
            uint4 v = smv[T::chunk(e, cv)];
            const uint32_t j = j0 + jj, low = low0 + 4u * lv;
            switch (p.epi_mode) {
                case EPI_TWIDDLE:
                    v = mul4(v, pow_lookup(p.epi, (j * e) << p.epi_shift));
                    break;
                case EPI_OUTIDX: {
                    const uint32_t k = (e << log_pfull) + low;
                    if (p.log_inner >= 2) {
                        v = mul4(v, pow_lookup(p.epi, k >> p.log_inner));
                    } else {
                        v.x = monty_mul(v.x, pow_lookup(p.epi, (k + 0) >> p.log_inner));
                        v.y = monty_mul(v.y, pow_lookup(p.epi, (k + 1) >> p.log_inner));
                        v.z = monty_mul(v.z, pow_lookup(p.epi, (k + 2) >> p.log_inner));
                        v.w = monty_mul(v.w, pow_lookup(p.epi, (k + 3) >> p.log_inner));
                    }
                    break;
                }
                case EPI_CONST:
                    v = mul4(v, p.epi_const);
                    break;
                default:
                    v = make_uint4(min(v.x, v.x - P), min(v.y, v.y - P), min(v.z, v.z - P), min(v.w, v.w - P));
                    break;
            }
            *reinterpret_cast<uint4*>(out + ((((size_t)j << LR) + e) << log_pfull) + low) = v;
        }
    } else {
        // pfull == 1 (first pass of a plain vector): column col is written as the contiguous run out[col*R + e];
        // lanes walk e, so each of the four scalar stores of a chunk is coalesced
#pragma unroll 2
        for (int i = tid; i < R * CV; i += NT) {
            const uint32_t e = i & (R - 1), cv = i >> LR;
            uint4 v = smv[T::chunk(e, cv)];
            const uint32_t col = col0 + 4u * cv;
            if (p.epi_mode == EPI_TWIDDLE) {
                v.x = monty_mul(v.x, pow_lookup(p.epi, ((col + 0) * e) << p.epi_shift));
                v.y = monty_mul(v.y, pow_lookup(p.epi, ((col + 1) * e) << p.epi_shift));
                v.z = monty_mul(v.z, pow_lookup(p.epi, ((col + 2) * e) << p.epi_shift));
                v.w = monty_mul(v.w, pow_lookup(p.epi, ((col + 3) * e) << p.epi_shift));
            } else if (p.epi_mode == EPI_CONST) {
                v = mul4(v, p.epi_const);
            } else {
                v = make_uint4(min(v.x, v.x - P), min(v.y, v.y - P), min(v.z, v.z - P), min(v.w, v.w - P));
            }
            uint32_t* o = out + ((size_t)col << LR) + e;
            o[0] = v.x;
            o[(size_t)1 << LR] = v.y;
            o[(size_t)2 << LR] = v.z;
            o[(size_t)3 << LR] = v.w;
        }
    }
}

template <int LR, int LC>
void launch_pass_v4(const PassParams& p, dim3 grid, cudaStream_t s) {
    using T = V4<LR, LC>;
    static bool configured[64] = {};
    if (T::SMEM > 48 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            cudaFuncSetAttribute(ntt_pass_v4_kernel<LR, LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
            configured[dev & 63] = true;
        }
    }
    ntt_pass_v4_kernel<LR, LC><<<grid, T::NT, T::SMEM, s>>>(p);
}

}  // namespace bb
