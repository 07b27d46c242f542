// First pass of the blowup-32 coset LDE 2^20 -> 2^25 (BabyBearDomain::fft on the prover's shifted domain,
// src/math/domain.rs:107-123,154-162 as called from src/fibonacci.rs:124-128): results bit-exact with the CPU path.
//
// The 2^25-point transform of (at most 2^21) coefficients is a four-step 8192 x 4096: with i = 4096 i1 + j and
// m = k1 + 8192 k2,   X[m] = sum_j w_n^(j k1) w_4096^(j k2) Z[j][k1],   Z[j][k1] = sum_i1 a'[4096 i1 + j] w_8192^(i1 k1),
// a'[i] = shift^i a[i].  The second sum is the TMA-staged pass-2 kernel (ntt_pass_v7.cuh) on 8192 columns.  This file
// is the first sum, and it is where the zero padding pays: only i1 < 256 (+ the rows of a masked trace polynomial,
// which has 140 coefficients more than 2^20) are non-zero, so writing k1 = 32 m + c the 8192-point column transform is
// 32 coset transforms of 256 points,  Z[j][32 m + c] = sum_i1 (a'[i1] w_8192^(i1 c)) w_256^(i1 m)  — eight butterfly
// stages per output instead of thirteen, and no multiplication besides the butterflies': with i1 = 16 a + b and
// m = a' + 16 b' both radix-16 rounds are decimation-in-time transforms of a geometric-twisted input, whose twist
// (w_8192^(16 c))^a resp. (w_8192^(32 a' + c))^b folds into the butterfly twiddles — every twiddle is a power of
// w_8192 below 4096, one table look-up.
//
// A persistent CTA of 16 warps walks tiles of four adjacent columns j (one tile = 4 x 8192 results = 128 KB of shared
// memory, laid out [a'][b][c] as uint4 over the four columns):
//   round A  warp = b, lane = c: reads the staged inputs a'[16 a + b] (a warp-wide broadcast; rows >= 256 fold in with
//            w_32^c), 16-point DIT over a, writes tile[a'][b][c]                  (lanes contiguous: conflict free)
//   round B  warp = a', lane = c: reads tile[a'][b][c], 16-point DIT over b, stores Z[j][512 b' + 32 a' + c] straight from
//            registers — a warp stores one 128-byte line (the 32 cosets of one m) per column.
// The 16 KB of inputs of the next tile are fetched (and multiplied by shift^i) between the rounds; the input vector is
// 4 MB and stays in L2.  Values leave in [0, 2p): pass 2 starts with a Shoup multiplication that takes any u32.
#pragma once
#include "ntt_v7.cuh"

namespace bb {

// x[k] holds input digit brev4(k); afterwards x[k] = sum_d in_d (w^H w_16^k)^d.  TW(t, kp) returns the twiddle
// omega_8192^((H + 512 kp) << (3 - t)).
template <typename TW>
BB_D void lde_dit16(uint4 (&x)[16], TW tw) {
#pragma unroll
    for (int t = 0; t < 4; t++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k & (1 << t)) continue;
            const int kp = k & ((1 << t) - 1);
            bfly4(x[k], x[k + (1 << t)], tw(t, kp));
        }
    }
}

__global__ void __launch_bounds__(LDE::NT, 1) lde_expand_kernel(const LdeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    {
        uint2* tb = reinterpret_cast<uint2*>(smem + LDE::OFF_TB);
        for (uint32_t i = tid; i < 15u * 512u; i += LDE::NT) {  // entry (2^t - 1 + kp, H = 32 a' + c) = omega_8192^((H + 512 kp) << (3 - t))
            const uint32_t e = i >> 9, H = i & 511u;
            const uint32_t t = (e >= 1u) + (e >= 3u) + (e >= 7u), kp = e - ((1u << t) - 1u);
            tb[i] = __ldg(&p.tw[(H + 512u * kp) << (3u - t)]);
        }
        uint2* ta = reinterpret_cast<uint2*>(smem + LDE::OFF_TA);
        for (uint32_t i = tid; i < 15u * 32u; i += LDE::NT) {  // entry (2^t - 1 + kp, c) = omega_8192^((16 c + 512 kp) << (3 - t))
            const uint32_t e = i >> 5, c = i & 31u;
            const uint32_t t = (e >= 1u) + (e >= 3u) + (e >= 7u), kp = e - ((1u << t) - 1u);
            ta[i] = __ldg(&p.tw[(16u * c + 512u * kp) << (3u - t)]);
        }
        if (tid < 32u) {  // omega_32^c = omega_8192^(256 c); the table stops at 4095: the upper half is the negated lower half
            const uint32_t e = 256u * tid;
            uint2 w = __ldg(&p.tw[e & 4095u]);
            if (e >= 4096u) {
                w.x = P - w.x;
                w.y = shoup_companion_fast(w.x);
            }
            reinterpret_cast<uint2*>(smem + LDE::OFF_W32)[tid] = w;
        }
    }
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

    const uint32_t nrows = (p.n_coeffs + 4095u) >> 12;
    // Inputs of one tile: thread `tid` brings row i1 = tid, four columns, multiplied by shift^i.  The load is issued before
    // round A and used after it; the power of the shift is a running product from tile to tile (one table look-up per
    // thread and kernel: field arithmetic is exact, a running product gives the same bits).
    auto fetch = [&](uint32_t tile) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t idx = (tid << 12) + 4u * tile;
        if (tid < nrows) {
            if (idx + 3u < p.n_coeffs) {
                v = __ldg(reinterpret_cast<const uint4*>(p.in + idx));
            } else {
                if (idx < p.n_coeffs) v.x = __ldg(p.in + idx);
                if (idx + 1u < p.n_coeffs) v.y = __ldg(p.in + idx + 1u);
                if (idx + 2u < p.n_coeffs) v.z = __ldg(p.in + idx + 2u);
            }
        }
        return v;
    };
    uint32_t tile = blockIdx.x, it = 0;
    uint32_t pw = 0, pw_step = 0, s1 = 0;  // Montgomery forms of shift^(4096 tid + 4 tile), shift^(4 gridDim), shift
    if (p.has_shift && tid < nrows && tile < p.total_tiles) {
        pw = pow_lookup(p.shift, (tid << 12) + 4u * tile);
        pw_step = pow_lookup(p.shift, 4u * gridDim.x);
        s1 = pow_lookup(p.shift, 1u);
    }
    auto stage = [&](uint4 v, uint32_t buf) {
        if (p.has_shift && tid < nrows) {
            uint32_t q = pw;
            v.x = monty_mul(v.x, q);
            q = monty_mul(q, s1);
            v.y = monty_mul(v.y, q);
            q = monty_mul(q, s1);
            v.z = monty_mul(v.z, q);
            q = monty_mul(q, s1);
            v.w = monty_mul(v.w, q);
            pw = monty_mul(pw, pw_step);
        }
        reinterpret_cast<uint4*>(smem + LDE::OFF_IN)[buf * 512u + tid] = v;
    };

    uint4* tile_s = reinterpret_cast<uint4*>(smem + LDE::OFF_TILE);
    if (tile < p.total_tiles) stage(fetch(tile), 0u);
    __syncthreads();
    while (tile < p.total_tiles) {
        const uint32_t next = tile + gridDim.x;
        const bool have_next = next < p.total_tiles;
        uint4 nx = make_uint4(0u, 0u, 0u, 0u);
        if (have_next) nx = fetch(next);
        {   // round A: b = warp, c = lane
            const uint4* xin = reinterpret_cast<const uint4*>(smem + LDE::OFF_IN) + (it & 1u) * 512u;
            uint4 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = xin[16u * v7_brev4(k) + warp];
            if (nrows > 256u) {
                const uint2 w = reinterpret_cast<const uint2*>(smem + LDE::OFF_W32)[lane];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t row2 = 256u + 16u * v7_brev4(k) + warp;
                    if (row2 < nrows) {  // uniform over the warp
                        uint4 y = xin[row2];
                        y.x = shoup_mul_lazy(y.x, w.x, w.y);
                        y.y = shoup_mul_lazy(y.y, w.x, w.y);
                        y.z = shoup_mul_lazy(y.z, w.x, w.y);
                        y.w = shoup_mul_lazy(y.w, w.x, w.y);
                        y = canon4(y);
                        x[k] = make_uint4(x[k].x + y.x, x[k].y + y.y, x[k].z + y.z, x[k].w + y.w);  // < 2p
                    }
                }
            }
            const uint2* ta = reinterpret_cast<const uint2*>(smem + LDE::OFF_TA) + lane;
            lde_dit16(x, [&](int t, int kp) { return ta[32 * ((1 << t) - 1 + kp)]; });
#pragma unroll
            for (int k = 0; k < 16; k++) tile_s[(k * 16u + warp) * 32u + lane] = x[k];
        }
        if (have_next) stage(nx, (it + 1u) & 1u);
        __syncthreads();  // the tile is complete; the inputs of the next tile are staged
        {   // round B: a' = warp, c = lane
            uint4 x[16];
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = tile_s[(warp * 16u + v7_brev4(k)) * 32u + lane];
            // every warp holds its part of the tile in registers: the buffer is free for round A of the next tile, which a
            // warp starts as soon as its own stores below are issued — stores and arithmetic of neighbouring tiles overlap
            __syncthreads();
            const uint2* tb = reinterpret_cast<const uint2*>(smem + LDE::OFF_TB) + 32u * warp + lane;  // H = 32 a' + c
            lde_dit16(x, [&](int t, int kp) { return tb[512 * ((1 << t) - 1 + kp)]; });
            uint32_t* o = p.out + ((size_t)(4u * tile) << LDE::LOG_ROWS) + 32u * warp + lane;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                o[512 * k + 0 * 8192] = x[k].x;
                o[512 * k + 1 * 8192] = x[k].y;
                o[512 * k + 2 * 8192] = x[k].z;
                o[512 * k + 3 * 8192] = x[k].w;
            }
        }
        tile = next;
        it++;
    }
}

}  // namespace bb
