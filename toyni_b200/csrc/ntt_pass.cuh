// One pass of the multi-pass (Stockham / four-step) BabyBear NTT.
//
// Replaces, from scratch, the reference's bit-reverse kernel + one-launch-per-stage radix-2 butterflies +
// scale pass (cuda/ntt_kernel.cu:103-143, 249-292).  A transform of n = R_1*...*R_m points runs as m passes
// (m <= 3).  Pass i splits the index as (d_i | rest): for every "column" (all other digits fixed) it does an
// R_i-point DFT over d_i inside shared memory, multiplies by the inter-pass twiddle w_n^(P*j*e) and writes
// the result so that the already transformed digits end up least significant:
//     read   in [ d * ncols + col ]                     col = j * pfull + low
//     write  out[ (j * R + e) * pfull + low ]
// After the last pass the array is the natural-order DFT (bit-exact with src/ntt.rs:24-53).  `pfull` also
// carries an interleave factor (4 for AoS extension-field arrays, src/math/domain.rs:140-151 becomes one
// launch instead of four transforms plus two transposes).
//
// Inside the tile: rows are stored bit-reversed at load time (rows are separate global segments, so the
// permutation is free), then radix-16 decimation-in-time rounds run in registers with strides 1, 16, 256.
// Butterflies are lazy Shoup butterflies (3 IMAD + 4 ALU, values in [0,2p)); the first round's twiddles are
// compile-time-indexed constants.  The epilogue canonicalises, so stores are always in [0,p).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bb_field.cuh"

namespace bb {

constexpr int LOG_TW = 12;  // the master Shoup twiddle table covers omega_4096 (largest in-tile transform)

enum : uint32_t { PRO_NONE = 0, PRO_INIDX = 1 };
enum : uint32_t { EPI_NONE = 0, EPI_TWIDDLE = 1, EPI_OUTIDX = 2, EPI_CONST = 3, EPI_FOURSTEP = 4 };

// g^t = hi[t >> lo_bits] * lo[t & mask]; entries are Shoup pairs (w, floor(w 2^32 / p)) of plain values
struct PowTable {
    const uint2* lo;
    const uint2* hi;
    uint32_t lo_bits;
};

struct PassParams {
    const uint32_t* in;
    uint32_t* out;
    unsigned long long in_batch_stride, out_batch_stride;  // u32 units per blockIdx.y
    unsigned long long n_in_limit;                         // zero padding: logical index >= limit reads as 0
    unsigned long long in_row_stride;                      // u32 units between consecutive d
    unsigned long long in_col_stride;                      // u32 units between consecutive columns
    uint32_t ncols;                                        // number of columns (u32 units, incl. interleave)
    uint32_t log_pfull;                                    // log2(pfull)
    uint32_t log_inner;                                    // log2(interleave factor)
    const uint2* tw;                                       // Shoup table: (w, w') of omega_4096^i, i < 2048
    uint2 tw16[8];                                         // (w, w') of omega_16^i (i<8) in this direction
    uint32_t transposed;                                   // 1: rows contiguous in memory (batch of vectors)
    uint32_t pro_mode, epi_mode;
    PowTable pro, epi;
    uint32_t epi_const;  // Montgomery-form constant for EPI_CONST
    uint32_t epi_unscale;  // Montgomery form of the inverse of the constant factor folded into epi.lo (R mod p if none)
    uint32_t epi_shift;  // EPI_TWIDDLE exponent = (j*e) << epi_shift
    // EPI_FOURSTEP (last pass of the column transforms of a sharded four-step NTT): multiply element (k1, c) by
    // w_n^((fs_col_offset + c) * k1) (table `epi`) and store row k1 into the buffer of the rank that owns it —
    // a peer pointer over NVLink for another rank — at [k1 mod rows_per_rank][fs_dst_col + c]
    uint32_t* fs_peer[8];
    uint32_t fs_log_rows_per_rank;
    uint32_t fs_dst_row_stride;
    uint32_t fs_dst_col;
    uint32_t fs_col_offset;
    uint32_t pdl;        // host only: launch this pass with programmatic stream serialization
    uint32_t prune_log;  // 5: blowup-32 zero padding, the first five stages of this pass are a plain copy (vector kernel)
};

__host__ __device__ constexpr int rpad(int r) { return r + (r >> 4); }
__host__ __device__ constexpr int col_pitch(int LR, int LC) {
    int rp = rpad(1 << LR) + 1;
    int want = (LC >= 5) ? 1 : (32 >> LC);  // pitch mod 32 that spreads a C-wide row segment over the banks
    int pitch = rp;
    while ((pitch & 31) != (want & 31)) pitch++;
    return pitch;
}
__host__ __device__ constexpr int pass_threads(int LR, int LC) {
    int groups = (1 << (LR + LC)) >> 4;  // radix-16 groups per round
    return groups < 64 ? 64 : (groups > 512 ? 512 : groups);
}
__host__ __device__ constexpr size_t pass_smem_bytes(int LR, int LC) { return (size_t)col_pitch(LR, LC) * (1u << LC) * 4u; }

// a + b on the ALU pipe: written as min(a + b, UINT_MAX) so that ptxas emits VIADDMNMX instead of turning the add
// into IMAD.IADD — the FMA-heavy pipe is the bottleneck of a butterfly (IMAD.HI costs 2.7 issue slots there)
BB_D uint32_t add_alu(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && !defined(BB_PLAIN_ADD)
    return __viaddmin_u32(a, b, 0xFFFFFFFFu);
#else
    return a + b;
#endif
}
BB_D void bfly(uint32_t& u, uint32_t& x, uint2 w) {
    uint32_t v = shoup_mul_lazy(x, w.x, w.y);
    v = min(v, v - P);
    uint32_t uu = min(u, u - P);
    u = add_alu(uu, v);
    x = uu - v + P;
}
BB_D void bfly_one(uint32_t& u, uint32_t& x) {  // twiddle == 1
    uint32_t v = min(x, x - P);
    uint32_t uu = min(u, u - P);
    u = add_alu(uu, v);
    x = uu - v + P;
}

// One radix-2^G_LOG DIT round over rows {blk*G*S + b + k*S}.
template <int LR, int LC, int NT, int S_LOG, int G_LOG>
BB_D void dit_round(uint32_t* __restrict__ sm, const PassParams& p) {
    static_assert(S_LOG == 0 || S_LOG == 4 || S_LOG == 8, "round strides are 1, 16, 256");
    constexpr int R = 1 << LR, C = 1 << LC, S = 1 << S_LOG, G = 1 << G_LOG;
    constexpr int RG = R / G;
    constexpr int RGT = RG < NT ? RG : NT;
    constexpr int CT = NT / RGT;
    constexpr int PITCH = col_pitch(LR, LC);
    constexpr int KOFF = (S_LOG == 0) ? 1 : (S + (S >> 4));  // padded distance between a group's rows
    const int tid = threadIdx.x;
    const int rg0 = tid % RGT, c0 = tid / RGT;
    const uint2* __restrict__ tw = p.tw;
    constexpr uint32_t log_tw = LOG_TW;
#pragma unroll 1
    for (int rg = rg0; rg < RG; rg += RGT) {
        const int b = rg & (S - 1), blk = rg >> S_LOG;
        const int row0 = blk * (G * S) + b;
        const int base = row0 + (row0 >> 4);
        uint2 w[G];  // entry (1<<t) + kp holds stage t's twiddle kp; entry 0 unused
        if (S_LOG > 0) {
#pragma unroll
            for (int t = 0; t < G_LOG; t++) {
#pragma unroll
                for (int kp = 0; kp < (1 << t); kp++) {
                    uint32_t idx = (uint32_t)(b + kp * S) << (log_tw - (S_LOG + t + 1));
                    w[(1 << t) + kp] = __ldg(&tw[idx]);
                }
            }
        }
#pragma unroll 1
        for (int c = c0; c < C; c += CT) {
            uint32_t* col = sm + c * PITCH + base;
            uint32_t x[G];
#pragma unroll
            for (int k = 0; k < G; k++) x[k] = col[k * KOFF];
#pragma unroll
            for (int t = 0; t < G_LOG; t++) {
#pragma unroll
                for (int k = 0; k < G; k++) {
                    if (k & (1 << t)) continue;
                    const int kp = k & ((1 << t) - 1);
                    if (S_LOG == 0) {
                        if (kp == 0)
                            bfly_one(x[k], x[k + (1 << t)]);
                        else
                            bfly(x[k], x[k + (1 << t)], p.tw16[kp << (3 - t)]);
                    } else {
                        bfly(x[k], x[k + (1 << t)], w[(1 << t) + kp]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < G; k++) col[k * KOFF] = x[k];
        }
    }
}

// Montgomery form of g^e (for a factor shared by several values)
BB_D uint32_t pow_lookup(const PowTable& t, uint32_t e) {
    const uint2 lo = __ldg(&t.lo[e & ((1u << t.lo_bits) - 1u)]);
    const uint2 hi = __ldg(&t.hi[e >> t.lo_bits]);
    const uint32_t w = shoup_mul_lazy(lo.x, hi.x, hi.y);  // lo*hi in [0,2p)
    return monty_mul(w, R2_MOD_P);                         // canonical, times R
}
// canonical g^e (times the table's constant factor), plain form
BB_D uint32_t pow_plain(const PowTable& t, uint32_t e) {
    const uint2 lo = __ldg(&t.lo[e & ((1u << t.lo_bits) - 1u)]);
    const uint2 hi = __ldg(&t.hi[e >> t.lo_bits]);
    const uint32_t w = shoup_mul_lazy(lo.x, hi.x, hi.y);
    return min(w, w - P);
}
// v * g^e, canonical, for any 32-bit v: two Shoup multiplications, no product twiddle formed
BB_D uint32_t pow_apply(const PowTable& t, uint32_t e, uint32_t v) {
    const uint2 lo = __ldg(&t.lo[e & ((1u << t.lo_bits) - 1u)]);
    const uint2 hi = __ldg(&t.hi[e >> t.lo_bits]);
    v = shoup_mul_lazy(v, lo.x, lo.y);
    v = shoup_mul_lazy(v, hi.x, hi.y);
    return min(v, v - P);
}

template <int LR, int LC>
__global__ void __launch_bounds__(pass_threads(LR, LC)) ntt_pass_kernel(const PassParams p) {
    constexpr int R = 1 << LR, C = 1 << LC, NT = pass_threads(LR, LC);
    constexpr int PITCH = col_pitch(LR, LC);
    constexpr int G1 = LR < 4 ? LR : 4;
    constexpr int G2 = (LR - G1) < 4 ? (LR - G1) : 4;
    constexpr int G3 = LR - G1 - G2;
    static_assert(G3 <= 4, "LR <= 12");
    extern __shared__ uint32_t sm[];

    const int tid = threadIdx.x;
    const uint32_t col0 = blockIdx.x * C;
    const uint32_t* __restrict__ in = p.in + (size_t)blockIdx.y * p.in_batch_stride;
    uint32_t* __restrict__ out = p.out + (size_t)blockIdx.y * p.out_batch_stride;

    // ---- load: tile rows go to bit-reversed positions
#pragma unroll 4
    for (int i = tid; i < R * C; i += NT) {
        int c, r, d;
        if (!p.transposed) {
            c = i & (C - 1);
            r = i >> LC;
            d = (LR == 0) ? 0 : (int)(__brev((uint32_t)r) >> (32 - LR));
        } else {
            d = i & (R - 1);
            c = i >> LR;
            r = (LR == 0) ? 0 : (int)(__brev((uint32_t)d) >> (32 - LR));
        }
        const uint32_t col = col0 + c;
        const unsigned long long lidx = p.transposed ? (unsigned long long)d : (unsigned long long)d * p.ncols + col;
        uint32_t v = 0;
        if (col < p.ncols && lidx < p.n_in_limit) {
            v = in[(size_t)d * p.in_row_stride + (size_t)col * p.in_col_stride];
            if (p.pro_mode == PRO_INIDX) v = pow_apply(p.pro, (uint32_t)(lidx >> p.log_inner), v);
        }
        sm[c * PITCH + r + (r >> 4)] = v;
    }
    __syncthreads();

    // ---- radix-16 DIT rounds
    if constexpr (G1 > 0) dit_round<LR, LC, NT, 0, G1>(sm, p);
    if constexpr (G2 > 0) {
        __syncthreads();
        dit_round<LR, LC, NT, 4, G2>(sm, p);
    }
    if constexpr (G3 > 0) {
        __syncthreads();
        dit_round<LR, LC, NT, 8, G3>(sm, p);
    }
    __syncthreads();

    // ---- epilogue + store, in global-address order
    const uint32_t log_cl = p.log_pfull < (uint32_t)LC ? p.log_pfull : (uint32_t)LC;
    const uint32_t pfull_mask = (1u << p.log_pfull) - 1u;
#pragma unroll 4
    for (int i = tid; i < R * C; i += NT) {
        const uint32_t l = i & ((1u << log_cl) - 1u);
        const uint32_t e = (i >> log_cl) & (R - 1);
        const uint32_t jj = i >> (log_cl + LR);
        const uint32_t c = (jj << log_cl) + l;
        const uint32_t col = col0 + c;
        if (col >= p.ncols) continue;
        uint32_t v = sm[c * PITCH + e + (e >> 4)];
        const uint32_t j = col >> p.log_pfull, low = col & pfull_mask;
        switch (p.epi_mode) {
            case EPI_TWIDDLE:
                v = pow_apply(p.epi, (j * e) << p.epi_shift, v);
                break;
            case EPI_OUTIDX:
                v = pow_apply(p.epi, ((e << p.log_pfull) + low) >> p.log_inner, v);
                break;
            case EPI_CONST:
                v = monty_mul(v, p.epi_const);
                break;
            default:
                v = min(v, v - P);
                break;
        }
        out[((((size_t)j << LR) + e) << p.log_pfull) + low] = v;
    }
}

// host-side launcher table entry
typedef void (*PassLaunchFn)(const PassParams&, dim3 grid, cudaStream_t);

template <int LR, int LC>
void launch_pass(const PassParams& p, dim3 grid, cudaStream_t s) {
    static bool configured[64] = {};  // per device: one opt-in for > 48 KB dynamic shared memory
    constexpr size_t smem = pass_smem_bytes(LR, LC);
    if (smem > 48 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            cudaFuncSetAttribute(ntt_pass_kernel<LR, LC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured[dev & 63] = true;
        }
    }
    ntt_pass_kernel<LR, LC><<<grid, pass_threads(LR, LC), smem, s>>>(p);
}

}  // namespace bb
