// TMA-staged NTT pass for the two-pass plan of a 2^24-point transform (4096 x 4096).
//
// Replaces, from scratch, the reference's bit-reverse + one-launch-per-stage + scale kernels
// (cuda/ntt_kernel.cu:103-143, 249-292); results are bit-exact with src/ntt.rs:24-66.
//
// One persistent CTA per SM owns ONE 128 KB tile buffer (4096 rows x 8 columns of u32) and walks its tiles with
// sixteen warps; warp w owns chunk w of the buffer (the 256 rows whose middle hex digit is w, one contiguous 8 KB
// region).  Per tile, three radix-16 decimation-in-time rounds:
//   round 1  reads chunk w as the TMA wrote it and writes it back in tile order         (warp-private, in place)
//   round 2  works across the chunks on the rows whose top digit is w                    (in place, between two barriers)
//   round 3  reads chunk w and stores its results straight from registers to global memory.
// Round 3 of tile t and round 1 of tile t+1 are fused in the same warp: as soon as the round-3 loads of chunk w sit
// in registers the region is free, and lane 0 re-fills it with chunk w of the NEXT tile — one cp.async.bulk.tensor
// (TMA) 5-D box {8 columns, 16 x d0, d1 = w, 16 x d2, batch}, i.e. the hex digits of the row index arrive in the order
// round 1 wants them (the digit swap of the Stockham recursion is the tensor map's dimension order), completion on an
// mbarrier per chunk.  The copy travels while the warp does its round-3 arithmetic and stores, so the single buffer is
// double-used, there is no producer warp and no landing ring, and loads, stores and arithmetic of adjacent tiles overlap.
//   * pass 1 stores transposed (out[col * 4096 + e], 64-byte runs per half warp); pass 2 multiplies its INPUT by the
//     inter-pass twiddle w_n^(d * col) = A[d2][c] * beta[d1, d0][c] (two Shoup multiplications from per-tile tables of
//     128 + 256 entries — no Montgomery epilogue, no running products), canonicalises and stores 32-byte row segments.
// Shared-memory accesses are conflict free: a lane owns (row mod 16, half row), so every LDS.128 / STS.128 of a warp
// covers 16 consecutive 32-byte rows; the one strided access (round-1 stores, row stride 16) is spread over the
// banks by XOR-ing the low two row bits with the low two bits of the top hex digit.
#pragma once
#include <cuda.h>

#include "ntt_pass_v4.cuh"

namespace bb {

constexpr int V7_LR = 12;
constexpr int V7_R = 1 << V7_LR;
constexpr int V7_C = 8;                       // columns per tile (32-byte row segments)
constexpr int V7_CW = 16;                     // warps; warp w owns chunk w
constexpr int V7_NT = V7_CW * 32;
constexpr uint32_t V7_CHUNK_BYTES = 256 * V7_C * 4;
constexpr uint32_t V7_OFF_TILE = 0;                                      // rows (q1, q2, q0): chunk q1 is contiguous
constexpr uint32_t V7_OFF_TW = V7_R * V7_C * 4;                          // w_4096^i, i < 2048 (Shoup pairs)
constexpr uint32_t V7_OFF_TW2 = V7_OFF_TW + 2048 * 8;                    // round-2 twiddles, per-stage compact
constexpr uint32_t V7_OFF_A = V7_OFF_TW2 + 256 * 8;                      // A[d2][c]  Shoup pairs
constexpr uint32_t V7_OFF_U = V7_OFF_A + 16 * 8 * 8;                     // U[k][c]   plain
constexpr uint32_t V7_OFF_V = V7_OFF_U + 16 * 8 * 4;                     // V[r][c]   Montgomery form (carries the scale)
constexpr uint32_t V7_OFF_BAR = V7_OFF_V + 16 * 8 * 4;                   // one mbarrier per chunk
constexpr uint32_t V7_SMEM = V7_OFF_BAR + 16 * 8;

struct V7Params {
    uint32_t* out;
    unsigned long long out_batch_stride;  // u32 units
    uint32_t tiles_x;                     // column tiles per vector (ncols / 8)
    uint32_t total_tiles;                 // tiles_x * batch
    uint32_t log_pfull;                   // row store: out[(((j << 12) + e) << log_pfull) + low], col = j * pfull + low
    uint32_t exp_mask;                    // n - 1
    const uint2* tw;                      // (w, w') of omega_4096^i, i < 2048, this direction
    uint2 tw16[8];                        // omega_16^i, i < 8
    PowTable tab;                         // w_n^t
    PowTable tab_scaled;                  // w_n^t * scale (n^-1 of an inverse transform, else the same table)
    uint32_t* err;                        // device words for time-outs / diagnostics (may be null)
    // diagnostics (TOYNI_V7_FLAGS): 1 = skip the butterflies (memory traffic only), 2 = no TMA traffic (arithmetic only);
    // results are garbage in both modes
    uint32_t flags;
    uint32_t skew;                        // cycles of start-up stagger between the four warps that share a scheduler
};

// ---------------------------------------------------------------- PTX wrappers
BB_D void v7_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
BB_D void v7_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
BB_D void v7_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }
BB_D uint32_t v7_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// bounded wait: a lost transfer traps (a sticky launch failure the host reports) instead of hanging the GPU
BB_D void v7_mbar_wait(uint32_t bar, uint32_t parity, uint32_t* err, uint32_t code) {
    if (v7_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!v7_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
BB_D void v7_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

__host__ __device__ constexpr uint32_t v7_brev4(uint32_t k) { return ((k & 1u) << 3) | ((k & 2u) << 1) | ((k & 4u) >> 1) | ((k & 8u) >> 3); }

// The four warps that share a scheduler (warp, warp + 4, ...) run the same unrolled code and leave every barrier in
// lockstep; their multiply bursts (FMA-heavy pipe) then coincide and so do their add / min bursts (ALU pipe), and the
// two pipes take turns instead of overlapping.  A stagger of a fraction of a burst after each barrier keeps one warp's
// multiplies next to another's additions.
BB_D void v7_stagger(uint32_t cycles) {
    if (cycles == 0u) return;
    const uint32_t t0 = (uint32_t)clock();
    while ((uint32_t)clock() - t0 < cycles) {
    }
}

BB_D uint4 v7_shoup4(uint4 v, uint4 a01, uint4 a23) {  // four columns, four different constants (w, w') pairs
    v.x = shoup_mul_lazy(v.x, a01.x, a01.y);
    v.y = shoup_mul_lazy(v.y, a01.z, a01.w);
    v.z = shoup_mul_lazy(v.z, a23.x, a23.y);
    v.w = shoup_mul_lazy(v.w, a23.z, a23.w);
    return v;
}

// per-tile tables of the inter-pass twiddle w_n^(d * col), d = 256 d2 + 16 d1 + d0, col = col0 + c
BB_D void v7_tables(uint8_t* smem, const V7Params& p, uint32_t col0, uint32_t ctid) {
    const uint32_t c = ctid & 7u, i = (ctid >> 3) & 15u, col = col0 + c;
    if (ctid < 128u) {
        const uint32_t w = pow_plain(p.tab, (256u * i * col) & p.exp_mask);
        reinterpret_cast<uint2*>(smem + V7_OFF_A)[i * 8 + c] = make_uint2(w, shoup_companion_fast(w));
    } else if (ctid < 256u) {
        reinterpret_cast<uint32_t*>(smem + V7_OFF_U)[i * 8 + c] = pow_plain(p.tab, (16u * i * col) & p.exp_mask);
        reinterpret_cast<uint32_t*>(smem + V7_OFF_V)[i * 8 + c] = pow_lookup(p.tab_scaled, (i * col) & p.exp_mask);
    }
}

// ---------------------------------------------------------------- the three rounds (one work item per lane:
// 16 rows x 4 columns held as uint4 x[16]; register k holds hex digit brev4(k) on load and digit k on store).
// Tile row of position (q1, q2, q0): q1 * 256 + q2 * 16 + (q0 ^ (q2 & 3)), 32 bytes per row.
template <bool PASS2>
BB_D void v7_round1(uint8_t* smem, const V7Params& p, uint32_t kc, uint32_t r, uint32_t h) {
    uint4 x[16];
    uint8_t* reg = smem + V7_OFF_TILE + kc * V7_CHUNK_BYTES;
    // as landed: row d2 * 16 + d0 of the chunk; this lane takes d0 = r
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = *reinterpret_cast<const uint4*>(reg + (v7_brev4(k) * 16u + r) * 32u + h * 16u);
    if constexpr (PASS2) {
        const uint8_t* at = smem + V7_OFF_A + h * 32u;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint4 a01 = *reinterpret_cast<const uint4*>(at + v7_brev4(k) * 64u);
            const uint4 a23 = *reinterpret_cast<const uint4*>(at + v7_brev4(k) * 64u + 16u);
            x[k] = v7_shoup4(x[k], a01, a23);
        }
    }
    if (!(p.flags & 1u)) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << t)) continue;
                const int kp = k & ((1 << t) - 1);
                if (kp == 0)
                    bfly4_one(x[k], x[k + (1 << t)]);
                else
                    bfly4(x[k], x[k + (1 << t)], p.tw16[kp << (3 - t)]);
            }
        }
    }
    if constexpr (PASS2) {
        const uint4 u = *reinterpret_cast<const uint4*>(smem + V7_OFF_U + kc * 32u + h * 16u);
        const uint4 v = *reinterpret_cast<const uint4*>(smem + V7_OFF_V + r * 32u + h * 16u);
        uint4 b01, b23;
        b01.x = monty_mul(u.x, v.x); b01.y = shoup_companion_fast(b01.x);
        b01.z = monty_mul(u.y, v.y); b01.w = shoup_companion_fast(b01.z);
        b23.x = monty_mul(u.z, v.z); b23.y = shoup_companion_fast(b23.x);
        b23.z = monty_mul(u.w, v.w); b23.w = shoup_companion_fast(b23.z);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = v7_shoup4(x[k], b01, b23);
    }
    // in place within the chunk: every lane's loads are in registers (and consumed) before any lane stores
    __syncwarp();
    // position (q1, q2, q0) = (kc, r, k)
    uint8_t* base = reg + r * 512u + h * 16u + ((r & 3u) << 5);
    const uint32_t bo = (uint32_t)(base - smem);
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<uint4*>(smem + (bo ^ ((uint32_t)k << 5))) = x[k];
}

BB_D void v7_round2(uint8_t* smem, uint32_t flags, uint32_t q2, uint32_t r, uint32_t h) {
    uint4 x[16];
    uint8_t* base = smem + V7_OFF_TILE + (q2 * 16u + (r ^ (q2 & 3u))) * 32u + h * 16u;
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = *reinterpret_cast<const uint4*>(base + v7_brev4(k) * V7_CHUNK_BYTES);
    const uint2* tws = reinterpret_cast<const uint2*>(smem + V7_OFF_TW2) + r;
    if (!(flags & 1u)) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & (1 << t)) continue;
                const int kp = k & ((1 << t) - 1);
                bfly4(x[k], x[k + (1 << t)], tws[16 * ((1 << t) - 1) + 16 * kp]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<uint4*>(base + k * V7_CHUNK_BYTES) = x[k];
}

// round 3 of chunk kc of tile `tile`; once the chunk sits in registers lane 0 re-fills the region with chunk kc of
// tile `next` (if any) and the copy travels during the arithmetic and the stores below
template <bool PASS2>
BB_D void v7_round3(uint8_t* smem, const V7Params& p, const CUtensorMap* tmap, uint32_t tile, uint32_t next, uint32_t kc, uint32_t r, uint32_t h,
                    uint32_t lane, uint32_t bar) {
    uint4 x[16];
    const uint32_t base = V7_OFF_TILE + kc * V7_CHUNK_BYTES + r * 32u + h * 16u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t q2 = v7_brev4(k);
        x[k] = *reinterpret_cast<const uint4*>(smem + ((base ^ ((q2 & 3u) << 5)) + q2 * 512u));
    }
    const uint2* tw = reinterpret_cast<const uint2*>(smem + V7_OFF_TW);
    const uint32_t b = r + 16u * kc;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        if (t == 1) {
            // stage 0 has consumed every loaded register of every lane: the region may be overwritten
            __syncwarp();
            if (lane == 0 && next < p.total_tiles && !(p.flags & 2u)) {
                const uint32_t bz = next / p.tiles_x, tx = next - bz * p.tiles_x;
                v7_mbar_expect_tx(bar, V7_CHUNK_BYTES);
                v7_tma_load_5d(smem_u32(smem + V7_OFF_TILE + kc * V7_CHUNK_BYTES), tmap, bar, tx * V7_C, 0u, kc, 0u, bz);
            }
        }
        if (p.flags & 1u) continue;
        const uint2* tws = tw + (b << (3 - t));
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k & (1 << t)) continue;
            const int kp = k & ((1 << t) - 1);
            bfly4(x[k], x[k + (1 << t)], tws[(256 * kp) << (3 - t)]);
        }
    }
    const uint32_t bz = tile / p.tiles_x, tx = tile - bz * p.tiles_x;
    const uint32_t col = tx * V7_C + 4u * h;
    uint32_t* out = p.out + (size_t)bz * p.out_batch_stride;
    if constexpr (PASS2) {
        const uint32_t j = col >> p.log_pfull, low = col & ((1u << p.log_pfull) - 1u);
        uint32_t* o = out + ((((size_t)j << V7_LR) + b) << p.log_pfull) + low;
        const size_t step = (size_t)256 << p.log_pfull;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            *reinterpret_cast<uint4*>(o) = canon4(x[k]);
            o += step;
        }
    } else {
        uint32_t* o = out + (size_t)col * V7_R + b;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            o[k * 256 + 0 * V7_R] = x[k].x;
            o[k * 256 + 1 * V7_R] = x[k].y;
            o[k * 256 + 2 * V7_R] = x[k].z;
            o[k * 256 + 3 * V7_R] = x[k].w;
        }
    }
}

template <bool PASS2>
__global__ void __launch_bounds__(V7_NT, 1) ntt_pass_v7_kernel(const CUtensorMap* __restrict__ tmap, const V7Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t bar = smem_u32(smem + V7_OFF_BAR) + 8u * warp;  // this warp's chunk barrier

    {   // twiddle tables of the in-tile rounds
        uint2* tw_s = reinterpret_cast<uint2*>(smem + V7_OFF_TW);
        for (uint32_t i = tid; i < 2048u; i += V7_NT) tw_s[i] = __ldg(&p.tw[i]);
        uint2* tw2_s = reinterpret_cast<uint2*>(smem + V7_OFF_TW2);
        for (uint32_t i = tid; i < 240u; i += V7_NT) {  // stage t: omega_(32 << t)^(r + 16 kp) at [16 (2^t - 1) + 16 kp + r]
            const uint32_t t = (i >= 16u) + (i >= 48u) + (i >= 112u);
            const uint32_t rem = i - 16u * ((1u << t) - 1u);
            tw2_s[i] = __ldg(&p.tw[rem << (7u - t)]);
        }
        if (lane == 0) v7_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // programmatic dependent launch: the previous kernel of the stream (the previous pass) is complete from here on
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

    const uint32_t r = lane >> 1, h = lane & 1u;
    uint32_t next = blockIdx.x, cur = 0, it = 0;
    bool have_cur = false;
    if (next < p.total_tiles) {
        if (lane == 0 && !(p.flags & 2u)) {  // chunk `warp` of the first tile
            const uint32_t bz = next / p.tiles_x, tx = next - bz * p.tiles_x;
            v7_mbar_expect_tx(bar, V7_CHUNK_BYTES);
            v7_tma_load_5d(smem_u32(smem + V7_OFF_TILE + warp * V7_CHUNK_BYTES), tmap, bar, tx * V7_C, 0u, warp, 0u, bz);
        }
        if constexpr (PASS2) v7_tables(smem, p, (next % p.tiles_x) * V7_C, tid);
    }
    __syncthreads();
    const uint32_t lag = (warp >> 2) * p.skew;
    while (true) {
        const bool have_next = next < p.total_tiles;
        v7_stagger(lag);
        if (have_cur) v7_round3<PASS2>(smem, p, tmap, cur, next, warp, r, h, lane, bar);
        if (!have_next) break;
        if (!(p.flags & 2u)) v7_mbar_wait(bar, it & 1u, p.err, 0x60000000u | (it << 4) | warp);
        v7_round1<PASS2>(smem, p, warp, r, h);
        __syncthreads();
        const uint32_t nn = next + gridDim.x;
        if constexpr (PASS2) {
            if (nn < p.total_tiles) v7_tables(smem, p, (nn % p.tiles_x) * V7_C, tid);
        }
        v7_stagger(lag);
        v7_round2(smem, p.flags, warp, r, h);
        __syncthreads();
        cur = next;
        have_cur = true;
        next = nn;
        it++;
    }
}

}  // namespace bb
