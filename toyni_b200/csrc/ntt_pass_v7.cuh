// TMA-staged NTT pass for the two-pass plan of a 2^24-point transform (4096 x 4096).
//
// Replaces, from scratch, the reference's bit-reverse + one-launch-per-stage + scale kernels
// (cuda/ntt_kernel.cu:103-143, 249-292); results are bit-exact with src/ntt.rs:24-66.
//
// A persistent CTA owns ONE tile buffer of 4096 rows x C columns of u32 (C = 8: 128 KB, sixteen warps, one CTA per SM;
// C = 4: 64 KB, eight warps, two independent CTAs per SM) and walks its tiles.  The buffer is sixteen chunks (the 256
// rows whose middle hex digit is q1, one contiguous region each); a lane owns (row mod 16, four columns) of one chunk,
// so a warp owns one chunk (C = 8) or two (C = 4).  Per tile, three radix-16 decimation-in-time rounds:
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

template <int C>
struct V7 {
    static_assert(C == 4 || C == 8, "4 or 8 columns per tile");
    static constexpr int WARPS = 2 * C;                  // 32 lanes x (16 rows x 4 columns) = one round of the tile
    static constexpr int NT = WARPS * 32;
    static constexpr int CTAS_PER_SM = C == 8 ? 1 : 2;
    static constexpr uint32_t RB = C * 4;                // bytes per tile row
    static constexpr uint32_t CHB = 256 * RB;            // bytes per chunk
    static constexpr uint32_t OFF_TILE = 0;              // rows (q1, q2, q0): chunk q1 is contiguous
    static constexpr uint32_t OFF_TW = V7_R * RB;        // w_4096^i, i < 2048 (Shoup pairs)
    static constexpr uint32_t OFF_TW2 = OFF_TW + 2048 * 8;   // round-2 twiddles, per-stage compact
    static constexpr uint32_t OFF_A = OFF_TW2 + 256 * 8;     // A[d2][c]  Shoup pairs
    static constexpr uint32_t OFF_U = OFF_A + 16 * C * 8;    // U[k][c]   plain
    static constexpr uint32_t OFF_V = OFF_U + 16 * C * 4;    // V[r][c]   Montgomery form (carries the scale)
    static constexpr uint32_t OFF_BAR = OFF_V + 16 * C * 4;  // one mbarrier per warp
    static constexpr uint32_t STAGE = 128 * RB;              // per warp: half a chunk of results on its way out (TMA store)
    static constexpr uint32_t OFF_STAGE = (OFF_BAR + WARPS * 8 + 127u) & ~127u;
    static constexpr uint32_t SMEM = OFF_STAGE + WARPS * STAGE;
    // physical q0 of position (q1, q2, q0) is q0 ^ swz(q2): the strided round-1 stores (row stride 16) then spread over
    // all banks — a quarter warp covers 128 bytes = 4 rows of 32 bytes (C = 8) or 8 rows of 16 bytes (C = 4)
    __host__ __device__ static constexpr uint32_t swz(uint32_t q2) { return q2 & (C == 8 ? 3u : 7u); }
};

struct V7Params {
    uint32_t* out;
    unsigned long long out_batch_stride;  // u32 units
    uint32_t tiles_x;                     // column tiles per vector (ncols / C)
    uint32_t total_tiles;                 // tiles_x * batch
    uint32_t log_pfull;                   // row store: out[(((j << 12) + e) << log_pfull) + low], col = j * pfull + low
    uint32_t exp_mask;                    // n - 1
    const uint2* tw;                      // (w, w') of omega_4096^i, i < 2048, this direction
    uint2 tw16[8];                        // omega_16^i, i < 8
    PowTable tab;                         // w_n^t
    PowTable tab_scaled;                  // w_n^t * scale (n^-1 of an inverse transform, else the same table)
    uint32_t* err;                        // device words for time-outs / diagnostics (may be null)
    // diagnostics (TOYNI_V7_FLAGS): 1 = skip the butterflies (memory traffic only), 2 = no TMA loads (arithmetic only),
    // 8 = no global stores; results are garbage in all three modes (tools/v7_time.py, profiles/v7_decomposition_r2.txt)
    uint32_t flags;
};

// ---------------------------------------------------------------- PTX wrappers
BB_D void v7_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory"); }
BB_D void v7_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
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
BB_D void v7_tma_store_5d(const CUtensorMap* map, uint32_t src, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
BB_D void v7_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

__host__ __device__ constexpr uint32_t v7_brev4(uint32_t k) { return ((k & 1u) << 3) | ((k & 2u) << 1) | ((k & 4u) >> 1) | ((k & 8u) >> 3); }

BB_D uint4 v7_shoup4(uint4 v, uint4 a01, uint4 a23) {  // four columns, four different constants (w, w') pairs
    v.x = shoup_mul_lazy(v.x, a01.x, a01.y);
    v.y = shoup_mul_lazy(v.y, a01.z, a01.w);
    v.z = shoup_mul_lazy(v.z, a23.x, a23.y);
    v.w = shoup_mul_lazy(v.w, a23.z, a23.w);
    return v;
}

// per-tile tables of the inter-pass twiddle w_n^(d * col), d = 256 d2 + 16 d1 + d0, col = col0 + c
template <int C>
BB_D void v7_tables(uint8_t* smem, const V7Params& p, uint32_t col0, uint32_t tid) {
    using T = V7<C>;
    const uint32_t c = tid & (C - 1u), i = (tid / C) & 15u, col = col0 + c;
    if (tid < 16u * C) {
        const uint32_t w = pow_plain(p.tab, (256u * i * col) & p.exp_mask);
        reinterpret_cast<uint2*>(smem + T::OFF_A)[i * C + c] = make_uint2(w, shoup_companion_fast(w));
    } else if (tid < 32u * C) {
        reinterpret_cast<uint32_t*>(smem + T::OFF_U)[i * C + c] = pow_plain(p.tab, (16u * i * col) & p.exp_mask);
        reinterpret_cast<uint32_t*>(smem + T::OFF_V)[i * C + c] = pow_lookup(p.tab_scaled, (i * col) & p.exp_mask);
    }
}

// this lane's place in the tile: row mod 16, chunk, byte offset of its four columns within a row
template <int C>
struct V7Lane {
    uint32_t r, kc, cq;
    BB_D V7Lane(uint32_t warp, uint32_t lane) {
        // C = 8: lane = 2 r + column quad, one chunk per warp; C = 4: lane = 16 (chunk select) + r, two chunks per warp.
        // Either way a quarter warp reads 128 contiguous bytes of one chunk.
        r = C == 8 ? lane >> 1 : lane & 15u;
        kc = C == 8 ? warp : 2u * warp + (lane >> 4);
        cq = C == 8 ? 16u * (lane & 1u) : 0u;
    }
};

// ---------------------------------------------------------------- the three rounds (one work item per lane:
// 16 rows x 4 columns held as uint4 x[16]; register k holds hex digit brev4(k) on load and digit k on store).
// Tile row of position (q1, q2, q0): q1 * 256 + q2 * 16 + (q0 ^ swz(q2)).
template <bool PASS2, int C>
BB_D void v7_round1(uint8_t* smem, const V7Params& p, const V7Lane<C>& ln) {
    using T = V7<C>;
    uint4 x[16];
    const uint32_t r = ln.r, kc = ln.kc;
    uint8_t* reg = smem + T::OFF_TILE + kc * T::CHB;
    // as landed: row d2 * 16 + d0 of the chunk; this lane takes d0 = r
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = *reinterpret_cast<const uint4*>(reg + (v7_brev4(k) * 16u + r) * T::RB + ln.cq);
    if constexpr (PASS2) {
        const uint8_t* at = smem + T::OFF_A + ln.cq * 2u;
#pragma unroll
        for (int k = 1; k < 16; k++) {
            const uint4 a01 = *reinterpret_cast<const uint4*>(at + v7_brev4(k) * (C * 8u));
            const uint4 a23 = *reinterpret_cast<const uint4*>(at + v7_brev4(k) * (C * 8u) + 16u);
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
        const uint4 u = *reinterpret_cast<const uint4*>(smem + T::OFF_U + kc * (C * 4u) + ln.cq);
        const uint4 v = *reinterpret_cast<const uint4*>(smem + T::OFF_V + r * (C * 4u) + ln.cq);
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
    const uint32_t bo = T::OFF_TILE + kc * T::CHB + r * 16u * T::RB + ln.cq + T::swz(r) * T::RB;
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<uint4*>(smem + (bo ^ ((uint32_t)k * T::RB))) = x[k];
}

// rows (q1 = 0..15, q2, q0 = r): in place
template <int C>
BB_D void v7_round2(uint8_t* smem, uint32_t flags, uint32_t q2, const V7Lane<C>& ln) {
    using T = V7<C>;
    uint4 x[16];
    const uint32_t r = ln.r;
    uint8_t* base = smem + T::OFF_TILE + (q2 * 16u + (r ^ T::swz(q2))) * T::RB + ln.cq;
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = *reinterpret_cast<const uint4*>(base + v7_brev4(k) * T::CHB);
    const uint2* tws = reinterpret_cast<const uint2*>(smem + T::OFF_TW2) + r;
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
    for (int k = 0; k < 16; k++) *reinterpret_cast<uint4*>(base + k * T::CHB) = x[k];
}

// round 3 of this lane's chunk of tile `tile`; once the warp's chunk(s) sit in registers lane 0 re-fills the region with
// the same chunk(s) of tile `next` (if any) and the copy travels during the arithmetic and the stores below
template <bool PASS2, int C>
BB_D void v7_round3(uint8_t* smem, const V7Params& p, const CUtensorMap* tmap, const CUtensorMap* omap, uint32_t tile, uint32_t next,
                    const V7Lane<C>& ln, uint32_t warp, uint32_t lane, uint32_t bar) {
    using T = V7<C>;
    uint4 x[16];
    const uint32_t r = ln.r, kc = ln.kc;
    const uint32_t base = T::OFF_TILE + kc * T::CHB + r * T::RB + ln.cq;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t q2 = v7_brev4(k);
        x[k] = *reinterpret_cast<const uint4*>(smem + ((base ^ (T::swz(q2) * T::RB)) + q2 * 16u * T::RB));
    }
    const uint2* tw = reinterpret_cast<const uint2*>(smem + T::OFF_TW);
    const uint32_t b = r + 16u * kc;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        if (t == 1) {
            // stage 0 has consumed every loaded register of every lane: the region may be overwritten
            __syncwarp();
            if (lane == 0 && next < p.total_tiles && !(p.flags & 2u)) {
                const uint32_t bz = next / p.tiles_x, tx = next - bz * p.tiles_x;
                constexpr uint32_t NCH = C == 8 ? 1u : 2u;  // chunks per warp
                v7_mbar_expect_tx(bar, NCH * T::CHB);
#pragma unroll
                for (uint32_t c = 0; c < NCH; c++) {
                    const uint32_t ch = NCH * warp + c;
                    v7_tma_load_5d(smem_u32(smem + T::OFF_TILE + ch * T::CHB), tmap, bar, tx * C, 0u, ch, 0u, bz);
                }
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
    if (p.flags & 8u) {  // diagnostic: no global stores (one predicated-off store keeps the results live)
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) acc ^= x[k].x ^ x[k].y ^ x[k].z ^ x[k].w;
        if (acc == 0x13572468u) p.out[0] = acc;
        return;
    }
    const uint32_t bz = tile / p.tiles_x, tx = tile - bz * p.tiles_x;
    const uint32_t col = tx * C + (ln.cq >> 2);
    uint32_t* out = p.out + (size_t)bz * p.out_batch_stride;
    if constexpr (PASS2 && C == 8) {
        if (omap) {
            // Row store through the TMA: the chunk's 256 result rows (32 bytes each, 16 KB apart in global memory) leave as
            // two boxes {8 columns, 16 x e0, e1 = kc, 8 x e2} from a per-warp staging buffer, so the scattered 32-byte
            // segments cost no LSU / L1 store slots (16 STG.128 per lane, 16 sectors each, were 13 us of a transform).
            uint8_t* stage = smem + T::OFF_STAGE + warp * T::STAGE;
            const uint32_t stage_s = smem_u32(stage);
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // previous tile's boxes have left
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int k = 0; k < 8; k++) *reinterpret_cast<uint4*>(stage + (k * 16u + r) * T::RB + ln.cq) = canon4(x[8 * half + k]);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the TMA
                __syncwarp();
                if (lane == 0) {
                    v7_tma_store_5d(omap, stage_s, tx * C, 0u, kc, 8u * half, bz);
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    if (half == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                }
                __syncwarp();
            }
            return;
        }
    }
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

template <bool PASS2, int C>
__global__ void __launch_bounds__(V7<C>::NT, V7<C>::CTAS_PER_SM) ntt_pass_v7_kernel(const CUtensorMap* __restrict__ tmap, const CUtensorMap* __restrict__ omap, const V7Params p) {
    using T = V7<C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t bar = smem_u32(smem + T::OFF_BAR) + 8u * warp;  // this warp's chunk barrier
    constexpr uint32_t NCH = C == 8 ? 1u : 2u;

    {   // twiddle tables of the in-tile rounds
        uint2* tw_s = reinterpret_cast<uint2*>(smem + T::OFF_TW);
        for (uint32_t i = tid; i < 2048u; i += T::NT) tw_s[i] = __ldg(&p.tw[i]);
        uint2* tw2_s = reinterpret_cast<uint2*>(smem + T::OFF_TW2);
        for (uint32_t i = tid; i < 240u; i += T::NT) {  // stage t: omega_(32 << t)^(r + 16 kp) at [16 (2^t - 1) + 16 kp + r]
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

    const V7Lane<C> ln(warp, lane);
    const uint32_t q2_r2 = C == 8 ? warp : 2u * warp + (lane >> 4);  // round 2: this lane's top digit
    uint32_t next = blockIdx.x, cur = 0, it = 0;
    bool have_cur = false;
    if (next < p.total_tiles) {
        if (lane == 0 && !(p.flags & 2u)) {  // this warp's chunk(s) of the first tile
            const uint32_t bz = next / p.tiles_x, tx = next - bz * p.tiles_x;
            v7_mbar_expect_tx(bar, NCH * T::CHB);
#pragma unroll
            for (uint32_t c = 0; c < NCH; c++) {
                const uint32_t ch = NCH * warp + c;
                v7_tma_load_5d(smem_u32(smem + T::OFF_TILE + ch * T::CHB), tmap, bar, tx * C, 0u, ch, 0u, bz);
            }
        }
        if constexpr (PASS2) v7_tables<C>(smem, p, (next % p.tiles_x) * C, tid);
    }
    __syncthreads();
    while (true) {
        const bool have_next = next < p.total_tiles;
        if (have_cur) v7_round3<PASS2, C>(smem, p, tmap, omap, cur, next, ln, warp, lane, bar);
        if (!have_next) break;
        if (!(p.flags & 2u)) v7_mbar_wait(bar, it & 1u, p.err, 0x60000000u | (it << 4) | warp);
        v7_round1<PASS2, C>(smem, p, ln);
        __syncthreads();
        const uint32_t nn = next + gridDim.x;
        if constexpr (PASS2) {
            if (nn < p.total_tiles) v7_tables<C>(smem, p, (nn % p.tiles_x) * C, tid);
        }
        v7_round2<C>(smem, p.flags, q2_r2, ln);
        __syncthreads();
        cur = next;
        have_cur = true;
        next = nn;
        it++;
    }
    if (PASS2 && omap && lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // the last boxes are in global memory
}

}  // namespace bb
