// Vectorised NTT pass: the hot kernel.  Same mathematics and parameters as ntt_pass.cuh (which stays as the
// scalar path for batches of contiguous short vectors), but the tile is kept row-major in shared memory as
// 16-byte chunks of four adjacent columns, so that
//   * global -> shared is one 16-byte cp.async per chunk (no register staging, the whole tile in flight at once),
//   * every radix-16 round moves data with LDS.128 / STS.128 and the four columns of a chunk share their twiddles,
//   * shared -> global is STG.128 (row-granular passes) or four coalesced STG.32 (the transposing first pass).
// Bank conflicts are removed by an XOR swizzle on the chunk index that is GF(2)-linear in (row, chunk), so the
// address of row (row0 | k*S) is base ^ constant_k; all loops are fully unrolled so those constants, the
// bit-reversed row numbers and the global strides fold at compile time.
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
#ifndef BB_VW
#define BB_VW 2
#endif
    // a work item of a round is 16 rows x VW adjacent columns; VW = 2 halves the register footprint of VW = 4
    // (32 instead of 64 live values), which doubles the warps per SM at the same tile size
    static constexpr int VW = ((R >> G1) * (C / BB_VW) <= 512) ? BB_VW : 4;
    static constexpr int LVW = (VW == 4) ? 2 : 1;
    static constexpr int ITEMS = (R >> G1) * (C / VW);  // work items of the first round
    static constexpr int NT = ITEMS < 32 ? 32 : (ITEMS > 512 ? 512 : ITEMS);
    static constexpr int CHUNKS = R * CV;
    static constexpr int RS = NT >> LCV;  // rows covered by one sweep of the CTA over chunks (when > 0)
    static constexpr size_t SMEM = (size_t)R * C * 4;
    static constexpr size_t TW_BYTES = (R >= 32) ? (size_t)(R / 2) * 8 : 0;  // omega_R^i, i < R/2, as Shoup pairs

    // physical 16-byte chunk index of (row, cv); GF(2)-linear in its arguments
    __host__ __device__ static constexpr uint32_t chunk(uint32_t row, uint32_t cv) {
        constexpr uint32_t qm = (1u << Q) - 1u;
        uint32_t rin = (row & qm) ^ ((row >> 4) & qm);
        uint32_t cvv = cv ^ ((row >> Q) & (CV - 1u));
        return ((row >> Q) << 3) | (rin << LCV) | cvv;
    }
    __host__ __device__ static constexpr uint32_t brev(uint32_t r) {  // LR-bit reversal
        uint32_t x = 0;
        for (int i = 0; i < LR; i++) x |= ((r >> i) & 1u) << (LR - 1 - i);
        return x;
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

template <int VW>
struct Vec;
template <>
struct Vec<4> {
    typedef uint4 type;
};
template <>
struct Vec<2> {
    typedef uint2 type;
};

template <int LR, int LC, int S_LOG, int G_LOG>
BB_D void dit_round_v4(uint4* __restrict__ sm, const uint2* __restrict__ stw, const PassParams& p) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, NT = T::NT, S = 1 << S_LOG, G = 1 << G_LOG, VW = T::VW;
    constexpr int CW = T::C / VW;  // items per row
    constexpr int ITEMS = (R / G) * CW;
    typedef typename Vec<VW>::type vec_t;
    vec_t* __restrict__ smw = reinterpret_cast<vec_t*>(sm);
#pragma unroll 1
    for (int it = threadIdx.x; it < ITEMS; it += NT) {
        const uint32_t cw = it & (CW - 1);
        const uint32_t cv = (VW == 4) ? cw : (cw >> 1), sub = (VW == 4) ? 0u : (cw & 1u);
        const uint32_t rg = it / CW;
        const uint32_t b = rg & (S - 1), blk = rg >> S_LOG;
        const uint32_t row0 = blk * (G * S) + b;
        const uint32_t base = (VW == 4) ? T::chunk(row0, cv) : ((T::chunk(row0, cv) << 1) | sub);
        uint32_t x[G][VW];
#pragma unroll
        for (int k = 0; k < G; k++) {
            const vec_t v = smw[base ^ (T::chunk((uint32_t)k << S_LOG, 0) << (VW == 4 ? 0 : 1))];
            if constexpr (VW == 4) {
                x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
            } else {
                x[k][0] = v.x; x[k][1] = v.y;
            }
        }
#pragma unroll
        for (int t = 0; t < G_LOG; t++) {
            // stage t twiddle kp is omega_R^((b + kp*S) << sh), read from the CTA's shared-memory copy of the table:
            // one base address per stage, kp folds into the immediate offset
            const int sh = (S_LOG == 0) ? 0 : (LR - (S_LOG + t + 1));
            const uint2* tws = stw + (b << sh);
#pragma unroll
            for (int k = 0; k < G; k++) {
                if (k & (1 << t)) continue;
                const int kp = k & ((1 << t) - 1);
                if (S_LOG == 0) {
                    if (kp == 0) {
#pragma unroll
                        for (int c = 0; c < VW; c++) bfly_one(x[k][c], x[k + (1 << t)][c]);
                    } else {
                        const uint2 w = p.tw16[kp << (3 - t)];
#pragma unroll
                        for (int c = 0; c < VW; c++) bfly(x[k][c], x[k + (1 << t)][c], w);
                    }
                } else {
                    const uint2 w = tws[(kp * S) << sh];
#pragma unroll
                    for (int c = 0; c < VW; c++) bfly(x[k][c], x[k + (1 << t)][c], w);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            vec_t v;
            if constexpr (VW == 4) {
                v.x = x[k][0]; v.y = x[k][1]; v.z = x[k][2]; v.w = x[k][3];
            } else {
                v.x = x[k][0]; v.y = x[k][1];
            }
            smw[base ^ (T::chunk((uint32_t)k << S_LOG, 0) << (VW == 4 ? 0 : 1))] = v;
        }
    }
}

BB_D uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

BB_D uint4 mul4(uint4 v, uint32_t w_m) {
    return make_uint4(monty_mul(v.x, w_m), monty_mul(v.y, w_m), monty_mul(v.z, w_m), monty_mul(v.w, w_m));
}
BB_D uint4 canon4(uint4 v) { return make_uint4(min(v.x, v.x - P), min(v.y, v.y - P), min(v.z, v.z - P), min(v.w, v.w - P)); }

// ---- epilogue + store for passes whose chunks stay whole in the output (log_pfull >= 2)
template <int LR, int LC, uint32_t EPI>
BB_D void store_rows_v4(const uint4* __restrict__ smv, uint32_t* __restrict__ out, const PassParams& p, uint32_t col0) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, CV = T::CV, LCV = T::LCV, NT = T::NT;
    const uint32_t tid = threadIdx.x;
    const uint32_t log_pfull = p.log_pfull;
    if (log_pfull >= (uint32_t)LC) {
        // the whole tile shares j: row e is one contiguous segment of C values
        const uint32_t j = col0 >> log_pfull, low = (col0 & ((1u << log_pfull) - 1u));
        if constexpr (T::RS > 0 && (T::CHUNKS % NT) == 0) {
            const uint32_t cv = tid & (CV - 1), e0 = tid >> LCV;
            const uint32_t sbase = T::chunk(e0, cv);
            uint32_t* o = out + ((((size_t)j << LR) + e0) << log_pfull) + low + 4u * cv;
            const size_t step = (size_t)T::RS << log_pfull;
            uint32_t tw = 0, twstep = 0;  // running inter-pass twiddle g^e, g = omega^(j << shift), Montgomery form
            if (EPI == EPI_TWIDDLE) {
                tw = pow_lookup(p.epi, (j * e0) << p.epi_shift);
                // g^RS; the table may carry a constant factor (n^-1 of an inverse transform), taken out again here
                twstep = monty_mul(pow_lookup(p.epi, (j * (uint32_t)T::RS) << p.epi_shift), p.epi_unscale);
            }
#pragma unroll
            for (int it = 0; it < T::CHUNKS / NT; it++) {
                uint4 v = smv[sbase ^ T::chunk((uint32_t)(it * T::RS), 0)];
                if (EPI == EPI_TWIDDLE) {
                    v = mul4(v, tw);
                    tw = monty_mul(tw, twstep);
                } else if (EPI == EPI_OUTIDX) {
                    const uint32_t k = (((uint32_t)(it * T::RS) + e0) << log_pfull) + low + 4u * cv;
                    if (p.log_inner >= 2) {
                        v = mul4(v, pow_lookup(p.epi, k >> p.log_inner));
                    } else {
                        v.x = pow_apply(p.epi, (k + 0) >> p.log_inner, v.x);
                        v.y = pow_apply(p.epi, (k + 1) >> p.log_inner, v.y);
                        v.z = pow_apply(p.epi, (k + 2) >> p.log_inner, v.z);
                        v.w = pow_apply(p.epi, (k + 3) >> p.log_inner, v.w);
                    }
                } else if (EPI == EPI_CONST) {
                    v = mul4(v, p.epi_const);
                } else {
                    v = canon4(v);
                }
                *reinterpret_cast<uint4*>(o) = v;
                o += step;
            }
            return;
        }
    }
    // general form: runs of min(C, pfull) columns are contiguous in the output
    const uint32_t pfull_mask = (1u << log_pfull) - 1u;
    const uint32_t log_clv = (log_pfull < (uint32_t)LC ? log_pfull : (uint32_t)LC) - 2;
    const uint32_t j0 = col0 >> log_pfull, low0 = col0 & pfull_mask;
#pragma unroll 2
    for (int i = tid; i < R * CV; i += NT) {
        const uint32_t lv = i & ((1u << log_clv) - 1u);
        const uint32_t e = (i >> log_clv) & (R - 1);
        const uint32_t jj = i >> (log_clv + LR);
        const uint32_t cv = (jj << log_clv) + lv;
        uint4 v = smv[T::chunk(e, cv)];
        const uint32_t j = j0 + jj, low = low0 + 4u * lv;
        if (EPI == EPI_TWIDDLE) {
            v = mul4(v, pow_lookup(p.epi, (j * e) << p.epi_shift));
        } else if (EPI == EPI_OUTIDX) {
            const uint32_t k = (e << log_pfull) + low;
            if (p.log_inner >= 2) {
                v = mul4(v, pow_lookup(p.epi, k >> p.log_inner));
            } else {
                v.x = pow_apply(p.epi, (k + 0) >> p.log_inner, v.x);
                v.y = pow_apply(p.epi, (k + 1) >> p.log_inner, v.y);
                v.z = pow_apply(p.epi, (k + 2) >> p.log_inner, v.z);
                v.w = pow_apply(p.epi, (k + 3) >> p.log_inner, v.w);
            }
        } else if (EPI == EPI_CONST) {
            v = mul4(v, p.epi_const);
        } else {
            v = canon4(v);
        }
        *reinterpret_cast<uint4*>(out + ((((size_t)j << LR) + e) << log_pfull) + low) = v;
    }
}

// ---- last pass of the column transforms of a sharded four-step NTT: inter-half twiddle + transpose in one go.
//      Row k1 of the local block belongs to rank k1 / rows_per_rank; it is written straight into that rank's
//      receive buffer (a peer mapping, i.e. stores that travel over NVLink) at the column range of this rank, so the
//      all-to-all exchange and the re-layout pass disappear into the NTT pass's own stores.
template <int LR, int LC>
BB_D void store_fourstep_v4(const uint4* __restrict__ smv, const PassParams& p, uint32_t col0) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, CV = T::CV, LCV = T::LCV, NT = T::NT;
    const uint32_t log_p = p.log_pfull - p.log_inner;
    const uint32_t low_p = col0 >> p.log_inner, c0 = col0 & ((1u << p.log_inner) - 1u);
    const uint32_t rpr_mask = (1u << p.fs_log_rows_per_rank) - 1u;
#pragma unroll 2
    for (int i = threadIdx.x; i < R * CV; i += NT) {
        const uint32_t cv = i & (CV - 1), e = i >> LCV;
        uint4 v = smv[T::chunk(e, cv)];
        const uint32_t k1 = (e << log_p) + low_p, c = c0 + 4u * cv;
        uint32_t tw = pow_lookup(p.epi, (p.fs_col_offset + c) * k1);  // geometric in c, ratio w^k1
        const uint32_t g = pow_lookup(p.epi, k1);
        if (p.epi_const) tw = monty_mul(tw, p.epi_const);                // n1^-1 of a single-pass inverse
        v.x = monty_mul(v.x, tw); tw = monty_mul(tw, g);
        v.y = monty_mul(v.y, tw); tw = monty_mul(tw, g);
        v.z = monty_mul(v.z, tw); tw = monty_mul(tw, g);
        v.w = monty_mul(v.w, tw);
        uint32_t* dst = p.fs_peer[k1 >> p.fs_log_rows_per_rank] + (size_t)(k1 & rpr_mask) * p.fs_dst_row_stride + p.fs_dst_col + c;
        *reinterpret_cast<uint4*>(dst) = v;
    }
}

// ---- epilogue + store for the first pass of a plain vector (pfull == 1): column col becomes the contiguous run
//      out[col*R + e]; lanes walk e, so each of the four scalar stores of a chunk is coalesced.  The inter-pass
//      twiddle w^(col*e) is generated on chip: for a fixed row e it is a geometric sequence in col, so a thread
//      that owns row e for all C columns needs two table lookups (w^(col0*e), w^e) and one multiply per value.
template <int LR, int LC, uint32_t EPI>
BB_D void store_cols_v4(const uint4* __restrict__ smv, uint32_t* __restrict__ out, const PassParams& p, uint32_t col0) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, CV = T::CV, NT = T::NT;
    const uint32_t tid = threadIdx.x;
    if constexpr (NT <= R && (R % NT) == 0) {
#pragma unroll 1
        for (uint32_t e = tid; e < (uint32_t)R; e += NT) {
            uint32_t tw = 0, g = 0;
            if (EPI == EPI_TWIDDLE) {
                tw = pow_lookup(p.epi, (col0 * e) << p.epi_shift);                       // carries the table's constant factor
                g = monty_mul(pow_lookup(p.epi, e << p.epi_shift), p.epi_unscale);     // w^e without it
            }
            uint32_t* o = out + ((size_t)col0 << LR) + e;
#pragma unroll
            for (int cv = 0; cv < CV; cv++) {
                uint4 v = smv[T::chunk(e, cv)];
                if (EPI == EPI_TWIDDLE) {
                    v.x = monty_mul(v.x, tw); tw = monty_mul(tw, g);
                    v.y = monty_mul(v.y, tw); tw = monty_mul(tw, g);
                    v.z = monty_mul(v.z, tw); tw = monty_mul(tw, g);
                    v.w = monty_mul(v.w, tw); tw = monty_mul(tw, g);
                } else if (EPI == EPI_CONST) {
                    v = mul4(v, p.epi_const);
                } else {
                    v = canon4(v);
                }
                o[(size_t)(4 * cv + 0) << LR] = v.x;
                o[(size_t)(4 * cv + 1) << LR] = v.y;
                o[(size_t)(4 * cv + 2) << LR] = v.z;
                o[(size_t)(4 * cv + 3) << LR] = v.w;
            }
        }
        return;
    }
#pragma unroll 4
    for (int i = tid; i < R * CV; i += NT) {
        const uint32_t e = i & (R - 1), cv = i >> LR;
        uint4 v = smv[T::chunk(e, cv)];
        const uint32_t col = col0 + 4u * cv;
        if (EPI == EPI_TWIDDLE) {
            const uint32_t t0 = (col * e) << p.epi_shift, dt = e << p.epi_shift;
            v.x = pow_apply(p.epi, t0, v.x);
            v.y = pow_apply(p.epi, t0 + dt, v.y);
            v.z = pow_apply(p.epi, t0 + 2 * dt, v.z);
            v.w = pow_apply(p.epi, t0 + 3 * dt, v.w);
        } else if (EPI == EPI_CONST) {
            v = mul4(v, p.epi_const);
        } else {
            v = canon4(v);
        }
        uint32_t* o = out + ((size_t)col << LR) + e;
        o[0] = v.x;
        o[(size_t)1 << LR] = v.y;
        o[(size_t)2 << LR] = v.z;
        o[(size_t)3 << LR] = v.w;
    }
}

// global -> shared for one tile (asynchronous; the caller commits / waits)
template <int LR, int LC>
BB_D void load_tile_v4(uint4* __restrict__ buf, const uint32_t* __restrict__ in, const PassParams& p, uint32_t col0) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, C = T::C, CV = T::CV, LCV = T::LCV, NT = T::NT;
    const uint32_t tid = threadIdx.x;
    const uint32_t ncols = p.ncols;
    const bool full = (unsigned long long)(R - 1) * ncols + col0 + C <= p.n_in_limit;  // no zero padding in this tile
    if constexpr (T::RS > 0 && (T::CHUNKS % NT) == 0) {
        if (full) {
            const uint32_t cv = tid & (CV - 1), r0 = tid >> LCV;
            const uint32_t sm0 = smem_u32(buf), cbase = T::chunk(r0, cv);
            const uint32_t d0 = (LR == 0) ? 0u : (__brev(r0) >> (32 - LR));
            // row r0 | it*RS (disjoint bits): the swizzle and the bit reversal split into thread part ^ constant.
            // brev(it*RS) runs over 0..NIT-1, so the NIT source rows are src0 + k*ncols: built by pointer increments
            // (64-bit multiplies would occupy the FMA-heavy pipe, which the butterflies need)
            constexpr int NIT = T::CHUNKS / NT;
            const uint32_t* rowp[NIT];
            rowp[0] = in + (size_t)d0 * ncols + col0 + 4u * cv;
#pragma unroll
            for (int k = 1; k < NIT; k++) rowp[k] = rowp[k - 1] + ncols;
#pragma unroll
            for (int it = 0; it < NIT; it++) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sm0 + ((cbase ^ T::chunk((uint32_t)(it * T::RS), 0)) << 4)),
                             "l"(rowp[T::brev((uint32_t)(it * T::RS))])
                             : "memory");
            }
            return;
        }
    }
    // zero-padded input (the coset LDE reads n_coeffs << size values): bytes past the limit are zero-filled
#pragma unroll 1
    for (int i = tid; i < R * CV; i += NT) {
        const uint32_t cv = i & (CV - 1), r = i >> LCV;
        const uint32_t d = (LR == 0) ? 0u : (__brev(r) >> (32 - LR));
        const unsigned long long lidx = (unsigned long long)d * ncols + col0 + 4u * cv;
        const unsigned long long remw = p.n_in_limit > lidx ? p.n_in_limit - lidx : 0ull;  // valid words from here
        const uint32_t bytes = remw >= 4ull ? 16u : (uint32_t)remw * 4u;
        const uint32_t* src = in + (bytes ? lidx : 0ull);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(&buf[T::chunk(r, cv)])), "l"(src), "r"(bytes)
                     : "memory");
    }
}

// Persistent CTAs walk the tiles (tile = batch index * tiles_per_vector + column tile).  With DB the next tile's
// cp.async traffic is in flight while the current one is transformed, so the DRAM latency of the load is hidden
// even at 2-4 warps per CTA.
template <int LR, int LC, bool DB>
__global__ void __launch_bounds__(V4<LR, LC>::NT) ntt_pass_v4_kernel(const PassParams p, uint32_t tiles_x, uint32_t total_tiles) {
    using T = V4<LR, LC>;
    constexpr int R = T::R, C = T::C, CV = T::CV, LCV = T::LCV, NT = T::NT;
    extern __shared__ uint4 smv_all[];

    const uint32_t tid = threadIdx.x;
    const uint32_t ncols = p.ncols;
    uint32_t tile = blockIdx.x;
    if (tile >= total_tiles) return;
    // compact twiddle table omega_R^i (i < R/2) in shared memory, after the tile buffer(s): every generic-round
    // twiddle of this CTA's whole life comes from here (LDS latency) instead of the L1/L2 path
    uint2* stw = reinterpret_cast<uint2*>(smv_all + (DB ? 2 : 1) * T::CHUNKS);
    if constexpr (T::TW_BYTES > 0) {
        for (uint32_t i = tid; i < (uint32_t)(R / 2); i += NT) stw[i] = __ldg(&p.tw[i << (LOG_TW - LR)]);
    }
    // Programmatic dependent launch: everything above ran while the previous kernel in the stream (the previous pass)
    // was still draining; its output is only touched from here on.  The next kernel may be scheduled as soon as
    // every CTA of this one has got this far (it waits for our completion at the same point).
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (DB) {
        const uint32_t bz = tile / tiles_x, tx = tile - bz * tiles_x;
        load_tile_v4<LR, LC>(smv_all, p.in + (size_t)bz * p.in_batch_stride, p, tx * C);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    for (uint32_t iter = 0; tile < total_tiles; tile += gridDim.x, iter++) {
        const uint32_t bz = tile / tiles_x, tx = tile - bz * tiles_x;
        const uint32_t col0 = tx * C;
        uint32_t* __restrict__ out = p.out + (size_t)bz * p.out_batch_stride;
        uint4* smv = smv_all + ((DB && (iter & 1)) ? T::CHUNKS : 0);
        if (DB) {
            const uint32_t nt = tile + gridDim.x;
            if (nt < total_tiles) {
                const uint32_t nbz = nt / tiles_x, ntx = nt - nbz * tiles_x;
                load_tile_v4<LR, LC>(smv_all + ((iter & 1) ? 0 : T::CHUNKS), p.in + (size_t)nbz * p.in_batch_stride, p, ntx * C);
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;\n" ::: "memory");
        } else {
            load_tile_v4<LR, LC>(smv, p.in + (size_t)bz * p.in_batch_stride, p, col0);
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        }
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
                    v.x = pow_apply(p.pro, (lidx + 0) >> p.log_inner, v.x);
                    v.y = pow_apply(p.pro, (lidx + 1) >> p.log_inner, v.y);
                    v.z = pow_apply(p.pro, (lidx + 2) >> p.log_inner, v.z);
                    v.w = pow_apply(p.pro, (lidx + 3) >> p.log_inner, v.w);
                }
                smv[T::chunk(r, cv)] = v;
            }
            __syncthreads();
        }

        bool pruned = false;
        if constexpr (LR >= 6 && LR <= 9) {
            // Blowup-32 LDE, first pass: only rows d < R/32 hold input, i.e. (bit-reversed) rows 32m.  The first five
            // DIT stages then only copy x_m over rows 32m..32m+31 (every butterfly has a zero odd input), so they are
            // replaced by that copy and ONE round of the remaining LR-5 stages (stride 32).
            if (p.prune_log == 5) {
                pruned = true;
#pragma unroll 1
                for (int i = tid; i < R * CV; i += NT) {
                    const uint32_t cv = i & (CV - 1), r = i >> LCV;
                    if (r & 31u) smv[T::chunk(r, cv)] = smv[T::chunk(r & ~31u, cv)];
                }
                __syncthreads();
                dit_round_v4<LR, LC, 5, LR - 5>(smv, stw, p);
            }
        }
        if (!pruned) {
            // ---- radix-16 DIT rounds, strides 1, 16, 256
            if constexpr (T::G1 > 0) dit_round_v4<LR, LC, 0, T::G1>(smv, stw, p);
            if constexpr (T::G2 > 0) {
                __syncthreads();
                dit_round_v4<LR, LC, 4, T::G2>(smv, stw, p);
            }
            if constexpr (T::G3 > 0) {
                __syncthreads();
                dit_round_v4<LR, LC, 8, T::G3>(smv, stw, p);
            }
        }
        __syncthreads();

        // ---- epilogue + store
        const uint32_t epi_mode = p.epi_mode;
        if (epi_mode == EPI_FOURSTEP) {
            store_fourstep_v4<LR, LC>(smv, p, col0);
        } else if (p.log_pfull >= 2) {
            switch (epi_mode) {
                case EPI_TWIDDLE: store_rows_v4<LR, LC, EPI_TWIDDLE>(smv, out, p, col0); break;
                case EPI_OUTIDX: store_rows_v4<LR, LC, EPI_OUTIDX>(smv, out, p, col0); break;
                case EPI_CONST: store_rows_v4<LR, LC, EPI_CONST>(smv, out, p, col0); break;
                default: store_rows_v4<LR, LC, EPI_NONE>(smv, out, p, col0); break;
            }
        } else {
            switch (epi_mode) {
                case EPI_TWIDDLE: store_cols_v4<LR, LC, EPI_TWIDDLE>(smv, out, p, col0); break;
                case EPI_CONST: store_cols_v4<LR, LC, EPI_CONST>(smv, out, p, col0); break;
                default: store_cols_v4<LR, LC, EPI_NONE>(smv, out, p, col0); break;
            }
        }
        __syncthreads();  // the buffer just read is the prefetch target of the next iteration
    }
}

extern int g_pdl;  // ntt_engine.cu: programmatic dependent launch between passes (TOYNI_NTT_PDL=0 switches it off)

template <int LR, int LC>
void launch_pass_v4(const PassParams& p, dim3 grid, cudaStream_t s) {
    using T = V4<LR, LC>;
    constexpr bool DB = (2 * T::SMEM <= 96 * 1024);  // double buffering only where >= 2 CTAs still fit per SM
    constexpr size_t smem = (DB ? 2 * T::SMEM : T::SMEM) + T::TW_BYTES;
    static int ctas_per_sm[64] = {};
    static int n_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (ctas_per_sm[dev] == 0) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(ntt_pass_v4_kernel<LR, LC, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ntt_pass_v4_kernel<LR, LC, DB>, T::NT, smem);
        cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
        ctas_per_sm[dev] = occ > 0 ? occ : 1;
    }
    const uint32_t tiles_x = grid.x, total = grid.x * grid.y;
    uint32_t ctas = (uint32_t)(ctas_per_sm[dev] * n_sm[dev]);
    if (ctas > total) ctas = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(T::NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_pdl && p.pdl) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, ntt_pass_v4_kernel<LR, LC, DB>, p, tiles_x, total);
}

}  // namespace bb
