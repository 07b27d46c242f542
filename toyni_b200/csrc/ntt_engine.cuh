// Host-side NTT engine: per-device twiddle cache, pass planner and executor.
// Replaces NttCtx / ntt_ctx_create / build_twiddles_device (cuda/ntt_kernel.cu:160-242): twiddles are
// generated on the device, are O(sqrt n) instead of 2(n-1) words, and are shared by every context.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "ntt_pass.cuh"

namespace bb {

constexpr int MAX_LOG_N = 27;  // two-adicity of BabyBear, src/babybear.rs:119
constexpr int MAX_LR = 12;     // largest in-tile transform

struct NttPlan {
    int npass;
    int lr[3];
    int lc[3];
};

// Sharded four-step NTT: fuse the inter-half twiddle and the transpose into the last pass of the column transforms.
struct FourStepScatter {
    uint32_t* peer[8];        // receive buffer of every rank (own buffer for this rank; peer mappings otherwise)
    int nranks, rank;
    int log_n;                // full transform length
    size_t dst_row_stride;    // n2
    size_t col_offset;        // first global column of this rank's block (= rank * cols)
};

struct NttDesc {
    int log_n;              // transform length 2^log_n
    int log_inner;          // interleaved columns per element (0 plain, 2 for AoS Ext)
    bool inverse;
    const uint32_t* in;     // device
    uint32_t* out;          // device (may equal `in`)
    size_t n_in;            // valid input elements per vector (<= n); the rest read as zero
    size_t batch;           // independent contiguous vectors
    size_t batch_stride_in;   // in elements*inner (u32 units)
    size_t batch_stride_out;
    uint32_t coset_shift;   // 0 or 1: none.  forward: x[i] *= s^i first; inverse: y[k] *= s^-k afterwards
    const FourStepScatter* scatter;  // non-null: column transforms of a four-step NTT (log_inner = log2 cols)
};

// All functions return cudaError_t as int (0 = success) and never synchronise the stream.
int ntt_execute(const NttDesc& d, cudaStream_t stream);
NttPlan ntt_plan_for(int log_n, int log_inner, size_t batch);
// Tuning hook: override the planner for one size ("lr1,lr2[,lr3]/lc1,lc2[,lc3]"); npass=0 clears it.
void ntt_plan_override(int log_n, const NttPlan& plan);
PassLaunchFn pass_launcher(int lr, int lc);     // scalar kernel; nullptr if that tile shape is not built
PassLaunchFn pass_launcher_v4(int lr, int lc);  // vectorised kernel (LC >= 2)
// 0: tile kernel (ntt_pass_v4.cuh) for every size; 1 (default): TMA-staged two-pass kernel (ntt_pass_v7.cuh) where it applies
void engine_select_kernel(int kernel);
int engine_warmup(int log_n, cudaStream_t stream);  // build tables / this stream's scratch ahead of time
void engine_drop_stream(cudaStream_t stream);       // free the scratch kept for a stream that is going away
size_t engine_scratch_bytes();
int engine_diag_words(uint32_t out[16]);           // time-out / diagnostic words of the TMA-staged kernel (synchronises)
void engine_release();                       // free every cached device allocation on the current device

}  // namespace bb
