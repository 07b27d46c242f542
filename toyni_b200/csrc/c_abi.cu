// C ABI of the library (include/toyni_ntt_cuda.h).  Section numbers follow the header.
#if __has_include("toyni_ntt_cuda.h")
#include "toyni_ntt_cuda.h"  // -I include (build.py) or the flat cuda/ directory of the toyni tree
#else
#include "../../include/toyni_ntt_cuda.h"
#endif

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "fri_fold.cuh"
#include "merkle.cuh"
#include "ntt_engine.cuh"
#include "prover_ew.cuh"

using namespace bb;

// ------------------------------------------------------------------ library state
static thread_local int g_last_error = 0;  // per host thread: one caller's clear cannot erase another's pending error
static std::atomic<unsigned long long> g_launches{0};
static thread_local cudaStream_t g_stream = nullptr;  // per host thread, as the header says

static inline int note(int rc) {
    if (rc != 0 && g_last_error == 0) g_last_error = rc;  // sticky: the first error since the last bb_clear_error()
    return rc;
}
#define CK(x)                                  \
    do {                                       \
        int rc_ = (int)(x);                    \
        if (rc_ != 0) return note(rc_);        \
    } while (0)

static inline cudaStream_t cur_stream() { return g_stream; }
static inline bool is_pow2(size_t v) { return v && !(v & (v - 1)); }
static inline uint32_t log2_of(size_t v) {
    uint32_t l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

// ------------------------------------------------------------------ width conversion kernels
__global__ void __launch_bounds__(256) narrow_kernel(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t v = src[i];
        // canonical inputs take the fast path; anything else is reduced like BabyBear::new (src/babybear.rs:26-30)
        dst[i] = (v < (uint64_t)P) ? (uint32_t)v : (uint32_t)(v % (uint64_t)P);
    }
}
__global__ void __launch_bounds__(256) widen_kernel(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (uint64_t)src[i];
}
static inline unsigned conv_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b ? b : 1));
}

extern "C" {
static int run_ntt(const uint32_t* in, uint32_t* out, uint32_t log_n, int log_inner, size_t n_in, size_t batch, int dir,
                   uint32_t shift, const FourStepScatter* scatter = nullptr);
}
// what the multi-device layer (mg.cu) needs from this file
namespace bb {
namespace abi {
int note_error(int rc) { return note(rc); }
cudaStream_t get_stream() { return g_stream; }
void set_stream(cudaStream_t s) { g_stream = s; }
void count_launches(unsigned n) { g_launches += n; }
int ntt(const uint32_t* in, uint32_t* out, uint32_t log_n, int log_inner, size_t n_in, size_t batch, int dir, const FourStepScatter* scatter) {
    return run_ntt(in, out, log_n, log_inner, n_in, batch, dir, 1, scatter);
}
}  // namespace abi
}  // namespace bb

extern "C" {

// ================================================================== 1. reference symbols
int cuda_malloc(uint64_t** d_ptr, size_t count) { return note((int)cudaMalloc((void**)d_ptr, count * sizeof(uint64_t))); }
int cuda_free(uint64_t* d_ptr) { return note((int)cudaFree(d_ptr)); }
int cuda_copy_to_device(uint64_t* d_dest, const uint64_t* h_src, size_t count) {
    return note((int)cudaMemcpy(d_dest, h_src, count * sizeof(uint64_t), cudaMemcpyHostToDevice));
}
int cuda_copy_from_device(uint64_t* h_dest, const uint64_t* d_src, size_t count) {
    return note((int)cudaMemcpy(h_dest, d_src, count * sizeof(uint64_t), cudaMemcpyDeviceToHost));
}
const char* cuda_get_error_string(int error) { return cudaGetErrorString((cudaError_t)error); }

struct NttCtx {
    uint32_t n, log_n;
    uint64_t* d64;  // staging in the reference's width
    uint32_t* d32;  // compute buffer
    cudaStream_t stream;
    std::mutex mu;  // the reference's context is not re-entrant (cuda/ntt_kernel.cu:246-248); this one serialises
};

void* ntt_ctx_create(uint32_t n) {
    if (!is_pow2(n)) {
        note((int)cudaErrorInvalidValue);
        return nullptr;
    }
    uint32_t log_n = log2_of(n);
    if (log_n > (uint32_t)MAX_LOG_N) {  // cuda/ntt_kernel.cu:220 returns NULL silently; here the reason is recorded
        note((int)cudaErrorInvalidValue);
        return nullptr;
    }
    if (!bb_device_ok()) {
        note((int)cudaErrorNoKernelImageForDevice);
        return nullptr;
    }
    NttCtx* c = new NttCtx();
    c->n = n;
    c->log_n = log_n;
    c->d64 = nullptr;
    c->d32 = nullptr;
    c->stream = nullptr;
    if (note((int)cudaMalloc(&c->d64, (size_t)n * 8)) || note((int)cudaMalloc(&c->d32, (size_t)n * 4)) ||
        note((int)cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) || note(engine_warmup((int)log_n, c->stream))) {
        ntt_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

void ntt_ctx_destroy(void* ctx) {
    NttCtx* c = (NttCtx*)ctx;
    if (!c) return;
    if (c->stream) {
        cudaStreamSynchronize(c->stream);
        engine_drop_stream(c->stream);
        cudaStreamDestroy(c->stream);
    }
    cudaFree(c->d64);
    cudaFree(c->d32);
    delete c;
}

static int run_host_inplace(NttCtx* c, uint64_t* h, bool inverse) {
    if (!c || !h) return note((int)cudaErrorInvalidValue);
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t n = c->n;
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->d64, h, n * 8, cudaMemcpyHostToDevice, s));
    narrow_kernel<<<conv_blocks(n), 256, 0, s>>>(c->d64, c->d32, n);
    NttDesc d{};
    d.log_n = (int)c->log_n;
    d.inverse = inverse;
    d.in = c->d32;
    d.out = c->d32;
    d.n_in = n;
    d.batch = 1;
    d.batch_stride_in = d.batch_stride_out = n;
    CK(ntt_execute(d, s));
    widen_kernel<<<conv_blocks(n), 256, 0, s>>>(c->d32, c->d64, n);
    g_launches += 2 + (unsigned)ntt_plan_for((int)c->log_n, 0, 1).npass;
    CK(cudaMemcpyAsync(h, c->d64, n * 8, cudaMemcpyDeviceToHost, s));
    return note((int)cudaStreamSynchronize(s));
}

void ntt_run_inplace(void* ctx, uint64_t* h_data) { run_host_inplace((NttCtx*)ctx, h_data, false); }
void intt_run_inplace(void* ctx, uint64_t* h_data) { run_host_inplace((NttCtx*)ctx, h_data, true); }
// the same with the outcome as a return value (the void symbols above are what src/ntt.rs:108-109 declares)
int ntt_run_inplace_rc(void* ctx, uint64_t* h_data) { return run_host_inplace((NttCtx*)ctx, h_data, false); }
int intt_run_inplace_rc(void* ctx, uint64_t* h_data) { return run_host_inplace((NttCtx*)ctx, h_data, true); }

// The same host-pointer transform on u32 values (4 bytes per element over PCIe instead of the reference's 8): NOT part of
// the drop-in surface — src/ntt.rs stores a BabyBear as u64 — but what a host that already keeps canonical u32 should call.
int bb_ntt_host_u32(void* ctx, uint32_t* h_data, int dir) {
    NttCtx* c = (NttCtx*)ctx;
    if (!c || !h_data || (dir != 0 && dir != 1)) return note((int)cudaErrorInvalidValue);
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t n = c->n;
    cudaStream_t s = c->stream;
    CK(cudaMemcpyAsync(c->d32, h_data, n * 4, cudaMemcpyHostToDevice, s));
    NttDesc d{};
    d.log_n = (int)c->log_n;
    d.inverse = dir == 1;
    d.in = c->d32;
    d.out = c->d32;
    d.n_in = n;
    d.batch = 1;
    d.batch_stride_in = d.batch_stride_out = n;
    CK(ntt_execute(d, s));
    g_launches += (unsigned)ntt_plan_for((int)c->log_n, 0, 1).npass;
    CK(cudaMemcpyAsync(h_data, c->d32, n * 4, cudaMemcpyDeviceToHost, s));
    return note((int)cudaStreamSynchronize(s));
}

// ================================================================== 2. device-resident API
int bb_last_error(void) { return g_last_error; }
const char* bb_last_error_string(void) { return cudaGetErrorString((cudaError_t)g_last_error); }
void bb_clear_error(void) {
    g_last_error = 0;
    cudaGetLastError();
}
int bb_device_ok(void) {
    // The library carries sm_100a SASS only and no PTX: arch-specific code does not run on another minor revision
    // (e.g. sm_103), so the gate is compute capability 10.0 AND a kernel image that actually loads.
    static std::atomic<int> cached[64];  // 0 unknown, 1 ok, 2 not ok
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int c = cached[dev & 63].load();
    if (c) return c == 1;
    bool ok = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
              cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess && major == 10 && minor == 0;
    if (ok) {
        cudaFuncAttributes fa;
        ok = cudaFuncGetAttributes(&fa, narrow_kernel) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    cached[dev & 63].store(ok ? 1 : 2);
    return ok ? 1 : 0;
}
void bb_set_stream(void* cuda_stream) { g_stream = (cudaStream_t)cuda_stream; }
int bb_sync(void) { return note((int)cudaStreamSynchronize(cur_stream())); }

int bb_dev_alloc(void** d_ptr, size_t bytes) { return note((int)cudaMalloc(d_ptr, bytes)); }
int bb_dev_free(void* d_ptr) { return note((int)cudaFree(d_ptr)); }
// Stream-ordered allocations from the device's default memory pool, which is told to keep what is freed: a prover that
// allocates its LDE-sized arrays per proof (toyni_prover.hpp) pays the driver once, not 10+ ms of cudaMalloc / cudaFree
// of multi-gigabyte buffers per proof.
static int pool_ready() {
    static std::atomic<int> done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (done[dev & 63].load()) return 0;
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return (int)e;
    uint64_t keep = UINT64_MAX;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return (int)e;
    done[dev & 63].store(1);
    return 0;
}
int bb_pool_alloc(void** d_ptr, size_t bytes) {
    int rc = pool_ready();
    if (rc) return note(rc);
    return note((int)cudaMallocAsync(d_ptr, bytes ? bytes : 1, cur_stream()));
}
int bb_pool_free(void* d_ptr) { return d_ptr ? note((int)cudaFreeAsync(d_ptr, cur_stream())) : 0; }
int bb_pool_trim(void) {
    int dev = 0;
    cudaMemPool_t pool;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cur_stream());
    if (e == cudaSuccess) e = cudaMemPoolTrimTo(pool, 0);
    return note((int)e);
}
int bb_h2d(void* d_dst, const void* h_src, size_t bytes) {
    return note((int)cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, cur_stream()));
}
int bb_d2h(void* h_dst, const void* d_src, size_t bytes) {
    return note((int)cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, cur_stream()));
}
int bb_d2d(void* d_dst, const void* d_src, size_t bytes) {
    return note((int)cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, cur_stream()));
}
int bb_narrow_u64_to_u32(const uint64_t* d_src, uint32_t* d_dst, size_t count) {
    if (count == 0) return 0;
    narrow_kernel<<<conv_blocks(count), 256, 0, cur_stream()>>>(d_src, d_dst, count);
    g_launches++;
    return note((int)cudaGetLastError());
}
int bb_widen_u32_to_u64(const uint32_t* d_src, uint64_t* d_dst, size_t count) {
    if (count == 0) return 0;
    widen_kernel<<<conv_blocks(count), 256, 0, cur_stream()>>>(d_src, d_dst, count);
    g_launches++;
    return note((int)cudaGetLastError());
}

static int run_ntt(const uint32_t* in, uint32_t* out, uint32_t log_n, int log_inner, size_t n_in, size_t batch, int dir,
                   uint32_t shift, const FourStepScatter* scatter) {
    if (log_n > (uint32_t)MAX_LOG_N || (dir != 0 && dir != 1)) return note((int)cudaErrorInvalidValue);
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    NttDesc d{};
    d.log_n = (int)log_n;
    d.log_inner = log_inner;
    d.inverse = dir == 1;
    d.in = in;
    d.out = out;
    d.n_in = n_in;
    d.batch = batch;
    d.batch_stride_in = d.batch_stride_out = ((size_t)1 << log_n) << log_inner;
    d.coset_shift = shift;
    d.scatter = scatter;
    CK(ntt_execute(d, cur_stream()));
    g_launches += (unsigned)ntt_plan_for((int)log_n, log_inner, batch).npass;
    return 0;
}

int bb_ntt_device(uint32_t* d_data, uint32_t log_n, int dir) {
    return run_ntt(d_data, d_data, log_n, 0, (size_t)1 << log_n, 1, dir, 1);
}
// Two halves of a batch on two helper streams: independent transforms overlap each other's memory-bound and
// integer-bound phases (two 2^24 transforms side by side take 101 us each instead of 120 us one after the other).
// The helper streams fork from and join back into the caller's stream, so the call stays stream-ordered.
namespace {
struct SplitStreams {
    cudaStream_t h[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    int dev = -1;
};
thread_local SplitStreams g_split[64];  // per host thread and device: nothing is dropped when the thread changes device
int split_streams(SplitStreams** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    SplitStreams& s = g_split[dev & 63];
    if (s.dev != dev) {
        for (int i = 0; i < 2; i++) {
            CK(cudaStreamCreateWithFlags(&s.h[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        s.dev = dev;
    }
    *out = &s;
    return 0;
}
}  // namespace

int bb_ntt_batch_device(uint32_t* d_data, uint32_t log_n, size_t batch, int dir) {
    static int split = -1;
    if (split < 0) {
        const char* e = getenv("TOYNI_NTT_SPLIT");
        split = e ? atoi(e) : 1;
    }
    const size_t n = (size_t)1 << log_n;
    // worth it once each half is a few hundred tiles of a multi-pass plan (and small enough to keep the launch overlap)
    if (!split || batch < 2 || log_n < 9 || log_n > (uint32_t)MAX_LOG_N || (batch / 2) * n < ((size_t)1 << 22) ||
        batch * n >= ((size_t)1 << 28) || (dir != 0 && dir != 1))
        return run_ntt(d_data, d_data, log_n, 0, n, batch, dir, 1);
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    SplitStreams* ss = nullptr;
    CK(split_streams(&ss));
    cudaStream_t s = cur_stream();
    CK(cudaEventRecord(ss->fork, s));
    const size_t b0 = batch / 2;
    for (int i = 0; i < 2; i++) {
        CK(cudaStreamWaitEvent(ss->h[i], ss->fork, 0));
        NttDesc d{};
        d.log_n = (int)log_n;
        d.inverse = dir == 1;
        d.in = d.out = d_data + (i ? b0 * n : 0);
        d.n_in = n;
        d.batch = i ? batch - b0 : b0;
        d.batch_stride_in = d.batch_stride_out = n;
        d.coset_shift = 1;
        CK(ntt_execute(d, ss->h[i]));
        g_launches += (unsigned)ntt_plan_for((int)log_n, 0, d.batch).npass;
        CK(cudaEventRecord(ss->join[i], ss->h[i]));
    }
    for (int i = 0; i < 2; i++) CK(cudaStreamWaitEvent(s, ss->join[i], 0));
    return 0;
}
int bb_ntt_ext_device(uint32_t* d_data, uint32_t log_n, int dir) {
    return run_ntt(d_data, d_data, log_n, 2, (size_t)1 << log_n, 1, dir, 1);
}

int bb_ntt_columns_device(uint32_t* d_block, uint32_t log_n1, size_t cols, int dir) {
    if (!is_pow2(cols)) return note((int)cudaErrorInvalidValue);
    return run_ntt(d_block, d_block, log_n1, (int)log2_of(cols), (size_t)1 << log_n1, 1, dir, 1);
}
int bb_fourstep_twiddle_device(uint32_t* d_block, uint32_t log_n, uint32_t log_n1, size_t cols, size_t col_offset, int dir) {
    CK(fourstep_twiddle(d_block, (int)log_n, (int)log_n1, cols, col_offset, dir == 1, cur_stream()));
    g_launches++;
    return 0;
}

int bb_ntt_columns_scatter_device(uint32_t* d_block, uint32_t log_n, uint32_t log_n1, size_t cols, int dir,
                                  void* const* peer_bufs, uint32_t nranks, uint32_t rank) {
    if (!is_pow2(cols) || !is_pow2(nranks) || nranks > 8 || rank >= nranks || log_n1 >= log_n || !peer_bufs)
        return note((int)cudaErrorInvalidValue);
    FourStepScatter fs{};
    for (uint32_t r = 0; r < nranks; r++) fs.peer[r] = (uint32_t*)peer_bufs[r];
    fs.nranks = (int)nranks;
    fs.rank = (int)rank;
    fs.log_n = (int)log_n;
    fs.dst_row_stride = (size_t)1 << (log_n - log_n1);  // n2
    fs.col_offset = (size_t)rank * cols;
    return run_ntt(d_block, d_block, log_n1, (int)log2_of(cols), (size_t)1 << log_n1, 1, dir, 1, &fs);
}
int bb_peer_signal_device(void* const* d_peer_flags, uint32_t nranks, uint32_t rank, uint32_t epoch) {
    CK(peer_signal((uint32_t* const*)d_peer_flags, nranks, rank, epoch, cur_stream()));
    g_launches++;
    return 0;
}
int bb_peer_wait_device(void* d_flags, uint32_t nranks, uint32_t epoch, void* d_err) {
    CK(peer_wait((uint32_t*)d_flags, nranks, epoch, (uint32_t*)d_err, cur_stream()));
    g_launches++;
    return 0;
}
int bb_ipc_get_handle(const void* d_ptr, uint8_t handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle_out, &h, 64);
    return 0;
}
int bb_ipc_open_handle(const uint8_t handle[64], void** d_ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    return note((int)cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
}
int bb_ipc_close_handle(void* d_ptr) { return note((int)cudaIpcCloseMemHandle(d_ptr)); }

int bb_coset_fft_device(const uint32_t* d_coeffs, size_t n_coeffs, uint32_t log_size, uint32_t shift, int limbs,
                        uint32_t* d_out) {
    if ((limbs != 1 && limbs != 4) || shift == 0 || shift >= P) return note((int)cudaErrorInvalidValue);
    size_t size = (size_t)1 << log_size;
    size_t take = n_coeffs < size ? n_coeffs : size;  // resize() truncates, src/math/domain.rs:108-109
    return run_ntt(d_coeffs, d_out, log_size, limbs == 4 ? 2 : 0, take, 1, 0, shift);
}
int bb_coset_ifft_device(uint32_t* d_evals, uint32_t log_size, uint32_t shift, int limbs) {
    if ((limbs != 1 && limbs != 4) || shift == 0 || shift >= P) return note((int)cudaErrorInvalidValue);
    return run_ntt(d_evals, d_evals, log_size, limbs == 4 ? 2 : 0, (size_t)1 << log_size, 1, 1, shift);
}

int bb_fri_fold_device(const uint32_t* d_evals, size_t m, uint32_t x0, const uint32_t beta[4], int limbs, uint32_t* d_out) {
    if (!is_pow2(m) || (limbs != 1 && limbs != 4)) return note((int)cudaErrorInvalidValue);
    CK(fri_fold_coset(d_evals, d_out, m, limbs, (int)log2_of(m), x0, beta, 1, 0, cur_stream()));
    g_launches++;
    return 0;
}
int bb_fri_fold_shard_device(const uint32_t* d_evals, size_t m_local, uint32_t log_m, uint32_t x0, const uint32_t beta[4],
                             int limbs, uint32_t nranks, uint32_t rank, uint32_t* d_out) {
    if ((limbs != 1 && limbs != 4) || nranks == 0 || rank >= nranks || m_local * nranks != ((size_t)1 << log_m))
        return note((int)cudaErrorInvalidValue);
    CK(fri_fold_coset(d_evals, d_out, m_local, limbs, (int)log_m, x0, beta, nranks, rank, cur_stream()));
    g_launches++;
    return 0;
}
int bb_fri_fold_chain_shard_device(const uint32_t* d_layer0, size_t m_local, uint32_t log_m, uint32_t shift, const uint32_t* betas,
                                   size_t nbetas, int limbs, uint32_t nranks, uint32_t rank, size_t until, uint32_t* d_layers,
                                   size_t* folds_out) {
    if ((limbs != 1 && limbs != 4) || nranks == 0 || rank >= nranks || m_local * nranks != ((size_t)1 << log_m) || !betas)
        return note((int)cudaErrorInvalidValue);
    size_t m = (size_t)1 << log_m, folds = 0;
    uint32_t x0 = shift % P;
    const uint32_t* src = d_layer0;
    uint32_t* dst = d_layers;
    cudaStream_t s = cur_stream();
    while (m > until && m / 2 >= nranks) {
        if (folds >= nbetas) return note((int)cudaErrorInvalidValue);
        const size_t ml = m / nranks;
        size_t left = 0;  // folds still to do from here
        for (size_t mm = m; mm > until && mm / 2 >= nranks; mm /= 2) left++;
        if (fri_fold_tail_applies(ml, left)) {  // the short end of the chain: one single-CTA launch
            if (folds + left > nbetas) return note((int)cudaErrorInvalidValue);
            CK(fri_fold_chain_tail(src, dst, ml, limbs, (int)log_m - (int)folds, x0, betas + folds * limbs, left, nranks, rank, s));
            g_launches++;
            folds += left;
            break;
        }
        CK(fri_fold_coset(src, dst, ml, limbs, (int)log_m - (int)folds, x0, betas + folds * limbs, nranks, rank, s));
        g_launches++;
        src = dst;
        dst += (ml / 2) * limbs;
        x0 = mul(x0, x0);
        m /= 2;
        folds++;
    }
    if (folds_out) *folds_out = folds;
    return 0;
}
int bb_fri_fold_xs_device(const uint32_t* d_evals, size_t m, const uint32_t* d_xs, const uint32_t beta[4], int limbs,
                          uint32_t* d_out) {
    if (limbs != 1 && limbs != 4) return note((int)cudaErrorInvalidValue);
    CK(fri_fold_xs(d_evals, d_xs, d_out, m, limbs, beta, cur_stream()));
    g_launches++;
    return 0;
}

size_t bb_merkle_node_count(size_t nleaves) { return merkle_node_count(nleaves); }

static int levels_of(size_t n) { return merkle_upper_launches(n); }  // kernel launches of merkle_upper_levels

static int root_to_host(const uint8_t* d_nodes, size_t n, uint8_t* root_out, cudaStream_t s) {
    CK(cudaMemcpyAsync(root_out, d_nodes + 32 * (merkle_node_count(n) - 1), 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int bb_merkle_commit_device(const uint32_t* d_vals, int limbs, size_t n, const uint8_t* d_salts, uint8_t* d_nodes,
                            uint8_t* root_out) {
    CK(merkle_commit(d_vals, limbs, n, d_salts, d_nodes, cur_stream()));
    g_launches += 1 + (unsigned)levels_of(n);
    if (root_out) return root_to_host(d_nodes, n, root_out, cur_stream());
    return 0;
}
int bb_merkle_build_bytes_device(const uint8_t* d_leaves, size_t n, size_t leaf_len, uint8_t* d_nodes, uint8_t* root_out) {
    CK(merkle_build_bytes(d_leaves, n, leaf_len, d_nodes, cur_stream()));
    g_launches += 1 + (unsigned)levels_of(n);
    if (root_out) return root_to_host(d_nodes, n, root_out, cur_stream());
    return 0;
}
int bb_merkle_open_device(const uint8_t* d_nodes, size_t nleaves, size_t index, uint8_t* path_out, uint8_t* pos_out,
                          size_t* depth_out) {
    uint8_t* d_path = nullptr;
    CK(cudaMalloc(&d_path, 32 * 64));
    size_t depth = 0;
    int rc = merkle_open(d_nodes, nleaves, index, d_path, pos_out, &depth, cur_stream());
    if (rc == 0 && depth > 0) {
        g_launches++;
        rc = (int)cudaMemcpyAsync(path_out, d_path, 32 * depth, cudaMemcpyDeviceToHost, cur_stream());
    }
    if (rc == 0) rc = (int)cudaStreamSynchronize(cur_stream());
    cudaFree(d_path);
    if (depth_out) *depth_out = depth;
    return note(rc);
}

// Small per-thread device scratch for the query-set entry points (indices, gathered paths / values): grow-only, so
// that an opening costs no cudaMalloc / cudaFree (each a device-wide synchronisation) after the first call.
namespace {
struct SmallScratch {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
};
thread_local SmallScratch g_small[64][2];  // per host thread and device
static thread_local std::vector<uint8_t> g_open_host;  // host landing buffer of bb_merkle_open_multi_device

int small_scratch(int slot, size_t bytes, void** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    SmallScratch& s = g_small[dev & 63][slot];
    if (s.dev != dev || s.cap < bytes) {
        if (s.p) {
            cudaStreamSynchronize(cur_stream());
            cudaFree(s.p);
        }
        s.p = nullptr;
        s.cap = 0;
        size_t want = bytes < 4096 ? 4096 : bytes;
        int rc = (int)cudaMalloc(&s.p, want);
        if (rc) return rc;
        s.cap = want;
        s.dev = dev;
    }
    *out = s.p;
    return 0;
}
}  // namespace

// ---- whole query sets at once (src/fibonacci.rs:250-295): paths, and the opened values / salts
int bb_merkle_open_batch_device(const uint8_t* d_nodes, size_t nleaves, const uint64_t* indices, size_t nq, uint8_t* paths_out,
                                uint8_t* pos_out, size_t* depth_out) {
    size_t depth = 0;
    for (size_t m = nleaves; m > 1; m = (m + 1) / 2) depth++;
    if (depth_out) *depth_out = depth;
    if (nq == 0) return 0;
    for (size_t q = 0; q < nq; q++) {
        if (indices[q] >= nleaves) return note((int)cudaErrorInvalidValue);
        size_t level_n = nleaves, cur = (size_t)indices[q], d = 0;
        while (level_n > 1) {  // position flags need no device data (src/merkle.rs:67-73)
            size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
            pos_out[q * depth + d++] = (sib >= level_n) ? 1 : (uint8_t)(cur % 2 == 1);
            cur /= 2;
            level_n = (level_n + 1) / 2;
        }
    }
    if (depth == 0) return 0;
    cudaStream_t s = cur_stream();
    unsigned long long* d_idx = nullptr;
    uint8_t* d_paths = nullptr;
    CK(small_scratch(0, nq * sizeof(unsigned long long), (void**)&d_idx));
    int rc = small_scratch(1, nq * depth * 32, (void**)&d_paths);
    if (rc == 0) rc = (int)cudaMemcpyAsync(d_idx, indices, nq * sizeof(unsigned long long), cudaMemcpyHostToDevice, s);
    if (rc == 0) rc = merkle_gather_paths(d_nodes, nleaves, d_idx, nq, (uint32_t)depth, d_paths, s);
    if (rc == 0) {
        g_launches++;
        rc = (int)cudaMemcpyAsync(paths_out, d_paths, nq * depth * 32, cudaMemcpyDeviceToHost, s);
    }
    if (rc == 0) rc = (int)cudaStreamSynchronize(s);
    return note(rc);
}
int bb_merkle_open_multi_device(const bb_open_request* reqs, size_t nreq, const uint64_t* indices, size_t nq, size_t elem_bytes,
                                uint8_t* paths_out, size_t paths_bytes, uint8_t* pos_out, uint8_t* vals_out, uint8_t* salts_out) {
    if (!reqs || !indices || !paths_out || !pos_out || !vals_out || !salts_out || elem_bytes == 0 || elem_bytes > 32)
        return note((int)cudaErrorInvalidValue);
    if (nq == 0) return 0;
    std::vector<OpenQuery> qs(nq);
    size_t path_off = 0, flags = 0, covered = 0;
    for (size_t r = 0; r < nreq; r++) {
        const bb_open_request& t = reqs[r];
        if (t.first != covered || t.first + t.count > nq || !t.d_nodes || !t.d_vals) return note((int)cudaErrorInvalidValue);
        size_t depth = 0;
        for (size_t m = t.nleaves; m > 1; m = (m + 1) / 2) depth++;
        for (size_t k = 0; k < t.count; k++) {
            const size_t q = t.first + k;
            if (indices[q] >= t.nleaves) return note((int)cudaErrorInvalidValue);
            qs[q] = OpenQuery{t.d_nodes, (const uint8_t*)t.d_vals, t.d_salts, t.nleaves, indices[q], path_off};
            size_t level_n = t.nleaves, cur = (size_t)indices[q];
            while (level_n > 1) {  // position flags need no device data (src/merkle.rs:67-73)
                const size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
                pos_out[flags++] = (sib >= level_n) ? 1 : (uint8_t)(cur % 2 == 1);
                cur /= 2;
                level_n = (level_n + 1) / 2;
            }
            path_off += depth * 32;
        }
        covered += t.count;
    }
    if (covered != nq || path_off != paths_bytes) return note((int)cudaErrorInvalidValue);
    // one device buffer [queries][paths][values][salts], one upload, one launch, one download, one synchronisation
    cudaStream_t s = cur_stream();
    const size_t qbytes = (nq * sizeof(OpenQuery) + 31) & ~(size_t)31, pbytes = (paths_bytes + 31) & ~(size_t)31;
    const size_t vbytes = (nq * elem_bytes + 31) & ~(size_t)31;
    uint8_t *d_q = nullptr, *d_out = nullptr;
    CK(small_scratch(0, qbytes, (void**)&d_q));
    CK(small_scratch(1, pbytes + vbytes + nq * 16, (void**)&d_out));
    int rc = (int)cudaMemcpyAsync(d_q, qs.data(), nq * sizeof(OpenQuery), cudaMemcpyHostToDevice, s);
    if (rc == 0) rc = merkle_open_multi((const OpenQuery*)d_q, nq, (uint32_t)elem_bytes, d_out, d_out + pbytes, d_out + pbytes + vbytes, s);
    if (rc == 0) {
        g_launches++;
        std::vector<uint8_t>& h = g_open_host;
        h.resize(pbytes + vbytes + nq * 16);
        rc = (int)cudaMemcpyAsync(h.data(), d_out, h.size(), cudaMemcpyDeviceToHost, s);
        if (rc == 0) rc = (int)cudaStreamSynchronize(s);
        if (rc == 0) {
            memcpy(paths_out, h.data(), paths_bytes);
            memcpy(vals_out, h.data() + pbytes, nq * elem_bytes);
            memcpy(salts_out, h.data() + pbytes + vbytes, nq * 16);
        }
    }
    return note(rc);
}
int bb_gather_device(const void* d_src, size_t elem_bytes, const uint64_t* indices, size_t nq, void* out) {
    if (nq == 0 || elem_bytes == 0) return 0;
    cudaStream_t s = cur_stream();
    unsigned long long* d_idx = nullptr;
    uint8_t* d_out = nullptr;
    CK(small_scratch(0, nq * sizeof(unsigned long long), (void**)&d_idx));
    int rc = small_scratch(1, nq * elem_bytes, (void**)&d_out);
    if (rc == 0) rc = (int)cudaMemcpyAsync(d_idx, indices, nq * sizeof(unsigned long long), cudaMemcpyHostToDevice, s);
    if (rc == 0) rc = gather_elems(d_src, (uint32_t)elem_bytes, d_idx, nq, d_out, s);
    if (rc == 0) {
        g_launches++;
        rc = (int)cudaMemcpyAsync(out, d_out, nq * elem_bytes, cudaMemcpyDeviceToHost, s);
    }
    if (rc == 0) rc = (int)cudaStreamSynchronize(s);
    return note(rc);
}

int bb_interleave_device(const uint32_t* d_src, uint32_t groups, size_t chunk, int limbs, uint32_t* d_dst) {
    if (groups == 0 || (limbs != 1 && limbs != 4)) return note((int)cudaErrorInvalidValue);
    CK(interleave(d_src, d_dst, groups, chunk, (uint32_t)limbs, cur_stream()));
    g_launches++;
    return 0;
}

// ---- element-wise stages of the Fibonacci prover between the LDE and the FRI commit loop
int bb_fib_constraint_device(const uint32_t* d_trace_lde, uint32_t log_n, uint32_t step, uint32_t shift, uint32_t b1, uint32_t b2,
                             uint32_t* d_out) {
    CK(fib_constraint(d_trace_lde, d_out, (int)log_n, step, shift, b1, b2, cur_stream()));
    g_launches++;
    return 0;
}
int bb_scale_periodic_device(uint32_t* d_vals, size_t n, const uint32_t* table, uint32_t period) {
    CK(scale_periodic(d_vals, n, table, period, cur_stream()));
    g_launches++;
    return 0;
}
int bb_fib_deep_device(const uint32_t* d_quotient, const uint32_t* d_trace_lde, uint32_t log_n, uint32_t step, uint32_t shift, uint32_t z,
                       uint32_t q_z, uint32_t t_z, uint32_t t_gz, uint32_t t_ggz, uint32_t* d_out) {
    CK(fib_deep(d_quotient, d_trace_lde, d_out, (int)log_n, step, shift, z, q_z, t_z, t_gz, t_ggz, cur_stream()));
    g_launches++;
    return 0;
}
int bb_poly_eval_device(const uint32_t* d_coeffs, size_t n, uint32_t z, uint32_t* value_out) {
    cudaStream_t s = cur_stream();
    unsigned long long* d_acc = nullptr;
    unsigned long long h = 0;
    CK(small_scratch(0, sizeof(unsigned long long), (void**)&d_acc));
    int rc = poly_eval(d_coeffs, n, z, d_acc, s);
    if (rc == 0) rc = (int)cudaMemcpyAsync(&h, d_acc, sizeof h, cudaMemcpyDeviceToHost, s);
    if (rc == 0) rc = (int)cudaStreamSynchronize(s);
    g_launches++;
    if (rc == 0 && value_out) *value_out = (uint32_t)(h % (unsigned long long)P);
    return note(rc);
}

int bb_fri_commit_device(const uint32_t* d_layer0, size_t n, uint32_t shift, size_t final_size, int limbs,
                         const uint8_t* d_salts, bb_challenge_fn challenge, void* user, const uint32_t* betas_in,
                         uint32_t* d_layers, uint8_t* d_nodes, uint8_t* roots_out, size_t* folds_out) {
    if (!is_pow2(n) || !is_pow2(final_size) || final_size > n || (limbs != 1 && limbs != 4) || shift == 0 || shift >= P)
        return note((int)cudaErrorInvalidValue);
    if (!challenge && !betas_in && n > final_size) return note((int)cudaErrorInvalidValue);
    if (challenge && !d_nodes) return note((int)cudaErrorInvalidValue);  // a transcript needs roots
    cudaStream_t s = cur_stream();
    const uint32_t* cur = d_layer0;  // layer 0 stays where the caller has it; d_layers receives layers 1, 2, ...
    uint32_t* next = d_layers;
    size_t cur_n = n, folds = 0;
    uint32_t x0 = shift, log_m = log2_of(n);
    const uint8_t* salt_ptr = d_salts;
    uint8_t* node_ptr = d_nodes;
    uint8_t root[32];
    if (d_nodes) {  // layer 0 = DEEP evaluations, src/fibonacci.rs:204-211
        CK(merkle_commit(cur, limbs, cur_n, cur_n == final_size ? nullptr : salt_ptr, node_ptr, s));
        g_launches += 1 + (unsigned)levels_of(cur_n);
        if (roots_out || challenge) {
            CK(root_to_host(node_ptr, cur_n, root, s));
            if (roots_out) memcpy(roots_out, root, 32);
        }
        if (salt_ptr && cur_n != final_size) salt_ptr += 16 * cur_n;
        node_ptr += 32 * merkle_node_count(cur_n);
    }
    while (cur_n > final_size) {  // src/fibonacci.rs:222
        uint32_t beta[4] = {0, 0, 0, 0};
        if (challenge)
            challenge(user, root, (uint32_t)folds, beta);  // absorb(root) happened for this root; squeeze beta (:223)
        else
            memcpy(beta, betas_in + (size_t)limbs * folds, sizeof(uint32_t) * (size_t)limbs);
        // fold (:225) and, when committing, the leaf level of the new layer's tree in the same kernel (:234-238:
        // salted, or unsalted for the final layer)
        const bool final_layer = (cur_n / 2 == final_size);
        const int hash_mode = d_nodes ? (final_layer || !salt_ptr ? 1 : 2) : 0;
        CK(fri_fold_coset(cur, next, cur_n, limbs, (int)log_m, x0, beta, 1, 0, s, hash_mode, salt_ptr, node_ptr));
        g_launches++;
        x0 = bb::mul(x0, x0);  // :228-231: x <- x^2
        log_m--;
        cur_n /= 2;
        folds++;
        if (d_nodes) {
            CK(merkle_upper_levels(node_ptr, cur_n, s));
            g_launches += (unsigned)levels_of(cur_n);
            if (roots_out || challenge) {
                CK(root_to_host(node_ptr, cur_n, root, s));
                if (roots_out) memcpy(roots_out + 32 * folds, root, 32);
            }
            if (salt_ptr && !final_layer) salt_ptr += 16 * cur_n;
            node_ptr += 32 * merkle_node_count(cur_n);
        }
        cur = next;
        next += cur_n * (size_t)limbs;
    }
    if (folds_out) *folds_out = folds;
    return 0;
}

int bb_ntt_set_plan(uint32_t log_n, int npass, const int* log_rows, const int* log_cols) {
    NttPlan pl{};
    pl.npass = npass;
    if (npass < 0 || npass > 3) return note((int)cudaErrorInvalidValue);
    int sum = 0;
    for (int i = 0; i < npass; i++) {
        pl.lr[i] = log_rows[i];
        pl.lc[i] = log_cols[i];
        sum += log_rows[i];
        if (!pass_launcher(pl.lr[i], pl.lc[i])) return note((int)cudaErrorInvalidValue);
    }
    if (npass > 0 && sum != (int)log_n) return note((int)cudaErrorInvalidValue);
    ntt_plan_override((int)log_n, pl);
    return 0;
}
int bb_ntt_get_plan(uint32_t log_n, int* log_rows, int* log_cols) {
    NttPlan pl = ntt_plan_for((int)log_n, 0, 1);
    for (int i = 0; i < pl.npass; i++) {
        if (log_rows) log_rows[i] = pl.lr[i];
        if (log_cols) log_cols[i] = pl.lc[i];
    }
    return pl.npass;
}
void bb_ntt_set_kernel(int kernel) { engine_select_kernel(kernel); }
int bb_ntt_launches(uint32_t log_n) { return ntt_plan_for((int)log_n, 0, 1).npass; }
unsigned long long bb_kernel_launch_count(void) { return g_launches.load(); }
int bb_warmup(uint32_t log_n) {
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    return note(engine_warmup((int)log_n, cur_stream()));
}
void bb_release(void) { engine_release(); }

}  // extern "C"

// out[i] = shift * w^i for i < 2^log_n, w = get_root_of_unity(log_n): BabyBearDomain::elements (src/math/domain.rs:61-69);
// shift = 1 gives roots_of_unity_domain (src/ntt.rs:69-81).  The reference builds both with n sequential multiplications.
template <typename T>
__global__ void __launch_bounds__(256) domain_elements_kernel(PowTable tab, uint32_t shift_m, T* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (T)monty_mul(pow_plain(tab, (uint32_t)i), shift_m);
}

extern "C" {

int bb_domain_elements_device(uint32_t log_n, uint32_t shift, uint32_t* d_out) {
    if (log_n > (uint32_t)MAX_LOG_N || shift == 0 || shift >= P || !d_out) return note((int)cudaErrorInvalidValue);
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    PowTable tab;
    CK(engine_pow_table(root_of_unity(log_n), (int)log_n, 1u, &tab));
    const size_t n = (size_t)1 << log_n;
    domain_elements_kernel<uint32_t><<<conv_blocks(n), 256, 0, cur_stream()>>>(tab, to_monty(shift), d_out, n);
    g_launches++;
    return note((int)cudaGetLastError());
}
int bb_ntt_diag(uint32_t words_out[16]) { return note(engine_diag_words(words_out)); }

}  // extern "C"

// ================================================================== 3. host-pointer forms
namespace {
struct Staging {  // grow-on-demand device staging for the host-pointer entry points
    uint64_t* d64 = nullptr;
    size_t n64 = 0;
    uint32_t* a32 = nullptr;
    size_t na = 0;
    uint32_t* b32 = nullptr;
    size_t nb = 0;
    uint8_t* bytes = nullptr;
    size_t nbytes = 0;
    uint8_t* nodes = nullptr;
    size_t nnodes = 0;
    std::mutex mu;
};
Staging g_stages[64];  // one set per device: a buffer grown on device 0 is never handed to a kernel on device 1
Staging& cur_stage() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_stages[dev & 63];
}

template <typename T>
int grow(T** p, size_t* have, size_t want) {
    if (*have >= want) return 0;
    if (*p) {
        cudaDeviceSynchronize();
        cudaFree(*p);
        *p = nullptr;
        *have = 0;
    }
    int rc = (int)cudaMalloc((void**)p, want * sizeof(T));
    if (rc == 0) *have = want;
    return rc;
}

// upload `count` u64 values and narrow them into dst32
int upload_narrow(const uint64_t* h, size_t count, uint32_t* dst32, cudaStream_t s) {
    if (count == 0) return 0;
    CK(grow(&cur_stage().d64, &cur_stage().n64, count));
    CK(cudaMemcpyAsync(cur_stage().d64, h, count * 8, cudaMemcpyHostToDevice, s));
    narrow_kernel<<<conv_blocks(count), 256, 0, s>>>(cur_stage().d64, dst32, count);
    g_launches++;
    return (int)cudaGetLastError();
}
int widen_download(const uint32_t* src32, size_t count, uint64_t* h, cudaStream_t s) {
    if (count == 0) return 0;
    CK(grow(&cur_stage().d64, &cur_stage().n64, count));
    widen_kernel<<<conv_blocks(count), 256, 0, s>>>(src32, cur_stage().d64, count);
    g_launches++;
    CK(cudaMemcpyAsync(h, cur_stage().d64, count * 8, cudaMemcpyDeviceToHost, s));
    return (int)cudaStreamSynchronize(s);
}

int domain_transform(const uint64_t* in, size_t n_in, size_t size, uint64_t shift, uint64_t* out, int limbs, bool inverse) {
    if (!is_pow2(size) || log2_of(size) > (uint32_t)MAX_LOG_N) return note((int)cudaErrorInvalidValue);
    if (inverse && n_in != size) return note((int)cudaErrorInvalidValue);  // assert_eq!, src/math/domain.rs:86
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    std::lock_guard<std::mutex> lk(cur_stage().mu);
    cudaStream_t s = cur_stream();
    size_t take = n_in < size ? n_in : size;
    CK(grow(&cur_stage().a32, &cur_stage().na, (take ? take : 1) * (size_t)limbs));
    CK(grow(&cur_stage().b32, &cur_stage().nb, size * (size_t)limbs));
    const uint32_t sh = (uint32_t)(shift % P);
    if (inverse) {
        CK(upload_narrow(in, size * (size_t)limbs, cur_stage().b32, s));
        CK(bb_coset_ifft_device(cur_stage().b32, log2_of(size), sh, limbs));
    } else {
        CK(upload_narrow(in, take * (size_t)limbs, cur_stage().a32, s));
        CK(bb_coset_fft_device(cur_stage().a32, take, log2_of(size), sh, limbs, cur_stage().b32));
    }
    return note(widen_download(cur_stage().b32, size * (size_t)limbs, out, s));
}

int fold_host(const uint64_t* evals, size_t m, const uint64_t* xs, const uint64_t* beta, uint64_t* out, int limbs) {
    if (m < 2 || (m & 1)) return note((int)cudaErrorInvalidValue);  // assert!, src/math/fri.rs:8,28
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    std::lock_guard<std::mutex> lk(cur_stage().mu);
    cudaStream_t s = cur_stream();
    const size_t half = m / 2;
    CK(grow(&cur_stage().a32, &cur_stage().na, m * (size_t)limbs + half));
    CK(grow(&cur_stage().b32, &cur_stage().nb, half * (size_t)limbs));
    uint32_t* d_xs = cur_stage().a32 + m * (size_t)limbs;
    CK(upload_narrow(evals, m * (size_t)limbs, cur_stage().a32, s));
    CK(upload_narrow(xs, half, d_xs, s));
    uint32_t b[4] = {0, 0, 0, 0};
    for (int k = 0; k < limbs; k++) b[k] = (uint32_t)(beta[k] % P);
    CK(bb_fri_fold_xs_device(cur_stage().a32, m, d_xs, b, limbs, cur_stage().b32));
    return note(widen_download(cur_stage().b32, half * (size_t)limbs, out, s));
}
}  // namespace

extern "C" {

int toyni_domain_fft(const uint64_t* coeffs, size_t n_coeffs, size_t size, uint64_t shift, uint64_t* evals_out) {
    return domain_transform(coeffs, n_coeffs, size, shift, evals_out, 1, false);
}
int toyni_domain_ifft(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* coeffs_out) {
    return domain_transform(evals, size, size, shift, coeffs_out, 1, true);
}
int toyni_domain_fft_ext(const uint64_t* coeffs, size_t n_coeffs, size_t size, uint64_t shift, uint64_t* evals_out) {
    return domain_transform(coeffs, n_coeffs, size, shift, evals_out, 4, false);
}
int toyni_domain_ifft_ext(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* coeffs_out) {
    return domain_transform(evals, size, size, shift, coeffs_out, 4, true);
}
int toyni_fri_fold(const uint64_t* evals, size_t m, const uint64_t* xs, uint64_t beta, uint64_t* out) {
    return fold_host(evals, m, xs, &beta, out, 1);
}
int toyni_fri_fold_ext(const uint64_t* evals, size_t m, const uint64_t* xs, const uint64_t beta[4], uint64_t* out) {
    return fold_host(evals, m, xs, beta, out, 4);
}
int toyni_merkle_commit(const uint64_t* values, size_t n, int limbs, const uint8_t* salts, uint8_t* nodes_out,
                        uint8_t root_out[32]) {
    if (n == 0 || (limbs != 1 && limbs != 4)) return note((int)cudaErrorInvalidValue);
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    std::lock_guard<std::mutex> lk(cur_stage().mu);
    cudaStream_t s = cur_stream();
    const size_t count = merkle_node_count(n);
    CK(grow(&cur_stage().a32, &cur_stage().na, n * (size_t)limbs));
    CK(grow(&cur_stage().nodes, &cur_stage().nnodes, count * 32));
    CK(upload_narrow(values, n * (size_t)limbs, cur_stage().a32, s));
    uint8_t* d_salts = nullptr;
    if (salts) {
        CK(grow(&cur_stage().bytes, &cur_stage().nbytes, n * 16));
        CK(cudaMemcpyAsync(cur_stage().bytes, salts, n * 16, cudaMemcpyHostToDevice, s));
        d_salts = cur_stage().bytes;
    }
    CK(bb_merkle_commit_device(cur_stage().a32, limbs, n, d_salts, cur_stage().nodes, root_out));
    if (nodes_out) CK(cudaMemcpyAsync(nodes_out, cur_stage().nodes, count * 32, cudaMemcpyDeviceToHost, s));
    return note((int)cudaStreamSynchronize(s));
}

/* roots_of_unity_domain (src/ntt.rs:69-81) / BabyBearDomain::elements (src/math/domain.rs:61-69): size values, u64 */
int toyni_domain_elements(size_t size, uint64_t shift, uint64_t* out) {
    if (!is_pow2(size) || log2_of(size) > (uint32_t)MAX_LOG_N || shift == 0 || !out) return note((int)cudaErrorInvalidValue);
    if (!bb_device_ok()) return note((int)cudaErrorNoKernelImageForDevice);
    std::lock_guard<std::mutex> lk(cur_stage().mu);
    cudaStream_t s = cur_stream();
    CK(grow(&cur_stage().d64, &cur_stage().n64, size));
    PowTable tab;
    const uint32_t log_n = log2_of(size);
    CK(engine_pow_table(root_of_unity(log_n), (int)log_n, 1u, &tab));
    domain_elements_kernel<uint64_t><<<conv_blocks(size), 256, 0, s>>>(tab, to_monty((uint32_t)(shift % P)), cur_stage().d64, size);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, cur_stage().d64, size * 8, cudaMemcpyDeviceToHost, s));
    return note((int)cudaStreamSynchronize(s));
}
int toyni_roots_of_unity_domain(size_t n, uint64_t* out) { return toyni_domain_elements(n, 1, out); }

}  // extern "C"
