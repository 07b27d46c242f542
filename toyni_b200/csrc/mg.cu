// C ABI, section 4: the sharded paths driven from ONE host process over the G GPUs of a box.
// Nothing like this exists in the reference (single device, default stream).  Peer access is enabled between all
// pairs, so a kernel on GPU q stores straight into a buffer of GPU r (NVLink / NVSwitch); ordering between devices is
// CUDA events on per-device streams — no NCCL, no IPC handles, no host synchronisation inside a transform.
#if __has_include("toyni_ntt_cuda.h")
#include "toyni_ntt_cuda.h"  // -I include (build.py) or the flat cuda/ directory of the toyni tree
#else
#include "../../include/toyni_ntt_cuda.h"
#endif

#include <cstring>
#include <vector>

#include "abi_internal.cuh"
#include "bb_field.cuh"
#include "fri_fold.cuh"

using namespace bb;

namespace {

#define MCK(x)                                        \
    do {                                              \
        int rc_ = (int)(x);                           \
        if (rc_ != 0) return abi::note_error(rc_);    \
    } while (0)

struct Mg {
    int G = 0;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev_cols, ev_done;
    // staging of the host-pointer form (grow-only, per device)
    std::vector<uint32_t*> blk, out;
    std::vector<uint64_t*> st64;
    size_t blk_words = 0, out_words = 0, st_words = 0;
};

struct DeviceGuard {  // the caller's current device and library stream come back whatever happens
    int dev = 0;
    cudaStream_t s;
    DeviceGuard() {
        cudaGetDevice(&dev);
        s = abi::get_stream();
    }
    ~DeviceGuard() {
        cudaSetDevice(dev);
        abi::set_stream(s);
    }
};

bool pow2(size_t v) { return v && !(v & (v - 1)); }

void split(uint32_t log_n, uint32_t* l1, uint32_t* l2) {  // n1 >= n2, as toyni_b200/multigpu.py::fourstep_split
    *l1 = (log_n + 1) / 2;
    *l2 = log_n - *l1;
}

// out64[c * rows + r] = in32[r * cols + c]: the (n1/G x n2) result block back to natural order, widened to the reference's u64
__global__ void __launch_bounds__(256) transpose_widen_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t rows, uint32_t cols) {
    __shared__ uint32_t tile[32][33];
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (uint32_t i = threadIdx.y; i < 32; i += 8) {
        const uint32_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0u;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.y; i < 32; i += 8) {
        const uint32_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = (uint64_t)tile[threadIdx.x][i];
    }
}
__global__ void __launch_bounds__(256) narrow64_kernel(const uint64_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t v = src[i];
        dst[i] = (v < (uint64_t)P) ? (uint32_t)v : (uint32_t)(v % (uint64_t)P);
    }
}

int fourstep(Mg* m, uint32_t log_n, int dir, uint32_t* const* d_blocks, uint32_t* const* d_outs) {
    const int G = m->G;
    uint32_t l1, l2;
    split(log_n, &l1, &l2);
    const size_t n1 = (size_t)1 << l1, n2 = (size_t)1 << l2;
    if (n1 % G || n2 % G || (n2 / G) < 16) return abi::note_error((int)cudaErrorInvalidValue);
    const size_t cw = n2 / G, rw = n1 / G;
    FourStepScatter fs{};
    for (int r = 0; r < 8; r++) fs.peer[r] = d_outs[r < G ? r : 0];
    fs.nranks = G;
    fs.log_n = (int)log_n;
    fs.dst_row_stride = n2;
    // column transforms of every device's block; their last pass multiplies by w_n^(j2 k1) and stores row k1 straight
    // into the buffer of the device that owns it.  A receive buffer is free once its device's previous row pass is done.
    for (int q = 0; q < G; q++) {
        MCK(cudaSetDevice(q));
        abi::set_stream(m->stream[q]);
        for (int r = 0; r < G; r++)
            if (r != q) MCK(cudaStreamWaitEvent(m->stream[q], m->ev_done[r], 0));
        fs.rank = q;
        fs.col_offset = (size_t)q * cw;
        MCK(abi::ntt(d_blocks[q], d_blocks[q], l1, (int)l2 - __builtin_ctz((unsigned)G), n1, 1, dir, &fs));
        MCK(cudaEventRecord(m->ev_cols[q], m->stream[q]));
    }
    // row transforms once every peer's stores have landed
    for (int r = 0; r < G; r++) {
        MCK(cudaSetDevice(r));
        abi::set_stream(m->stream[r]);
        for (int q = 0; q < G; q++)
            if (q != r) MCK(cudaStreamWaitEvent(m->stream[r], m->ev_cols[q], 0));
        MCK(abi::ntt(d_outs[r], d_outs[r], l2, 0, n2, rw, dir, nullptr));
        MCK(cudaEventRecord(m->ev_done[r], m->stream[r]));
    }
    return 0;
}

template <typename T>
int grow(std::vector<T*>& v, size_t* have, size_t want, int G) {
    if (*have >= want) return 0;
    for (int r = 0; r < G; r++) {
        MCK(cudaSetDevice(r));
        if (v[r]) {
            MCK(cudaDeviceSynchronize());
            MCK(cudaFree(v[r]));
            v[r] = nullptr;
        }
        MCK(cudaMalloc((void**)&v[r], want * sizeof(T)));
    }
    *have = want;
    return 0;
}

}  // namespace

extern "C" {

int bb_mg_init(int ngpus, void** mg_out) {
    if (!mg_out || ngpus < 1 || ngpus > 8 || !pow2((size_t)ngpus)) return abi::note_error((int)cudaErrorInvalidValue);
    int have = 0;
    MCK(cudaGetDeviceCount(&have));
    if (have < ngpus) return abi::note_error((int)cudaErrorInvalidDevice);
    DeviceGuard guard;
    Mg* m = new Mg();
    m->G = ngpus;
    m->stream.assign(ngpus, nullptr);
    m->ev_cols.assign(ngpus, nullptr);
    m->ev_done.assign(ngpus, nullptr);
    m->blk.assign(ngpus, nullptr);
    m->out.assign(ngpus, nullptr);
    m->st64.assign(ngpus, nullptr);
    for (int r = 0; r < ngpus; r++) {
        int rc = (int)cudaSetDevice(r);
        if (rc == 0 && !bb_device_ok()) rc = (int)cudaErrorNoKernelImageForDevice;
        for (int q = 0; rc == 0 && q < ngpus; q++) {
            if (q == r) continue;
            int can = 0;
            rc = (int)cudaDeviceCanAccessPeer(&can, r, q);
            if (rc == 0 && !can) rc = (int)cudaErrorPeerAccessUnsupported;
            if (rc == 0) {
                rc = (int)cudaDeviceEnablePeerAccess(q, 0);
                if (rc == (int)cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    rc = 0;
                }
            }
        }
        if (rc == 0) rc = (int)cudaStreamCreateWithFlags(&m->stream[r], cudaStreamNonBlocking);
        if (rc == 0) rc = (int)cudaEventCreateWithFlags(&m->ev_cols[r], cudaEventDisableTiming);
        if (rc == 0) rc = (int)cudaEventCreateWithFlags(&m->ev_done[r], cudaEventDisableTiming);
        if (rc) {
            bb_mg_destroy(m);
            return abi::note_error(rc);
        }
    }
    *mg_out = m;
    return 0;
}

void bb_mg_destroy(void* mg) {
    Mg* m = (Mg*)mg;
    if (!m) return;
    DeviceGuard guard;
    for (int r = 0; r < m->G; r++) {
        cudaSetDevice(r);
        if (m->stream[r]) {
            cudaStreamSynchronize(m->stream[r]);
            engine_drop_stream(m->stream[r]);
            cudaStreamDestroy(m->stream[r]);
        }
        if (m->ev_cols[r]) cudaEventDestroy(m->ev_cols[r]);
        if (m->ev_done[r]) cudaEventDestroy(m->ev_done[r]);
        cudaFree(m->blk[r]);
        cudaFree(m->out[r]);
        cudaFree(m->st64[r]);
    }
    delete m;
}

int bb_mg_ngpus(void* mg) { return mg ? ((Mg*)mg)->G : 0; }

int bb_mg_sync(void* mg) {
    Mg* m = (Mg*)mg;
    if (!m) return abi::note_error((int)cudaErrorInvalidValue);
    DeviceGuard guard;
    for (int r = 0; r < m->G; r++) {
        MCK(cudaSetDevice(r));
        MCK(cudaStreamSynchronize(m->stream[r]));
    }
    return 0;
}

void* bb_mg_stream(void* mg, int device) {
    Mg* m = (Mg*)mg;
    return (m && device >= 0 && device < m->G) ? (void*)m->stream[device] : nullptr;
}

int bb_mg_ntt_fourstep(void* mg, uint32_t log_n, int dir, uint32_t* const* d_blocks, uint32_t* const* d_outs) {
    Mg* m = (Mg*)mg;
    if (!m || !d_blocks || !d_outs || log_n > (uint32_t)MAX_LOG_N || (dir != 0 && dir != 1)) return abi::note_error((int)cudaErrorInvalidValue);
    DeviceGuard guard;
    return fourstep(m, log_n, dir, d_blocks, d_outs);
}

int bb_mg_ntt_batch(void* mg, uint32_t log_n, int dir, uint32_t* const* d_cols, const size_t* ncols) {
    Mg* m = (Mg*)mg;
    if (!m || !d_cols || !ncols || log_n > (uint32_t)MAX_LOG_N || (dir != 0 && dir != 1)) return abi::note_error((int)cudaErrorInvalidValue);
    DeviceGuard guard;
    for (int r = 0; r < m->G; r++) {
        if (ncols[r] == 0) continue;
        MCK(cudaSetDevice(r));
        abi::set_stream(m->stream[r]);
        MCK(bb_ntt_batch_device(d_cols[r], log_n, ncols[r], dir));
    }
    return 0;
}

int bb_mg_fri_chain(void* mg, uint32_t log_m, uint32_t shift, int limbs, size_t final_size, const uint32_t* betas, const uint32_t* const* d_shards,
                    uint32_t* const* d_layers_out, size_t* folds_out) {
    Mg* m = (Mg*)mg;
    if (!m || !betas || !d_shards || !d_layers_out || (limbs != 1 && limbs != 4) || log_m > 31 || shift == 0 || shift >= P)
        return abi::note_error((int)cudaErrorInvalidValue);
    const int G = m->G;
    if (((size_t)1 << log_m) < (size_t)2 * G) return abi::note_error((int)cudaErrorInvalidValue);
    DeviceGuard guard;
    size_t folds = 0;
    for (int r = 0; r < G; r++) {
        MCK(cudaSetDevice(r));
        abi::set_stream(m->stream[r]);
        size_t mm = (size_t)1 << log_m;
        uint32_t x0 = shift, k = 0;
        const uint32_t* cur = d_shards[r];
        uint32_t* next = d_layers_out[r];
        // the fold partner i + m/2 is on the same device while m/2 >= G (cyclic layout): no exchange
        while (mm > final_size && (mm / 2) >= (size_t)G) {
            size_t left = 0;  // folds still to do: the short end of the chain goes in one single-CTA launch
            for (size_t t = mm; t > final_size && (t / 2) >= (size_t)G; t /= 2) left++;
            if (fri_fold_tail_applies(mm / G, left)) {
                MCK(fri_fold_chain_tail(cur, next, mm / G, limbs, (int)(log_m - k), x0, betas + (size_t)limbs * k, left, (uint32_t)G, (uint32_t)r,
                                        m->stream[r]));
                abi::count_launches(1);
                k += (uint32_t)left;
                break;
            }
            MCK(fri_fold_coset(cur, next, mm / G, limbs, (int)(log_m - k), x0, betas + (size_t)limbs * k, (uint32_t)G, (uint32_t)r, m->stream[r]));
            abi::count_launches(1);
            cur = next;
            mm /= 2;
            next += (mm / G) * (size_t)limbs;
            x0 = bb::mul(x0, x0);
            k++;
        }
        folds = k;
    }
    if (folds_out) *folds_out = folds;
    return 0;
}

int bb_mg_ntt_host(void* mg, uint64_t* h_data, uint32_t log_n, int dir) {
    Mg* m = (Mg*)mg;
    if (!m || !h_data || log_n > (uint32_t)MAX_LOG_N || (dir != 0 && dir != 1)) return abi::note_error((int)cudaErrorInvalidValue);
    const int G = m->G;
    uint32_t l1, l2;
    split(log_n, &l1, &l2);
    const size_t n1 = (size_t)1 << l1, n2 = (size_t)1 << l2;
    if (n1 % G || n2 % G || (n2 / G) < 16) return abi::note_error((int)cudaErrorInvalidValue);
    const size_t cw = n2 / G, rw = n1 / G;
    DeviceGuard guard;
    MCK(grow(m->blk, &m->blk_words, n1 * cw, G));
    MCK(grow(m->out, &m->out_words, rw * n2, G));
    MCK(grow(m->st64, &m->st_words, n1 * cw > rw * n2 ? n1 * cw : rw * n2, G));
    for (int r = 0; r < G; r++) {  // column block r of the n1 x n2 input matrix -> device r, narrowed to u32
        MCK(cudaSetDevice(r));
        MCK(cudaMemcpy2DAsync(m->st64[r], cw * 8, h_data + (size_t)r * cw, n2 * 8, cw * 8, n1, cudaMemcpyHostToDevice, m->stream[r]));
        narrow64_kernel<<<2048, 256, 0, m->stream[r]>>>(m->st64[r], m->blk[r], n1 * cw);
        MCK(cudaGetLastError());
    }
    MCK(fourstep(m, log_n, dir, m->blk.data(), m->out.data()));
    for (int r = 0; r < G; r++) {  // out_r[k1_local][k2] = X[k1 + n1 k2]  ->  natural order in the caller's array
        MCK(cudaSetDevice(r));
        dim3 grid((unsigned)((n2 + 31) / 32), (unsigned)((rw + 31) / 32));
        transpose_widen_kernel<<<grid, dim3(32, 8), 0, m->stream[r]>>>(m->out[r], m->st64[r], (uint32_t)rw, (uint32_t)n2);
        MCK(cudaGetLastError());
        MCK(cudaMemcpy2DAsync(h_data + (size_t)r * rw, n1 * 8, m->st64[r], rw * 8, rw * 8, n2, cudaMemcpyDeviceToHost, m->stream[r]));
    }
    abi::count_launches(2u * (unsigned)G);
    for (int r = 0; r < G; r++) {
        MCK(cudaSetDevice(r));
        MCK(cudaStreamSynchronize(m->stream[r]));
    }
    return 0;
}

}  // extern "C"
