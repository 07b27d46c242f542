// Element-wise helpers of the multi-GPU four-step NTT.
#include "fri_fold.cuh"
#include "ntt_pass.cuh"

namespace bb {

// d[k1][c] *= w_n^((col_offset + c) * k1) for a local block of `cols` columns (step 2 of the four-step NTT:
// the twiddle between the column transforms and the transpose).  One thread per 4 adjacent columns.
__global__ void __launch_bounds__(256) fourstep_twiddle_kernel(uint32_t* __restrict__ d, uint32_t rows, uint32_t cols,
                                                               uint32_t col_offset, PowTable tw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cq = cols >> 2;
    if (i >= (size_t)rows * cq) return;
    const uint32_t k1 = (uint32_t)(i / cq), c = (uint32_t)(i % cq) * 4u;
    uint4* p = reinterpret_cast<uint4*>(d + (size_t)k1 * cols + c);
    uint4 v = *p;
    // geometric in the column: w^((c0+c) k1), ratio w^k1
    uint32_t t = pow_lookup(tw, (col_offset + c) * k1);
    const uint32_t g = pow_lookup(tw, k1);
    v.x = monty_mul(v.x, t); t = monty_mul(t, g);
    v.y = monty_mul(v.y, t); t = monty_mul(t, g);
    v.z = monty_mul(v.z, t); t = monty_mul(t, g);
    v.w = monty_mul(v.w, t);
    *p = v;
}

int fourstep_twiddle(uint32_t* d, int log_n, int log_n1, size_t cols, size_t col_offset, bool inverse, cudaStream_t s) {
    if (log_n < 2 || log_n > MAX_LOG_N || log_n1 < 1 || log_n1 >= log_n || (cols & 3)) return (int)cudaErrorInvalidValue;
    uint32_t omega = root_of_unity(log_n);
    if (inverse) omega = bb::inv(omega);
    PowTable tw;
    int rc = engine_pow_table(omega, log_n, 1u, &tw);
    if (rc) return rc;
    const size_t rows = (size_t)1 << log_n1;
    const size_t work = rows * (cols >> 2);
    fourstep_twiddle_kernel<<<(unsigned)((work + 255) / 256), 256, 0, s>>>(d, (uint32_t)rows, (uint32_t)cols, (uint32_t)col_offset, tw);
    return (int)cudaGetLastError();
}

// ---- device-side rendezvous between the ranks of one box (one process per GPU, flags in CUDA-IPC peer memory)
// signal: tell every rank that this rank's stores of epoch `epoch` are done (runs after the scatter kernels in stream
// order; the system-scope fence orders those peer stores before the flag).
__global__ void peer_signal_kernel(uint32_t* const* peer_flags, uint32_t nranks, uint32_t rank, uint32_t epoch) {
    if (threadIdx.x < nranks) {
        __threadfence_system();
        volatile uint32_t* f = peer_flags[threadIdx.x] + 32u * rank;  // one 128-byte line per writer
        *f = epoch;
    }
}
// wait: spin until every rank has signalled `epoch` into this rank's flags.  Each GPU runs its own waiter, the
// writers run on OTHER GPUs, so nothing here depends on co-scheduling; a clock bound turns a lost peer into an error
// word instead of a hang.
__global__ void peer_wait_kernel(volatile uint32_t* flags, uint32_t nranks, uint32_t epoch, uint32_t* err) {
    if (threadIdx.x < nranks) {
        const long long t0 = clock64();
        volatile uint32_t* f = flags + 32u * threadIdx.x;
        while ((int32_t)(*f - epoch) < 0) {
            if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
                *err = 1u + threadIdx.x;
                break;
            }
            __nanosleep(200);
        }
        __threadfence_system();
    }
}

int peer_signal(uint32_t* const* d_peer_flags, uint32_t nranks, uint32_t rank, uint32_t epoch, cudaStream_t s) {
    peer_signal_kernel<<<1, 32, 0, s>>>(d_peer_flags, nranks, rank, epoch);
    return (int)cudaGetLastError();
}
int peer_wait(uint32_t* d_flags, uint32_t nranks, uint32_t epoch, uint32_t* d_err, cudaStream_t s) {
    peer_wait_kernel<<<1, 32, 0, s>>>(d_flags, nranks, epoch, d_err);
    return (int)cudaGetLastError();
}

}  // namespace bb
