// NTT engine: on-device twiddle generation + cache, pass planner, executor.  See ntt_engine.cuh.
#include "ntt_engine.cuh"

#include "fri_fold.cuh"
#include "ntt_v7.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace bb {

// ------------------------------------------------------------------ table generation kernels
__device__ __forceinline__ uint32_t monty_pow_dev(uint32_t g_m, uint32_t e) {
    uint32_t r = R_MOD_P, b = g_m;
    while (e) {
        if (e & 1u) r = monty_mul(r, b);
        b = monty_mul(b, b);
        e >>= 1;
    }
    return r;
}

// out[i] = Shoup pair of g^i * scale (plain form)
__global__ void gen_pow_kernel(uint2* out, uint32_t count, uint32_t g_m, uint32_t scale) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        uint32_t w = monty_mul(monty_pow_dev(g_m, i), scale);  // (g^i R) * scale / R
        out[i] = make_uint2(w, shoup_companion(w));
    }
}

// out[i] = (w, floor(w 2^32 / p)) with w = g^i, plain form
__global__ void gen_shoup_kernel(uint2* out, uint32_t count, uint32_t g_m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        uint32_t w = from_monty(monty_pow_dev(g_m, i));
        out[i] = make_uint2(w, shoup_companion(w));
    }
}

// ------------------------------------------------------------------ per-device state
struct PowTab {
    uint2* lo = nullptr;
    uint2* hi = nullptr;
    uint32_t lo_bits = 0;
};

struct DeviceState {
    uint2* tw[2] = {nullptr, nullptr};  // omega_4096^(+-i), i < 2048
    uint2 tw16[2][8];
    uint2* tw8192 = nullptr;            // omega_8192^i, i < 4096 (forward): the LDE expansion pass, built on first use
    std::map<std::tuple<uint32_t, int, uint32_t>, PowTab> pow_tabs;  // (g, log_total, scale) -> tables
    // one scratch buffer per stream: transforms on different streams (e.g. two legacy contexts used from two host
    // threads) may execute concurrently and must not share the intermediate array
    std::map<cudaStream_t, std::pair<uint32_t*, size_t>> scratch;
    uint32_t* err_word = nullptr;  // time-out diagnostics of the TMA-staged kernel
    bool ready = false;
};

static std::mutex g_mu;
static std::map<int, DeviceState> g_states;
static std::map<int, NttPlan> g_plan_override;
static bool g_force_scalar = false;  // test hook: run every pass on the scalar kernel
int g_pdl = 1;
static int g_v7 = -1;                // TMA-staged two-pass kernel for plain 2^24-point vectors (TOYNI_NTT_V7=0 switches it off)

#define BB_CK(x)                          \
    do {                                  \
        cudaError_t e_ = (x);             \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

static int state_get(DeviceState** out) {
    int dev = 0;
    BB_CK(cudaGetDevice(&dev));
    DeviceState& st = g_states[dev];
    if (!st.ready) {
        for (int inv = 0; inv < 2; inv++) {
            uint32_t w = root_of_unity(LOG_TW);
            if (inv) w = bb::inv(w);
            BB_CK(cudaMalloc(&st.tw[inv], sizeof(uint2) << (LOG_TW - 1)));
            gen_shoup_kernel<<<(1 << (LOG_TW - 1)) / 256, 256>>>(st.tw[inv], 1u << (LOG_TW - 1), to_monty(w));
            BB_CK(cudaGetLastError());
            uint32_t w16 = root_of_unity(4);
            if (inv) w16 = bb::inv(w16);
            uint32_t cur = 1;
            for (int i = 0; i < 8; i++) {
                st.tw16[inv][i] = make_uint2(cur, shoup_companion(cur));
                cur = bb::mul(cur, w16);
            }
        }
        BB_CK(cudaMalloc(&st.err_word, 16 * sizeof(uint32_t)));
        BB_CK(cudaMemset(st.err_word, 0, 16 * sizeof(uint32_t)));
        BB_CK(cudaDeviceSynchronize());
        st.ready = true;
    }
    *out = &st;
    return 0;
}

// tables for g^t, t < 2^log_total, every entry pre-multiplied by `scale` (through the lo table)
static int pow_table_get(DeviceState& st, uint32_t g, int log_total, uint32_t scale, PowTable* out) {
    auto key = std::make_tuple(g, log_total, scale);
    auto it = st.pow_tabs.find(key);
    if (it == st.pow_tabs.end()) {
        PowTab t;
        t.lo_bits = (uint32_t)((log_total + 1) / 2);
        uint32_t n_lo = 1u << t.lo_bits, n_hi = 1u << (log_total - (int)t.lo_bits);
        BB_CK(cudaMalloc(&t.lo, sizeof(uint2) * n_lo));
        BB_CK(cudaMalloc(&t.hi, sizeof(uint2) * n_hi));
        gen_pow_kernel<<<(n_lo + 255) / 256, 256>>>(t.lo, n_lo, to_monty(g), scale);
        gen_pow_kernel<<<(n_hi + 255) / 256, 256>>>(t.hi, n_hi, to_monty(bb::pow(g, n_lo)), 1u);
        BB_CK(cudaGetLastError());
        // generated on the legacy default stream: make them visible to any (possibly non-blocking) stream
        BB_CK(cudaDeviceSynchronize());
        it = st.pow_tabs.emplace(key, t).first;
    }
    out->lo = it->second.lo;
    out->hi = it->second.hi;
    out->lo_bits = it->second.lo_bits;
    return 0;
}

static int scratch_get(DeviceState& st, cudaStream_t stream, size_t words, uint32_t** out) {
    auto& slot = st.scratch[stream];
    if (slot.second < words) {
        if (slot.first) {
            BB_CK(cudaStreamSynchronize(stream));
            BB_CK(cudaFree(slot.first));
            slot.first = nullptr;
            slot.second = 0;
        }
        BB_CK(cudaMalloc(&slot.first, words * sizeof(uint32_t)));
        slot.second = words;
    }
    *out = slot.first;
    return 0;
}

int engine_pow_table(uint32_t g, int log_total, uint32_t scale, PowTable* out) {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState* st;
    int rc = state_get(&st);
    if (rc) return rc;
    return pow_table_get(*st, g, log_total, scale, out);
}

int engine_diag_words(uint32_t out[16]) {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState* st;
    int rc = state_get(&st);
    if (rc) return rc;
    return (int)cudaMemcpy(out, st->err_word, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost);
}

size_t engine_scratch_bytes() {
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    auto it = g_states.find(dev);
    size_t total = 0;
    if (it != g_states.end())
        for (auto& kv : it->second.scratch) total += kv.second.second * sizeof(uint32_t);
    return total;
}

void engine_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    auto it = g_states.find(dev);
    if (it == g_states.end()) return;
    cudaDeviceSynchronize();
    DeviceState& st = it->second;
    for (int i = 0; i < 2; i++) cudaFree(st.tw[i]);
    cudaFree(st.tw8192);
    st.tw8192 = nullptr;
    for (auto& kv : st.pow_tabs) {
        cudaFree(kv.second.lo);
        cudaFree(kv.second.hi);
    }
    for (auto& kv : st.scratch) cudaFree(kv.second.first);
    cudaFree(st.err_word);
    g_states.erase(it);
}

// ------------------------------------------------------------------ planner
static int max_lc_for(int lr) { return lr >= 12 ? 3 : (lr == 11 ? 4 : 5); }

static int pick_lc(int lr, int want) {  // largest built LC <= want (built: 0,2,3,4,5 within the smem limit)
    int lc = want < max_lc_for(lr) ? want : max_lc_for(lr);
    if (lc == 1) lc = 0;
    if (lc < 0) lc = 0;
    return lc;
}

static int ceil_log2(size_t v) {
    int l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

void ntt_plan_override(int log_n, const NttPlan& plan) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (plan.npass == 0)
        g_plan_override.erase(log_n);
    else
        g_plan_override[log_n] = plan;
}

static bool v7_enabled() {
    if (g_v7 < 0) {
        const char* e = getenv("TOYNI_NTT_V7");
        g_v7 = e ? atoi(e) : 1;
        const char* e2 = getenv("TOYNI_NTT_PDL");
        if (e2) g_pdl = atoi(e2);
    }
    return g_v7 != 0 && !g_force_scalar;
}

static NttPlan plan_locked(int log_n, int log_inner, size_t batch, bool allow_v7 = true) {
    auto it = g_plan_override.find(log_n);
    if (it != g_plan_override.end()) return it->second;
    NttPlan pl;
    memset(&pl, 0, sizeof pl);
    if (allow_v7 && log_n == 2 * V7_LR && log_inner == 0 && v7_enabled()) {
        // 4096 x 4096 in two TMA-staged passes (ntt_pass_v7.cuh): 16 instead of 24 bytes of traffic per element
        pl.npass = 2;
        pl.lr[0] = pl.lr[1] = V7_LR;
        pl.lc[0] = pl.lc[1] = 3;
        return pl;
    }
    // Measured on B200 (tools/plan_sweep.py): passes of about 2^8 rows x 16 columns (16 KB tiles, double buffered,
    // 7 CTAs per SM) beat fewer, larger passes — the kernels are instruction-bound, not DRAM-bound, and the
    // intermediate arrays of a <= 2^24 transform mostly stay in the 126 MB L2.
    if (log_n <= 8 && log_inner == 0) {
        // short plain vectors: one pass of the scalar kernel, the vectors of the batch are the tile columns
        pl.npass = 1;
        pl.lr[0] = log_n;
        int want = MAX_LR - log_n;
        int lb = ceil_log2(batch);
        pl.lc[0] = pick_lc(log_n, want < lb ? want : lb);
        return pl;
    }
    if (log_n <= 6 || (log_inner > 0 && log_n <= 8)) {
        pl.npass = 1;
        pl.lr[0] = log_n;
        pl.lc[0] = pick_lc(log_n, log_inner > 0 ? log_inner : 0);
        return pl;
    }
    // 2^17..2^20 in two passes of 256..1024 rows when there is enough work to fill the persistent grid with such tiles
    // (measured with tools/plan_sweep_batch.py on 2^24 words: 5-30 % faster than three passes); single short vectors
    // keep three passes of small tiles
    const bool two_long = log_n >= 17 && log_n <= 20 && (((size_t)batch << (log_n + log_inner)) >= ((size_t)1 << 23));
    if (log_n <= 16 || two_long) {
        pl.npass = 2;
        pl.lr[0] = (log_n + 1) / 2;
        pl.lr[1] = log_n - pl.lr[0];
        if (log_n == 17) { pl.lr[0] = 8; pl.lr[1] = 9; }
    } else {
        pl.npass = 3;
        pl.lr[0] = (log_n + 2) / 3;
        pl.lr[1] = (log_n - pl.lr[0] + 1) / 2;
        pl.lr[2] = log_n - pl.lr[0] - pl.lr[1];
        // measured on B200 (tools/plan_sweep.py): the largest sizes prefer an uneven split by 1-6 %
        if (log_n == 25) { pl.lr[0] = 8; pl.lr[1] = 8; pl.lr[2] = 9; }
        if (log_n == 26) { pl.lr[0] = 10; pl.lr[1] = 8; pl.lr[2] = 8; }
        if (log_n == 27) { pl.lr[0] = 8; pl.lr[1] = 10; pl.lr[2] = 9; }
    }
    for (int i = 0; i < pl.npass; i++) {
        int ncols_log = log_n - pl.lr[i] + log_inner;
        int want = 4;  // 16 columns = 64-byte row segments
        if (want > ncols_log) want = ncols_log;
        pl.lc[i] = pick_lc(pl.lr[i], want);
    }
    return pl;
}

void engine_select_kernel(int kernel) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_v7 = kernel != 0;  // 0: tile kernel everywhere; otherwise the TMA-staged kernel where it applies (default)
}

// plain 2^24-point vectors: two passes of the TMA-staged kernel.  Tables and scratch are looked up under the engine
// lock (v7_prepare); the launches themselves run without it.
struct V7Job {
    V7Params p;
    uint32_t* scratch;
    bool pdl;
};

static int v7_prepare(DeviceState& st, const NttDesc& d, cudaStream_t stream, V7Job* job) {
    const int inv = d.inverse ? 1 : 0;
    const size_t n = (size_t)1 << d.log_n;
    uint32_t omega = root_of_unity(d.log_n);
    if (inv) omega = bb::inv(omega);
    V7Params& p = job->p;
    memset(&p, 0, sizeof p);
    int rc = pow_table_get(st, omega, d.log_n, 1u, &p.tab);
    if (rc) return rc;
    p.tab_scaled = p.tab;
    if (inv) {
        rc = pow_table_get(st, omega, d.log_n, bb::inv((uint32_t)(n % P)), &p.tab_scaled);
        if (rc) return rc;
    }
    rc = scratch_get(st, stream, n * d.batch, &job->scratch);
    if (rc) return rc;
    p.tw = st.tw[inv];
    memcpy(p.tw16, st.tw16[inv], sizeof p.tw16);
    p.exp_mask = (uint32_t)(n - 1);
    p.err = st.err_word;
    job->pdl = n * d.batch < ((size_t)1 << 28);
    return 0;
}

static int v7_launch(const NttDesc& d, V7Job& job, cudaStream_t stream) {
    const size_t n = (size_t)1 << d.log_n;
    const size_t ncols = n >> V7_LR;
    V7Params p = job.p;
    // pass 1: columns of in[4096][ncols] -> scratch[col][e]
    p.out = job.scratch;
    p.out_batch_stride = n;
    p.log_pfull = 0;
    int rc = launch_pass_v7(false, d.in, ncols, d.batch, d.batch_stride_in, p, job.pdl, stream);
    if (rc) return rc;
    // pass 2: columns of scratch[ncols][4096] (rows = columns of pass 1) -> out[e2 * 4096 + e1]
    p.out = d.out;
    p.out_batch_stride = d.batch_stride_out;
    p.log_pfull = V7_LR;
    return launch_pass_v7(true, job.scratch, (size_t)V7_R, d.batch, n, p, job.pdl, stream);
}

// Blowup-32 LDE 2^20 -> 2^25 (zero-padded forward transform of at most 2^21 coefficients, optional coset shift):
// the expansion pass (lde_expand.cuh: 32 coset transforms of 256 points per column) and pass 2 of the TMA-staged kernel
// on 8192 columns.  13 + 12 = 25 butterfly stages shrink to 8 + 12 and the data cross HBM twice instead of three times.
struct LdeJob {
    LdeParams lp;
    V7Params p;
    uint32_t* scratch;
};
static bool lde25_applies(const NttDesc& d) {
    return d.log_n == LDE::LOG_ROWS + LDE::LOG_COLS && !d.inverse && d.log_inner == 0 && d.batch == 1 && !d.scatter && d.n_in > 0 &&
           d.n_in <= ((size_t)1 << 21) && v7_enabled() && g_plan_override.find(d.log_n) == g_plan_override.end() &&
           ((((uintptr_t)d.in | (uintptr_t)d.out) & 15u) == 0);
}
static int lde25_prepare(DeviceState& st, const NttDesc& d, cudaStream_t stream, LdeJob* job) {
    const size_t n = (size_t)1 << d.log_n;
    if (!st.tw8192) {
        BB_CK(cudaMalloc(&st.tw8192, sizeof(uint2) * 4096));
        gen_shoup_kernel<<<4096 / 256, 256>>>(st.tw8192, 4096u, to_monty(root_of_unity(LDE::LOG_ROWS)));
        BB_CK(cudaGetLastError());
        BB_CK(cudaDeviceSynchronize());
    }
    V7Params& p = job->p;
    memset(&p, 0, sizeof p);
    int rc = pow_table_get(st, root_of_unity(d.log_n), d.log_n, 1u, &p.tab);
    if (rc) return rc;
    p.tab_scaled = p.tab;
    LdeParams& lp = job->lp;
    memset(&lp, 0, sizeof lp);
    lp.has_shift = d.coset_shift > 1;
    if (lp.has_shift) {
        rc = pow_table_get(st, d.coset_shift, d.log_n, 1u, &lp.shift);
        if (rc) return rc;
    }
    rc = scratch_get(st, stream, n, &job->scratch);
    if (rc) return rc;
    p.tw = st.tw[0];
    memcpy(p.tw16, st.tw16[0], sizeof p.tw16);
    p.exp_mask = (uint32_t)(n - 1);
    p.err = st.err_word;
    p.out = d.out;
    p.out_batch_stride = n;
    p.log_pfull = LDE::LOG_ROWS;
    lp.in = d.in;
    lp.n_coeffs = (uint32_t)d.n_in;
    lp.out = job->scratch;
    lp.tw = st.tw8192;
    return 0;
}
static int lde25_launch(LdeJob& job, cudaStream_t stream) {
    const size_t n = (size_t)1 << (LDE::LOG_ROWS + LDE::LOG_COLS);
    int rc = launch_lde_expand(job.lp, true, stream);
    if (rc) return rc;
    // Z[j][k1] is a row-major [4096][8192] matrix: 4096-point transforms down its 8192 columns, rows out[k2 * 8192 + k1]
    return launch_pass_v7(true, job.scratch, (size_t)1 << LDE::LOG_ROWS, 1, n, job.p, true, stream);
}

NttPlan ntt_plan_for(int log_n, int log_inner, size_t batch) {
    std::lock_guard<std::mutex> lk(g_mu);
    return plan_locked(log_n, log_inner, batch);
}

// ------------------------------------------------------------------ executor
void engine_drop_stream(cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    int dev = 0;
    cudaGetDevice(&dev);
    auto it = g_states.find(dev);
    if (it == g_states.end()) return;
    auto sc = it->second.scratch.find(stream);
    if (sc == it->second.scratch.end()) return;
    cudaStreamSynchronize(stream);
    cudaFree(sc->second.first);
    it->second.scratch.erase(sc);
}

int engine_warmup(int log_n, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState* st;
    int rc = state_get(&st);
    if (rc) return rc;
    if (log_n > MAX_LR) {
        uint32_t* s;
        rc = scratch_get(*st, stream, (size_t)1 << log_n, &s);
        if (rc) return rc;
        for (int inv = 0; inv < 2; inv++) {
            uint32_t w = root_of_unity(log_n);
            PowTable t;
            if (inv) {
                w = bb::inv(w);
                rc = pow_table_get(*st, w, log_n, bb::inv((uint32_t)(((uint64_t)1 << log_n) % P)), &t);
                if (rc) return rc;
            }
            rc = pow_table_get(*st, w, log_n, 1, &t);
            if (rc) return rc;
        }
    }
    return 0;
}

int ntt_execute(const NttDesc& d, cudaStream_t stream) {
    if (d.log_n < 0 || d.log_n > MAX_LOG_N) return (int)cudaErrorInvalidValue;
    if (d.batch == 0) return 0;
    // the lock covers planning, table and scratch look-ups; the kernel launches below run without it, so transforms
    // issued from several host threads (or for several devices) do not serialise on the host
    std::unique_lock<std::mutex> lk(g_mu);
    DeviceState* st;
    int rc = state_get(&st);
    if (rc) return rc;

    const int inv = d.inverse ? 1 : 0;
    const size_t n = (size_t)1 << d.log_n;
    const size_t inner = (size_t)1 << d.log_inner;
    const bool coset = d.coset_shift > 1;
    const bool use_v7 = d.log_n == 2 * V7_LR && d.log_inner == 0 && !coset && !d.scatter && d.n_in == n && v7_enabled() &&
                        g_plan_override.find(d.log_n) == g_plan_override.end() && ((((uintptr_t)d.in | (uintptr_t)d.out) & 15u) == 0) &&
                        (d.batch == 1 || (d.batch_stride_in % 4 == 0 && d.batch_stride_out % 4 == 0));
    if (lde25_applies(d)) {
        LdeJob job;
        rc = lde25_prepare(*st, d, stream, &job);
        lk.unlock();
        return rc ? rc : lde25_launch(job, stream);
    }
    if (use_v7) {
        V7Job job;
        rc = v7_prepare(*st, d, stream, &job);
        lk.unlock();
        return rc ? rc : v7_launch(d, job, stream);
    }
    NttPlan pl = plan_locked(d.log_n, d.log_inner, d.batch, false);
    if (d.scatter && pl.npass >= 1 && g_plan_override.find(d.log_n) == g_plan_override.end()) {
        // the scattering pass stores rows of 2^lc values to peer GPUs: 32 columns make them full 128-byte lines on NVLink
        const int last = pl.npass - 1, lc5 = pick_lc(pl.lr[last], 5);
        if (d.log_inner >= lc5) pl.lc[last] = lc5;
    }

    const uint32_t n_inv = bb::inv((uint32_t)(n % P));
    uint32_t omega = root_of_unity(d.log_n);
    if (inv) omega = bb::inv(omega);

    PowTable tw_first{}, tw_rest{};
    if (pl.npass > 1) {
        // inverse transforms fold n^-1 into the first inter-pass twiddle unless the coset epilogue carries it
        uint32_t scale_first = (inv && !coset) ? n_inv : 1u;
        rc = pow_table_get(*st, omega, d.log_n, scale_first, &tw_first);
        if (rc) return rc;
        if (pl.npass > 2) {
            rc = pow_table_get(*st, omega, d.log_n, 1u, &tw_rest);
            if (rc) return rc;
        }
    }
    PowTable coset_tab{};
    if (coset) {
        if (!inv)
            rc = pow_table_get(*st, d.coset_shift, d.log_n, 1u, &coset_tab);
        else
            rc = pow_table_get(*st, bb::inv(d.coset_shift), d.log_n, n_inv, &coset_tab);
        if (rc) return rc;
    }

    PowTable fs_tab{};
    if (d.scatter) {
        if (coset || d.batch != 1 || d.log_inner < 4) return (int)cudaErrorInvalidValue;
        uint32_t wn = root_of_unity(d.scatter->log_n);
        if (inv) wn = bb::inv(wn);
        rc = pow_table_get(*st, wn, d.scatter->log_n, 1u, &fs_tab);
        if (rc) return rc;
    }

    uint32_t* scratch = nullptr;
    if (pl.npass > 1) {
        rc = scratch_get(*st, stream, n * inner * d.batch, &scratch);
        if (rc) return rc;
    }

    const bool transposed = (pl.npass == 1 && d.log_inner == 0);
    if (transposed && d.batch > 1 && (d.batch_stride_out != n)) return (int)cudaErrorInvalidValue;

    const uint2* tw_dir = st->tw[inv];
    uint2 tw16_dir[8];
    memcpy(tw16_dir, st->tw16[inv], sizeof tw16_dir);
    const bool force_scalar = g_force_scalar;
    lk.unlock();

    int log_p = 0;  // log2(R_1 ... R_{i-1})
    for (int i = 0; i < pl.npass; i++) {
        const int lr = pl.lr[i], lc = pl.lc[i];
        const bool first = (i == 0), last = (i == pl.npass - 1);

        PassParams p;
        memset(&p, 0, sizeof p);
        // buffer chain: 1 pass in->out; 2 passes in->scratch->out; 3 passes in->scratch->out->out
        const uint32_t* src;
        uint32_t* dst;
        size_t src_bs, dst_bs;
        if (first) {
            src = d.in;
            src_bs = d.batch_stride_in;
        } else if (i == 1) {
            src = scratch;
            src_bs = n * inner;
        } else {
            src = d.out;
            src_bs = d.batch_stride_out;
        }
        if (last || i == 1) {
            dst = d.out;
            dst_bs = d.batch_stride_out;
        } else {
            dst = scratch;
            dst_bs = n * inner;
        }
        p.in = src;
        p.out = dst;
        p.tw = tw_dir;
        memcpy(p.tw16, tw16_dir, sizeof p.tw16);
        p.log_inner = (uint32_t)d.log_inner;
        // Programmatic dependent launch hides the launch gap and prologue between passes (2^24: 127 -> 120 us) but was
        // measured to slow transforms of >= 2^28 words in total by up to 27 % (64 x 2^22: 1.70 -> 2.15 ms), so it is
        // used below that size only (tools/batch22_time.py)
        p.pdl = (n * inner * d.batch < ((size_t)1 << 28)) ? 1u : 0u;
        p.n_in_limit = first ? (unsigned long long)d.n_in * inner : ~0ull;

        dim3 grid;
        if (transposed) {
            p.transposed = 1;
            p.ncols = (uint32_t)d.batch;
            p.in_row_stride = 1;
            p.in_col_stride = d.batch_stride_in;
            p.log_pfull = 0;
            p.in_batch_stride = p.out_batch_stride = 0;
            p.n_in_limit = d.n_in;
            grid = dim3((unsigned)((d.batch + ((size_t)1 << lc) - 1) >> lc), 1, 1);
        } else {
            const size_t ncols = (n >> lr) * inner;
            p.transposed = 0;
            p.ncols = (uint32_t)ncols;
            p.in_row_stride = ncols;
            p.in_col_stride = 1;
            p.log_pfull = (uint32_t)(log_p + d.log_inner);
            p.in_batch_stride = src_bs;
            p.out_batch_stride = dst_bs;
            grid = dim3((unsigned)((ncols + ((size_t)1 << lc) - 1) >> lc), (unsigned)d.batch, 1);
        }

        if (first && coset && !inv) {
            p.pro_mode = PRO_INIDX;
            p.pro = coset_tab;
        }
        if (!last) {
            p.epi_mode = EPI_TWIDDLE;
            p.epi = first ? tw_first : tw_rest;
            p.epi_shift = (uint32_t)log_p;
            p.epi_unscale = (first && inv && !coset) ? to_monty((uint32_t)(n % P)) : R_MOD_P;
        } else if (d.scatter) {
            p.epi_mode = EPI_FOURSTEP;
            p.epi = fs_tab;
            p.epi_const = (inv && pl.npass == 1) ? to_monty(n_inv) : 0u;
            for (int r = 0; r < 8; r++) p.fs_peer[r] = d.scatter->peer[r < d.scatter->nranks ? r : 0];
            p.fs_log_rows_per_rank = (uint32_t)(d.log_n - ceil_log2((size_t)d.scatter->nranks));
            p.fs_dst_row_stride = (uint32_t)d.scatter->dst_row_stride;
            p.fs_dst_col = (uint32_t)d.scatter->col_offset;
            p.fs_col_offset = (uint32_t)d.scatter->col_offset;
        } else if (coset && inv) {
            p.epi_mode = EPI_OUTIDX;
            p.epi = coset_tab;
        } else if (inv && pl.npass == 1) {
            p.epi_mode = EPI_CONST;
            p.epi_const = to_monty(n_inv);
        } else {
            p.epi_mode = EPI_NONE;
        }
        const bool aligned = ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) && (src_bs % 4 == 0) && (dst_bs % 4 == 0);
        // vectorised kernel whenever chunks of four columns stay whole; scalar kernel otherwise
        PassLaunchFn fn = nullptr;
        if (!transposed && lc >= 2 && aligned && p.log_pfull != 1 && (p.ncols & ((1u << lc) - 1u)) == 0 && !force_scalar)
            fn = pass_launcher_v4(lr, lc);
        if (fn && first && !inv && lr >= 6 && lr <= 9 && d.n_in * 32 == n && p.in_batch_stride % 4 == 0) {
            // exactly the first R/32 rows of every tile hold input (blowup 32): prune the copy-only stages
            p.prune_log = 5;
        }
        if (!fn && p.epi_mode == EPI_FOURSTEP) return (int)cudaErrorInvalidConfiguration;
        if (!fn) fn = pass_launcher(lr, lc);
        if (!fn) return (int)cudaErrorInvalidConfiguration;
        fn(p, grid, stream);
        BB_CK(cudaGetLastError());
        log_p += lr;
    }
    return 0;
}

}  // namespace bb
