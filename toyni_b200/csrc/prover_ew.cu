// Element-wise stages of the Fibonacci prover on the device: with these, a committed trace
// never leaves HBM between the LDE and the query phase.
//
//   constraint evaluation over the shifted domain            src/fibonacci.rs:133-143
//   quotient by the vanishing polynomial (32 distinct values of Z_H on the blowup-32 coset)   :147-150
//   DEEP composition with 1/(x - z) by batched inversion      :186-198
//   polynomial evaluation at one point (the out-of-domain evaluations)   :164-167, src/math/polynomial.rs:134-144
//   Merkle openings for a whole query set in one launch       :250-295, src/merkle.rs:50-80
//
// The reference walks these per point with Horner evaluations and one Fermat inverse each; field arithmetic is
// exact, so the same canonical values come out of the closed forms used here: T(g x_i) = lde[(i + 32) mod N]
// (the identity the verifier relies on, src/verifier.rs:128-129), x_i = 7 w^i from the twiddle cache, Z_H(x_i)
// periodic in i, and Montgomery's trick (one inversion per 8 points).
#include "fri_fold.cuh"
#include "merkle.cuh"
#include "ntt_engine.cuh"
#include "ntt_pass_v4.cuh"
#include "prover_ew.cuh"

namespace bb {

// c[i] = (T[i + 2 step] - T[i + step] - T[i]) (x_i - b1) (x_i - b2), indices mod n = 2^log_n
__global__ void __launch_bounds__(256) fib_constraint_kernel(const uint32_t* __restrict__ t, uint32_t* __restrict__ out, uint32_t log_n,
                                                             uint32_t step, uint32_t b1, uint32_t b2, PowTable xs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, mask = (1u << log_n) - 1u;
    if (i > mask) return;
    const uint32_t t0 = t[i], t1 = t[(i + step) & mask], t2 = t[(i + 2u * step) & mask];
    const uint32_t x = pow_plain(xs, i);
    uint32_t c = sub(t2, add(t1, t0));
    c = mul(c, sub(x, b1));
    c = mul(c, sub(x, b2));
    out[i] = c;
}

// v[i] *= tab[i mod period]   (period a power of two <= 64; the table holds Montgomery forms)
struct PeriodicTable {
    uint32_t m[64];
};
__global__ void __launch_bounds__(256) scale_periodic_kernel(uint32_t* __restrict__ v, size_t n, PeriodicTable tab, uint32_t pmask) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = monty_mul(v[i], tab.m[i & pmask]);
}

// d[i] = ((q[i] - q_z) + (T[i + 2 step] - t_ggz) + (T[i + step] - t_gz) + (T[i] - t_z)) / (x_i - z); 8 points per thread
// share one inversion.  x_i - z is never zero: z is drawn outside the shifted domain (src/fibonacci.rs:378-399).
__global__ void __launch_bounds__(128) fib_deep_kernel(const uint32_t* __restrict__ q, const uint32_t* __restrict__ t, uint32_t* __restrict__ out,
                                                       uint32_t log_n, uint32_t step, uint32_t z, uint32_t q_z, uint32_t t_z,
                                                       uint32_t t_gz, uint32_t t_ggz, PowTable xs, uint2 w) {
    constexpr int K = 8;
    const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * K, mask = (1u << log_n) - 1u;
    if (i0 > mask) return;
    uint32_t a[K], pre[K];
    uint32_t x = pow_plain(xs, i0);
#pragma unroll
    for (int k = 0; k < K; k++) {
        a[k] = to_monty(sub(x, z));
        pre[k] = k ? monty_mul(pre[k - 1], a[k]) : a[k];  // Montgomery forms throughout
        const uint32_t nx = shoup_mul_lazy(x, w.x, w.y);
        x = min(nx, nx - P);
    }
    // inverse of the running product: (pre R)^(p-2) in Montgomery form
    uint32_t inv_all = R_MOD_P, b = pre[K - 1];
    for (uint32_t e = P - 2u; e; e >>= 1) {
        if (e & 1u) inv_all = monty_mul(inv_all, b);
        b = monty_mul(b, b);
    }
#pragma unroll
    for (int k = K - 1; k >= 0; k--) {
        const uint32_t inv_k = k ? monty_mul(inv_all, pre[k - 1]) : inv_all;  // 1 / a[k], Montgomery form
        inv_all = monty_mul(inv_all, a[k]);
        const uint32_t i = i0 + (uint32_t)k;
        uint32_t num = sub(q[i], q_z);
        num = add(num, sub(t[(i + 2u * step) & mask], t_ggz));
        num = add(num, sub(t[(i + step) & mask], t_gz));
        num = add(num, sub(t[i], t_z));
        out[i] = monty_mul(num, inv_k);  // plain * Montgomery form = plain
    }
}

// sum_k c[k] z^k: 64 coefficients per thread by Horner, times z^(64 j), added into a 64-bit accumulator (canonical
// terms < 2^31, at most 2^21 of them, so it cannot overflow); the caller reduces mod p.
__global__ void __launch_bounds__(128) poly_eval_kernel(const uint32_t* __restrict__ c, size_t n, uint32_t z, unsigned long long* acc) {
    constexpr int CH = 64;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x, base = j * CH;
    if (base >= n) return;
    const uint32_t zm = to_monty(z);
    uint32_t v = 0;
    const int cnt = (n - base) < (size_t)CH ? (int)(n - base) : CH;
    for (int k = cnt - 1; k >= 0; k--) v = add(monty_mul(v, zm), c[base + k]);  // plain * Montgomery = plain
    uint32_t pw = R_MOD_P, bsq = zm;                                            // z^base, Montgomery form
    for (size_t e = base; e; e >>= 1) {
        if (e & 1) pw = monty_mul(pw, bsq);
        bsq = monty_mul(bsq, bsq);
    }
    atomicAdd(acc, (unsigned long long)monty_mul(v, pw));
}

// One block per query: the sibling digests of `index` level by level, exactly as src/merkle.rs:59-77.
__global__ void __launch_bounds__(32) gather_paths_kernel(const uint8_t* __restrict__ nodes, size_t nleaves, const unsigned long long* __restrict__ idx,
                                                          uint32_t depth, uint8_t* __restrict__ paths) {
    size_t level_off = 0, level_n = nleaves, cur = idx[blockIdx.x];
    uint8_t* dst = paths + (size_t)blockIdx.x * depth * 32;
    for (uint32_t d = 0; level_n > 1; d++) {
        const size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
        const size_t src = (sib >= level_n) ? cur : sib;
        dst[32 * d + threadIdx.x] = nodes[32 * (level_off + src) + threadIdx.x];
        cur /= 2;
        level_off += level_n;
        level_n = (level_n + 1) / 2;
    }
}
// out[q] = src[idx[q]] for fixed-size elements (values, salts)
__global__ void gather_elems_kernel(const uint8_t* __restrict__ src, uint32_t elem_bytes, const unsigned long long* __restrict__ idx,
                                    uint8_t* __restrict__ out) {
    for (uint32_t b = threadIdx.x; b < elem_bytes; b += blockDim.x)
        out[(size_t)blockIdx.x * elem_bytes + b] = src[(size_t)idx[blockIdx.x] * elem_bytes + b];
}

// The openings of a whole proof in one launch: one block per query, each with its own tree (src/fibonacci.rs:250-295 over
// src/merkle.rs:59-77).  The block copies the sibling digests level by level, then the opened value and its salt.
__global__ void __launch_bounds__(32) open_multi_kernel(const OpenQuery* __restrict__ qs, uint32_t val_bytes, uint8_t* __restrict__ paths,
                                                        uint8_t* __restrict__ vals, uint8_t* __restrict__ salts) {
    const OpenQuery q = qs[blockIdx.x];
    size_t level_off = 0, level_n = q.nleaves, cur = q.index;
    uint8_t* dst = paths + q.path_off;
    for (uint32_t d = 0; level_n > 1; d++) {
        const size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
        const size_t src = (sib >= level_n) ? cur : sib;
        dst[32 * d + threadIdx.x] = q.nodes[32 * (level_off + src) + threadIdx.x];
        cur /= 2;
        level_off += level_n;
        level_n = (level_n + 1) / 2;
    }
    if (threadIdx.x < val_bytes) vals[(size_t)blockIdx.x * val_bytes + threadIdx.x] = q.vals[q.index * val_bytes + threadIdx.x];
    if (threadIdx.x < 16u) salts[(size_t)blockIdx.x * 16 + threadIdx.x] = q.salts ? q.salts[q.index * 16 + threadIdx.x] : (uint8_t)0;
}

// dst[(j*G + r)] = src[r*c + j], elements of `limbs` words (re-layout after the cyclic -> block exchange)
__global__ void __launch_bounds__(256) interleave_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t groups,
                                                         size_t chunk, uint32_t limbs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // destination element
    if (i >= (size_t)groups * chunk) return;
    const size_t j = i / groups, r = i - j * groups;
    for (uint32_t l = 0; l < limbs; l++) dst[i * limbs + l] = src[(r * chunk + j) * limbs + l];
}

// ---------------------------------------------------------------- host-side launchers
int interleave(const uint32_t* d_src, uint32_t* d_dst, uint32_t groups, size_t chunk, uint32_t limbs, cudaStream_t s) {
    const size_t n = (size_t)groups * chunk;
    if (n) interleave_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_src, d_dst, groups, chunk, limbs);
    return (int)cudaGetLastError();
}

static int coset_table(int log_n, uint32_t shift, PowTable* xs) {
    if (log_n < 1 || log_n > MAX_LOG_N) return (int)cudaErrorInvalidValue;
    return engine_pow_table(root_of_unity((uint32_t)log_n), log_n, shift % P, xs);
}

int fib_constraint(const uint32_t* d_t, uint32_t* d_out, int log_n, uint32_t step, uint32_t shift, uint32_t b1, uint32_t b2,
                   cudaStream_t s) {
    PowTable xs;
    int rc = coset_table(log_n, shift, &xs);
    if (rc) return rc;
    const size_t n = (size_t)1 << log_n;
    fib_constraint_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_t, d_out, (uint32_t)log_n, step, b1 % P, b2 % P, xs);
    return (int)cudaGetLastError();
}

int scale_periodic(uint32_t* d_v, size_t n, const uint32_t* h_table, uint32_t period, cudaStream_t s) {
    if (period == 0 || period > 64 || (period & (period - 1))) return (int)cudaErrorInvalidValue;
    PeriodicTable tab;
    for (uint32_t i = 0; i < 64; i++) tab.m[i] = to_monty(h_table[i % period] % P);
    if (n) scale_periodic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_v, n, tab, period - 1);
    return (int)cudaGetLastError();
}

int fib_deep(const uint32_t* d_q, const uint32_t* d_t, uint32_t* d_out, int log_n, uint32_t step, uint32_t shift, uint32_t z,
             uint32_t q_z, uint32_t t_z, uint32_t t_gz, uint32_t t_ggz, cudaStream_t s) {
    if (log_n < 3) return (int)cudaErrorInvalidValue;
    PowTable xs;
    int rc = coset_table(log_n, shift, &xs);
    if (rc) return rc;
    const uint32_t w = root_of_unity((uint32_t)log_n);
    const size_t threads = ((size_t)1 << log_n) / 8;
    fib_deep_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(d_q, d_t, d_out, (uint32_t)log_n, step, z % P, q_z % P, t_z % P,
                                                                      t_gz % P, t_ggz % P, xs, make_uint2(w, shoup_companion(w)));
    return (int)cudaGetLastError();
}

int poly_eval(const uint32_t* d_c, size_t n, uint32_t z, unsigned long long* d_acc, cudaStream_t s) {
    int rc = (int)cudaMemsetAsync(d_acc, 0, sizeof(unsigned long long), s);
    if (rc || n == 0) return rc;
    const size_t threads = (n + 63) / 64;
    poly_eval_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(d_c, n, z % P, d_acc);
    return (int)cudaGetLastError();
}

int merkle_gather_paths(const uint8_t* d_nodes, size_t nleaves, const unsigned long long* d_idx, size_t nq, uint32_t depth, uint8_t* d_paths,
                        cudaStream_t s) {
    if (nq && depth) gather_paths_kernel<<<(unsigned)nq, 32, 0, s>>>(d_nodes, nleaves, d_idx, depth, d_paths);
    return (int)cudaGetLastError();
}
int merkle_open_multi(const OpenQuery* d_queries, size_t nq, uint32_t val_bytes, uint8_t* d_paths, uint8_t* d_vals, uint8_t* d_salts,
                      cudaStream_t s) {
    if (val_bytes == 0 || val_bytes > 32) return (int)cudaErrorInvalidValue;
    if (nq) open_multi_kernel<<<(unsigned)nq, 32, 0, s>>>(d_queries, val_bytes, d_paths, d_vals, d_salts);
    return (int)cudaGetLastError();
}
int gather_elems(const void* d_src, uint32_t elem_bytes, const unsigned long long* d_idx, size_t nq, void* d_out, cudaStream_t s) {
    if (nq && elem_bytes) gather_elems_kernel<<<(unsigned)nq, 32, 0, s>>>((const uint8_t*)d_src, elem_bytes, d_idx, (uint8_t*)d_out);
    return (int)cudaGetLastError();
}

}  // namespace bb
