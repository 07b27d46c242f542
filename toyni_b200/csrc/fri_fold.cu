// FRI layer folding on the device.
//   out[i] = (a+b)/2 + (a-b)/2 * beta * x_i^-1,  a = evals[i], b = evals[i+half]
// follows src/math/fri.rs:27-48 (base field) and :7-25 (extension field).  The reference spends one Fermat
// inverse (61 modmuls) per output; here x_i = x0 * omega_m^i is generated on chip from the cached
// omega_N^-t power table, so a fold reads 2 values, writes 1 and needs no xs array.
// A general-xs form (arbitrary evaluation points, per-element Fermat inverse) keeps the reference signature.
#include "fri_fold.cuh"

#include <cstring>

#include "merkle.cuh"
#include "ntt_pass.cuh"
#include "sha256.cuh"

namespace bb {

// exponent of omega_N^-1 for local index t:  ((t * idx_mul + idx_add) << shift)
struct FoldIdx {
    uint32_t idx_mul, idx_add, shift;
};

// HASH: 0 fold only; 1 also write the unsalted leaf digest of the new value; 2 the salted one ("fold and next layer's
// leaves" in one kernel: the folded value is hashed while it is still in registers)
template <int HASH>
__global__ void __launch_bounds__(256) fold_base_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        size_t half, PowTable winv, FoldIdx fi, uint32_t c_m,
                                                        const uint4* __restrict__ salts, uint8_t* __restrict__ leaf_nodes) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    uint32_t a = in[i], b = in[i + half];
    uint32_t tw = monty_mul(pow_lookup(winv, ((uint32_t)i * fi.idx_mul + fi.idx_add) << fi.shift), c_m);  // Montgomery form
    const uint32_t r = add(halve(add(a, b)), monty_mul(sub(a, b), tw));
    out[i] = r;
    if (HASH) {
        uint32_t v[1] = {r};
        Sha s;
        if (HASH == 2)
            leaf_digest<1, true>(v, salts[i], s);
        else
            leaf_digest<1, false>(v, make_uint4(0, 0, 0, 0), s);
        store_digest(leaf_nodes + 32 * i, s);
    }
}

// Ext * (Ext constant): out_k = sum_j d_j * C[k][j] with C the 4x4 multiplication matrix of the constant
// (W-scaled where X^4 wraps, src/ext.rs:186-189), entries in Montgomery form.  The four 62-bit products of a row are
// accumulated in 64 bits (4 p^2 < 2^64) and reduced ONCE: 4 IMAD.WIDE + 1 Montgomery reduction per limb instead of
// 4-5 full modular multiplications.
struct ExtMat {
    uint32_t m[4][4];
};

__device__ __forceinline__ uint32_t redc64(unsigned long long t) {  // t < 2^64 -> t * 2^-32 mod p, canonical
    const uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    const uint32_t mm = lo * P_INV;
    const uint32_t u = __umulhi(mm, P);
    uint32_t r = hi - u;              // true value in (-p, 2p): hi < 2^32 < 2.14 p
    if (hi < u) r += P;               // negative -> + p  (then in [0, p))
    return min(r, r - P);             // [0, 2p) -> [0, p)
}

__device__ __forceinline__ Ext ext_mul_const(const Ext& d, const ExtMat& c) {
    Ext r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned long long acc = (unsigned long long)d.c[0] * c.m[k][0];
        acc += (unsigned long long)d.c[1] * c.m[k][1];
        acc += (unsigned long long)d.c[2] * c.m[k][2];
        acc += (unsigned long long)d.c[3] * c.m[k][3];
        r.c[k] = redc64(acc);
    }
    return r;
}

// PER_THREAD outputs per thread: more independent 16-byte loads in flight per warp
template <int PER_THREAD, int HASH>
__global__ void __launch_bounds__(256) fold_ext_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t half,
                                                       PowTable winv, FoldIdx fi, ExtMat cm, const uint4* __restrict__ salts,
                                                       uint8_t* __restrict__ leaf_nodes) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x) * PER_THREAD + threadIdx.x;
    uint4 av[PER_THREAD], bv[PER_THREAD];
#pragma unroll
    for (int u = 0; u < PER_THREAD; u++) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < half) {
            av[u] = in[i];
            bv[u] = in[i + half];
        }
    }
#pragma unroll
    for (int u = 0; u < PER_THREAD; u++) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i >= half) continue;
        Ext a{{av[u].x, av[u].y, av[u].z, av[u].w}}, b{{bv[u].x, bv[u].y, bv[u].z, bv[u].w}};
        const uint32_t tw = pow_lookup(winv, ((uint32_t)i * fi.idx_mul + fi.idx_add) << fi.shift);  // omega^-i, Montgomery form
        Ext s = ext_add(a, b), d = ext_sub(a, b);
#pragma unroll
        for (int k = 0; k < 4; k++) d.c[k] = monty_mul(d.c[k], tw);  // (a-b) * omega^-i, plain
        const Ext t = ext_mul_const(d, cm);                             // * (beta/2 * x0^-1)
        uint4 r;
        r.x = add(halve(s.c[0]), t.c[0]);
        r.y = add(halve(s.c[1]), t.c[1]);
        r.z = add(halve(s.c[2]), t.c[2]);
        r.w = add(halve(s.c[3]), t.c[3]);
        out[i] = r;
        if (HASH) {
            uint32_t v[4] = {r.x, r.y, r.z, r.w};
            Sha h;
            if (HASH == 2)
                leaf_digest<4, true>(v, salts[i], h);
            else
                leaf_digest<4, false>(v, make_uint4(0, 0, 0, 0), h);
            store_digest(leaf_nodes + 32 * i, h);
        }
    }
}

// general evaluation points: x^-1 by Fermat, exactly as src/babybear.rs:111-114 (but in Montgomery form)
__device__ __forceinline__ uint32_t inv_monty_dev(uint32_t x) {  // returns Montgomery form of x^-1
    uint32_t b = to_monty(x), r = R_MOD_P;
    uint32_t e = P - 2;
#pragma unroll 1
    while (e) {
        if (e & 1u) r = monty_mul(r, b);
        b = monty_mul(b, b);
        e >>= 1;
    }
    return r;
}

__global__ void __launch_bounds__(256) fold_base_xs_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ xs,
                                                           uint32_t* __restrict__ out, size_t half, uint32_t halfbeta_m) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    uint32_t a = in[i], b = in[i + half];
    uint32_t tw = monty_mul(inv_monty_dev(xs[i]), halfbeta_m);
    out[i] = add(halve(add(a, b)), monty_mul(sub(a, b), tw));
}

__global__ void __launch_bounds__(256) fold_ext_xs_kernel(const uint4* __restrict__ in, const uint32_t* __restrict__ xs,
                                                          uint4* __restrict__ out, size_t half, Ext halfbeta_m) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    uint4 av = in[i], bv = in[i + half];
    Ext a{{av.x, av.y, av.z, av.w}}, b{{bv.x, bv.y, bv.z, bv.w}};
    uint32_t xi = inv_monty_dev(xs[i]);
    Ext g;
#pragma unroll
    for (int k = 0; k < 4; k++) g.c[k] = monty_mul(halfbeta_m.c[k], xi);
    Ext s = ext_add(a, b), d = ext_sub(a, b);
    Ext t = ext_mul_monty(d, g);
    uint4 r;
    r.x = add(halve(s.c[0]), t.c[0]);
    r.y = add(halve(s.c[1]), t.c[1]);
    r.z = add(halve(s.c[2]), t.c[2]);
    r.w = add(halve(s.c[3]), t.c[3]);
    out[i] = r;
}

static inline unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

int fri_fold_coset(const uint32_t* d_in, uint32_t* d_out, size_t m_local, int limbs, int log_m_global, uint32_t x0,
                   const uint32_t beta[4], uint32_t idx_mul, uint32_t idx_add, cudaStream_t s, int hash_mode,
                   const uint8_t* d_salts, uint8_t* d_leaf_nodes) {
    if (m_local < 2 || (m_local & 1) || log_m_global < 1 || log_m_global > MAX_LOG_N || x0 == 0) return (int)cudaErrorInvalidValue;
    PowTable winv;
    // omega_m^-t table; shared with the inverse NTT of the same size
    int rc = engine_pow_table(bb::inv(root_of_unity(log_m_global)), log_m_global, 1u, &winv);
    if (rc) return rc;
    const size_t half = m_local / 2;
    const uint32_t hx = bb::mul(HALF, bb::inv(x0));  // (1/2) * x0^-1
    FoldIdx fi{idx_mul, idx_add, 0};
    if (limbs == 1) {
        const uint32_t cmul = to_monty(bb::mul(hx, beta[0]));
        const uint4* sl = reinterpret_cast<const uint4*>(d_salts);
        if (hash_mode == 2)
            fold_base_kernel<2><<<blocks_for(half), 256, 0, s>>>(d_in, d_out, half, winv, fi, cmul, sl, d_leaf_nodes);
        else if (hash_mode == 1)
            fold_base_kernel<1><<<blocks_for(half), 256, 0, s>>>(d_in, d_out, half, winv, fi, cmul, sl, d_leaf_nodes);
        else
            fold_base_kernel<0><<<blocks_for(half), 256, 0, s>>>(d_in, d_out, half, winv, fi, cmul, sl, d_leaf_nodes);
    } else {
        // multiplication matrix of c = beta * (1/2 x0^-1): (d*c)_k = sum_j d_j c_(k-j), wrapped terms times W = 11
        uint32_t c[4];
        for (int k = 0; k < 4; k++) c[k] = bb::mul(hx, beta[k]);
        ExtMat cm;
        for (int k = 0; k < 4; k++)
            for (int j = 0; j < 4; j++) {
                uint32_t v = (k >= j) ? c[k - j] : bb::mul(EXT_W, c[k - j + 4]);
                cm.m[k][j] = to_monty(v);
            }
        constexpr int PT = 2;
        const unsigned blocks = (unsigned)((half + 256 * PT - 1) / (256 * PT));
        const uint4* sl = reinterpret_cast<const uint4*>(d_salts);
        if (hash_mode == 2)
            fold_ext_kernel<PT, 2><<<blocks, 256, 0, s>>>((const uint4*)d_in, (uint4*)d_out, half, winv, fi, cm, sl, d_leaf_nodes);
        else if (hash_mode == 1)
            fold_ext_kernel<PT, 1><<<blocks, 256, 0, s>>>((const uint4*)d_in, (uint4*)d_out, half, winv, fi, cm, sl, d_leaf_nodes);
        else
            fold_ext_kernel<PT, 0><<<blocks, 256, 0, s>>>((const uint4*)d_in, (uint4*)d_out, half, winv, fi, cm, sl, d_leaf_nodes);
    }
    return (int)cudaGetLastError();
}

// The short end of a fold chain in ONE launch.  Below ~2^13 local values a fold is a few microseconds of launch and drain
// around almost no work, and a 2^25 -> 16 chain has a dozen of them: a single CTA walks those layers with a block barrier
// between them.  Layer f of the tail has evaluation points omega_M^(t << f) (M = the first tail layer's global size), so
// one power table serves all layers; the per-fold constant beta_f / (2 x0_f) arrives as a kernel parameter.
constexpr int FOLD_TAIL_MAX_FOLDS = 27, FOLD_TAIL_THREADS = 1024;
constexpr size_t FOLD_TAIL_MAX_LOCAL = (size_t)1 << 13;
struct FoldTailParams {
    uint32_t nfolds, idx_mul, idx_add;
    uint32_t c[FOLD_TAIL_MAX_FOLDS][16];  // limbs = 4: the ExtMat of the fold; limbs = 1: c[f][0] = Montgomery form of the constant
};

template <int LIMBS>
__global__ void __launch_bounds__(FOLD_TAIL_THREADS) fold_tail_kernel(const uint32_t* in, uint32_t* out, size_t half, PowTable winv,
                                                                      const FoldTailParams p) {
    const uint32_t* src = in;
    uint32_t* dst = out;
    for (uint32_t f = 0; f < p.nfolds; f++) {
        for (size_t i = threadIdx.x; i < half; i += FOLD_TAIL_THREADS) {
            const uint32_t tw = pow_lookup(winv, ((uint32_t)i * p.idx_mul + p.idx_add) << f);  // omega_m^-i of this layer, Montgomery form
            if (LIMBS == 1) {
                const uint32_t a = __ldcg(src + i), b = __ldcg(src + i + half);  // L2: written by other threads of this CTA
                dst[i] = add(halve(add(a, b)), monty_mul(sub(a, b), monty_mul(tw, p.c[f][0])));
            } else {
                const uint4 av = __ldcg(reinterpret_cast<const uint4*>(src) + i), bv = __ldcg(reinterpret_cast<const uint4*>(src) + i + half);
                Ext a{{av.x, av.y, av.z, av.w}}, b{{bv.x, bv.y, bv.z, bv.w}};
                Ext s = ext_add(a, b), d = ext_sub(a, b);
#pragma unroll
                for (int k = 0; k < 4; k++) d.c[k] = monty_mul(d.c[k], tw);
                ExtMat cm;
#pragma unroll
                for (int k = 0; k < 16; k++) cm.m[k >> 2][k & 3] = p.c[f][k];
                const Ext t = ext_mul_const(d, cm);
                reinterpret_cast<uint4*>(dst)[i] =
                    make_uint4(add(halve(s.c[0]), t.c[0]), add(halve(s.c[1]), t.c[1]), add(halve(s.c[2]), t.c[2]), add(halve(s.c[3]), t.c[3]));
            }
        }
        __syncthreads();  // block-scope ordering of the stores above with the (L2) loads of the next layer
        src = dst;
        dst += half * LIMBS;
        half /= 2;
    }
}

static void fold_constant(uint32_t x0, const uint32_t* beta, int limbs, uint32_t out[16]) {
    const uint32_t hx = bb::mul(HALF, bb::inv(x0));  // (1/2) * x0^-1
    if (limbs == 1) {
        out[0] = to_monty(bb::mul(hx, beta[0]));
        return;
    }
    uint32_t c[4];
    for (int k = 0; k < 4; k++) c[k] = bb::mul(hx, beta[k]);
    for (int k = 0; k < 4; k++)
        for (int j = 0; j < 4; j++) out[4 * k + j] = to_monty((k >= j) ? c[k - j] : bb::mul(EXT_W, c[k - j + 4]));
}

bool fri_fold_tail_applies(size_t m_local, size_t nfolds) { return m_local <= FOLD_TAIL_MAX_LOCAL && nfolds >= 2 && nfolds <= (size_t)FOLD_TAIL_MAX_FOLDS; }

int fri_fold_chain_tail(const uint32_t* d_in, uint32_t* d_out, size_t m_local, int limbs, int log_m_global, uint32_t x0,
                        const uint32_t* betas, size_t nfolds, uint32_t idx_mul, uint32_t idx_add, cudaStream_t s) {
    if (!fri_fold_tail_applies(m_local, nfolds) || (m_local >> nfolds) == 0 || (m_local & (m_local - 1)) || (limbs != 1 && limbs != 4) ||
        log_m_global < 1 || log_m_global > MAX_LOG_N || x0 == 0)
        return (int)cudaErrorInvalidValue;
    PowTable winv;
    int rc = engine_pow_table(bb::inv(root_of_unity(log_m_global)), log_m_global, 1u, &winv);
    if (rc) return rc;
    FoldTailParams p;
    memset(&p, 0, sizeof p);
    p.nfolds = (uint32_t)nfolds;
    p.idx_mul = idx_mul;
    p.idx_add = idx_add;
    for (size_t f = 0; f < nfolds; f++) {
        fold_constant(x0, betas + f * (size_t)limbs, limbs, p.c[f]);
        x0 = bb::mul(x0, x0);
    }
    if (limbs == 1)
        fold_tail_kernel<1><<<1, FOLD_TAIL_THREADS, 0, s>>>(d_in, d_out, m_local / 2, winv, p);
    else
        fold_tail_kernel<4><<<1, FOLD_TAIL_THREADS, 0, s>>>(d_in, d_out, m_local / 2, winv, p);
    return (int)cudaGetLastError();
}

int fri_fold_xs(const uint32_t* d_in, const uint32_t* d_xs, uint32_t* d_out, size_t m, int limbs, const uint32_t beta[4],
                cudaStream_t s) {
    if (m < 2 || (m & 1)) return (int)cudaErrorInvalidValue;
    const size_t half = m / 2;
    if (limbs == 1) {
        fold_base_xs_kernel<<<blocks_for(half), 256, 0, s>>>(d_in, d_xs, d_out, half, to_monty(bb::mul(HALF, beta[0])));
    } else {
        Ext c;
        for (int k = 0; k < 4; k++) c.c[k] = to_monty(bb::mul(HALF, beta[k]));
        fold_ext_xs_kernel<<<blocks_for(half), 256, 0, s>>>((const uint4*)d_in, d_xs, (uint4*)d_out, half, c);
    }
    return (int)cudaGetLastError();
}

}  // namespace bb
