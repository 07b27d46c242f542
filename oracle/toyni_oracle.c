/*
 * toyni_oracle.c — CPU restatement of the jonas089/toyni prover hot path (see toyni_oracle.h).
 * TEST INFRASTRUCTURE ONLY: the checker and the CPU baseline, never the product path.
 * Loop order and arithmetic follow the cited reference lines so that timings are a fair
 * stand-in for the reference's own single-threaded CPU path (no Rust toolchain exists here).
 */
#include "toyni_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ BabyBear */

uint64_t to_bb_new(uint64_t v) { return v % TO_P; } /* src/babybear.rs:26-30 */

static inline uint64_t reduce_wide(uint64_t v) { /* src/babybear.rs:80-89: two conditional subtracts */
    if (v >= TO_P) v -= TO_P;
    if (v >= TO_P) v -= TO_P;
    return v;
}

uint64_t to_bb_add(uint64_t a, uint64_t b) { return reduce_wide(a + b); } /* :133-138 */

uint64_t to_bb_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + TO_P - b; } /* :151-158 */

uint64_t to_bb_mul(uint64_t a, uint64_t b) { /* :173-177: widen to 128 bits, then % p */
    u128 prod = (u128)a * (u128)b;
    return (uint64_t)(prod % (u128)TO_P);
}

uint64_t to_bb_neg(uint64_t a) { return a == 0 ? 0 : TO_P - a; } /* :197-205 */

uint64_t to_bb_pow(uint64_t a, uint64_t e) { /* :91-108 */
    if (e == 0) return 1;
    uint64_t base = a, result = 1;
    while (e > 0) {
        if (e & 1) result = to_bb_mul(result, base);
        base = to_bb_mul(base, base);
        e >>= 1;
    }
    return result;
}

uint64_t to_bb_inverse(uint64_t a) { /* :111-114: Fermat, a^(p-2) */
    if (a == 0) return 0;             /* the reference asserts; callers here never pass 0 */
    return to_bb_pow(a, TO_P - 2);
}

uint64_t to_bb_root_of_unity(uint32_t log_n) { /* :118-126 */
    if (log_n > 27) return 0;
    return to_bb_pow(to_bb_new(440564289ULL), 1ULL << (27 - log_n));
}

/* ----------------------------------------------------------------------- Ext */

void to_ext_add(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
    for (int k = 0; k < 4; k++) r[k] = to_bb_add(a[k], b[k]);
}

void to_ext_sub(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
    for (int k = 0; k < 4; k++) r[k] = to_bb_sub(a[k], b[k]);
}

void to_ext_mul_base(const uint64_t a[4], uint64_t s, uint64_t r[4]) { /* src/ext.rs:76-78 */
    for (int k = 0; k < 4; k++) r[k] = to_bb_mul(a[k], s);
}

void to_ext_mul(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) { /* src/ext.rs:178-192 */
#define M(x, y) to_bb_mul((x), (y))
#define A(x, y) to_bb_add((x), (y))
    const uint64_t w = to_bb_new(11); /* X^4 = 11, src/ext.rs:20 */
    uint64_t r0 = A(M(a[0], b[0]), M(w, A(A(M(a[1], b[3]), M(a[2], b[2])), M(a[3], b[1]))));
    uint64_t r1 = A(A(M(a[0], b[1]), M(a[1], b[0])), M(w, A(M(a[2], b[3]), M(a[3], b[2]))));
    uint64_t r2 = A(A(A(M(a[0], b[2]), M(a[1], b[1])), M(a[2], b[0])), M(w, M(a[3], b[3])));
    uint64_t r3 = A(A(A(M(a[0], b[3]), M(a[1], b[2])), M(a[2], b[1])), M(a[3], b[0]));
#undef M
#undef A
    r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3;
}

void to_ext_inverse(const uint64_t a[4], uint64_t r[4]) { /* src/ext.rs:107-127: a^(p^4-2) */
    u128 p = TO_P;
    u128 e = p * p * p * p - 2;
    uint64_t base[4] = {a[0], a[1], a[2], a[3]};
    uint64_t res[4] = {1, 0, 0, 0};
    while (e > 0) {
        if (e & 1) to_ext_mul(res, base, res);
        to_ext_mul(base, base, base);
        e >>= 1;
    }
    memcpy(r, res, sizeof res);
}

/* ----------------------------------------------------------------------- NTT */

static inline size_t bit_reverse(size_t x, unsigned log_n) { /* src/ntt.rs:14-21 */
    size_t r = 0;
    for (unsigned i = 0; i < log_n; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

static unsigned log2_exact(size_t n) {
    unsigned l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

void to_ntt(uint64_t* v, size_t n, uint64_t omega) { /* src/ntt.rs:24-53 */
    unsigned log_n = log2_exact(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = bit_reverse(i, log_n);
        if (i < j) {
            uint64_t t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    }
    for (size_t len = 2; len <= n; len *= 2) {
        size_t step = n / len;
        uint64_t w_len = to_bb_pow(omega, step);
        for (size_t i = 0; i < n; i += len) {
            uint64_t w = 1;
            for (size_t j = 0; j < len / 2; j++) {
                uint64_t u = v[i + j];
                uint64_t t = to_bb_mul(v[i + j + len / 2], w);
                v[i + j] = to_bb_add(u, t);
                v[i + j + len / 2] = to_bb_sub(u, t);
                w = to_bb_mul(w, w_len);
            }
        }
    }
}

/* Thread count for the element-wise loops below (coset shift, folds, leaf / node hashing) and for the transforms inside
 * to_domain_fft / to_domain_ifft.  1 (the default) runs the reference's serial loops verbatim; > 1 cuts the same loops
 * into chunks over OpenMP — identical arithmetic per element (a chunk restarts a running power at shift^i0, exactly the
 * value the serial loop reaches there), so the results are the same bits.  Used for full-size parity checks. */
static int g_threads = 1;
void to_set_threads(int threads) { g_threads = threads > 1 ? threads : 1; }
int to_get_threads(void) { return g_threads; }

static void scale_by_powers(uint64_t* v, size_t n, uint64_t g) { /* v[i] *= g^i, the loops of src/math/domain.rs:154-174 */
    if (g_threads <= 1 || n < 8192) {
        uint64_t sp = 1;
        for (size_t i = 0; i < n; i++) {
            v[i] = to_bb_mul(v[i], sp);
            sp = to_bb_mul(sp, g);
        }
        return;
    }
    const size_t chunk = 4096, nchunks = (n + chunk - 1) / chunk;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (size_t c = 0; c < nchunks; c++) {
        size_t i0 = c * chunk, i1 = i0 + chunk < n ? i0 + chunk : n;
        uint64_t sp = to_bb_pow(g, (uint64_t)i0);
        for (size_t i = i0; i < i1; i++) {
            v[i] = to_bb_mul(v[i], sp);
            sp = to_bb_mul(sp, g);
        }
    }
}

void to_intt(uint64_t* v, size_t n, uint64_t omega) { /* src/ntt.rs:56-66 */
    uint64_t inv_omega = to_bb_pow(omega, (uint64_t)n - 1);
    to_ntt(v, n, inv_omega);
    uint64_t inv_n = to_bb_inverse(to_bb_new((uint64_t)n));
    for (size_t i = 0; i < n; i++) v[i] = to_bb_mul(v[i], inv_n);
}

/* All-cores variant: identical arithmetic per butterfly; each stage's work is cut into
 * contiguous chunks of the (group, j) iteration space and a chunk restarts the running twiddle
 * at w_len^j (exactly the value the serial loop reaches there). */
void to_ntt_mt(uint64_t* v, size_t n, uint64_t omega, int threads) {
    if (threads <= 1 || n < 4096) {
        to_ntt(v, n, omega);
        return;
    }
    unsigned log_n = log2_exact(n);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        size_t j = bit_reverse(i, log_n);
        if (i < j) {
            uint64_t t = v[i];
            v[i] = v[j];
            v[j] = t;
        }
    }
    for (size_t len = 2; len <= n; len *= 2) {
        size_t half = len / 2;
        uint64_t w_len = to_bb_pow(omega, n / len);
        size_t chunk = half < 1024 ? half : 1024; /* butterflies per restart of the running twiddle */
        size_t chunks_per_group = half / chunk;
        size_t total_chunks = (n / len) * chunks_per_group;
#pragma omp parallel for num_threads(threads) schedule(static)
        for (size_t c = 0; c < total_chunks; c++) {
            size_t g = c / chunks_per_group, j0 = (c % chunks_per_group) * chunk;
            size_t i = g * len;
            uint64_t w = to_bb_pow(w_len, j0);
            for (size_t j = j0; j < j0 + chunk; j++) {
                uint64_t u = v[i + j];
                uint64_t t = to_bb_mul(v[i + j + half], w);
                v[i + j] = to_bb_add(u, t);
                v[i + j + half] = to_bb_sub(u, t);
                w = to_bb_mul(w, w_len);
            }
        }
    }
}

void to_intt_mt(uint64_t* v, size_t n, uint64_t omega, int threads) {
    uint64_t inv_omega = to_bb_pow(omega, (uint64_t)n - 1);
    to_ntt_mt(v, n, inv_omega, threads);
    uint64_t inv_n = to_bb_inverse(to_bb_new((uint64_t)n));
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (size_t i = 0; i < n; i++) v[i] = to_bb_mul(v[i], inv_n);
}

void to_roots_of_unity_domain(uint64_t* out, size_t n) { /* src/ntt.rs:69-81 */
    uint64_t omega = to_bb_root_of_unity(log2_exact(n));
    uint64_t cur = 1;
    for (size_t i = 0; i < n; i++) {
        out[i] = cur;
        cur = to_bb_mul(cur, omega);
    }
}

/* -------------------------------------------------------------------- domain */

void to_domain_elements(uint64_t* out, size_t size, uint64_t shift) { /* src/math/domain.rs:61-69 */
    uint64_t omega = to_bb_root_of_unity(log2_exact(size));
    uint64_t cur = shift;
    for (size_t i = 0; i < size; i++) {
        out[i] = cur;
        cur = to_bb_mul(cur, omega);
    }
}

void to_domain_fft(const uint64_t* coeffs, size_t ncoeffs, size_t size, uint64_t shift, uint64_t* out) {
    /* src/math/domain.rs:107-123: to_vec + resize (zero-pad or truncate) */
    size_t take = ncoeffs < size ? ncoeffs : size;
    memcpy(out, coeffs, take * sizeof(uint64_t));
    memset(out + take, 0, (size - take) * sizeof(uint64_t));
    if (shift != 1) scale_by_powers(out, size, shift); /* apply_coset_shift, :154-162 */
    to_ntt_mt(out, size, to_bb_root_of_unity(log2_exact(size)), g_threads);
}

void to_domain_ifft(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* out) {
    /* src/math/domain.rs:85-102 */
    memcpy(out, evals, size * sizeof(uint64_t));
    if (g_threads > 1)
        to_intt_mt(out, size, to_bb_root_of_unity(log2_exact(size)), g_threads);
    else
        to_intt(out, size, to_bb_root_of_unity(log2_exact(size)));
    if (shift != 1) scale_by_powers(out, size, to_bb_inverse(shift)); /* undo_coset_shift, :165-174 */
}

static void transform_ext(const uint64_t* in, size_t nin, size_t size, uint64_t shift, uint64_t* out, int inverse) {
    /* src/math/domain.rs:140-151: split into four coordinate vectors, transform each, re-interleave */
    uint64_t* coord = (uint64_t*)malloc(nin * sizeof(uint64_t));
    uint64_t* res = (uint64_t*)malloc(size * sizeof(uint64_t));
    for (int k = 0; k < 4; k++) {
        for (size_t i = 0; i < nin; i++) coord[i] = in[4 * i + k];
        if (inverse)
            to_domain_ifft(coord, size, shift, res);
        else
            to_domain_fft(coord, nin, size, shift, res);
        for (size_t i = 0; i < size; i++) out[4 * i + k] = res[i];
    }
    free(coord);
    free(res);
}

void to_domain_fft_ext(const uint64_t* coeffs, size_t ncoeffs, size_t size, uint64_t shift, uint64_t* out) {
    transform_ext(coeffs, ncoeffs, size, shift, out, 0);
}

void to_domain_ifft_ext(const uint64_t* evals, size_t size, uint64_t shift, uint64_t* out) {
    transform_ext(evals, size, size, shift, out, 1);
}

/* ------------------------------------------------------------------ FRI fold */

void to_fri_fold(const uint64_t* evals, size_t m, const uint64_t* xs, uint64_t beta, uint64_t* out) {
    /* src/math/fri.rs:27-48 */
    size_t half = m / 2;
    uint64_t half_inv = to_bb_inverse(to_bb_new(2));
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && half >= 4096)
    for (size_t i = 0; i < half; i++) {
        uint64_t a = evals[i], b = evals[i + half], x = xs[i];
        uint64_t avg = to_bb_mul(to_bb_add(a, b), half_inv);
        uint64_t diff = to_bb_mul(to_bb_sub(a, b), half_inv);
        out[i] = to_bb_add(avg, to_bb_mul(to_bb_mul(diff, beta), to_bb_inverse(x)));
    }
}

void to_fri_fold_ext(const uint64_t* evals, size_t m, const uint64_t* xs, const uint64_t beta[4], uint64_t* out) {
    /* src/math/fri.rs:7-25 */
    size_t half = m / 2;
    uint64_t half_inv = to_bb_inverse(to_bb_new(2));
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && half >= 4096)
    for (size_t i = 0; i < half; i++) {
        const uint64_t* a = evals + 4 * i;
        const uint64_t* b = evals + 4 * (i + half);
        uint64_t x_inv = to_bb_inverse(xs[i]);
        uint64_t s[4], d[4], avg[4], diff[4], t[4], xe[4] = {x_inv, 0, 0, 0};
        to_ext_add(a, b, s);
        to_ext_sub(a, b, d);
        to_ext_mul_base(s, half_inv, avg);
        to_ext_mul_base(d, half_inv, diff);
        to_ext_mul(diff, beta, t); /* diff * beta ... */
        to_ext_mul(t, xe, t);      /* ... * Ext::from_base(x_inv) */
        to_ext_add(avg, t, out + 4 * i);
    }
}

/* ------------------------------------------------------------------- SHA-256 */

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t h[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

typedef struct {
    uint32_t h[8];
    uint8_t buf[64];
    size_t buflen;
    uint64_t total;
} sha_ctx;

static void sha_init(sha_ctx* c) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(c->h, iv, sizeof iv);
    c->buflen = 0;
    c->total = 0;
}

static void sha_update(sha_ctx* c, const uint8_t* d, size_t n) {
    c->total += n;
    while (n > 0) {
        size_t take = 64 - c->buflen;
        if (take > n) take = n;
        memcpy(c->buf + c->buflen, d, take);
        c->buflen += take;
        d += take;
        n -= take;
        if (c->buflen == 64) {
            sha256_block(c->h, c->buf);
            c->buflen = 0;
        }
    }
}

static void sha_final(sha_ctx* c, uint8_t out[32]) {
    uint64_t bits = c->total * 8;
    uint8_t pad = 0x80;
    sha_update(c, &pad, 1);
    uint8_t z = 0;
    while (c->buflen != 56) sha_update(c, &z, 1);
    uint8_t lenb[8];
    for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(c, lenb, 8);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(c->h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(c->h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c->h[i] >> 8);
        out[4 * i + 3] = (uint8_t)c->h[i];
    }
}

void to_sha256(const uint8_t* data, size_t len, uint8_t out[32]) { /* src/lib.rs:14-18 */
    sha_ctx c;
    sha_init(&c);
    sha_update(&c, data, len);
    sha_final(&c, out);
}

/* -------------------------------------------------------------------- Merkle */

void to_hash_leaf(const uint8_t* leaf, size_t len, uint8_t out[32]) { /* src/merkle.rs:105,109-114 */
    sha_ctx c;
    uint8_t tag = 0x00;
    sha_init(&c);
    sha_update(&c, &tag, 1);
    sha_update(&c, leaf, len);
    sha_final(&c, out);
}

void to_hash_node(const uint8_t l[32], const uint8_t r[32], uint8_t out[32]) { /* src/merkle.rs:106,117-123 */
    sha_ctx c;
    uint8_t tag = 0x01;
    sha_init(&c);
    sha_update(&c, &tag, 1);
    sha_update(&c, l, 32);
    sha_update(&c, r, 32);
    sha_final(&c, out);
}

size_t to_merkle_node_count(size_t nleaves) {
    size_t total = nleaves, cur = nleaves;
    while (cur > 1) {
        cur = (cur + 1) / 2;
        total += cur;
    }
    return total;
}

static void merkle_upper_levels(uint8_t* nodes, size_t nleaves, uint8_t root_out[32]) {
    /* src/merkle.rs:34-47: pair adjacent nodes; an odd level pairs its last node with itself */
    uint8_t* cur = nodes;
    size_t cur_n = nleaves;
    while (cur_n > 1) {
        uint8_t* next = cur + 32 * cur_n;
        size_t next_n = (cur_n + 1) / 2;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && cur_n >= 4096)
        for (size_t i = 0; i < cur_n; i += 2) {
            const uint8_t* l = cur + 32 * i;
            const uint8_t* r = (i + 1 < cur_n) ? cur + 32 * (i + 1) : l;
            to_hash_node(l, r, next + 32 * (i / 2));
        }
        cur = next;
        cur_n = next_n;
    }
    if (root_out) memcpy(root_out, cur, 32); /* :82-84 */
}

void to_merkle_build(const uint8_t* leaves, size_t nleaves, size_t leaf_len, uint8_t* nodes_out, uint8_t root_out[32]) {
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && nleaves >= 4096)
    for (size_t i = 0; i < nleaves; i++) to_hash_leaf(leaves + i * leaf_len, leaf_len, nodes_out + 32 * i); /* :29-31 */
    merkle_upper_levels(nodes_out, nleaves, root_out);
}

void to_commit_values(const uint64_t* values, size_t n, int limbs, const uint8_t* salts, uint8_t* nodes_out,
                      uint8_t root_out[32]) {
    /* src/fibonacci.rs:340-363: leaf = salt || value.to_bytes(), or just value.to_bytes() */
    size_t vbytes = 8 * (size_t)limbs;
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && n >= 4096)
    for (size_t i = 0; i < n; i++) {
        uint8_t leaf[16 + 32];
        size_t off = 0;
        if (salts) {
            memcpy(leaf, salts + 16 * i, 16);
            off = 16;
        }
        for (int k = 0; k < limbs; k++) {
            uint64_t v = values[(size_t)limbs * i + k];
            for (int b = 0; b < 8; b++) leaf[off + 8 * k + b] = (uint8_t)(v >> (8 * b)); /* src/babybear.rs:53-55 */
        }
        to_hash_leaf(leaf, off + vbytes, nodes_out + 32 * i);
    }
    merkle_upper_levels(nodes_out, n, root_out);
}

size_t to_merkle_open(const uint8_t* nodes, size_t nleaves, size_t index, uint8_t* path_out, uint8_t* pos_out) {
    /* src/merkle.rs:50-80 */
    const uint8_t* level = nodes;
    size_t level_n = nleaves, cur = index, depth = 0;
    while (level_n > 1) {
        size_t sib = (cur % 2 == 0) ? cur + 1 : cur - 1;
        if (sib >= level_n) {
            memcpy(path_out + 32 * depth, level + 32 * cur, 32);
            pos_out[depth] = 1;
        } else {
            memcpy(path_out + 32 * depth, level + 32 * sib, 32);
            pos_out[depth] = (cur % 2 == 1);
        }
        depth++;
        cur /= 2;
        level += 32 * level_n;
        level_n = (level_n + 1) / 2;
    }
    return depth;
}

int to_merkle_verify(const uint8_t* leaf, size_t leaf_len, const uint8_t* path, const uint8_t* pos, size_t depth,
                     const uint8_t root[32]) {
    /* src/merkle.rs:86-101 */
    uint8_t cur[32], nxt[32];
    to_hash_leaf(leaf, leaf_len, cur);
    for (size_t d = 0; d < depth; d++) {
        if (pos[d])
            to_hash_node(path + 32 * d, cur, nxt);
        else
            to_hash_node(cur, path + 32 * d, nxt);
        memcpy(cur, nxt, 32);
    }
    return memcmp(cur, root, 32) == 0;
}

/* ---------------------------------------------------------------- transcript */

void to_transcript_init(to_transcript* t) { /* src/transcript.rs:12-16 */
    static const char tag[] = "toyni-stark-v1";
    t->cap = 256;
    t->state = (uint8_t*)malloc(t->cap);
    t->len = sizeof(tag) - 1;
    memcpy(t->state, tag, t->len);
}

void to_transcript_free(to_transcript* t) {
    free(t->state);
    t->state = NULL;
    t->len = t->cap = 0;
}

void to_transcript_absorb(to_transcript* t, const uint8_t* data, size_t n) { /* :19-21: append */
    if (t->len + n > t->cap) {
        while (t->len + n > t->cap) t->cap *= 2;
        t->state = (uint8_t*)realloc(t->state, t->cap);
    }
    memcpy(t->state + t->len, data, n);
    t->len += n;
}

static uint64_t squeeze_u64(to_transcript* t) { /* state <- SHA256(state); first 8 bytes little-endian */
    uint8_t h[32];
    to_sha256(t->state, t->len, h);
    memcpy(t->state, h, 32);
    t->len = 32;
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)h[i] << (8 * i);
    return v;
}

uint64_t to_transcript_squeeze(to_transcript* t) { /* :34-39 + src/babybear.rs:65-71 */
    return squeeze_u64(t) % TO_P;
}

void to_transcript_squeeze_ext(to_transcript* t, uint64_t out[4]) { /* :43-50 */
    for (int k = 0; k < 4; k++) out[k] = to_transcript_squeeze(t);
}

void to_transcript_squeeze_indices(to_transcript* t, size_t count, size_t max, uint64_t* out) { /* :58-72 */
    size_t got = 0;
    while (got < count) {
        uint64_t idx = squeeze_u64(t) % (uint64_t)max;
        int seen = 0;
        for (size_t i = 0; i < got; i++)
            if (out[i] == idx) seen = 1;
        if (!seen) out[got++] = idx;
    }
}

/* ------------------------------------------------------------ FRI commit loop */

static size_t fri_commit_impl(const uint64_t* layer0, size_t n, uint64_t shift, size_t final_size, const uint8_t* salts,
                              to_transcript* t, uint64_t* layers_out, uint8_t* roots_out, uint64_t* betas_out, int limbs) {
    /* src/fibonacci.rs:200-247 (base field); the Ext form swaps in fri_fold_ext / squeeze_ext_challenge */
    uint64_t* xs = (uint64_t*)malloc(n * sizeof(uint64_t));
    to_domain_elements(xs, n, shift); /* :214 xs = shifted_elements */
    uint8_t* nodes = (uint8_t*)malloc(32 * to_merkle_node_count(n));
    uint64_t* cur = layers_out;
    memcpy(cur, layer0, n * limbs * sizeof(uint64_t)); /* :204 layer 0 = DEEP evaluations */
    size_t cur_n = n, folds = 0;
    const uint8_t* salt_ptr = salts;
    uint8_t root[32];
    to_commit_values(cur, cur_n, limbs, salt_ptr, nodes, root); /* :206-211 */
    salt_ptr += 16 * cur_n;
    to_transcript_absorb(t, root, 32);
    memcpy(roots_out, root, 32);
    while (cur_n > final_size) { /* :222 */
        uint64_t beta[4] = {0, 0, 0, 0};
        if (limbs == 1)
            beta[0] = to_transcript_squeeze(t); /* :223 */
        else
            to_transcript_squeeze_ext(t, beta);
        memcpy(betas_out + (size_t)limbs * folds, beta, limbs * sizeof(uint64_t));
        uint64_t* next = cur + cur_n * limbs;
        if (limbs == 1)
            to_fri_fold(cur, cur_n, xs, beta[0], next); /* :225 */
        else
            to_fri_fold_ext(cur, cur_n, xs, beta, next);
        cur_n /= 2;
        for (size_t i = 0; i < cur_n; i++) xs[i] = to_bb_mul(xs[i], xs[i]); /* :228-231 */
        folds++;
        if (cur_n == final_size) { /* :234-238: the last layer is unsalted */
            to_commit_values(next, cur_n, limbs, NULL, nodes, root);
        } else {
            to_commit_values(next, cur_n, limbs, salt_ptr, nodes, root);
            salt_ptr += 16 * cur_n;
        }
        to_transcript_absorb(t, root, 32); /* :239-242 */
        memcpy(roots_out + 32 * folds, root, 32);
        cur = next;
    }
    free(xs);
    free(nodes);
    return folds;
}

size_t to_fri_commit(const uint64_t* layer0, size_t n, uint64_t shift, size_t final_size, const uint8_t* salts,
                     to_transcript* t, uint64_t* layers_out, uint8_t* roots_out, uint64_t* betas_out) {
    return fri_commit_impl(layer0, n, shift, final_size, salts, t, layers_out, roots_out, betas_out, 1);
}

size_t to_fri_commit_ext(const uint64_t* layer0, size_t n, uint64_t shift, size_t final_size, const uint8_t* salts,
                         to_transcript* t, uint64_t* layers_out, uint8_t* roots_out, uint64_t* betas_out) {
    return fri_commit_impl(layer0, n, shift, final_size, salts, t, layers_out, roots_out, betas_out, 4);
}

/* ------------------------------------------------------------ synthetic data */

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void to_fill_random(uint64_t* out, size_t n, uint64_t seed) {
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) out[i] = splitmix64(&s) % TO_P;
}

void to_fill_random_bytes(uint8_t* out, size_t n, uint64_t seed) {
    uint64_t s = seed;
    for (size_t i = 0; i < n; i += 8) {
        uint64_t v = splitmix64(&s);
        for (size_t b = 0; b < 8 && i + b < n; b++) out[i + b] = (uint8_t)(v >> (8 * b));
    }
}
