// toyni_prover.hpp — StarkProver::generate_proof (src/fibonacci.rs:99-310) as compiled host code over the C ABI:
// the loop a Rust maintainer would write around `mod cuda` once the LDE-sized vectors stay on the device.
//
//   device (include/toyni_ntt_cuda.h):  trace interpolation (INTT), blowup-32 coset LDE, both coset IFFTs, the
//       salted / unsalted Merkle trees, constraint / quotient / DEEP element-wise formulas, out-of-domain
//       evaluations, the FRI commit loop (transcript as its callback) and the openings of the whole query set;
//   host (this file):  the Fiat-Shamir transcript (src/transcript.rs), the trace_len + 140 coefficients of the masked
//       trace polynomial (src/fibonacci.rs:110-121) and the proof object (src/fibonacci.rs:44-86).
//
// The reference draws the mask and the salts from thread_rng(); here they are explicit inputs, so a proof is
// reproducible and comparable byte for byte (toyni::serialize_proof) with the CPU oracle's and with
// toyni_b200/prover.py, which is the same loop in Python.  No CPU fallback: without a device every call throws.
#pragma once
#include <chrono>
#include <cstring>

#include "toyni.hpp"

namespace toyni {

constexpr size_t NUM_QUERIES = 44, BLOWUP = 32;      // src/fibonacci.rs:11-13
constexpr uint64_t COSET_SHIFT = 7;                  // src/fibonacci.rs:16
constexpr size_t MASK_DEGREE = 3 * NUM_QUERIES + 8;  // src/fibonacci.rs:19

namespace detail {

/// SHA-256 (FIPS 180-4) for the transcript: a few hundred bytes per proof, host side as in the reference (sha2 crate).
class Sha256 {
  public:
    static std::array<uint8_t, 32> hash(const uint8_t* data, size_t len) {
        uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        size_t full = len / 64;
        for (size_t b = 0; b < full; b++) block(h, data + 64 * b);
        uint8_t tail[128] = {0};
        size_t rem = len - 64 * full;
        std::memcpy(tail, data + 64 * full, rem);
        tail[rem] = 0x80;
        size_t tl = rem + 9 <= 64 ? 64 : 128;
        uint64_t bits = (uint64_t)len * 8;
        for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
        for (size_t b = 0; b < tl / 64; b++) block(h, tail + 64 * b);
        std::array<uint8_t, 32> out;
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(h[i] >> (24 - 8 * j));
        return out;
    }

  private:
    static uint32_t rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
    static void block(uint32_t h[8], const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
            0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
            0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
            0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
            0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
            0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
};

/// Device allocation through the ABI (bb_pool_alloc / bb_pool_free: stream ordered, cached between proofs), freed on
/// scope exit.
template <typename T>
class DevBuf {
  public:
    DevBuf() = default;
    explicit DevBuf(size_t count) : n_(count) {
        void* p = nullptr;
        check(bb_pool_alloc(&p, (count ? count : 1) * sizeof(T)), "bb_pool_alloc");
        p_ = static_cast<T*>(p);
    }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            reset();
            p_ = o.p_;
            n_ = o.n_;
            o.p_ = nullptr;
        }
        return *this;
    }
    ~DevBuf() { reset(); }
    void reset() {
        if (p_) bb_pool_free(p_);
        p_ = nullptr;
    }
    T* get() const { return p_; }
    size_t size() const { return n_; }
    void upload(const T* h, size_t count) {
        check(bb_h2d(p_, h, count * sizeof(T)), "bb_h2d");
        check(bb_sync(), "bb_sync");
    }

  private:
    T* p_ = nullptr;
    size_t n_ = 0;
};

inline uint64_t mulm(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % BABYBEAR_PRIME); }
inline uint64_t addm(uint64_t a, uint64_t b) { return (a + b) % BABYBEAR_PRIME; }
inline uint64_t subm(uint64_t a, uint64_t b) { return (a + BABYBEAR_PRIME - b % BABYBEAR_PRIME) % BABYBEAR_PRIME; }
inline uint32_t log2_exact(size_t n, const char* what) {
    if (n == 0 || (n & (n - 1))) throw std::logic_error(std::string(what) + " must be a power of two");
    uint32_t l = 0;
    while ((size_t(1) << l) < n) l++;
    return l;
}

/// One committed layer on the device: values, tree nodes, salts (nullptr = unsalted), root.
struct DeviceTree {
    const uint32_t* vals = nullptr;
    const uint8_t* nodes = nullptr;
    const uint8_t* salts = nullptr;
    size_t n = 0;

    /// open_merkle for a whole index list: src/fibonacci.rs:366-374 over MerkleTree::get_proof, src/merkle.rs:50-80
    std::vector<MerkleOpening> open_many(const std::vector<uint64_t>& idx) const {
        size_t nq = idx.size(), depth = 0;
        size_t d = 0;
        while ((size_t(1) << d) < n) d++;
        std::vector<uint8_t> paths(nq * d * 32 + 1), pos(nq * d + 1);
        check(bb_merkle_open_batch_device(nodes, n, idx.data(), nq, paths.data(), pos.data(), &depth), "bb_merkle_open_batch_device");
        if (depth != d) throw std::runtime_error("unexpected tree depth");
        std::vector<uint32_t> vals_h(nq);
        check(bb_gather_device(vals, 4, idx.data(), nq, vals_h.data()), "bb_gather_device");
        std::vector<uint8_t> salts_h;
        if (salts) {
            salts_h.resize(nq * 16);
            check(bb_gather_device(salts, 16, idx.data(), nq, salts_h.data()), "bb_gather_device");
        }
        std::vector<MerkleOpening> out(nq);
        for (size_t q = 0; q < nq; q++) {
            MerkleOpening& o = out[q];
            o.index = idx[q];
            o.value = BabyBear{vals_h[q]};
            if (salts) o.salt.assign(salts_h.begin() + 16 * q, salts_h.begin() + 16 * (q + 1));
            o.path.resize(d);
            o.position.resize(d);
            for (size_t k = 0; k < d; k++) {
                std::memcpy(o.path[k].data(), &paths[(q * d + k) * 32], 32);
                o.position[k] = pos[q * d + k] != 0;
            }
        }
        return out;
    }
};

/// The openings of many trees in one call (bb_merkle_open_multi_device): add(tree, indices) per tree, run() returns the
/// openings tree by tree, in the order they were added.
class OpeningBatch {
  public:
    void add(const DeviceTree& t, const std::vector<uint64_t>& idx) {
        bb_open_request r;
        r.d_nodes = t.nodes;
        r.nleaves = t.n;
        r.d_vals = t.vals;
        r.d_salts = t.salts;
        r.first = indices_.size();
        r.count = idx.size();
        reqs_.push_back(r);
        indices_.insert(indices_.end(), idx.begin(), idx.end());
    }
    std::vector<std::vector<MerkleOpening>> run() const {
        const size_t nq = indices_.size();
        size_t path_bytes = 0;
        std::vector<size_t> depth(reqs_.size());
        for (size_t r = 0; r < reqs_.size(); r++) {
            size_t d = 0;
            while ((size_t(1) << d) < reqs_[r].nleaves) d++;
            depth[r] = d;
            path_bytes += reqs_[r].count * d * 32;
        }
        std::vector<uint8_t> paths(path_bytes + 1), pos(path_bytes / 32 + 1), salts(16 * nq + 1);
        std::vector<uint32_t> vals(nq + 1);
        check(bb_merkle_open_multi_device(reqs_.data(), reqs_.size(), indices_.data(), nq, 4, paths.data(), path_bytes, pos.data(),
                                          reinterpret_cast<uint8_t*>(vals.data()), salts.data()),
              "bb_merkle_open_multi_device");
        std::vector<std::vector<MerkleOpening>> out(reqs_.size());
        size_t po = 0;  // digests consumed so far
        for (size_t r = 0; r < reqs_.size(); r++) {
            out[r].resize(reqs_[r].count);
            for (size_t k = 0; k < reqs_[r].count; k++) {
                const size_t q = reqs_[r].first + k;
                MerkleOpening& o = out[r][k];
                o.index = indices_[q];
                o.value = BabyBear{vals[q]};
                if (reqs_[r].d_salts) o.salt.assign(salts.begin() + 16 * q, salts.begin() + 16 * (q + 1));
                o.path.resize(depth[r]);
                o.position.resize(depth[r]);
                for (size_t d = 0; d < depth[r]; d++, po++) {
                    std::memcpy(o.path[d].data(), &paths[32 * po], 32);
                    o.position[d] = pos[po] != 0;
                }
            }
        }
        return out;
    }

  private:
    std::vector<bb_open_request> reqs_;
    std::vector<uint64_t> indices_;
};

}  // namespace detail

/// src/transcript.rs: state = label, absorb appends, squeeze hashes the state and replaces it with the digest.
class FiatShamirTranscript {
  public:
    FiatShamirTranscript() {
        static const char label[] = "toyni-stark-v1";
        state_.assign(label, label + sizeof(label) - 1);
    }
    void absorb(const uint8_t* data, size_t len) { state_.insert(state_.end(), data, data + len); }
    void absorb(const std::array<uint8_t, 32>& d) { absorb(d.data(), 32); }
    void absorb_field(BabyBear v) {  // BabyBear::to_bytes, src/babybear.rs:53-55
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v.value >> (8 * i));
        absorb(b, 8);
    }
    BabyBear squeeze_challenge() { return BabyBear{squeeze_u64() % BABYBEAR_PRIME}; }
    Ext squeeze_ext_challenge() {  // src/transcript.rs:43-50
        Ext e;
        for (auto& c : e.c) c = squeeze_challenge();
        return e;
    }
    std::vector<uint64_t> squeeze_indices(size_t count, uint64_t max) {  // distinct, in order of first appearance
        std::vector<uint64_t> out;
        while (out.size() < count) {
            uint64_t idx = squeeze_u64() % max;
            bool seen = false;
            for (uint64_t v : out) seen |= v == idx;
            if (!seen) out.push_back(idx);
        }
        return out;
    }

  private:
    uint64_t squeeze_u64() {
        auto h = detail::Sha256::hash(state_.data(), state_.size());
        state_.assign(h.begin(), h.end());
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)h[i] << (8 * i);
        return v;
    }
    std::vector<uint8_t> state_;
};

/// src/fibonacci.rs:89-97
inline std::vector<BabyBear> fibonacci_trace(size_t n) {
    std::vector<BabyBear> t(n);
    for (size_t i = 0; i < n; i++) t[i] = BabyBear{i < 2 ? 1 : (t[i - 1].value + t[i - 2].value) % BABYBEAR_PRIME};
    return t;
}

/// Bytes of salt the commit loop consumes for an lde_size codeword folded down to final_size (16 per leaf of every
/// salted layer; the final layer is unsalted, src/fibonacci.rs:234-238).
inline size_t fri_salt_bytes(size_t lde_size, size_t final_size) {
    size_t total = 0;
    for (size_t m = lde_size; m > final_size; m /= 2) total += 16 * m;
    return total;
}

/// StarkProver (src/fibonacci.rs:99-310) with the randomness passed in.
///   mask        : MASK_DEGREE coefficients of R in T + Z_H * R (:117-120)
///   salts_trace : 16 bytes per LDE point for the trace tree (:129), salts_quot likewise for the quotient tree (:153)
///   salts_fri   : fri_salt_bytes(lde, final) bytes, layers back to back, layer 0 first (:206)
class StarkProver {
  public:
    explicit StarkProver(std::vector<BabyBear> trace_column) : trace_(std::move(trace_column)) {}

    /// When set, generate_proof* synchronises the device after every stage and records (stage, milliseconds) here.
    mutable std::vector<std::pair<const char*, double>>* stage_timings = nullptr;

    StarkProof generate_proof(const std::vector<BabyBear>& mask, const std::vector<uint8_t>& salts_trace,
                              const std::vector<uint8_t>& salts_quot, const std::vector<uint8_t>& salts_fri) const {
        using namespace detail;
        if (!cuda_available()) throw std::runtime_error("CUDA not available");
        const size_t lde = trace_.size() * BLOWUP;
        if (salts_trace.size() != 16 * lde || salts_quot.size() != 16 * lde) throw std::logic_error("one 16-byte salt per LDE point");
        DevBuf<uint8_t> d_t(salts_trace.size()), d_q(salts_quot.size()), d_f(salts_fri.size());
        d_t.upload(salts_trace.data(), salts_trace.size());
        d_q.upload(salts_quot.data(), salts_quot.size());
        d_f.upload(salts_fri.data(), salts_fri.size());
        return generate_proof_device_salts(mask, d_t.get(), d_q.get(), d_f.get(), salts_fri.size());
    }

    /// The same with the three salt arrays already on the device (16 * lde, 16 * lde and salts_fri_bytes bytes): a caller
    /// that draws them there, or keeps them across proofs, moves nothing LDE-sized over PCIe.
    StarkProof generate_proof_device_salts(const std::vector<BabyBear>& mask, const uint8_t* d_salts_trace, const uint8_t* d_salts_quot,
                                           const uint8_t* d_salts_fri, size_t salts_fri_bytes) const {
        using namespace detail;
        if (!cuda_available()) throw std::runtime_error("CUDA not available");
        const size_t trace_len = trace_.size();
        const uint32_t log_t = log2_exact(trace_len, "trace length");
        const size_t lde = trace_len * BLOWUP;
        const uint32_t log_lde = log_t + 5;
        if (log_lde > 27) throw std::logic_error("BabyBear only supports NTT up to 2^27");
        if (mask.size() != MASK_DEGREE) throw std::logic_error("mask: MASK_DEGREE coefficients required");
        size_t bound = 1;
        while (bound < trace_len + MASK_DEGREE) bound *= 2;  // next_power_of_two of the degree bound, :201-203
        const size_t final_size = lde / bound;
        if (final_size == 0) throw std::logic_error("trace too short for the blowup");
        if (salts_fri_bytes < fri_salt_bytes(lde, final_size)) throw std::logic_error("salts_fri too short");
        const uint64_t g = get_root_of_unity(log_t).value;
        const uint32_t shift = (uint32_t)COSET_SHIFT;
        auto t_last = std::chrono::steady_clock::now();
        auto mark = [&](const char* stage) {
            if (!stage_timings) return;
            check(bb_sync(), "bb_sync");
            auto now = std::chrono::steady_clock::now();
            stage_timings->emplace_back(stage, std::chrono::duration<double, std::milli>(now - t_last).count());
            t_last = now;
        };

        // 1. trace polynomial: INTT on the device, then T + Z_H * R (:110-121).  Z_H * R = X^n R - R touches the first and
        //    the last MASK_DEGREE coefficients only: those 2 x 140 values are patched from the host, the 2^k coefficients
        //    in between never leave the device.
        std::vector<uint32_t> h32(trace_len);
        for (size_t i = 0; i < trace_len; i++) h32[i] = (uint32_t)(trace_[i].value % BABYBEAR_PRIME);
        DevBuf<uint32_t> d_tpoly(trace_len + MASK_DEGREE);
        d_tpoly.upload(h32.data(), trace_len);
        check(bb_coset_ifft_device(d_tpoly.get(), log_t, 1, 1), "bb_coset_ifft_device");
        size_t n_tp = trace_len + MASK_DEGREE;
        if (trace_len >= MASK_DEGREE && mask[MASK_DEGREE - 1].value % BABYBEAR_PRIME != 0) {
            uint32_t head[MASK_DEGREE], tail[MASK_DEGREE];
            check(bb_d2h(head, d_tpoly.get(), sizeof head), "bb_d2h");
            check(bb_sync(), "bb_sync");
            for (size_t i = 0; i < MASK_DEGREE; i++) {
                const uint64_t m = mask[i].value % BABYBEAR_PRIME;
                head[i] = (uint32_t)subm(head[i], m);
                tail[i] = (uint32_t)m;
            }
            check(bb_h2d(d_tpoly.get(), head, sizeof head), "bb_h2d");
            check(bb_h2d(d_tpoly.get() + trace_len, tail, sizeof tail), "bb_h2d");
            check(bb_sync(), "bb_sync");  // head / tail live on this stack frame
        } else {  // short traces (the two ranges overlap) or a zero leading mask coefficient (the polynomial is trimmed)
            check(bb_d2h(h32.data(), d_tpoly.get(), trace_len * 4), "bb_d2h");
            check(bb_sync(), "bb_sync");
            std::vector<uint32_t> tp(trace_len + MASK_DEGREE, 0);
            for (size_t i = 0; i < trace_len; i++) tp[i] = h32[i];
            for (size_t i = 0; i < MASK_DEGREE; i++) {
                const uint64_t m = mask[i].value % BABYBEAR_PRIME;
                tp[trace_len + i] = (uint32_t)addm(tp[trace_len + i], m);
                tp[i] = (uint32_t)subm(tp[i], m);
            }
            while (n_tp > 0 && tp[n_tp - 1] == 0) n_tp--;  // Polynomial::new trims, src/math/polynomial.rs:11-16
            d_tpoly.upload(tp.data(), n_tp);
        }
        // LDE over the shifted domain + salted commit (:124-130)
        DevBuf<uint32_t> d_tlde(lde);
        check(bb_coset_fft_device(d_tpoly.get(), n_tp, log_lde, shift, 1, d_tlde.get()), "bb_coset_fft_device");
        StarkProof proof;
        proof.trace_len = trace_len;
        proof.lde_size = lde;
        const size_t nodes_lde = bb_merkle_node_count(lde);
        DevBuf<uint8_t> d_nodes_t(nodes_lde * 32), d_nodes_q(nodes_lde * 32);
        check(bb_merkle_commit_device(d_tlde.get(), 1, lde, d_salts_trace, d_nodes_t.get(), proof.trace_commitment.data()), "trace commit");

        mark("interpolate + LDE + trace commit");

        // 2. constraint and quotient (:133-153): Z_H over the coset takes BLOWUP values, 7^n (w_N^n)^i - 1
        const uint64_t b1 = pow_mod(g, trace_len - 1), b2 = pow_mod(g, trace_len - 2);
        DevBuf<uint32_t> d_q(lde);
        check(bb_fib_constraint_device(d_tlde.get(), log_lde, (uint32_t)BLOWUP, shift, (uint32_t)b1, (uint32_t)b2, d_q.get()), "bb_fib_constraint_device");
        const uint64_t g_ext = get_root_of_unity(log_lde).value;
        const uint64_t sn = pow_mod(COSET_SHIFT, trace_len), wn = pow_mod(g_ext, trace_len);
        uint32_t zh_inv[BLOWUP];
        for (size_t i = 0; i < BLOWUP; i++) zh_inv[i] = (uint32_t)pow_mod(subm(mulm(sn, pow_mod(wn, i)), 1), BABYBEAR_PRIME - 2);
        check(bb_scale_periodic_device(d_q.get(), lde, zh_inv, (uint32_t)BLOWUP), "bb_scale_periodic_device");
        DevBuf<uint32_t> d_qcoef(lde);
        copy_device(d_qcoef.get(), d_q.get(), lde);
        check(bb_coset_ifft_device(d_qcoef.get(), log_lde, shift, 1), "bb_coset_ifft_device");
        check(bb_merkle_commit_device(d_q.get(), 1, lde, d_salts_quot, d_nodes_q.get(), proof.quotient_commitment.data()), "quotient commit");

        mark("constraint + quotient + quotient commit");

        // 3. z outside both domains (:156-161, :378-399): z^N != 1 and (z / 7)^N != 1
        FiatShamirTranscript tr;
        tr.absorb(proof.trace_commitment);
        tr.absorb(proof.quotient_commitment);
        const uint64_t inv_shift = pow_mod(COSET_SHIFT, BABYBEAR_PRIME - 2);
        uint64_t z;
        do {
            z = tr.squeeze_challenge().value;
        } while (pow_mod(z, lde) == 1 || pow_mod(mulm(z, inv_shift), lde) == 1);

        // 4. out-of-domain evaluations and the constraint check at z (:164-183)
        uint32_t t_z, t_gz, t_ggz, q_z;
        check(bb_poly_eval_device(d_tpoly.get(), n_tp, (uint32_t)z, &t_z), "bb_poly_eval_device");
        check(bb_poly_eval_device(d_tpoly.get(), n_tp, (uint32_t)mulm(g, z), &t_gz), "bb_poly_eval_device");
        check(bb_poly_eval_device(d_tpoly.get(), n_tp, (uint32_t)mulm(mulm(g, g), z), &t_ggz), "bb_poly_eval_device");
        check(bb_poly_eval_device(d_qcoef.get(), lde, (uint32_t)z, &q_z), "bb_poly_eval_device");
        d_qcoef.reset();
        const uint64_t c_z = mulm(mulm(subm(t_ggz, addm(t_gz, t_z)), subm(z, b1)), subm(z, b2));
        if (c_z != mulm(q_z, subm(pow_mod(z, trace_len), 1))) throw std::logic_error("Constraint check at z failed");  // :173-177
        proof.t_z = BabyBear{t_z};
        proof.t_gz = BabyBear{t_gz};
        proof.t_ggz = BabyBear{t_ggz};
        proof.q_z = BabyBear{q_z};
        for (BabyBear v : {proof.t_z, proof.t_gz, proof.t_ggz, proof.q_z}) tr.absorb_field(v);

        mark("z + out-of-domain evaluations");

        // 5. DEEP polynomial (:186-198)
        DevBuf<uint32_t> d_deep(lde);
        check(bb_fib_deep_device(d_q.get(), d_tlde.get(), log_lde, (uint32_t)BLOWUP, shift, (uint32_t)z, q_z, t_z, t_gz, t_ggz, d_deep.get()),
              "bb_fib_deep_device");

        // 6. FRI commit loop on the device, the transcript as its callback (:200-247)
        std::vector<size_t> sizes;
        for (size_t m = lde;; m /= 2) {
            sizes.push_back(m);
            if (m <= final_size) break;
        }
        size_t folded = 0, node_total = 0;
        for (size_t k = 0; k < sizes.size(); k++) {
            if (k) folded += sizes[k];
            node_total += bb_merkle_node_count(sizes[k]);
        }
        DevBuf<uint32_t> d_layers(folded);
        DevBuf<uint8_t> d_nodes_f(node_total * 32);
        std::vector<uint8_t> roots(32 * sizes.size());
        size_t folds = 0;
        check(bb_fri_commit_device(d_deep.get(), lde, shift, final_size, 1, d_salts_fri, &challenge_cb, &tr, nullptr, d_layers.get(),
                                   d_nodes_f.get(), roots.data(), &folds),
              "bb_fri_commit_device");
        if (folds + 1 != sizes.size()) throw std::runtime_error("unexpected number of FRI folds");
        for (size_t k = 0; k < sizes.size(); k++) {
            std::array<uint8_t, 32> r;
            std::memcpy(r.data(), &roots[32 * k], 32);
            proof.fri_commitments.push_back(r);
        }
        tr.absorb(proof.fri_commitments.back());  // the callback absorbed every root but the last, :239-242
        std::vector<DeviceTree> trees(sizes.size());
        {
            size_t voff = 0, noff = 0, soff = 0;
            for (size_t k = 0; k < sizes.size(); k++) {
                trees[k].n = sizes[k];
                trees[k].vals = k == 0 ? d_deep.get() : d_layers.get() + voff;
                trees[k].nodes = d_nodes_f.get() + noff * 32;
                trees[k].salts = k + 1 < sizes.size() ? d_salts_fri + soff : nullptr;
                if (k) voff += sizes[k];
                noff += bb_merkle_node_count(sizes[k]);
                if (k + 1 < sizes.size()) soff += 16 * sizes[k];
            }
        }

        mark("DEEP + FRI commit loop");

        // 7. query phase (:250-295): every opening of the proof — 24 trees, ~2100 leaves at 2^20 rows — in ONE launch, one
        //    copy and one synchronisation (bb_merkle_open_multi_device)
        const std::vector<uint64_t> queries = tr.squeeze_indices(NUM_QUERIES, lde / 2);
        const DeviceTree trace_tree{d_tlde.get(), d_nodes_t.get(), d_salts_trace, lde};
        const DeviceTree quot_tree{d_q.get(), d_nodes_q.get(), d_salts_quot, lde};
        OpeningBatch batch;
        {
            std::vector<uint64_t> idx;
            for (uint64_t q : queries) {
                idx.push_back(q);
                idx.push_back(q + lde / 2);
            }
            batch.add(trees[0], idx);
            idx.clear();
            for (uint64_t q : queries)
                for (uint64_t k = 0; k < 3; k++) idx.push_back((q + k * BLOWUP) % lde);
            batch.add(trace_tree, idx);
            batch.add(quot_tree, queries);
            std::vector<uint64_t> cur = queries;
            for (size_t k = 1; k + 1 < trees.size(); k++) {  // :270-283
                const uint64_t half = trees[k].n / 2;
                idx.clear();
                for (uint64_t& i : cur) {
                    i %= half;
                    idx.push_back(i);
                    idx.push_back(i + half);
                }
                batch.add(trees[k], idx);
            }
        }
        std::vector<std::vector<MerkleOpening>> opened = batch.run();
        std::vector<MerkleOpening>&deep = opened[0], &trc = opened[1], &quo = opened[2];
        proof.query_proofs.resize(NUM_QUERIES);
        for (size_t n = 0; n < NUM_QUERIES; n++) {
            QueryProof& qp = proof.query_proofs[n];
            qp.index = queries[n];
            qp.deep_opening = std::move(deep[2 * n]);
            qp.deep_opening_pair = std::move(deep[2 * n + 1]);
            qp.trace_opening = std::move(trc[3 * n]);
            qp.trace_opening_g = std::move(trc[3 * n + 1]);
            qp.trace_opening_gg = std::move(trc[3 * n + 2]);
            qp.quotient_opening = std::move(quo[n]);
            for (size_t k = 3; k < opened.size(); k++) qp.fri_openings.emplace_back(std::move(opened[k][2 * n]), std::move(opened[k][2 * n + 1]));
        }
        std::vector<uint32_t> fin(sizes.back());
        check(bb_d2h(fin.data(), trees.back().vals, fin.size() * 4), "bb_d2h");
        check(bb_sync(), "bb_sync");
        for (uint32_t v : fin) proof.fri_final_layer.push_back(BabyBear{v});
        mark("queries: openings + proof object");
        return proof;
    }

  private:
    static void challenge_cb(void* user, const uint8_t root[32], uint32_t /*layer*/, uint32_t* beta_out) {
        auto* tr = static_cast<FiatShamirTranscript*>(user);
        tr->absorb(root, 32);
        beta_out[0] = (uint32_t)tr->squeeze_challenge().value;
    }
    static void copy_device(uint32_t* dst, const uint32_t* src, size_t n) {
        check(bb_d2d(dst, src, n * sizeof(uint32_t)), "bb_d2d");
    }
    std::vector<BabyBear> trace_;
};

}  // namespace toyni
