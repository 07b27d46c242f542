// toyni.hpp — C++ host-side mirror of the reference's Rust interfaces for the GPU hot path, over the C ABI in
// include/toyni_ntt_cuda.h.  The reference is compiled code (Rust) whose toolchain is absent from this image, so
// this header is what a compiled host links instead of `mod cuda` of src/ntt.rs; names, argument meaning and error
// behaviour follow the reference:
//   src/ntt.rs:144-251        cuda_available, CudaBuffer, ntt_cuda, intt_cuda (Result<(), String> -> std::runtime_error;
//                             assert! -> std::logic_error)
//   src/math/domain.rs:9-175  BabyBearDomain::{new_, get_coset, with_gpu, fft, ifft, fft_ext, ifft_ext, elements}
//   src/math/fri.rs:7-48      fri_fold, fri_fold_ext
//   src/fibonacci.rs:325-363  SaltedTree, build_merkle_tree, build_unsalted_tree (salts explicit)
// There is no CPU fallback: use_gpu = false throws (that branch is the reference's own CPU code).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "toyni_ntt_cuda.h"

#if !defined(__CUDACC__) && !defined(__CUDA_RUNTIME_H__)
extern "C" int cudaGetDeviceCount(int*);  // libcudart, as src/ntt.rs:102 (a plain C++ host needs no CUDA headers)
#endif

namespace toyni {

constexpr uint64_t BABYBEAR_PRIME = 2013265921ull;  // src/babybear.rs:8

struct BabyBear {  // #[repr(C)] struct { value: u64 }, src/babybear.rs:10-14
    uint64_t value;
};
static_assert(sizeof(BabyBear) == 8, "layout of src/babybear.rs:10-14");
struct Ext {  // #[repr(C)] struct { c: [BabyBear; 4] }, src/ext.rs:22-26
    std::array<BabyBear, 4> c;
};
static_assert(sizeof(Ext) == 32, "layout of src/ext.rs:22-26");

inline uint64_t pow_mod(uint64_t b, uint64_t e) {
    unsigned __int128 r = 1, x = b % BABYBEAR_PRIME;
    while (e) {
        if (e & 1) r = r * x % BABYBEAR_PRIME;
        x = x * x % BABYBEAR_PRIME;
        e >>= 1;
    }
    return (uint64_t)r;
}
inline BabyBear get_root_of_unity(uint32_t log_n) {  // src/babybear.rs:118-126
    if (log_n > 27) throw std::logic_error("BabyBear only supports NTT up to 2^27");
    return BabyBear{pow_mod(440564289ull, 1ull << (27 - log_n))};
}

inline std::string cuda_error(int e) { return std::string(cuda_get_error_string(e)); }
inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + cuda_error(rc));
}

/// src/ntt.rs:144-150 (plus: the device must be able to run the sm_100a-only library)
inline bool cuda_available() {
    int count = 0;
    return cudaGetDeviceCount(&count) == 0 && count > 0 && bb_device_ok() == 1;
}

/// src/ntt.rs:153-212
class CudaBuffer {
  public:
    explicit CudaBuffer(size_t size) : size_(size) {
        int err = cuda_malloc(&ptr_, size);
        if (err != 0) throw std::runtime_error("CUDA malloc failed: " + cuda_error(err));
    }
    CudaBuffer(const CudaBuffer&) = delete;
    CudaBuffer& operator=(const CudaBuffer&) = delete;
    ~CudaBuffer() { cuda_free(ptr_); }
    void copy_from_host(const std::vector<uint64_t>& data) {
        if (data.size() != size_) throw std::logic_error("Size mismatch");
        int err = cuda_copy_to_device(ptr_, data.data(), size_);
        if (err != 0) throw std::runtime_error("CUDA copy to device failed: " + cuda_error(err));
    }
    void copy_to_host(std::vector<uint64_t>& data) const {
        if (data.size() != size_) throw std::logic_error("Size mismatch");
        int err = cuda_copy_from_device(data.data(), ptr_, size_);
        if (err != 0) throw std::runtime_error("CUDA copy from device failed: " + cuda_error(err));
    }
    uint64_t* as_ptr() const { return ptr_; }

  private:
    uint64_t* ptr_ = nullptr;
    size_t size_;
};

namespace detail {
inline void* get_or_create_ctx(size_t n) {  // src/ntt.rs:128-141: per-size cache for the life of the process
    static std::mutex mu;
    static std::map<size_t, void*> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    void* ctx = ntt_ctx_create((uint32_t)n);
    if (!ctx) throw std::runtime_error("ntt_ctx_create failed: " + std::string(bb_last_error_string()));
    cache[n] = ctx;
    return ctx;
}
inline void run(std::vector<BabyBear>& values, bool inverse) {
    if (!cuda_available()) throw std::runtime_error("CUDA not available");  // src/ntt.rs:225-227
    size_t n = values.size();
    if (n == 0 || (n & (n - 1))) throw std::logic_error("NTT size must be power of 2");  // :229
    if (n > (1ull << 27)) throw std::logic_error("BabyBear only supports NTT up to 2^27");  // :230
    void* ctx = get_or_create_ctx(n);
    bb_clear_error();
    uint64_t* raw = reinterpret_cast<uint64_t*>(values.data());  // same cast as src/ntt.rs:233
    if (inverse)
        intt_run_inplace(ctx, raw);
    else
        ntt_run_inplace(ctx, raw);
    if (int e = bb_last_error()) throw std::runtime_error("CUDA NTT failed: " + cuda_error(e));
}
}  // namespace detail

inline void ntt_cuda(std::vector<BabyBear>& values) { detail::run(values, false); }   // src/ntt.rs:224-236
inline void intt_cuda(std::vector<BabyBear>& values) { detail::run(values, true); }   // src/ntt.rs:239-251

/// src/math/domain.rs:9-175
class BabyBearDomain {
  public:
    size_t size;
    uint32_t log_size;
    BabyBear omega, shift;
    bool use_gpu;

    static BabyBearDomain new_(size_t size) {
        if (size == 0 || (size & (size - 1))) throw std::logic_error("Domain size must be power of 2");
        uint32_t l = 0;
        while ((1ull << l) < size) l++;
        return BabyBearDomain{size, l, get_root_of_unity(l), BabyBear{1}, true};
    }
    BabyBearDomain get_coset(BabyBear s) const { return BabyBearDomain{size, log_size, omega, s, use_gpu}; }
    BabyBearDomain with_gpu(bool g) const { return BabyBearDomain{size, log_size, omega, shift, g}; }
    BabyBear group_gen() const { return omega; }

    std::vector<BabyBear> fft(const std::vector<BabyBear>& coeffs) const {  // :107-123
        require_gpu();
        std::vector<BabyBear> out(size);
        check(toyni_domain_fft(u64(coeffs), coeffs.size(), size, shift.value, u64(out)), "CUDA NTT");
        return out;
    }
    std::vector<BabyBear> ifft(const std::vector<BabyBear>& evals) const {  // :85-102
        if (evals.size() != size) throw std::logic_error("Evaluation count must match domain size");  // :86
        require_gpu();
        std::vector<BabyBear> out(size);
        check(toyni_domain_ifft(u64(evals), size, shift.value, u64(out)), "CUDA INTT");
        return out;
    }
    std::vector<Ext> fft_ext(const std::vector<Ext>& coeffs) const {  // :135-137
        require_gpu();
        std::vector<Ext> out(size);
        check(toyni_domain_fft_ext(u64(coeffs), coeffs.size(), size, shift.value, u64(out)), "CUDA NTT");
        return out;
    }
    std::vector<Ext> ifft_ext(const std::vector<Ext>& evals) const {  // :130-132
        if (evals.size() != size) throw std::logic_error("Evaluation count must match domain size");
        require_gpu();
        std::vector<Ext> out(size);
        check(toyni_domain_ifft_ext(u64(evals), size, shift.value, u64(out)), "CUDA INTT");
        return out;
    }
    std::vector<BabyBear> elements() const {  // :61-69 == the coset evaluation of the polynomial X
        if (size == 1) return {shift};
        return fft({BabyBear{0}, BabyBear{1}});
    }

  private:
    template <typename T>
    static const uint64_t* u64(const std::vector<T>& v) { return reinterpret_cast<const uint64_t*>(v.data()); }
    template <typename T>
    static uint64_t* u64(std::vector<T>& v) { return reinterpret_cast<uint64_t*>(v.data()); }
    void require_gpu() const {
        if (!use_gpu) throw std::runtime_error("use_gpu = false is the reference's CPU path; this mirror has no CPU fallback");
        if (!cuda_available()) throw std::runtime_error("CUDA not available");
    }
};

/// src/math/fri.rs:27-48
inline std::vector<BabyBear> fri_fold(const std::vector<BabyBear>& evals, const std::vector<BabyBear>& xs, BabyBear beta) {
    if (evals.size() % 2) throw std::logic_error("Evaluations length must be even");
    if (xs.size() < evals.size() / 2) throw std::logic_error("xs shorter than evals/2");
    std::vector<BabyBear> out(evals.size() / 2);
    check(toyni_fri_fold(reinterpret_cast<const uint64_t*>(evals.data()), evals.size(), reinterpret_cast<const uint64_t*>(xs.data()),
                         beta.value, reinterpret_cast<uint64_t*>(out.data())), "fri_fold");
    return out;
}
/// src/math/fri.rs:7-25
inline std::vector<Ext> fri_fold_ext(const std::vector<Ext>& evals, const std::vector<BabyBear>& xs, Ext beta) {
    if (evals.size() % 2) throw std::logic_error("Evaluations length must be even");
    if (xs.size() < evals.size() / 2) throw std::logic_error("xs shorter than evals/2");
    std::vector<Ext> out(evals.size() / 2);
    check(toyni_fri_fold_ext(reinterpret_cast<const uint64_t*>(evals.data()), evals.size(), reinterpret_cast<const uint64_t*>(xs.data()),
                             reinterpret_cast<const uint64_t*>(beta.c.data()), reinterpret_cast<uint64_t*>(out.data())), "fri_fold_ext");
    return out;
}

/// src/fibonacci.rs:325-337: tree + salts; `nodes` holds every level (leaf level first), 32 bytes per digest
struct SaltedTree {
    size_t nleaves = 0;
    std::vector<uint8_t> nodes;
    std::array<uint8_t, 32> root_{};
    std::vector<std::array<uint8_t, 16>> salts;  // empty for an unsalted tree
    const std::array<uint8_t, 32>& root() const { return root_; }
};
inline SaltedTree commit(const std::vector<BabyBear>& evals, const std::vector<std::array<uint8_t, 16>>* salts) {
    SaltedTree t;
    t.nleaves = evals.size();
    t.nodes.resize(bb_merkle_node_count(t.nleaves) * 32);
    if (salts) {
        if (salts->size() != evals.size()) throw std::logic_error("one salt per leaf");
        t.salts = *salts;
    }
    check(toyni_merkle_commit(reinterpret_cast<const uint64_t*>(evals.data()), evals.size(), 1,
                              salts ? reinterpret_cast<const uint8_t*>(salts->data()) : nullptr, t.nodes.data(), t.root_.data()),
          "merkle commit");
    return t;
}
/// src/fibonacci.rs:340-353 with the salts passed in (the reference draws them from thread_rng)
inline SaltedTree build_merkle_tree(const std::vector<BabyBear>& evals, const std::vector<std::array<uint8_t, 16>>& salts) {
    return commit(evals, &salts);
}
/// src/fibonacci.rs:357-363
inline SaltedTree build_unsalted_tree(const std::vector<BabyBear>& evals) { return commit(evals, nullptr); }


// ------------------------------------------------------------------------------------------------------------------
// Proof object and its canonical byte form.  The structs mirror src/fibonacci.rs:44-86; the reference has no
// serialization (`#[derive(Debug)]` only), so the format is defined by this library (toyni_b200/proof.py is the same
// format in Python): u64 little-endian integers / field values (src/babybear.rs:53-55), raw 32-byte digests, u64 length
// prefixes, fields in declaration order; a salt is prefixed by one length byte, a path entry is digest + 1 byte
// (1 = sibling on the right).
struct MerkleOpening {  // src/fibonacci.rs:44-51
    uint64_t index = 0;
    BabyBear value{0};
    std::vector<uint8_t> salt;
    std::vector<std::array<uint8_t, 32>> path;
    std::vector<bool> position;
};
struct QueryProof {  // src/fibonacci.rs:53-60
    uint64_t index = 0;
    MerkleOpening deep_opening, deep_opening_pair, trace_opening, trace_opening_g, trace_opening_gg, quotient_opening;
    std::vector<std::pair<MerkleOpening, MerkleOpening>> fri_openings;
};
struct StarkProof {  // src/fibonacci.rs:62-86
    uint64_t trace_len = 0, lde_size = 0;
    std::array<uint8_t, 32> trace_commitment{}, quotient_commitment{};
    BabyBear t_z{0}, t_gz{0}, t_ggz{0}, q_z{0};
    std::vector<std::array<uint8_t, 32>> fri_commitments;
    std::vector<BabyBear> fri_final_layer;
    std::vector<QueryProof> query_proofs;
};

namespace detail {
inline void put_u64(std::vector<uint8_t>& out, uint64_t v) {
    for (int b = 0; b < 8; b++) out.push_back((uint8_t)(v >> (8 * b)));
}
inline void put_opening(std::vector<uint8_t>& out, const MerkleOpening& o) {
    put_u64(out, o.index);
    put_u64(out, o.value.value);
    if (o.salt.size() > 255 || o.path.size() != o.position.size()) throw std::logic_error("malformed opening");
    out.push_back((uint8_t)o.salt.size());
    out.insert(out.end(), o.salt.begin(), o.salt.end());
    put_u64(out, o.path.size());
    for (size_t i = 0; i < o.path.size(); i++) {
        out.insert(out.end(), o.path[i].begin(), o.path[i].end());
        out.push_back(o.position[i] ? 1 : 0);
    }
}
struct Reader {
    const uint8_t* p;
    size_t n, o = 0;
    const uint8_t* take(size_t k) {
        if (o + k > n) throw std::runtime_error("truncated proof");
        const uint8_t* r = p + o;
        o += k;
        return r;
    }
    uint64_t u64() {
        const uint8_t* b = take(8);
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)b[i] << (8 * i);
        return v;
    }
    std::array<uint8_t, 32> digest() {
        std::array<uint8_t, 32> d;
        const uint8_t* b = take(32);
        for (int i = 0; i < 32; i++) d[i] = b[i];
        return d;
    }
    MerkleOpening opening() {
        MerkleOpening m;
        m.index = u64();
        m.value = BabyBear{u64()};
        size_t sl = *take(1);
        const uint8_t* sp = take(sl);
        m.salt.assign(sp, sp + sl);
        uint64_t len = u64();
        if (len > 64) throw std::runtime_error("implausible path length");
        for (uint64_t i = 0; i < len; i++) {
            m.path.push_back(digest());
            m.position.push_back(*take(1) != 0);
        }
        return m;
    }
};
}  // namespace detail

inline std::vector<uint8_t> serialize_proof(const StarkProof& p) {
    std::vector<uint8_t> out;
    detail::put_u64(out, p.trace_len);
    detail::put_u64(out, p.lde_size);
    out.insert(out.end(), p.trace_commitment.begin(), p.trace_commitment.end());
    out.insert(out.end(), p.quotient_commitment.begin(), p.quotient_commitment.end());
    for (BabyBear v : {p.t_z, p.t_gz, p.t_ggz, p.q_z}) detail::put_u64(out, v.value);
    detail::put_u64(out, p.fri_commitments.size());
    for (const auto& r : p.fri_commitments) out.insert(out.end(), r.begin(), r.end());
    detail::put_u64(out, p.fri_final_layer.size());
    for (BabyBear v : p.fri_final_layer) detail::put_u64(out, v.value);
    detail::put_u64(out, p.query_proofs.size());
    for (const QueryProof& q : p.query_proofs) {
        detail::put_u64(out, q.index);
        for (const MerkleOpening* o : {&q.deep_opening, &q.deep_opening_pair, &q.trace_opening, &q.trace_opening_g, &q.trace_opening_gg,
                                       &q.quotient_opening})
            detail::put_opening(out, *o);
        detail::put_u64(out, q.fri_openings.size());
        for (const auto& pr : q.fri_openings) {
            detail::put_opening(out, pr.first);
            detail::put_opening(out, pr.second);
        }
    }
    return out;
}

inline StarkProof deserialize_proof(const uint8_t* data, size_t len) {
    detail::Reader r{data, len};
    StarkProof p;
    p.trace_len = r.u64();
    p.lde_size = r.u64();
    p.trace_commitment = r.digest();
    p.quotient_commitment = r.digest();
    p.t_z = BabyBear{r.u64()};
    p.t_gz = BabyBear{r.u64()};
    p.t_ggz = BabyBear{r.u64()};
    p.q_z = BabyBear{r.u64()};
    for (uint64_t i = 0, n = r.u64(); i < n; i++) p.fri_commitments.push_back(r.digest());
    for (uint64_t i = 0, n = r.u64(); i < n; i++) p.fri_final_layer.push_back(BabyBear{r.u64()});
    for (uint64_t i = 0, n = r.u64(); i < n; i++) {
        QueryProof q;
        q.index = r.u64();
        for (MerkleOpening* o : {&q.deep_opening, &q.deep_opening_pair, &q.trace_opening, &q.trace_opening_g, &q.trace_opening_gg, &q.quotient_opening})
            *o = r.opening();
        for (uint64_t j = 0, m = r.u64(); j < m; j++) {
            MerkleOpening a = r.opening();
            MerkleOpening b = r.opening();
            q.fri_openings.emplace_back(std::move(a), std::move(b));
        }
        p.query_proofs.push_back(std::move(q));
    }
    if (r.o != len) throw std::runtime_error("trailing bytes after the proof");
    return p;
}

}  // namespace toyni
