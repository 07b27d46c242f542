// The reference's own GPU tests (src/ntt.rs:253-311) and domain / fold / Merkle checks, written against the C++ host
// mirror (toyni_b200/host/toyni.hpp) with the CPU oracle (oracle/toyni_oracle.h) as the reference side.
#include <cstdio>
#include <cstring>
#include <vector>

#include "toyni.hpp"
#include "toyni_oracle.h"

using namespace toyni;
static int failures = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) {                                     \
            failures++;                                    \
            printf("FAIL %s:%d: ", __FILE__, __LINE__);    \
            printf(__VA_ARGS__);                           \
            printf("\n");                                  \
        }                                                  \
    } while (0)

static std::vector<BabyBear> seq(size_t n) {  // (i * 7 + 3), src/ntt.rs:270-272
    std::vector<BabyBear> v(n);
    for (size_t i = 0; i < n; i++) v[i].value = (i * 7 + 3) % BABYBEAR_PRIME;
    return v;
}

static void test_cuda_available() { printf("CUDA available: %d\n", (int)cuda_available()); }  // src/ntt.rs:258-261

static void test_cuda_ntt_vs_cpu() {  // src/ntt.rs:263-287
    for (size_t n : {256ul, 1ul << 12, 1ul << 17}) {
        auto cpu = seq(n), gpu = seq(n);
        to_ntt(reinterpret_cast<uint64_t*>(cpu.data()), n, get_root_of_unity(__builtin_ctzl(n)).value);
        ntt_cuda(gpu);
        for (size_t i = 0; i < n; i++) EXPECT(cpu[i].value == gpu[i].value, "Mismatch at index %zu: CPU=%lu, GPU=%lu", i, cpu[i].value, gpu[i].value);
    }
}

static void test_cuda_intt_roundtrip() {  // src/ntt.rs:289-310
    auto original = seq(256), values = seq(256);
    ntt_cuda(values);
    intt_cuda(values);
    for (size_t i = 0; i < 256; i++) EXPECT(original[i].value == values[i].value, "Roundtrip failed at index %zu", i);
}

static void test_bad_sizes() {
    std::vector<BabyBear> v(12);
    bool threw = false;
    try { ntt_cuda(v); } catch (const std::logic_error&) { threw = true; }
    EXPECT(threw, "non power of two must be rejected like the assert at src/ntt.rs:229");
}

static void test_cuda_buffer() {
    CudaBuffer buf(1000);
    std::vector<uint64_t> src(1000), dst(1000);
    to_fill_random(src.data(), 1000, 5);
    buf.copy_from_host(src);
    buf.copy_to_host(dst);
    EXPECT(src == dst, "CudaBuffer round trip");
}

static void test_domain() {  // src/math/domain.rs:194-278 on the GPU branch
    auto dom = BabyBearDomain::new_(1 << 11).with_gpu(true).get_coset(BabyBear{7});
    std::vector<BabyBear> c(64);
    to_fill_random(reinterpret_cast<uint64_t*>(c.data()), 64, 1);
    auto ev = dom.fft(c);
    std::vector<uint64_t> ref(1 << 11);
    to_domain_fft(reinterpret_cast<uint64_t*>(c.data()), 64, 1 << 11, 7, ref.data());
    EXPECT(memcmp(ev.data(), ref.data(), ref.size() * 8) == 0, "coset fft");
    auto back = dom.ifft(ev);
    for (size_t i = 0; i < back.size(); i++) EXPECT(back[i].value == (i < 64 ? c[i].value : 0), "coset ifft at %zu", i);
    std::vector<uint64_t> el(1 << 11);
    to_domain_elements(el.data(), 1 << 11, 7);
    auto mine = dom.elements();
    EXPECT(memcmp(mine.data(), el.data(), el.size() * 8) == 0, "elements");
    std::vector<Ext> ce(32);
    to_fill_random(reinterpret_cast<uint64_t*>(ce.data()), 128, 2);
    auto ee = dom.fft_ext(ce);
    std::vector<uint64_t> refe(4 << 11);
    to_domain_fft_ext(reinterpret_cast<uint64_t*>(ce.data()), 32, 1 << 11, 7, refe.data());
    EXPECT(memcmp(ee.data(), refe.data(), refe.size() * 8) == 0, "fft_ext");
    bool threw = false;
    try { ev.pop_back(); dom.ifft(ev); } catch (const std::logic_error&) { threw = true; }
    EXPECT(threw, "ifft length assert, src/math/domain.rs:86");
}

static void test_fold_and_merkle() {
    const size_t m = 1 << 10;
    std::vector<BabyBear> ev(m), xs(m);
    to_fill_random(reinterpret_cast<uint64_t*>(ev.data()), m, 3);
    to_domain_elements(reinterpret_cast<uint64_t*>(xs.data()), m, 7);
    auto f = fri_fold(ev, xs, BabyBear{123456789});
    std::vector<uint64_t> ref(m / 2);
    to_fri_fold(reinterpret_cast<uint64_t*>(ev.data()), m, reinterpret_cast<uint64_t*>(xs.data()), 123456789, ref.data());
    EXPECT(memcmp(f.data(), ref.data(), ref.size() * 8) == 0, "fri_fold");
    std::vector<Ext> ee(m);
    to_fill_random(reinterpret_cast<uint64_t*>(ee.data()), 4 * m, 4);
    Ext beta{{BabyBear{5}, BabyBear{6}, BabyBear{7}, BabyBear{8}}};
    auto fe = fri_fold_ext(ee, xs, beta);
    std::vector<uint64_t> refe(2 * m);
    uint64_t b4[4] = {5, 6, 7, 8};
    to_fri_fold_ext(reinterpret_cast<uint64_t*>(ee.data()), m, reinterpret_cast<uint64_t*>(xs.data()), b4, refe.data());
    EXPECT(memcmp(fe.data(), refe.data(), refe.size() * 8) == 0, "fri_fold_ext");
    std::vector<std::array<uint8_t, 16>> salts(m);
    to_fill_random_bytes(reinterpret_cast<uint8_t*>(salts.data()), 16 * m, 9);
    auto tree = build_merkle_tree(ev, salts);
    std::vector<uint8_t> nodes(to_merkle_node_count(m) * 32);
    uint8_t root[32];
    to_commit_values(reinterpret_cast<uint64_t*>(ev.data()), m, 1, reinterpret_cast<uint8_t*>(salts.data()), nodes.data(), root);
    EXPECT(memcmp(tree.root().data(), root, 32) == 0 && tree.nodes == nodes, "salted tree");
    auto ut = build_unsalted_tree(ev);
    to_commit_values(reinterpret_cast<uint64_t*>(ev.data()), m, 1, nullptr, nodes.data(), root);
    EXPECT(memcmp(ut.root().data(), root, 32) == 0, "unsalted tree");
}

int main() {
    test_cuda_available();
    if (!cuda_available()) { printf("CUDA not available, skipping test\n"); return 77; }  // src/ntt.rs:265-268
    test_cuda_ntt_vs_cpu();
    test_cuda_intt_roundtrip();
    test_bad_sizes();
    test_cuda_buffer();
    test_domain();
    test_fold_and_merkle();
    printf(failures ? "%d FAILURES\n" : "all C++ host-mirror tests passed\n", failures);
    return failures ? 1 : 0;
}
