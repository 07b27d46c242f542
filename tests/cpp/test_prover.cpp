// test_prover.cpp — toyni::StarkProver (toyni_b200/host/toyni_prover.hpp) end to end over the C ABI.
//   test_prover.bin <inputs.bin> <proof_out.bin>
//   test_prover.bin --bench <log2 trace_len> <reps> <proof_out.bin>   (salts drawn once and kept on the device)
// inputs.bin (little-endian, written by tests/test_cpp_host_mirror.py): u64 trace_len, trace_len u64 trace values,
// MASK_DEGREE u64 mask coefficients, 16*lde trace salts, 16*lde quotient salts, u64 nfri, nfri FRI salt bytes.
// Writes the canonical proof bytes; the Python test compares them with the CPU oracle's proof of the same inputs.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "toyni_prover.hpp"

static std::vector<uint8_t> read_all(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// Timing mode: a Fibonacci trace of 2^log_t rows, salts from SplitMix64 uploaded once, `reps` proofs timed one by one.
static int bench(int log_t, int reps, const char* out_path) {
    using namespace toyni;
    const size_t trace_len = size_t(1) << log_t, lde = trace_len * BLOWUP;
    size_t bound = 1;
    while (bound < trace_len + MASK_DEGREE) bound *= 2;
    const size_t nfri = fri_salt_bytes(lde, lde / bound);
    uint64_t st = 0x70796E69ull;
    auto next = [&st]() {
        uint64_t z = (st += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    std::vector<BabyBear> mask(MASK_DEGREE);
    for (auto& m : mask) m = BabyBear{next() % BABYBEAR_PRIME};
    const size_t total = 32 * lde + nfri;
    std::vector<uint64_t> host((total + 7) / 8);
    for (auto& w : host) w = next();
    detail::DevBuf<uint8_t> d_salts(host.size() * 8);
    d_salts.upload(reinterpret_cast<const uint8_t*>(host.data()), host.size() * 8);
    StarkProver prover(fibonacci_trace(trace_len));
    StarkProof proof;
    double best = 1e30, sum = 0;
    for (int r = 0; r <= reps; r++) {  // run 0 builds tables and scratch, untimed
        auto t0 = std::chrono::steady_clock::now();
        proof = prover.generate_proof_device_salts(mask, d_salts.get(), d_salts.get() + 16 * lde, d_salts.get() + 32 * lde, nfri);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r) {
            best = ms < best ? ms : best;
            sum += ms;
        }
    }
    std::vector<std::pair<const char*, double>> stages;  // one more proof with a device synchronisation after every stage
    prover.stage_timings = &stages;
    auto t0 = std::chrono::steady_clock::now();
    proof = prover.generate_proof_device_salts(mask, d_salts.get(), d_salts.get() + 16 * lde, d_salts.get() + 32 * lde, nfri);
    double staged_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    prover.stage_timings = nullptr;
    std::vector<uint8_t> bytes = serialize_proof(proof);
    std::ofstream(out_path, std::ios::binary).write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
    std::printf("{\"trace_len\": %zu, \"lde_size\": %zu, \"reps\": %d, \"prove_ms_best\": %.3f, \"prove_ms_mean\": %.3f, \"proof_bytes\": %zu, "
                "\"staged_total_ms\": %.3f, \"stages_ms\": {",
                trace_len, lde, reps, best, sum / reps, bytes.size(), staged_ms);
    for (size_t i = 0; i < stages.size(); i++) std::printf("%s\"%s\": %.3f", i ? ", " : "", stages[i].first, stages[i].second);
    std::printf("}}\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "--bench") {
        try {
            return bench(std::atoi(argv[2]), std::atoi(argv[3]), argv[4]);
        } catch (const std::exception& e) {
            std::cerr << "FAILED: " << e.what() << "\n";
            return 1;
        }
    }
    if (argc != 3) {
        std::cerr << "usage: test_prover.bin inputs.bin proof_out.bin\n";
        return 2;
    }
    try {
        // the transcript's hash against the FIPS 180-4 "abc" vector before anything depends on it
        const uint8_t abc[3] = {'a', 'b', 'c'};
        auto h = toyni::detail::Sha256::hash(abc, 3);
        if (h[0] != 0xba || h[1] != 0x78 || h[31] != 0xad) throw std::runtime_error("SHA-256 known answer failed");
        std::vector<uint8_t> in = read_all(argv[1]);
        toyni::detail::Reader r{in.data(), in.size()};
        const size_t trace_len = r.u64(), lde = trace_len * toyni::BLOWUP;
        std::vector<toyni::BabyBear> trace(trace_len), mask(toyni::MASK_DEGREE);
        for (auto& v : trace) v = toyni::BabyBear{r.u64()};
        for (auto& v : mask) v = toyni::BabyBear{r.u64()};
        const uint8_t* p = r.take(16 * lde);
        std::vector<uint8_t> st(p, p + 16 * lde);
        p = r.take(16 * lde);
        std::vector<uint8_t> sq(p, p + 16 * lde);
        const size_t nfri = r.u64();
        p = r.take(nfri);
        std::vector<uint8_t> sf(p, p + nfri);
        std::vector<toyni::BabyBear> fib = toyni::fibonacci_trace(trace_len);
        for (size_t i = 0; i < trace_len; i++)
            if (fib[i].value != trace[i].value) throw std::runtime_error("fibonacci_trace differs from the input trace");

        toyni::StarkProver prover(trace);
        toyni::StarkProof proof = prover.generate_proof(mask, st, sq, sf);  // first call builds tables and scratch
        auto t0 = std::chrono::steady_clock::now();
        toyni::StarkProof again = prover.generate_proof(mask, st, sq, sf);
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::vector<uint8_t> bytes = toyni::serialize_proof(proof);
        if (bytes != toyni::serialize_proof(again)) throw std::runtime_error("two runs of the prover differ");
        toyni::StarkProof back = toyni::deserialize_proof(bytes.data(), bytes.size());
        if (toyni::serialize_proof(back) != bytes) throw std::runtime_error("proof does not round-trip");
        std::ofstream(argv[2], std::ios::binary).write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
        std::printf("trace_len=%zu lde_size=%zu fri_roots=%zu final=%zu queries=%zu bytes=%zu prove_ms=%.2f\n", trace_len, lde,
                    proof.fri_commitments.size(), proof.fri_final_layer.size(), proof.query_proofs.size(), bytes.size(), ms);
        std::cout << "C++ prover ok\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "FAILED: " << e.what() << "\n";
        return 1;
    }
}
