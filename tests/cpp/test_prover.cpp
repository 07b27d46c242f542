// test_prover.cpp — toyni::StarkProver (toyni_b200/host/toyni_prover.hpp) end to end over the C ABI.
//   test_prover.bin <inputs.bin> <proof_out.bin>
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

int main(int argc, char** argv) {
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
