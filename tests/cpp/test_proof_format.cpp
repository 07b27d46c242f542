// The C++ proof object of toyni.hpp reads and re-emits the canonical byte form (no GPU involved):
//   test_proof_format <proof.bin>   exits 0 iff deserialize + serialize reproduces the file byte for byte, and prints
//   the shape it read (trace_len, lde_size, FRI roots, final layer, queries, openings of the first query).
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>

#include "toyni.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
        toyni::StarkProof p = toyni::deserialize_proof(blob.data(), blob.size());
        std::vector<uint8_t> again = toyni::serialize_proof(p);
        std::printf("trace_len=%llu lde_size=%llu fri_roots=%zu final=%zu queries=%zu fri_pairs=%zu path=%zu\n", (unsigned long long)p.trace_len,
                    (unsigned long long)p.lde_size, p.fri_commitments.size(), p.fri_final_layer.size(), p.query_proofs.size(),
                    p.query_proofs.empty() ? 0 : p.query_proofs[0].fri_openings.size(),
                    p.query_proofs.empty() ? 0 : p.query_proofs[0].trace_opening.path.size());
        if (again != blob) {
            std::printf("MISMATCH: %zu bytes in, %zu bytes out\n", blob.size(), again.size());
            return 1;
        }
        // a truncated file must be rejected, not mis-read
        bool threw = false;
        try {
            toyni::deserialize_proof(blob.data(), blob.size() - 1);
        } catch (const std::runtime_error&) {
            threw = true;
        }
        if (!threw) return 1;
    } catch (const std::exception& e) {
        std::printf("ERROR: %s\n", e.what());
        return 1;
    }
    std::printf("round trip ok\n");
    return 0;
}
