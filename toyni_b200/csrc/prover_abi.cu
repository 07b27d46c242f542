// C ABI section 5: the whole prover loop behind one call.  The loop itself is toyni::StarkProver
// (toyni_b200/host/toyni_prover.hpp: StarkProver::generate_proof of src/fibonacci.rs:99-310 over the device entry points
// of this library, transcript on the host); this file only gives it a C face so that a host in any language — the Rust
// crate, Python through ctypes — gets a serialized proof without a per-stage trip through its own runtime.
#if __has_include("toyni_ntt_cuda.h")
#include "toyni_ntt_cuda.h"  // -I include (build.py) or the flat cuda/ directory of the toyni tree
#else
#include "../../include/toyni_ntt_cuda.h"
#endif

#if __has_include("toyni_prover.hpp")
#include "toyni_prover.hpp"  // flat layout: cuda/ of the toyni tree
#else
#include "../host/toyni_prover.hpp"
#endif

#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "abi_internal.cuh"

namespace {
thread_local std::string g_prover_error;

size_t final_size_for(size_t trace_len) {
    size_t bound = 1;
    while (bound < trace_len + toyni::MASK_DEGREE) bound *= 2;
    return trace_len * toyni::BLOWUP / bound;
}
}  // namespace

extern "C" {

size_t toyni_fri_salt_bytes(size_t trace_len) {
    const size_t fin = final_size_for(trace_len);
    return fin ? toyni::fri_salt_bytes(trace_len * toyni::BLOWUP, fin) : 0;
}

const char* toyni_prover_error(void) { return g_prover_error.c_str(); }

int toyni_prove_fibonacci(const uint64_t* trace, size_t trace_len, const uint64_t* mask, const uint8_t* salts_trace,
                          const uint8_t* salts_quot, const uint8_t* salts_fri, size_t salts_fri_bytes, int salts_on_device,
                          uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    g_prover_error.clear();
    if (!trace || !mask || !salts_trace || !salts_quot || !salts_fri || !proof_len) {
        g_prover_error = "null argument";
        return bb::abi::note_error((int)cudaErrorInvalidValue);
    }
    try {
        std::vector<toyni::BabyBear> t(trace_len), m(toyni::MASK_DEGREE);
        for (size_t i = 0; i < trace_len; i++) t[i] = toyni::BabyBear{trace[i]};
        for (size_t i = 0; i < toyni::MASK_DEGREE; i++) m[i] = toyni::BabyBear{mask[i]};
        toyni::StarkProver prover(std::move(t));
        toyni::StarkProof proof;
        if (salts_on_device) {
            proof = prover.generate_proof_device_salts(m, salts_trace, salts_quot, salts_fri, salts_fri_bytes);
        } else {
            const size_t lde = trace_len * toyni::BLOWUP;
            proof = prover.generate_proof(m, std::vector<uint8_t>(salts_trace, salts_trace + 16 * lde),
                                          std::vector<uint8_t>(salts_quot, salts_quot + 16 * lde),
                                          std::vector<uint8_t>(salts_fri, salts_fri + salts_fri_bytes));
        }
        const std::vector<uint8_t> bytes = toyni::serialize_proof(proof);
        *proof_len = bytes.size();
        if (!proof_out || proof_cap < bytes.size()) {
            g_prover_error = "proof buffer too small: " + std::to_string(bytes.size()) + " bytes needed";
            return bb::abi::note_error((int)cudaErrorInvalidValue);
        }
        std::memcpy(proof_out, bytes.data(), bytes.size());
        return 0;
    } catch (const std::logic_error& e) {  // the reference's assert! / panic paths
        g_prover_error = e.what();
        return bb::abi::note_error((int)cudaErrorInvalidValue);
    } catch (const std::exception& e) {
        g_prover_error = e.what();
        const int last = bb_last_error();
        return last ? last : bb::abi::note_error((int)cudaErrorUnknown);
    }
}

}  // extern "C"
