"""The C++ host mirror (toyni_b200/host/toyni.hpp) over the C ABI: compiles everywhere, runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.bin")


def _build():
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "toyni_b200", "host"),
           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
           "-L", os.path.join(ROOT, "toyni_b200"), "-lntt_cuda", "-L", os.path.join(ROOT, "oracle"), "-loracle",
           "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "toyni_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-o", BIN]
    subprocess.check_call(cmd)


def test_cpp_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_host_mirror_runs_the_reference_gpu_tests():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all C++ host-mirror tests passed" in out.stdout


def test_cpp_proof_format_round_trips_an_oracle_proof(tmp_path):
    """toyni.hpp's StarkProof reads the canonical bytes of a proof made by the oracle prover and writes the same bytes back
    (the C++ counterpart of toyni_b200/proof.py; the reference has no serialization, src/fibonacci.rs:62-86)."""
    from oracle import fibonacci as F
    exe = os.path.join(ROOT, "tests", "cpp", "test_proof_format.bin")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "toyni_b200", "host"),
                           os.path.join(ROOT, "tests", "cpp", "test_proof_format.cpp"), "-o", exe])
    p = F.generate_proof(F.fibonacci_trace(64), *F.proof_randomness(64), interpolate="intt")
    path = tmp_path / "proof.bin"
    path.write_bytes(F.serialize_proof(p))
    out = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "trace_len=64 lde_size=2048 fri_roots=9 final=8 queries=44" in out.stdout and "round trip ok" in out.stdout


PROVER_BIN = os.path.join(ROOT, "tests", "cpp", "test_prover.bin")


def _build_prover():
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "toyni_b200", "host"), os.path.join(ROOT, "tests", "cpp", "test_prover.cpp"),
                           "-L", os.path.join(ROOT, "toyni_b200"), "-lntt_cuda", "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath," + os.path.join(ROOT, "toyni_b200"), "-Wl,-rpath,/usr/local/cuda/lib64", "-o", PROVER_BIN])


def test_cpp_prover_compiles_and_links():
    _build_prover()
    assert os.path.exists(PROVER_BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("trace_len", [64, 1024])
def test_cpp_prover_proof_bytes_equal_the_oracle_provers(tmp_path, trace_len):
    """toyni::StarkProver::generate_proof (toyni_prover.hpp: StarkProver::generate_proof, src/fibonacci.rs:99-310, as
    compiled host code over the C ABI) on the same trace, mask and salts as the CPU oracle's prover: the canonical proof
    bytes are identical, and the restated verifier (src/verifier.rs) accepts the deserialized proof."""
    import numpy as np
    from oracle import fibonacci as F
    _build_prover()
    trace = F.fibonacci_trace(trace_len)
    mask, st, sq, sf = F.proof_randomness(trace_len)
    blob = b"".join([np.uint64(trace_len).tobytes(), np.asarray(trace, np.uint64).tobytes(), np.asarray(mask, np.uint64).tobytes(),
                     np.ascontiguousarray(st).tobytes(), np.ascontiguousarray(sq).tobytes(),
                     np.uint64(np.asarray(sf).size).tobytes(), np.ascontiguousarray(sf).tobytes()])
    inp, outp = tmp_path / "in.bin", tmp_path / "proof.bin"
    inp.write_bytes(blob)
    out = subprocess.run([PROVER_BIN, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C++ prover ok" in out.stdout
    ref = F.generate_proof(trace, mask, st, sq, sf, interpolate="intt")
    got = outp.read_bytes()
    assert got == F.serialize_proof(ref)
    from toyni_b200.proof import deserialize_proof
    assert F.verify(deserialize_proof(got))
