"""The C++ prover (toyni_b200/host/toyni_prover.hpp over the C ABI, no Python in the proof path) at a large trace length:
tests/cpp/test_prover.bin --bench times `reps` proofs with the salts resident on the device; this script only builds the
binary, runs it and puts the proof it wrote through the restated verifier (oracle, src/verifier.rs).
usage: python tools/prove_large_cpp.py [log2 trace_len = 20] [reps = 5]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fibonacci as F  # noqa: E402  (checker only)
from toyni_b200.proof import deserialize_proof  # noqa: E402

log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
exe = os.path.join(ROOT, "tests", "cpp", "test_prover.bin")
subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "toyni_b200", "host"),
                       os.path.join(ROOT, "tests", "cpp", "test_prover.cpp"), "-L", os.path.join(ROOT, "toyni_b200"), "-lntt_cuda",
                       "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.join(ROOT, "toyni_b200"),
                       "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
out_path = "/tmp/proof_cpp.bin"
run = subprocess.run([exe, "--bench", str(log_t), str(reps), out_path], capture_output=True, text=True, timeout=900)
if run.returncode:
    sys.exit(run.stdout + run.stderr)
rec = json.loads(run.stdout.strip().splitlines()[-1])
t0 = time.perf_counter()
rec["verifier_accepts"] = bool(F.verify(deserialize_proof(open(out_path, "rb").read())))
rec["verify_s"] = round(time.perf_counter() - t0, 2)
rec["host"] = "C++ (toyni::StarkProver::generate_proof_device_salts), wall clock per proof incl. every synchronisation"
print(json.dumps(rec))
