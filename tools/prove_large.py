"""Fibonacci proof at a large trace length, every LDE-sized array on the device; the restated verifier checks it.
usage: python tools/prove_large.py [log2 trace_len = 20] [reps = 3]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import fibonacci as F
from oracle import oracle as O
from toyni_b200 import prover

log_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
trace_len = 1 << log_t
lde = trace_len * 32
t0 = time.perf_counter()
tr = F.fibonacci_trace(trace_len)
t_trace = time.perf_counter() - t0
g = torch.Generator(device="cuda")
g.manual_seed(7)
salts = [torch.randint(0, 256, (m, 16), dtype=torch.uint8, device="cuda", generator=g) for m in (lde, lde, 2 * lde)]
mask = O.random_field(prover.MASK_DEGREE, 3)
times = []
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = prover.generate_proof(tr, mask, *salts)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
stages = {}
prover.generate_proof(tr, mask, *salts, timings=stages)
# the same proof from ONE call into the library (toyni_prove_fibonacci: the loop as compiled host code)
native_times = []
for r in range(reps + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    blob = prover.generate_proof_native(tr, mask, *salts, as_bytes=True)
    native_times.append(time.perf_counter() - t0)
from toyni_b200 import proof as product_proof
same = blob == product_proof.serialize_proof(p)
t0 = time.perf_counter()
ok = F.verify(p, algebraic=True)
t_verify = time.perf_counter() - t0
out = {"trace_len": trace_len, "lde_size": lde, "fri_layers": len(p["fri_commitments"]), "prove_s": [round(t, 4) for t in times],
       "prove_best_ms": round(min(times) * 1e3, 1), "trace_generation_s": round(t_trace, 3), "verify_s": round(t_verify, 3),
       "verifier_accepts": bool(ok), "one_call_native_best_ms": round(min(native_times[1:]) * 1e3, 1),
       "one_call_native_bytes_equal": bool(same), "stages_ms": {k: round(v * 1e3, 2) for k, v in stages.items()}}
print(json.dumps(out))
json.dump(out, open(f"gpurun_out/prove_2^{log_t}.json", "w"), indent=1)
sys.exit(0 if ok and same else 1)
