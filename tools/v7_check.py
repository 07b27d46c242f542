"""TMA-staged two-pass 2^24 kernel (ntt_pass_v7.cuh) against the tile kernel (ntt_pass_v4.cuh) and the CPU oracle:
bit-exact outputs on the same inputs (forward, inverse, batches), then timings on one and two streams
(CUDA events, inputs rotating over four 64 MiB buffers, i.e. a working set larger than L2)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

L = lib()
LOG_N = 24
n = 1 << LOG_N
out = {"parity": [], "timing": {}}


def set_v7(on):
    L.bb_ntt_set_kernel(1 if on else 0)


def fail(msg):
    out["error"] = msg
    print(json.dumps(out), flush=True)
    sys.exit(1)


g = torch.Generator(device="cuda").manual_seed(0x70796E69)
x = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g)
for inverse in (False, True):
    a, b = x.clone(), x.clone()
    set_v7(False)
    D.ntt_(a, inverse)
    set_v7(True)
    D.ntt_(b, inverse)
    torch.cuda.synchronize()
    eq = bool(torch.equal(a, b))
    canon = bool(((b >= 0) & (b < P)).all())
    nbad = int((a != b).sum())
    first_bad = int(torch.nonzero(a != b)[0]) if nbad else -1
    out["parity"].append({"case": "single", "inverse": inverse, "equal_v4": eq, "canonical": canon, "mismatches": nbad, "first_bad": first_bad})
    print(out["parity"][-1], flush=True)
    if not eq:
        bad = torch.nonzero(a != b).flatten()[:16].tolist()
        print("bad idx", bad, "v4", a[bad].tolist(), "v7", b[bad].tolist(), flush=True)

# round trip and batch
set_v7(True)
for batch in (2, 3):
    xb = torch.randint(0, P, (batch, n), dtype=torch.int32, device="cuda", generator=g)
    yb = xb.clone()
    D.ntt_batch_(yb, False)
    set_v7(False)
    yr = xb.clone()
    D.ntt_batch_(yr, False)
    set_v7(True)
    eq = bool(torch.equal(yb, yr))
    D.ntt_batch_(yb, True)
    rt = bool(torch.equal(yb, xb))
    out["parity"].append({"case": f"batch{batch}", "equal_v4": eq, "round_trip": rt})
    print(out["parity"][-1], flush=True)
    del xb, yb, yr

# oracle (multi-threaded C restatement of src/ntt.rs:24-53), every element
if "--oracle" in sys.argv:
    from oracle import oracle as O
    xh = D.to_host(x)
    import os
    want = O.ntt(xh, threads=os.cpu_count() or 1)
    b = x.clone()
    D.ntt_(b, False)
    got = D.to_host(b)
    out["parity"].append({"case": "oracle_forward", "equal": bool(np.array_equal(got, want))})
    print(out["parity"][-1], flush=True)

ok = all(all(v for k, v in rec.items() if isinstance(v, bool)) for rec in out["parity"])
out["parity_ok"] = ok

bufs = [torch.randint(0, P, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
for name, v7 in (("v4", False), ("v7", True)):
    set_v7(v7)
    for nstreams in (1, 2):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        for i in range(8):
            with torch.cuda.stream(streams[i % nstreams]):
                D.ntt_(bufs[i % 4])
        torch.cuda.synchronize()
        reps = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(reps):
            with torch.cuda.stream(streams[i % nstreams]):
                D.ntt_(bufs[i % 4])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / reps
        out["timing"][f"{name}_{nstreams}stream_us"] = round(us, 2)
        print(name, nstreams, "streams:", round(us, 2), "us per transform", flush=True)
set_v7(True)
print(json.dumps(out), flush=True)
sys.exit(0 if ok else 1)
