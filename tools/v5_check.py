"""Warp-private NTT pass kernel (ntt_pass_v5.cuh) against the tile kernel (ntt_pass_v4.cuh): bit-exact outputs on the
same inputs, then timings of both (CUDA events, inputs rotating over buffers larger than L2)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

L = lib()
MODE = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 1
out = {"parity": [], "timing": []}


def run(kind, t, inverse):
    if kind == "vec":
        D.ntt_(t, inverse)
    elif kind == "batch":
        D.ntt_batch_(t, inverse)
    else:
        D.ntt_ext_(t, inverse)


def parity(kind, shape, inverse):
    g = torch.Generator(device="cuda").manual_seed(0x70796E69)
    x = torch.randint(0, P, shape, dtype=torch.int32, device="cuda", generator=g)
    a, b = x.clone(), x.clone()
    L.bb_ntt_set_kernel(0, 0)
    run(kind, a, inverse)
    L.bb_ntt_set_kernel(MODE, 0)
    run(kind, b, inverse)
    torch.cuda.synchronize()
    ok = bool(torch.equal(a, b))
    canon = bool(((b >= 0) & (b < P)).all())
    out["parity"].append({"kind": kind, "shape": list(shape), "inverse": inverse, "equal": ok, "canonical": canon})
    print(out["parity"][-1], flush=True)
    return ok and canon


def timing(kind, shape, reps=200):
    nbuf = max(2, min(8, (1 << 28) // (4 * int(torch.tensor(shape).prod())) + 1))
    bufs = [torch.randint(0, P, shape, dtype=torch.int32, device="cuda") for _ in range(nbuf)]
    res = {"kind": kind, "shape": list(shape)}
    for v5 in (0, MODE):
        L.bb_ntt_set_kernel(v5, 0)
        for i in range(5):
            run(kind, bufs[i % nbuf], False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(kind, bufs[i % nbuf], False)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / reps
        n = 1
        for s in shape:
            n *= s
        res["v5" if v5 else "v4"] = {"us": round(us, 2), "gelem_s": round(n / us / 1e3, 1)}
    out["timing"].append(res)
    print(res, flush=True)


ok = True
ok &= parity("batch", (256, 1 << 16), False)
ok &= parity("batch", (256, 1 << 16), True)
ok &= parity("vec", (1 << 24,), False)
ok &= parity("vec", (1 << 24,), True)
ok &= parity("vec", (1 << 22,), False)
ok &= parity("vec", (1 << 23,), True)
ok &= parity("ext", (1 << 24, 4), False)
ok &= parity("ext", (1 << 24, 4), True)
ok &= parity("vec", (1 << 25,), False)
ok &= parity("batch", (3, 1 << 24), True)
out["ok"] = bool(ok)
if ok and "--time" in sys.argv:
    timing("vec", (1 << 24,))
    timing("batch", (256, 1 << 16))
    timing("batch", (4, 1 << 24), reps=50)
    timing("ext", (1 << 24, 4), reps=50)
    timing("vec", (1 << 22,))
    timing("vec", (1 << 25,), reps=100)
json.dump(out, open("gpurun_out/v5_check.json", "w"), indent=1)
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
