"""Batched transforms with and without the two-stream split (TOYNI_NTT_SPLIT), back to back (development aid)."""
import sys, torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
for shape in ((4, 1 << 24), (2, 1 << 24), (16, 1 << 20), (256, 1 << 16), (4096, 1 << 12), (4, 1 << 22), (64, 1 << 21), (8, 1 << 24)):
    bufs = [torch.randint(0, P, shape, dtype=torch.int32, device="cuda") for _ in range(2)]
    for i in range(3): D.ntt_batch_(bufs[i % 2])
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): D.ntt_batch_(bufs[i % 2])
    e1.record(); torch.cuda.synchronize()
    print(shape, round(e0.elapsed_time(e1) * 1000 / reps, 1), "us", flush=True)
    del bufs
