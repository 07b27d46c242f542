"""Programmatic dependent launch on / off (TOYNI_NTT_PDL) over a few transform shapes (development aid)."""
import sys, torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
for shape in ((64, 1 << 22), (16, 1 << 24), (4, 1 << 22), (1024, 1 << 18), (1, 1 << 26), (256, 1 << 20), (64, 1 << 21), (64, 1 << 23)):
    b = torch.randint(0, P, shape, dtype=torch.int32, device="cuda")
    for _ in range(2): D.ntt_batch_(b, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.ntt_batch_(b, False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(shape, round(sorted(ts)[2], 1), "us", flush=True)
    del b
