"""Time device-resident forward NTTs (2^24 single and batch of 4) with the pass kernel selected by argv[1]
(0 tile kernel, 1 warp-private 8-column strips, 2 warp-private 16-column strips)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib
L = lib()
v5 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.bb_ntt_set_kernel(v5, 0)
for shape, reps in (((1 << 24,), 400), ((4, 1 << 24), 50), ((1 << 22,), 400), ((1 << 27,), 20)):
    bufs = [torch.randint(0, P, shape, dtype=torch.int32, device="cuda") for _ in range(4)]
    f = D.ntt_ if len(shape) == 1 else D.ntt_batch_
    for i in range(5):
        f(bufs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        f(bufs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    print(shape, "kernel", v5, round(e0.elapsed_time(e1) * 1000 / reps, 2), "us", flush=True)
    del bufs
