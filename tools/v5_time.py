"""Time device-resident forward NTTs with the library named by TOYNI_NTT_LIB (tuning builds)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib
L = lib()
v5 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L.bb_ntt_set_kernel(v5, 0)
for shape, reps in (((1 << 24,), 200), ((4, 1 << 24), 50)):
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
    print(shape, "v5" if v5 else "v4", round(e0.elapsed_time(e1) * 1000 / reps, 2), "us", flush=True)
