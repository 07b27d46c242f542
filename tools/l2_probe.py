"""How much of a 2^24 transform's traffic is served by L2: time with 1 vs 4 rotating buffers (argv[1] = buffers)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib
L = lib()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
bufs = [torch.randint(0, P, (1 << log_n,), dtype=torch.int32, device="cuda") for _ in range(nb)]
for i in range(5):
    D.ntt_(bufs[i % nb])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    D.ntt_(bufs[i % nb])
e1.record()
torch.cuda.synchronize()
print("buffers", nb, "log_n", log_n, round(e0.elapsed_time(e1) * 1000 / reps, 2), "us", flush=True)
