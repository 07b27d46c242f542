"""A few 2^24 transforms for ncu captures of the TMA-staged kernel (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
n = 1 << 24
bufs = [torch.randint(0, P, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    D.ntt_(bufs[i % 4])
torch.cuda.synchronize()
print("ok")
