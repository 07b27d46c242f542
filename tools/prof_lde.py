"""ncu / timing target: the blowup-32 coset LDE 2^20 -> 2^25 (expansion pass + TMA-staged pass 2), a few calls.
usage: python tools/prof_lde.py [calls = 6] [n_coeffs = 2^20]"""
import sys

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_coeffs = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
g = torch.Generator(device="cuda")
g.manual_seed(5)
c = torch.randint(0, P, (n_coeffs,), dtype=torch.int32, device="cuda", generator=g)
outs = [torch.empty(1 << 25, dtype=torch.int32, device="cuda") for _ in range(3)]
for i in range(3):
    D.coset_fft(c, 1 << 25, 7, out=outs[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(calls):
    D.coset_fft(c, 1 << 25, 7, out=outs[i % 3])
e1.record()
torch.cuda.synchronize()
print(f"LDE 2^20 -> 2^25 ({n_coeffs} coefficients): {e0.elapsed_time(e1) / calls * 1e3:.1f} us per call")
