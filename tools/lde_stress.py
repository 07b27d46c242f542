"""Stress of the blowup-32 LDE plan (expansion pass + TMA-staged pass 2, programmatic dependent launch between them)
under concurrency: two streams run LDEs of two different coefficient vectors back to back, with 2^24 transforms of the
TMA-staged kernel interleaved on the same streams; every result is compared with the general three-pass plan's.
Counts mismatches.  usage: python tools/lde_stress.py [iterations = 60]"""
import sys

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

L = lib()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
g = torch.Generator(device="cuda").manual_seed(3)
coeffs = [torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g) for n in ((1 << 20), (1 << 20) + 140)]
x24 = torch.randint(0, P, (1 << 24,), dtype=torch.int32, device="cuda", generator=g)
L.bb_ntt_set_kernel(0)
refs = [D.coset_fft(c, 1 << 25, 7) for c in coeffs]
ref24 = D.ntt_(x24.clone(), False)
L.bb_ntt_set_kernel(1)
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
outs = [[torch.empty(1 << 25, dtype=torch.int32, device="cuda") for _ in range(2)] for _ in range(2)]
bad = {"lde": 0, "ntt24": 0}
for it in range(iters):
    ys = []
    for s in range(2):
        with torch.cuda.stream(streams[s]):
            a = D.coset_fft(coeffs[s], 1 << 25, 7, out=outs[s][0])
            y = D.ntt_(x24.clone(), False)
            b = D.coset_fft(coeffs[s ^ (it & 1)], 1 << 25, 7, out=outs[s][1])
            ys.append((a, y, b, s ^ (it & 1)))
    torch.cuda.synchronize()
    for s, (a, y, b, which) in enumerate(ys):
        bad["lde"] += int(not torch.equal(a, refs[s])) + int(not torch.equal(b, refs[which]))
        bad["ntt24"] += int(not torch.equal(y, ref24))
print(f"{iters} iterations x 2 streams x (LDE, 2^24 NTT, LDE): mismatches {bad}")
sys.exit(1 if any(bad.values()) else 0)
