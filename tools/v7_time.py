"""Time the TMA-staged 2^24 transform (one stream, rotating buffers); TOYNI_V7_FLAGS selects the diagnostic modes of
ntt_pass_v7.cuh (1 memory traffic only, 2 arithmetic only).  Development aid."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

L = lib()
n = 1 << 24
bufs = [torch.randint(0, P, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
for i in range(8):
    D.ntt_(bufs[i % 4])
torch.cuda.synchronize()
reps = 100
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    D.ntt_(bufs[i % 4])
e1.record()
torch.cuda.synchronize()
w = (C.c_uint32 * 16)()
L.bb_ntt_diag(w)
print("flags", os.environ.get("TOYNI_V7_FLAGS"), "promo", os.environ.get("TOYNI_V7_L2PROMO"), "pdl", os.environ.get("TOYNI_NTT_PDL"),
      round(e0.elapsed_time(e1) * 1000 / reps, 2), "us per transform; diag", list(w)[:9], flush=True)
