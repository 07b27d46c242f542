"""Two forward 2^24 NTTs through the reference's own CUDA path (cuda/ntt_kernel.cu rebuilt for sm_100a as
oracle/_ref/libntt_cuda_ref.so) and two through this library's device-resident entry point, for an ncu launch list:
the device-only comparator BASELINE.md section 3 asks for (the host-pointer form is PCIe-bound for both)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P

n = 1 << 24
ref = ctypes.CDLL(os.path.join("oracle", "_ref", "libntt_cuda_ref.so"), mode=ctypes.RTLD_LOCAL)
ref.ntt_ctx_create.restype = ctypes.c_void_p
ref.ntt_ctx_create.argtypes = [ctypes.c_uint32]
ref.ntt_run_inplace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
host = torch.empty(n, dtype=torch.int64).pin_memory()
hv = host.numpy().view(np.uint64)
hv[:] = (np.arange(n, dtype=np.uint64) * 7 + 3) % P
ctx = ref.ntt_ctx_create(n)
for _ in range(2):
    ref.ntt_run_inplace(ctx, hv.ctypes.data)
torch.cuda.synchronize()
x = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda")
for _ in range(2):
    D.ntt_(x)
torch.cuda.synchronize()
print("ok")
