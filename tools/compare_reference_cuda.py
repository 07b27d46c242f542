"""Time the reference's own CUDA NTT (cuda/ntt_kernel.cu rebuilt for sm_100a as oracle/_ref/libntt_cuda_ref.so)
against this library on the same B200, through the identical host-pointer ABI (ntt_run_inplace on pinned u64)."""
import ctypes, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from toyni_b200 import ntt as ours
from toyni_b200.lib import lib

ref = ctypes.CDLL(os.path.join("oracle", "_ref", "libntt_cuda_ref.so"), mode=ctypes.RTLD_LOCAL)
ref.ntt_ctx_create.restype = ctypes.c_void_p
ref.ntt_ctx_create.argtypes = [ctypes.c_uint32]
ref.ntt_run_inplace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
res = []
for log_n in (16, 20, 22, 24):
    n = 1 << log_n
    host = torch.empty(n, dtype=torch.int64).pin_memory()
    hv = host.numpy().view(np.uint64)
    hv[:] = (np.arange(n, dtype=np.uint64) * 7 + 3) % 2013265921
    t0 = time.perf_counter(); ctx = ref.ntt_ctx_create(n); torch.cuda.synchronize(); t_ctx_ref = time.perf_counter() - t0
    for _ in range(2): ref.ntt_run_inplace(ctx, hv.ctypes.data)
    t0 = time.perf_counter()
    for _ in range(5): ref.ntt_run_inplace(ctx, hv.ctypes.data)
    t_ref = (time.perf_counter() - t0) / 5
    for _ in range(2): ours.ntt_cuda(hv)
    t0 = time.perf_counter()
    for _ in range(5): ours.ntt_cuda(hv)
    t_ours = (time.perf_counter() - t0) / 5
    res.append({"log_n": log_n, "reference_cuda_ms": t_ref * 1e3, "ours_ms": t_ours * 1e3, "speedup_e2e": t_ref / t_ours,
                "reference_ctx_create_s": t_ctx_ref})
print(json.dumps({"host_pointer_ntt_run_inplace": res}, indent=1))
