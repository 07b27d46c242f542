"""Stress of the TMA-staged kernel under concurrency (two streams / split batches / PDL on and off): counts round-trip
and v4-equality failures.  Development aid."""
import os
import sys

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

L = lib()
n = 1 << 24


def set_v7(on):
    L.bb_ntt_set_kernel(1 if on else 0)


g = torch.Generator(device="cuda").manual_seed(1)
xb = torch.randint(0, P, (2, n), dtype=torch.int32, device="cuda", generator=g)
set_v7(False)
fwd_ref = xb.clone()
D.ntt_batch_(fwd_ref, False)
inv_ref = xb.clone()
D.ntt_batch_(inv_ref, True)
set_v7(True)
res = {"fwd_bad": 0, "inv_bad": 0, "rt_bad": 0, "two_stream_fwd_bad": 0, "two_stream_inv_bad": 0}
for it in range(6):
    y = xb.clone()
    D.ntt_batch_(y, False)
    if not torch.equal(y, fwd_ref):
        res["fwd_bad"] += 1
        bad = torch.nonzero(y != fwd_ref)
        print("fwd bad", it, bad.shape[0], bad[:4].tolist(), bad[-2:].tolist(), flush=True)
    D.ntt_batch_(y, True)
    if not torch.equal(y, xb):
        res["rt_bad"] += 1
    z = xb.clone()
    D.ntt_batch_(z, True)
    if not torch.equal(z, inv_ref):
        res["inv_bad"] += 1
        bad = torch.nonzero(z != inv_ref)
        print("inv bad", it, bad.shape[0], bad[:4].tolist(), bad[-2:].tolist(), flush=True)
# two python-level streams, single transforms
streams = [torch.cuda.Stream() for _ in range(2)]
for it in range(6):
    for inverse, ref, key in ((False, fwd_ref, "two_stream_fwd_bad"), (True, inv_ref, "two_stream_inv_bad")):
        y = xb.clone()
        torch.cuda.synchronize()
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                D.ntt_(y[k], inverse)
        torch.cuda.synchronize()
        if not torch.equal(y, ref):
            res[key] += 1
            bad = torch.nonzero(y != ref)
            print(key, it, bad.shape[0], bad[:4].tolist(), bad[-2:].tolist(), flush=True)
import ctypes as C
w = (C.c_uint32 * 16)()
L.bb_ntt_diag(w)
print(os.environ.get("TOYNI_NTT_SPLIT"), os.environ.get("TOYNI_NTT_PDL"), os.environ.get("TOYNI_V7_FLAGS"), res, "diag", list(w)[:9], flush=True)
