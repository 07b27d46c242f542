"""Run a few device-resident transforms of one size/plan (target for ncu)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import lib

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
plan = sys.argv[2] if len(sys.argv) > 2 else ""
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L = lib()
if plan:
    rows, cols = plan.split("/")
    lrs = [int(x) for x in rows.split(",")]
    lcs = [int(x) for x in cols.split(",")]
    ar = (C.c_int * 3)(*lrs, *([0] * (3 - len(lrs))))
    ac = (C.c_int * 3)(*lcs, *([0] * (3 - len(lcs))))
    assert L.bb_ntt_set_plan(log_n, len(lrs), ar, ac) == 0
n = 1 << log_n
bufs = [torch.randint(0, 2013265921, (n,), dtype=torch.int32, device="cuda") for _ in range(4)]
for i in range(reps):
    D.ntt_(bufs[i % 4])
torch.cuda.synchronize()
print("done", log_n, plan)
