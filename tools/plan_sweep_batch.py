"""Single-pass (vectors as tile columns) vs two-pass plans for batches of short vectors (development aid)."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import lib
L = lib()

def time_batch(log_n, plan, reps=20):
    n = 1 << log_n
    B = (1 << 24) // n
    if plan is None:
        z = (C.c_int * 3)(0, 0, 0)
        L.bb_ntt_set_plan(log_n, 0, z, z)
    else:
        lrs, lcs = plan
        ar = (C.c_int * 3)(*lrs, *([0] * (3 - len(lrs)))); ac = (C.c_int * 3)(*lcs, *([0] * (3 - len(lcs))))
        if L.bb_ntt_set_plan(log_n, len(lrs), ar, ac) != 0:
            L.bb_clear_error(); return None
    bufs = [torch.randint(0, 2013265921, (B, n), dtype=torch.int32, device="cuda") for _ in range(4)]
    ref = None
    for i in range(3):
        D.ntt_batch_(bufs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        D.ntt_batch_(bufs[i % 4])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

import os
SIZES = [int(a) for a in sys.argv[1:]] or [9, 10, 11, 12, 13, 14]
for log_n in SIZES:
    if log_n <= 14:
        plans = [None] + [([log_n], [lc]) for lc in (2, 3, 4, 5)]
    else:
        a = (log_n + 1) // 2
        plans = [None] + [([a, log_n - a], [lc, lc]) for lc in (3, 4, 5)] + [([log_n - a, a], [4, 4])]
    for p in plans:
        t = time_batch(log_n, p)
        print(f"log_n={log_n} plan={'default' if p is None else p}: " + (f"{t:8.1f} us {(1<<24)/t/1e3:7.1f} Gelem/s" if t else "unavailable"), flush=True)
    z = (C.c_int * 3)(0, 0, 0); L.bb_ntt_set_plan(log_n, 0, z, z)
