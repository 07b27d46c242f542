"""Time alternative pass plans for one transform size (development aid)."""
import ctypes as C, sys, itertools
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import lib

def time_plan(log_n, lrs, lcs, reps=10, nb=4):
    L = lib()
    n = 1 << log_n
    arr_r = (C.c_int * 3)(*lrs, *([0] * (3 - len(lrs))))
    arr_c = (C.c_int * 3)(*lcs, *([0] * (3 - len(lcs))))
    if L.bb_ntt_set_plan(log_n, len(lrs), arr_r, arr_c) != 0:
        L.bb_clear_error()
        return None
    bufs = [torch.randint(0, 2013265921, (n,), dtype=torch.int32, device="cuda") for _ in range(nb)]
    for i in range(3):
        D.ntt_(bufs[i % nb])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        D.ntt_(bufs[i % nb])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us

if __name__ == "__main__":
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    plans = []
    if log_n == 24:
        plans = [([12, 12], [3, 3]), ([12, 12], [2, 2]), ([8, 8, 8], [5, 5, 5]), ([8, 8, 8], [4, 4, 4]), ([8, 8, 8], [3, 3, 3]),
                 ([9, 8, 7], [5, 5, 5]), ([10, 10, 4], [3, 3, 5]), ([10, 10, 4], [4, 4, 5]), ([10, 10, 4], [5, 5, 5]), ([11, 11, 2], [3, 3, 5]),
                 ([11, 11, 2], [4, 4, 5]), ([9, 9, 6], [5, 5, 5]), ([9, 9, 6], [4, 4, 5])]
    elif log_n == 27:
        plans = [([9, 9, 9], [4, 4, 4]), ([9, 9, 9], [5, 5, 5]), ([9, 9, 9], [3, 3, 3]), ([10, 9, 8], [4, 4, 4]), ([8, 9, 10], [4, 4, 4]),
                 ([10, 10, 7], [4, 4, 4]), ([7, 10, 10], [4, 4, 4]), ([11, 8, 8], [4, 4, 4]), ([8, 8, 11], [4, 4, 4]), ([8, 10, 9], [4, 4, 4]),
                 ([10, 9, 8], [5, 5, 5]), ([9, 10, 8], [4, 4, 4])]
    elif log_n == 26:
        plans = [([9, 9, 8], [4, 4, 4]), ([8, 9, 9], [4, 4, 4]), ([9, 8, 9], [4, 4, 4]), ([10, 8, 8], [4, 4, 4]), ([8, 8, 10], [4, 4, 4]), ([9, 9, 8], [5, 5, 5])]
    elif log_n == 25:
        plans = [([9, 8, 8], [4, 4, 4]), ([8, 9, 8], [4, 4, 4]), ([8, 8, 9], [4, 4, 4]), ([9, 8, 8], [5, 5, 5])]
    elif log_n == 20:
        plans = [([10, 10], [3, 3]), ([10, 10], [4, 4]), ([10, 10], [5, 5]), ([7, 7, 6], [5, 5, 5]), ([8, 8, 4], [5, 5, 5]), ([12, 8], [3, 5])]
    elif log_n == 16:
        plans = [([8, 8], [5, 5]), ([8, 8], [3, 3]), ([8, 8], [4, 4])]
    for lrs, lcs in plans:
        if sum(lrs) != log_n:
            continue
        t = time_plan(log_n, lrs, lcs)
        print(f"log_n={log_n} plan rows={lrs} cols={lcs}: " + (f"{t:9.1f} us  {((1<<log_n)/t/1e3):8.1f} Gelem/s" if t else "unavailable"), flush=True)
