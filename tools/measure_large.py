"""Single-GPU timings of the large transforms of BASELINE config 5: 2^25..2^27 direct, 64 x 2^22 batched."""
import json, sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0

def timeit(fn, reps=8, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

out = []
for log_n in (25, 26, 27):
    n = 1 << log_n
    x = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda")
    for inv in (False, True):
        t = timeit(lambda: D.ntt_(x, inv))
        out.append({"what": f"{'inverse' if inv else 'forward'} NTT 2^{log_n}", "us": t, "gelem_s": n / t / 1e3,
                    "frac_of_hbm_roofline": 8.0 * n / (t * 1e-6) / 1e9 / HBM, "kernels": lib().bb_ntt_launches(log_n)})
    del x
b = torch.randint(0, P, (64, 1 << 22), dtype=torch.int32, device="cuda")
t = timeit(lambda: D.ntt_batch_(b, False), reps=5)
out.append({"what": "64 x 2^22 batched forward NTT", "us": t, "gelem_s": 64 * (1 << 22) / t / 1e3,
            "frac_of_hbm_roofline": 8.0 * 64 * (1 << 22) / (t * 1e-6) / 1e9 / HBM})
print(json.dumps(out, indent=1))
