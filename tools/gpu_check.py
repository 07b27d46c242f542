"""Ad-hoc GPU parity sweep (development aid; the real tests are tests/ -m gpu)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import oracle as O
from toyni_b200 import device as D
from toyni_b200.lib import lib as _lib

def check_ntt(log_n, inverse=False):
    n = 1 << log_n
    x = O.random_field(n, seed=log_n + 7)
    t = D.to_device(x)
    D.ntt_(t, inverse)
    got = D.to_host(t)
    ref = O.intt(x, threads=8) if inverse else O.ntt(x, threads=8)
    ok = np.array_equal(got, ref)
    if not ok:
        bad = np.nonzero(got != ref)[0]
        print(f"  mismatch log_n={log_n} inv={inverse}: {bad.size} bad, first {bad[:5]} got {got[bad[:3]]} ref {ref[bad[:3]]}")
    return ok

if __name__ == "__main__":
    maxlog = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    allok = True
    for log_n in range(0, maxlog + 1):
        for inv in (False, True):
            t0 = time.time()
            ok = check_ntt(log_n, inv)
            allok &= ok
            print(f"ntt log_n={log_n:2d} inv={int(inv)} {'OK' if ok else 'FAIL'} plan={[_lib().bb_ntt_launches(log_n)]} ({time.time()-t0:.2f}s)", flush=True)
    print("ALL OK" if allok else "FAILURES")
    sys.exit(0 if allok else 1)
