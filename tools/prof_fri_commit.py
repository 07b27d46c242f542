"""ncu / timing target: the prover's FRI commit loop (bb_fri_commit_device) on a 2^log_m Ext codeword with a host
transcript, as bench.py's sharded.fri_commit_ext at N = 1.
usage: python tools/prof_fri_commit.py [log_m = 25] [reps = 3]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
from toyni_b200.prover import FiatShamirTranscript

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, final = 1 << log_m, 16
g = torch.Generator(device="cuda")
g.manual_seed(11)
full = torch.randint(0, P, (m, 4), dtype=torch.int32, device="cuda", generator=g)
salts = torch.randint(0, 256, (32 * m,), dtype=torch.uint8, device="cuda", generator=g)


def once():
    tr = FiatShamirTranscript()

    def challenge(root, layer):
        tr.absorb(root)
        return [tr.squeeze_challenge() for _ in range(4)]
    return D.fri_commit(full, 7, final, salts, challenge=challenge)


once()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    once()
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"FRI commit loop 2^{log_m} Ext -> {final}: best {min(ts) * 1e3:.3f} ms, all {[round(t * 1e3, 3) for t in ts]}")
