"""A few salted Merkle commits of 2^log_n base-field leaves (target for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << log_n
vals = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda")
salts = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device="cuda")
nodes = torch.empty((D.merkle_node_count(n), 32), dtype=torch.uint8, device="cuda")
for _ in range(2):
    D.merkle_commit(vals, salts, nodes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    D.merkle_commit(vals, salts, nodes)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"commit 2^{log_n}: {ms:.3f} ms, {3 * n / ms / 1e6:.2f} G compressions/s")
