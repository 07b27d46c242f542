"""One Ext fold 2^25 -> 2^24 and one salted commit of 2^24 base-field leaves (target for ncu)."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
n = 1 << 25
ev = torch.randint(0, P, (n, 4), dtype=torch.int32, device="cuda")
out = torch.empty((n // 2, 4), dtype=torch.int32, device="cuda")
m = 1 << 24
vals = torch.randint(0, P, (m,), dtype=torch.int32, device="cuda")
salts = torch.randint(0, 256, (m, 16), dtype=torch.uint8, device="cuda")
nodes = torch.empty((D.merkle_node_count(m), 32), dtype=torch.uint8, device="cuda")
for _ in range(2):
    D.fri_fold(ev, 7, [1, 2, 3, 4], out=out)
    D.merkle_commit(vals, salts, nodes, want_root=False)
torch.cuda.synchronize()
print("done")
