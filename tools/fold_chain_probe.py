import sys, time, torch
sys.path.insert(0, ".")
from toyni_b200 import multigpu as MG
from toyni_b200.lib import P
for log_m, world in ((25, 1), (22, 1), (25, 8)):
    m = 1 << log_m
    local = torch.randint(0, P, (m // world, 4), dtype=torch.int32, device="cuda")
    betas = [[(7 * k + j + 1) % P for j in range(4)] for k in range(log_m)]
    ch = MG.FoldChain(log_m, 7, betas, 0, world, 4, until=16)
    for fn, name in ((lambda: ch.run(local), "FoldChain.run"), (lambda: MG.fold_chain_cuda(local, log_m, 7, betas, 0, world, until=16), "fold_chain_cuda")):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"log_m={log_m} shard 1/{world}: {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
