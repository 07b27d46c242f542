"""Sharded FRI commit loop under torchrun (one process per GPU): parity of every root with the single-GPU loop at
2^18 (base field and Ext), then timing of the 2^25 Ext loop (BASELINE config 4 (ii)): fold + per-layer salted commit
+ replicated host transcript.  usage: torchrun ... tools/mg_fri_commit.py [log_m = 25] [reps = 5]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200 import multigpu as MG
from toyni_b200.lib import P
from toyni_b200.prover import FiatShamirTranscript

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
log_big = int(sys.argv[1]) if len(sys.argv) > 1 else 25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = MG.CudaFriBackend(dev)
FINAL, SHIFT = 16, 7


def make_challenge(limbs):
    tr = FiatShamirTranscript()

    def challenge(root, layer):
        tr.absorb(root)
        return tr.squeeze_challenge() if limbs == 1 else [tr.squeeze_challenge() for _ in range(4)]
    return challenge


def salts_table(m, seed):
    """All ranks build the same global salt array (parity check only), layer after layer."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    sizes, mm = [], m
    while mm > FINAL:
        sizes.append(mm)
        mm //= 2
    allsalt = torch.randint(0, 256, (sum(sizes), 16), dtype=torch.uint8, device=dev, generator=g)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return allsalt, (lambda k, lo, hi: allsalt[offs[k] + lo: offs[k] + hi])


ok = True
for limbs in (1, 4):
    log_m = 18
    m = 1 << log_m
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + limbs)
    full = torch.randint(0, P, (m, 4) if limbs == 4 else (m,), dtype=torch.int32, device=dev, generator=g)
    allsalt, salts_for = salts_table(m, 99)
    ref_layers, _, ref_roots = D.fri_commit(full, SHIFT, FINAL, allsalt.reshape(-1), challenge=make_challenge(limbs))
    roots, layers, nodes, tail = MG.fri_commit_sharded(full[rank::world].contiguous(), log_m, SHIFT, FINAL, salts_for,
                                                       make_challenge(limbs), rank, world, B)
    good = roots == ref_roots and torch.equal(tail[-1], ref_layers[-1])
    good &= all(torch.equal(lay, ref_layers[k][rank::world]) for k, lay in enumerate(layers))
    ok &= bool(good)
    if rank == 0:
        print(f"sharded fri commit limbs={limbs} world={world} layers={len(roots)} sharded={len(nodes)}: {'OK' if good else 'FAIL'}", flush=True)

# ---- timing: 2^log_big Ext codeword, salts drawn per rank (any salts do for a timing)
m = 1 << log_big
g = torch.Generator(device=dev)
g.manual_seed(5 + rank)
local0 = torch.randint(0, P, (m // world, 4), dtype=torch.int32, device=dev, generator=g)
pool = torch.randint(0, 256, (m // world, 16), dtype=torch.uint8, device=dev, generator=g)
salts_for = lambda k, lo, hi: pool[: hi - lo]
times = []
for r in range(reps + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    roots, layers, nodes, tail = MG.fri_commit_sharded(local0, log_big, SHIFT, FINAL, salts_for, make_challenge(4), rank, world, B)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if r:
        times.append(float(dt.item()))
    del layers, nodes, tail
if rank == 0:
    out = {"workload": f"FRI commit loop, Ext codeword 2^{log_big} -> 16, salted tree per layer, replicated host transcript",
           "n_gpus": world, "ms_best": round(min(times) * 1e3, 3), "ms_median": round(sorted(times)[len(times) // 2] * 1e3, 3),
           "layers": len(roots), "parity_2^18": bool(ok)}
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/fri_commit_g{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
