"""Multi-GPU parity check (run under torchrun): four-step NTT, cyclic fold shards, column-sharded batch."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from oracle import oracle as O
from toyni_b200 import device as D, multigpu as MG

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for log_n in (14, 16, 20):
    n = 1 << log_n
    x = O.random_field(n, seed=log_n)
    ref = O.ntt(x, threads=4)
    for inv in (False, True):
        blk = D.to_device(MG.fourstep_scatter(x, rank, world))
        out = MG.fourstep_ntt_cuda(blk, log_n, rank, world, inverse=inv)
        n1, n2 = MG.fourstep_split(log_n, world)
        k1 = np.arange(rank * n1 // world, (rank + 1) * n1 // world)
        want = (O.intt(x, threads=4) if inv else ref).reshape(n2, n1)[:, k1].T
        good = np.array_equal(D.to_host(out), want)
        ok &= good
        if rank == 0: print(f"fourstep log_n={log_n} inv={inv} world={world}: {'OK' if good else 'FAIL'}", flush=True)
# fused exchange (peer stores over NVLink inside the NTT pass)
for log_n in (14, 16, 20, 22):
    n = 1 << log_n
    x = O.random_field(n, seed=100 + log_n)
    fs = MG.FourStepFused(log_n, rank, world)
    for inv in (False, True, False):
        blk = D.to_device(MG.fourstep_scatter(x, rank, world))
        out = fs.run(blk, inverse=inv)
        n1, n2 = MG.fourstep_split(log_n, world)
        k1 = np.arange(rank * n1 // world, (rank + 1) * n1 // world)
        want = (O.intt(x, threads=4) if inv else O.ntt(x, threads=4)).reshape(n2, n1)[:, k1].T
        good = np.array_equal(D.to_host(out), want)
        ok &= good
        if rank == 0: print(f"fused fourstep log_n={log_n} inv={inv} world={world}: {'OK' if good else 'FAIL'}", flush=True)
    fs.check_peers()
    fs.close()
# two transforms in flight (run_async): six different vectors, outputs read in issue order
for log_n in (16, 22):
    n = 1 << log_n
    fs = MG.FourStepFused(log_n, rank, world, nbuf=3)
    n1, n2 = MG.fourstep_split(log_n, world)
    k1 = np.arange(rank * n1 // world, (rank + 1) * n1 // world)
    xs_ = [O.random_field(n, seed=500 + log_n + i) for i in range(6)]
    wants = [O.ntt(x, threads=4).reshape(n2, n1)[:, k1].T for x in xs_]
    pend, good, got_n = [], True, 0
    for i in range(6):
        pend.append(fs.run_async(D.to_device(MG.fourstep_scatter(xs_[i], rank, world))))
        if len(pend) == 2:
            o, ev = pend.pop(0)
            torch.cuda.current_stream().wait_event(ev)
            good &= np.array_equal(D.to_host(o), wants[got_n]); got_n += 1
    for o, ev in pend:
        torch.cuda.current_stream().wait_event(ev)
        good &= np.array_equal(D.to_host(o), wants[got_n]); got_n += 1
    fs.join()
    out = fs.run(D.to_device(MG.fourstep_scatter(xs_[0], rank, world)))  # back to the serial form on the same object
    good &= np.array_equal(D.to_host(out), wants[0])
    fs.check_peers()
    fs.close()
    ok &= good
    if rank == 0: print(f"pipelined fused fourstep log_n={log_n} world={world}: {'OK' if good else 'FAIL'}", flush=True)
# cyclic fold chain
m_log = 14
ee = O.random_field(4 << m_log, seed=3).reshape(1 << m_log, 4)
betas = [[k + 1, k + 2, k + 3, k + 4] for k in range(m_log)]
layers = MG.fold_chain_cuda(D.to_device(np.ascontiguousarray(ee[rank::world])), m_log, 7, betas, rank, world, until=16)
cur, xs = ee, O.domain_elements(1 << m_log, 7)
good = True
for k in range(1, len(layers)):
    cur = O.fri_fold_ext(cur, xs, betas[k - 1])
    xs = (xs[: cur.shape[0]].astype(object) ** 2 % O.P).astype(np.uint64)
    good &= np.array_equal(D.to_host(layers[k]), cur[rank::world])
ok &= good
if rank == 0: print(f"cyclic fold chain 2^{m_log} -> {cur.shape[0]} over {world} ranks: {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("ALL OK" if int(t.item()) else "FAILURES", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
