"""NVLink evidence and a per-phase timeline for the fused four-step 2^27 transform (run under torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
      tools/nvlink_fourstep.py [K]

1. link counters: `nvidia-smi nvlink -gt d` (data KiB sent / received per link, summed per GPU) read by rank 0 before
   and after K transforms; the delta per transform per GPU is compared with the bytes the algorithm stores to peers,
   (n / G) * (G - 1) / G * 4.
2. timeline: CUDA events between the phases of one transform (column passes incl. the peer stores, signal + wait,
   row passes), averaged over K transforms, max over ranks.
Prints one JSON object."""
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from toyni_b200 import multigpu as MG  # noqa: E402
from toyni_b200.device import _bind_stream, _chk, ntt_batch_  # noqa: E402
from toyni_b200.lib import P, check, lib  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
LOG_N = 27
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def link_counters():
    """{gpu: (tx_kib, rx_kib)} summed over the links of each GPU."""
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True, timeout=60).stdout
    res, gpu = {}, None
    for line in out.splitlines():
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            res[gpu] = [0, 0]
            continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            res[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return res, out


def max_over_ranks(v):
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


n = 1 << LOG_N
n1, n2 = MG.fourstep_split(LOG_N, world)
cw = n2 // world
g = torch.Generator(device="cuda")
g.manual_seed(1234 + rank)
blocks = [torch.randint(0, P, (n1, cw), dtype=torch.int32, device="cuda", generator=g) for _ in range(3)]
fs = MG.FourStepFused(LOG_N, rank, world)
for i in range(3):
    fs.run(blocks[i % 3])
torch.cuda.synchronize()
fs.check_peers()
dist.barrier()

# 1. link counters around K transforms
before = raw_before = None
if rank == 0:
    before, raw_before = link_counters()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(K):
    fs.run(blocks[i % 3])
e1.record()
torch.cuda.synchronize()
ms = max_over_ranks(e0.elapsed_time(e1) / K)
fs.check_peers()
dist.barrier()
links = None
if rank == 0:
    after, raw_after = link_counters()
    expect = (n // world) * (world - 1) // world * 4
    links = {"expected_bytes_stored_to_peers_per_gpu_per_transform": expect, "per_gpu": {}}
    for gpu in sorted(after):
        if gpu in before:
            tx = (after[gpu][0] - before[gpu][0]) * 1024 / K
            rx = (after[gpu][1] - before[gpu][1]) * 1024 / K
            links["per_gpu"][str(gpu)] = {"tx_bytes_per_transform": tx, "rx_bytes_per_transform": rx,
                                          "tx_over_expected": tx / expect if expect else None}
    links["raw_sample"] = raw_after[:700]

# 2. per-phase timeline (same calls as FourStepFused.run, events between them)
import ctypes as C  # noqa: E402

L = lib()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
dist.barrier()
torch.cuda.synchronize()
for i in range(K):
    block = blocks[i % 3]
    _bind_stream()
    fs.epoch += 1
    k = fs.epoch % fs.NBUF
    ev[i][0].record()
    check(L.bb_ntt_columns_scatter_device(_chk(block), LOG_N, n1.bit_length() - 1, cw, 0, fs.peers[k], world, rank),
          "bb_ntt_columns_scatter_device")
    ev[i][1].record()
    check(L.bb_peer_signal_device(C.c_void_p(fs.d_peer_flags.data_ptr()), world, rank, fs.epoch), "bb_peer_signal_device")
    check(L.bb_peer_wait_device(fs.flags_ptr, world, fs.epoch, fs.err_ptr), "bb_peer_wait_device")
    ev[i][2].record()
    ntt_batch_(fs.out[k], False)
    ev[i][3].record()
torch.cuda.synchronize()
fs.check_peers()
ph = [sum(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(5, K)) / (K - 5) for j in range(3)]
phases = {"column_passes_with_peer_stores_ms": max_over_ranks(ph[0]), "signal_and_wait_ms": max_over_ranks(ph[1]),
          "row_passes_ms": max_over_ranks(ph[2]),
          "note": "max over ranks of each phase's mean; the wait absorbs the skew between ranks and the drain of the peer stores"}
# 3. two transforms in flight (run_async)
ms_pipe = None
if fs.NBUF >= 3:
    for i in range(4):
        fs.run_async(blocks[i % 3])
    fs.join()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        fs.run_async(blocks[i % 3])
    fs.join()
    e1.record()
    torch.cuda.synchronize()
    ms_pipe = max_over_ranks(e0.elapsed_time(e1) / K)
fs.check_peers()
fs.close()
if rank == 0:
    sent = (n // world) * (world - 1) // world * 4
    print(json.dumps({"world": world, "nbuf": fs.NBUF, "log_n": LOG_N, "n1": n1, "n2": n2, "transforms": K, "ms_per_transform": ms, "ms_per_transform_two_in_flight": ms_pipe,
                      "nvlink_gbs_per_gpu_over_whole_transform": sent / (ms * 1e-3) / 1e9,
                      "nvlink_gbs_per_gpu_over_column_phase": sent / (phases["column_passes_with_peer_stores_ms"] * 1e-3) / 1e9,
                      "links": links, "phases": phases}))
dist.destroy_process_group()
