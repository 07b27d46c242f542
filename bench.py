#!/usr/bin/env python
"""bench.py — BabyBear NTT throughput on B200 (BASELINE.json metric: "BabyBear NTT Gelem/s @2^24").

One step = one forward 2^24-point NTT (the configuration the metric is quoted on, BASELINE.json configs[1]) on
synthetic random field elements.  `value` is device-resident throughput (inputs already in HBM, CUDA-event
timed, rotating over more buffers than fit in L2); `e2e` is the same transform through the reference-facing
C ABI (`ntt_run_inplace`, src/ntt.rs:108) on pinned HOST u64 buffers with both PCIe copies inside the timed
region.  With --gpus N every rank transforms its own column (independent NTTs shard with no collective: weak
scaling); --workload fourstep27 runs ONE 2^27 transform sharded over the ranks with an all-to-all instead.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ntt24|fourstep27]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "babybear_ntt_2^24_throughput"
UNIT = "Gelem/s"
LOG_N = 24


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _traffic(log_n):
    """dram bytes per transform from the committed ncu capture (profiles/ncu_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(f"ntt_2^{log_n}_dram_bytes_per_transform")
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """The reference's own CPU implementation of the path (src/ntt.rs:24-53).  No Rust toolchain exists in this
    image, so it is the C port in oracle/ (same loop order, same u128 % p multiply), with every host thread the
    algorithm can use (independent butterfly groups per stage over OpenMP)."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    # one step = one forward NTT of 2^24; if K steps of that would not end within a few minutes, each step becomes
    # a smaller bounded sample of the same workload (a 2^22 / 2^20 transform) and the line says so
    log_s = LOG_N
    x = O.random_field(1 << log_s)
    t0 = time.perf_counter()
    O.ntt_inplace(x, threads=cores)
    probe = time.perf_counter() - t0
    while log_s > 20 and probe * (args.steps + args.warmup) > 150.0:
        log_s -= 2
        probe /= 4.4
        x = O.random_field(1 << log_s)
    n = 1 << log_s
    for _ in range(args.warmup):
        O.ntt_inplace(x, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.ntt_inplace(x, threads=cores)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "forward BabyBear NTT, n=2^24, one vector per step (BASELINE configs[1])",
                   "note": "reference CPU algorithm (src/ntt.rs:24-53) as the C port oracle/toyni_oracle.c; "
                           "the reference is Rust and no Rust toolchain exists here"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} forward NTTs of 2^{log_s} after {args.warmup} warm-ups, OpenMP over {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from toyni_b200 import device as D
    from toyni_b200 import ntt as host_ntt
    from toyni_b200.lib import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = lib()
    if not L.bb_device_ok():
        raise SystemExit("bench.py needs a compute-capability 10.x device (no CPU fallback)")
    n = 1 << LOG_N

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "fourstep27":
        from toyni_b200 import multigpu
        return multigpu.bench_fourstep(args, rank, world, dev)

    # Inputs: NB distinct device-resident vectors (NB * 64 MiB > the 126 MB L2), uniformly random canonical values.
    NB = 4
    g = torch.Generator(device=dev)
    g.manual_seed(0x70796E69 + rank)
    bufs = [torch.randint(0, 2013265921, (n,), dtype=torch.int32, device=dev, generator=g) for _ in range(NB)]
    L.bb_warmup(LOG_N)

    # Steps are independent transforms of different vectors, so they are issued alternately on NS CUDA streams through
    # the library's stream-ordered API (bb_set_stream): while one transform is in the memory-bound part of a pass the
    # other one's arithmetic fills the SMs (120 -> 101 us per transform on one B200).  Buffer i % NB always goes to
    # stream i % NS (NB is a multiple of NS), so no buffer is ever touched from two streams.
    NS = 2
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]

    def step(i):
        with torch.cuda.stream(streams[i % NS]):
            D.ntt_(bufs[i % NB], inverse=False)

    for i in range(max(args.warmup, NB)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)  # nvidia-smi needs ~0.2 s before its first line: have it sampling when the timed region starts
    for i in range(NB):  # back under load after the pause
        step(i)
    barrier()
    launches0 = L.bb_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for st in streams:
        st.wait_event(ev0)
    for i in range(args.steps):
        step(i)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    ev1.record()
    barrier()
    launches = L.bb_kernel_launch_count() - launches0
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = ms_total / args.steps  # the pass kernels are the whole timed region: K transforms x 3 launches
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the reference-facing C ABI on pinned host buffers (every rank, max over ranks).
    # Serial: one caller, one context (what src/ntt.rs:224-236 does) - H2D, kernels, D2H strictly one after the other.
    # Pipelined: NTHR host threads, each with its own context and pinned buffer (ntt_ctx_create once per thread); one
    # thread's H2D overlaps another's D2H on the full-duplex PCIe link and a third's kernels.  Both forms move
    # 8 B/element each way inside the timing.
    import ctypes
    import threading
    host = torch.empty(n, dtype=torch.int64).pin_memory()
    hv = host.numpy().view(np.uint64)
    hv[:] = (np.arange(n, dtype=np.uint64) * np.uint64(7) + np.uint64(3)) % np.uint64(2013265921)
    NTHR = 3
    e2e_steps = max(NTHR * 2, min(args.steps, 24)) // NTHR * NTHR
    for _ in range(2):
        host_ntt.ntt_cuda(hv)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_ntt.ntt_cuda(hv)  # H2D (u64) + narrow + NTT passes + widen + D2H (u64), synchronous
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t0

    hosts = [hv] + [torch.empty(n, dtype=torch.int64).pin_memory().numpy().view(np.uint64) for _ in range(NTHR - 1)]
    for h in hosts[1:]:
        h[:] = hv
    ctxs = [L.ntt_ctx_create(n) for _ in range(NTHR)]
    assert all(ctxs), "ntt_ctx_create failed"

    def pump(k, count):
        torch.cuda.set_device(local_rank)  # the CUDA current device is per host thread
        for _ in range(count):
            L.ntt_run_inplace(ctypes.c_void_p(ctxs[k]), hosts[k].ctypes.data)

    for k in range(NTHR):
        pump(k, 1)
    barrier()
    threads = [threading.Thread(target=pump, args=(k, e2e_steps // NTHR)) for k in range(NTHR)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_err = int(L.bb_last_error())  # never raise here: the other ranks would wait in the reduction below forever
    for c in ctxs:
        L.ntt_ctx_destroy(ctypes.c_void_p(c))
    t = torch.tensor([e2e_s, e2e_serial_s, float(e2e_err)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * n * e2e_steps / float(t[0].item()) / 1e9
    e2e_serial_val = world * n * e2e_steps / float(t[1].item()) / 1e9
    if t[2].item() != 0:
        raise SystemExit(f"bench.py: CUDA error {int(t[2].item())} in the end-to-end path")

    if rank != 0:
        return
    peak, peak_kind = _peaks()
    npass = L.bb_ntt_launches(LOG_N)
    achieved = 8.0 * n / (kernel_ms * 1e-3) / 1e9  # algorithmic bytes: 4 B read + 4 B written per element per transform
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "forward BabyBear NTT, n=2^24, one vector per GPU per step (BASELINE configs[1])",
                   "l2": f"inputs rotate over {NB} x 64 MiB device buffers (larger than the 126 MB L2)",
                   "kernels_per_transform": npass, "streams": NS,
                   "parallelism": f"independent columns x{world}, no collective; steps alternate over {NS} CUDA streams per GPU"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _traffic(LOG_N), "peak_source": peak_kind,
                     "note": "algorithmic 8 B/element over the transform's pass kernels (CUDA events around the timed region, K transforms = 3K launches on two streams); "
                             "three passes move 24 B/element (a strided 64 MB-in / 64 MB-out pass alone costs 30-37 us, "
                             "tools/ubench_strided.cu) next to ~27 us of integer work per pass, see DESIGN.md and profiles/"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "steps": e2e_steps, "api": f"ntt_run_inplace (src/ntt.rs:108) on pinned host u64, {NTHR} host threads with "
                                           "one context each (H2D of one overlaps D2H and kernels of the others)",
                "serial_value": e2e_serial_val, "serial_api": "one caller, one context: ntt_cuda as in src/ntt.rs:224-236"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        x = O.random_field(n)
        t0 = time.perf_counter()
        O.ntt_inplace(x, threads=1)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "one forward NTT of 2^24, single thread (the reference has no threading), "
                                          "C port of src/ntt.rs:24-53"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ntt24", choices=["ntt24", "fourstep27"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="fourstep27: fused peer stores or NCCL all-to-all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps is None:
        args.steps = 2000 if args.impl == "ours" else 10

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
