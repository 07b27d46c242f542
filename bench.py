#!/usr/bin/env python
"""bench.py — BabyBear NTT throughput on B200 (BASELINE.json metric: "BabyBear NTT Gelem/s @2^24, 2^27 (1/2/4/8 B200);
LDE+FRI commit ms").

Headline (unchanged between rounds): one step = one forward 2^24-point NTT per GPU (BASELINE configs[1]) on synthetic
random field elements.  `value` is device-resident throughput (inputs already in HBM, CUDA-event timed, rotating over
more buffers than fit in L2); `e2e` is the same transform through the reference-facing C ABI (`ntt_run_inplace`,
src/ntt.rs:108) on pinned HOST u64 buffers with both PCIe copies inside the timed region.  With --gpus N every rank
transforms its own vector (weak scaling, no collective).

The rest of the metric rides on the same JSON line:
  `sharded` (every N) — ONE 2^27 four-step transform over all ranks (peer stores over NVLink fused into the NTT pass),
            64 x 2^22 column-sharded NTTs, the 2^25 Ext fold chain on cyclic shards and the sharded FRI commit loop
            (BASELINE configs[3], [4]); every record carries `parity_ok` = outputs of ALL ranks compared with the CPU
            oracle (checker only, outside the timed regions).
  `extras`  (N = 1)   — coset LDE 2^20 -> 2^25 and the salted Merkle commit of its 2^25 leaves (configs[2]), each with
            its own roofline (the FRI commit loop of configs[3] is `sharded.fri_commit_ext_2^25`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-sharded] [--no-extras]
"""
import argparse
import hashlib
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
P = 2013265921
SEED = 0x70796E69
NVLINK_MEASURED_GBS = 770.0  # peer copy per direction per GPU measured on this pool (B200_PROFILING.md); nominal 900


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _traffic(log_n):
    """dram bytes per transform from the committed steady-state ncu capture (profiles/ncu_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(f"ntt_2^{log_n}_dram_bytes_per_transform")
    except Exception:
        return None


def _pin_to_gpu_numa(local_rank):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU's PCIe root."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "numa_node": open(path.replace("local_cpulist", "numa_node")).read().strip()}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:80]}
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


# ----------------------------------------------------------------------------------------------- helpers (our arm)
class Ctx:
    """What every measurement needs: torch, the library, this rank's place in the job."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        from toyni_b200 import device as D
        from toyni_b200.lib import lib
        self.torch, self.dist, self.D, self.L = torch, dist, D, lib()
        self.rank, self.world, self.local = rank, world, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.cores = os.cpu_count() or 1

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def min_over_ranks(self, flag):
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def time_ms(self, fn, reps, warm=2, join=None):
        """CUDA-event time of `reps` calls on the current stream, max over ranks, per call.  `join` (optional) makes the
        current stream wait for work `fn` issued on other streams, before the closing event."""
        torch = self.torch
        for _ in range(warm):
            fn()
        if join:
            join()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        if join:
            join()
        e1.record()
        self.barrier()
        return self.max_over_ranks([e0.elapsed_time(e1)])[0] / reps

    def gather_digests(self, digests):
        """digests: list of 32-byte strings of this rank -> on every rank, list per rank of those lists."""
        torch = self.torch
        mine = torch.tensor(list(b"".join(digests)), dtype=torch.uint8, device=self.dev)
        if self.world == 1:
            return [digests]
        allt = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allt, mine)
        out = []
        for t in allt:
            b = bytes(t.cpu().tolist())
            out.append([b[i:i + 32] for i in range(0, len(b), 32)])
        return out


def _sha(t):
    """SHA-256 of a device tensor's bytes (int32 / uint8, contiguous)."""
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).digest()


def _sha_np_i32(a_u64):
    return hashlib.sha256(np.ascontiguousarray(a_u64).astype(np.int32).tobytes()).digest()


def _seeded_field(c, shape, seed):
    g = c.torch.Generator(device=c.dev)
    g.manual_seed(seed)
    return c.torch.randint(0, P, shape, dtype=c.torch.int32, device=c.dev, generator=g)


# ----------------------------------------------------------------------------------------------- sharded workloads
def sharded_fourstep27(c, steps):
    """ONE 2^27 forward NTT over all ranks (BASELINE configs[4]): four-step, the inter-half twiddle and the transpose
    inside the last column pass, rows stored straight into the owning rank's buffer over NVLink (CUDA IPC)."""
    from toyni_b200 import multigpu as MG
    torch, D = c.torch, c.D
    log_n = 27
    n = 1 << log_n
    n1, n2 = MG.fourstep_split(log_n, c.world)
    cw, rw = n2 // c.world, n1 // c.world
    full = _seeded_field(c, (n,), SEED + 27)  # every rank draws the same vector and keeps its column block
    block = full.view(n1, n2)[:, c.rank * cw:(c.rank + 1) * cw].contiguous()
    fused = MG.FourStepFused(log_n, c.rank, c.world)
    out = fused.run(block.clone())
    torch.cuda.synchronize()
    fused.check_peers()
    digests = c.gather_digests([_sha(out)])
    parity = None
    if c.rank == 0:  # checker: the whole vector through the CPU oracle (src/ntt.rs:24-53), slab by slab
        from oracle import oracle as O
        ref = O.ntt(full.cpu().numpy().astype(np.uint64), threads=c.cores).reshape(n2, n1)  # [k2][k1] = X[k1 + n1 k2]
        parity = all(_sha_np_i32(ref[:, r * rw:(r + 1) * rw].T) == digests[r][0] for r in range(c.world))
        del ref
    # best single-GPU transform of the same length on this box (plain three-pass plan), for the strong-scaling ratio
    t_single = c.time_ms(lambda: D.ntt_(full), 10)
    del full
    work = [block.clone() for _ in range(3)]
    it = [0]

    def step():
        fused.run(work[it[0] % 3])
        it[0] += 1
    ms = c.time_ms(step, steps, warm=3)
    fused.check_peers()
    fused.check_peers()
    fused.close()
    sent = (n // c.world) * (c.world - 1) // c.world * 4
    gbs = sent / (ms * 1e-3) / 1e9
    return {"workload": "one forward 2^27 NTT, four-step over column blocks, peer stores over NVLink fused into the NTT pass",
            "ms": ms, "gelem_s": n / (ms * 1e-3) / 1e9, "scaling": "strong", "n1": n1, "n2": n2,
            "nvlink_bytes_sent_per_gpu": sent, "nvlink_gbs_per_gpu": gbs,
            "nvlink_frac_of_measured_770": gbs / NVLINK_MEASURED_GBS, "nvlink_frac_of_nominal_900": gbs / 900.0,
            "nvlink_note": "bytes this GPU stores to its peers / time of the WHOLE transform (not a link counter)",
            "single_gpu_2^27_ms": t_single, "speedup_vs_single_gpu": t_single / ms, "steps": steps, "parity_ok": parity,
            "parity": "every output slab of every rank, SHA-256 against the CPU oracle on the same seeded input"}


def sharded_columns(c, steps):
    """64 independent 2^22-point column NTTs (BASELINE configs[4]), column j on rank j mod G: no communication."""
    torch, D = c.torch, c.D
    cols, log_n = 64, 22
    n = 1 << log_n
    mine = list(range(c.rank, cols, c.world))
    batch = torch.stack([_seeded_field(c, (n,), SEED + 1000 + j) for j in mine])
    got = D.ntt_batch_(batch.clone(), False)
    from oracle import oracle as O
    ok = True
    for idx in sorted({0, len(mine) - 1}):  # checker: first and last local column, every element
        ref = O.ntt(batch[idx].cpu().numpy().astype(np.uint64), threads=c.cores)
        ok &= _sha_np_i32(ref) == _sha(got[idx])
    parity = c.min_over_ranks(ok)
    work = [batch.clone() for _ in range(2)]
    it = [0]

    def step():
        D.ntt_batch_(work[it[0] % 2], False)
        it[0] += 1
    ms = c.time_ms(step, steps, warm=2)
    return {"workload": f"64 columns x 2^22 forward NTT, {len(mine)} columns per GPU, no collective", "ms": ms,
            "gelem_s": cols * n / (ms * 1e-3) / 1e9, "scaling": "strong", "steps": steps, "parity_ok": parity,
            "parity": "first and last column of every rank, every element against the CPU oracle"}


def _oracle_fold_chain(O, full_u64, shift, betas, final):
    """src/math/fri.rs:7-25 layer after layer with xs = the prover's squared domain (src/fibonacci.rs:214,228-231)."""
    m = full_u64.shape[0]
    xs = O.domain_elements(m, shift)
    cur, layers, k = full_u64, [], 0
    while cur.shape[0] > final:
        cur = O.fri_fold_ext(cur, xs, betas[k])
        xs = (xs[: cur.shape[0]] * xs[: cur.shape[0]]) % np.uint64(P)  # p^2 < 2^62: exact in uint64
        layers.append(cur)
        k += 1
    return layers


def sharded_fold_chain(c, steps):
    """Ext fold chain 2^25 -> 16 with the betas supplied up front (BASELINE configs[3] (i)): cyclic shards, no exchange."""
    from toyni_b200 import multigpu as MG
    log_m, shift, final = 25, 7, 16
    m = 1 << log_m
    full = _seeded_field(c, (m, 4), SEED + 25)
    local = full[c.rank::c.world].contiguous()
    betas = [[(7 * k + j + 1) % P for j in range(4)] for k in range(log_m)]
    layers = MG.fold_chain_cuda(local, log_m, shift, betas, c.rank, c.world, until=final)
    digests = c.gather_digests([_sha(l) for l in layers[1:]])
    parity = None
    if c.rank == 0:
        from oracle import oracle as O
        O.set_threads(c.cores)
        ref = _oracle_fold_chain(O, full.cpu().numpy().astype(np.uint64), shift, betas, final)
        O.set_threads(1)
        parity = all(_sha_np_i32(ref[k][r::c.world]) == digests[r][k] for r in range(c.world) for k in range(len(digests[r])))
        parity &= len(digests[0]) >= 1
        del ref
    nl = len(layers) - 1
    del full, layers
    chain = MG.FoldChain(log_m, shift, betas, c.rank, c.world, 4, until=final)  # sizes, betas and arena prepared once
    ms = c.time_ms(lambda: chain.run(local), steps, warm=2)
    algo = sum(24 * (m >> k) for k in range(nl))  # 16 B read + 8 B written per Ext input
    peak, kind = _peaks()
    return {"workload": f"Ext fold chain 2^25 -> {m >> nl} ({nl} folds on cyclic shards, betas up front), no exchange",
            "ms": ms, "folds": nl, "scaling": "strong", "steps": steps,
            "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak * c.world, "unit": "GB/s",
                         "frac": algo / (ms * 1e-3) / 1e9 / (peak * c.world), "peak_source": kind + " x n_gpus"},
            "parity_ok": parity, "parity": "every layer of every rank, SHA-256 against the CPU oracle (OpenMP over the fold loop)"}


def sharded_fri_commit(c, steps):
    """The prover's FRI commit loop (src/fibonacci.rs:200-247) on a 2^25 Ext codeword (BASELINE configs[3] (ii)): fold +
    per-layer salted Merkle tree + host transcript; cyclic shards, one cyclic -> block exchange per layer."""
    from toyni_b200 import multigpu as MG
    from toyni_b200.prover import FiatShamirTranscript
    torch = c.torch
    log_m, shift, final = 25, 7, 16
    m = 1 << log_m
    full = _seeded_field(c, (m, 4), SEED + 26)
    sizes, mm = [], m
    while mm > final:
        sizes.append(mm)
        mm //= 2
    g = torch.Generator(device=c.dev)
    g.manual_seed(SEED + 2)
    allsalt = torch.randint(0, 256, (sum(sizes), 16), dtype=torch.uint8, device=c.dev, generator=g)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    salts_for = lambda k, lo, hi: allsalt[offs[k] + lo: offs[k] + hi]  # noqa: E731
    B = MG.CudaFriBackend(c.dev)

    def make_challenge():
        tr = FiatShamirTranscript()

        def challenge(root, layer):
            tr.absorb(root)
            return [tr.squeeze_challenge() for _ in range(4)]
        return challenge
    local0 = full[c.rank::c.world].contiguous()
    roots, layers, nodes, tail = MG.fri_commit_sharded(local0, log_m, shift, final, salts_for, make_challenge(), c.rank, c.world, B)
    parity = None
    if c.rank == 0:  # checker: the oracle's commit loop on the same codeword, salts and transcript
        from oracle import oracle as O
        O.set_threads(c.cores)
        _, ref_roots, _ = O.fri_commit(full.cpu().numpy().astype(np.uint64).reshape(-1), shift, final, allsalt.cpu().numpy().reshape(-1), ext=True)
        O.set_threads(1)
        parity = [bytes(r) for r in roots] == [bytes(r) for r in ref_roots]
    del layers, nodes, tail, full
    times = []
    for _ in range(steps):
        c.barrier()
        t0 = time.perf_counter()
        MG.fri_commit_sharded(local0, log_m, shift, final, salts_for, make_challenge(), c.rank, c.world, B)
        torch.cuda.synchronize()
        times.append(c.max_over_ranks([time.perf_counter() - t0])[0])
    ms = min(times) * 1e3
    leaves = sum(sizes) + final
    comps = sum(3 * s - 2 for s in sizes) + (3 * final - 2)
    return {"workload": "FRI commit loop, Ext codeword 2^25 -> 16: fold + salted Merkle tree per layer + host transcript",
            "ms": ms, "ms_median": statistics.median(times) * 1e3, "layers": len(roots), "scaling": "strong", "steps": steps,
            "leaves_hashed": leaves, "g_sha256_compressions_s": comps / (ms * 1e-3) / 1e9,
            "timing": "host wall clock around the loop (it synchronises once per layer for the transcript), best of steps, max over ranks",
            "parity_ok": parity, "parity": "all layer roots against the CPU oracle's commit loop (same salts, same transcript)"}


def run_sharded(c, args):
    out = {}
    for name, fn, steps in (("fourstep_2^27", sharded_fourstep27, 30), ("columns_64x2^22", sharded_columns, 10),
                            ("fold_chain_ext_2^25", sharded_fold_chain, 10), ("fri_commit_ext_2^25", sharded_fri_commit, 3)):
        try:
            out[name] = fn(c, steps)
        except Exception as e:  # noqa: BLE001 - a failed workload must not take the headline down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        c.torch.cuda.empty_cache()
        c.L.bb_release()
        c.barrier()
    return out


# ----------------------------------------------------------------------------------------------- extras (N = 1)
def run_extras(c):
    torch, D, L = c.torch, c.D, c.L
    peak, kind = _peaks()
    out = {}
    # coset LDE 2^20 -> 2^25, blowup 32, shift 7 (src/math/domain.rs:107-123)
    n, size = 1 << 20, 1 << 25
    coeffs = [_seeded_field(c, (n,), SEED + 40 + i) for i in range(4)]
    outs = [torch.empty(size, dtype=torch.int32, device=c.dev) for _ in range(4)]  # 4 x 128 MiB > L2
    it = [0]

    def lde():
        D.coset_fft(coeffs[it[0] % 4], size, 7, out=outs[it[0] % 4])
        it[0] += 1
    ms = c.time_ms(lde, 40, warm=4)
    algo = 4 * n + 4 * size
    out["lde_2^20_to_2^25"] = {"ms": ms, "g_outputs_s": size / (ms * 1e-3) / 1e9,
                               "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                            "frac": algo / (ms * 1e-3) / 1e9 / peak, "peak_source": kind,
                                            "note": "algorithmic 4 n + 4 * 32 n bytes (SURVEY 8d)"}}
    # salted Merkle commit of the 2^25 evaluations (src/fibonacci.rs:340-353)
    salts = torch.randint(0, 256, (size, 16), dtype=torch.uint8, device=c.dev)
    vals = outs[0]
    ms = c.time_ms(lambda: D.merkle_commit(vals, salts, want_root=False), 5, warm=1)
    comps = 3 * size - 2
    algo = 4 * size + 16 * size + 32 * (2 * size - 1) + 32 * (2 * size - 2)
    out["commit_2^25"] = {"ms": ms, "g_sha256_compressions_s": comps / (ms * 1e-3) / 1e9,
                          "roofline": {"bound": "alu pipe (SHA-256: rotates / LOP3 only run there); HBM figure for reference",
                                       "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": algo / (ms * 1e-3) / 1e9 / peak, "peak_source": kind}}
    out["lde_plus_commit_ms"] = out["lde_2^20_to_2^25"]["ms"] + ms
    del salts, outs, coeffs
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    affinity = _pin_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    c = Ctx(rank, world, local_rank)
    D, L, dev = c.D, c.L, c.dev
    from toyni_b200 import ntt as host_ntt
    if not L.bb_device_ok():
        raise SystemExit("bench.py needs a compute-capability 10.0 device (no CPU fallback)")
    n = 1 << LOG_N
    barrier = c.barrier

    # Inputs: NB distinct device-resident vectors (NB * 64 MiB > the 126 MB L2), uniformly random canonical values.
    NB = 4
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + rank)
    bufs = [torch.randint(0, P, (n,), dtype=torch.int32, device=dev, generator=g) for _ in range(NB)]
    L.bb_warmup(LOG_N)

    # Steps are independent transforms of different vectors, so they are issued alternately on NS CUDA streams through
    # the library's stream-ordered API (bb_set_stream): one transform's tail (the last tiles of a pass leave SMs idle)
    # is filled by the other's kernels.  Buffer i % NB always goes to stream i % NS (NB is a multiple of NS), so no
    # buffer is ever touched from two streams.
    NS = int(os.environ.get("TOYNI_BENCH_STREAMS", 2))  # 2 measured best (4: same within noise)
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
    kernel_ms = ms_total / args.steps  # the pass kernels are the whole timed region
    ms_total = c.max_over_ranks([ms_total])[0]
    value = world * n * args.steps / (ms_total * 1e-3) / 1e9
    # one transform alone on one stream (latency, not throughput): reported next to the headline
    it = [0]

    def one():
        D.ntt_(bufs[it[0] % NB], inverse=False)
        it[0] += 1
    single_ms = c.time_ms(one, 200, warm=4)

    # ---- end to end through the reference-facing C ABI on pinned host buffers (every rank, max over ranks).
    # Serial: one caller, one context (what src/ntt.rs:224-236 does) - H2D, kernels, D2H strictly one after the other.
    # Pipelined: NTHR host threads, each with its own context and pinned buffer (ntt_ctx_create once per thread); one
    # thread's H2D overlaps another's D2H on the full-duplex PCIe link and a third's kernels.  Both forms move
    # 8 B/element each way inside the timing.
    import ctypes
    import threading
    host = torch.empty(n, dtype=torch.int64).pin_memory()
    hv = host.numpy().view(np.uint64)
    hv[:] = (np.arange(n, dtype=np.uint64) * np.uint64(7) + np.uint64(3)) % np.uint64(P)
    # PCIe peaks of this box, measured the plain way: one 128 MiB pinned copy per direction, best of 5
    dbuf = torch.empty(n, dtype=torch.int64, device=dev)
    pcie = {}
    for name, (dst, src) in (("h2d", (dbuf, host)), ("d2h", (host, dbuf))):
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        pcie[name] = 8 * n / best / 1e9
    hv[:] = (np.arange(n, dtype=np.uint64) * np.uint64(7) + np.uint64(3)) % np.uint64(P)
    del dbuf
    NTHR = int(os.environ.get("TOYNI_E2E_THREADS", 3))
    e2e_steps = max(NTHR * 2, min(args.steps, 96)) // NTHR * NTHR  # 32 per host thread: pipeline fill and drain are 1/32 of the run
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
    errs = [0] * NTHR

    def pump(k, count):
        torch.cuda.set_device(local_rank)  # the CUDA current device is per host thread
        for _ in range(count):
            rc = L.ntt_run_inplace_rc(ctypes.c_void_p(ctxs[k]), hosts[k].ctypes.data)
            errs[k] = errs[k] or rc

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
    # the same pipeline through the u32 host entry point (bb_ntt_host_u32: NOT the reference's u64 signature, reported
    # beside the drop-in number): 4 B/element each way
    hosts32 = [torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32) for _ in range(NTHR)]
    for h in hosts32:
        h[:] = hv.astype(np.uint32)

    def pump32(k, count):
        torch.cuda.set_device(local_rank)
        for _ in range(count):
            rc = L.bb_ntt_host_u32(ctypes.c_void_p(ctxs[k]), hosts32[k].ctypes.data, 0)
            errs[k] = errs[k] or rc

    for k in range(NTHR):
        pump32(k, 1)
    barrier()
    threads = [threading.Thread(target=pump32, args=(k, e2e_steps // NTHR)) for k in range(NTHR)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    torch.cuda.synchronize()
    e2e_u32_s = c.max_over_ranks([time.perf_counter() - t0])[0]
    del hosts32
    e2e_err = max(errs)  # never raise here: the other ranks would wait in the reduction below forever
    for cx in ctxs:
        L.ntt_ctx_destroy(ctypes.c_void_p(cx))
    e2e_s, e2e_serial_s, e2e_err = c.max_over_ranks([e2e_s, e2e_serial_s, float(e2e_err)])
    e2e_val = world * n * e2e_steps / e2e_s / 1e9
    e2e_serial_val = world * n * e2e_steps / e2e_serial_s / 1e9
    if e2e_err != 0:
        raise SystemExit(f"bench.py: CUDA error {int(e2e_err)} in the end-to-end path")
    del hosts, host
    for b in bufs:
        del b
    bufs = None
    torch.cuda.empty_cache()

    sharded = run_sharded(c, args) if not args.no_sharded else None
    extras = run_extras(c) if (world == 1 and not args.no_extras) else None

    if rank != 0:
        return
    peak, peak_kind = _peaks()
    npass = L.bb_ntt_launches(LOG_N)
    achieved = 8.0 * n / (kernel_ms * 1e-3) / 1e9  # algorithmic bytes: 4 B read + 4 B written per element per transform
    per_gpu_bytes_s = 8 * n * e2e_steps / e2e_s  # per direction, per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "forward BabyBear NTT, n=2^24, one vector per GPU per step (BASELINE configs[1])",
                   "l2": f"inputs rotate over {NB} x 64 MiB device buffers (larger than the 126 MB L2)",
                   "kernels_per_transform": npass, "streams": NS, "single_stream_ms_per_transform": single_ms,
                   "parallelism": f"independent vectors x{world}, no collective; steps alternate over {NS} CUDA streams per GPU",
                   "cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": _traffic(LOG_N), "peak_source": peak_kind,
                     "single_stream_frac": 8.0 * n / (single_ms * 1e-3) / 1e9 / peak,
                     "note": f"algorithmic 8 B/element over the transform's {npass} pass kernels (CUDA events around the timed region, "
                             f"K transforms = {npass}K launches on two streams); the two passes move 16 B/element, see DESIGN.md and profiles/"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * n,
                "steps": e2e_steps, "api": f"ntt_run_inplace (src/ntt.rs:108) on pinned host u64, {NTHR} host threads with "
                                           "one context each (H2D of one overlaps D2H and kernels of the others)",
                "serial_value": e2e_serial_val, "serial_api": "one caller, one context: ntt_cuda as in src/ntt.rs:224-236",
                "u32_host_value": world * n * e2e_steps / e2e_u32_s / 1e9,
                "u32_host_api": "bb_ntt_host_u32: the same pipeline on canonical u32 host values (4 B/element each way); not the "
                                "reference's signature (src/ntt.rs stores u64), reported beside the drop-in number, not instead of it",
                "pcie_peak_gbs": pcie, "pcie_frac": per_gpu_bytes_s / 1e9 / min(pcie.values()),
                "pcie_note": "bytes per direction per GPU / time / the slower of the two measured one-way pinned-copy rates "
                             "(max over ranks); PCIe moves 16 B per element here against 8 B of algorithmic HBM traffic"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if sharded is not None:
        line["sharded"] = sharded
    if extras is not None:
        line["extras"] = extras
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded 2^27 / columns / fold / commit records")
    ap.add_argument("--no-extras", action="store_true", help="skip the LDE / commit records (N = 1)")
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
