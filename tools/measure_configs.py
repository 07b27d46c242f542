"""Measure BASELINE.json configs 2-4 on one B200 (CUDA events, device-resident data) and print one JSON document.
   python tools/measure_configs.py > profiles/configs_r2.json"""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P, lib

HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0x70796E69)
rnd = lambda *shape: torch.randint(0, P, shape, dtype=torch.int32, device=dev, generator=g)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return {"best_us": ts[0], "median_us": ts[len(ts) // 2]}

out = {"hbm_peak_gbs_measured": HBM, "device": torch.cuda.get_device_name(0)}
# ---- config 2: batched forward / inverse NTT sweep, B = max(1, 2^24 / n) columns, > L2 by rotating 4 buffers
sweep = []
for log_n in (12, 14, 16, 18, 20, 22, 24):
    n = 1 << log_n; B = max(1, (1 << 24) // n)
    bufs = [rnd(B, n) for _ in range(4)]
    it = [0]
    def fwd(): it[0] += 1; D.ntt_batch_(bufs[it[0] % 4], False)
    def inv(): it[0] += 1; D.ntt_batch_(bufs[it[0] % 4], True)
    tf, ti = timeit(fwd), timeit(inv)
    sweep.append({"log_n": log_n, "batch": B, "forward": tf, "inverse": ti,
                  "forward_gelem_s": B * n / tf["median_us"] / 1e3, "inverse_gelem_s": B * n / ti["median_us"] / 1e3,
                  "forward_frac_of_hbm_roofline": 8.0 * B * n / (tf["median_us"] * 1e-6) / 1e9 / HBM,
                  "kernels_per_transform": lib().bb_ntt_launches(log_n)})
    del bufs
out["config2_ntt_sweep"] = sweep
# ---- config 3: coset LDE 2^20 -> 2^25 (blowup 32, shift 7) + salted Merkle commit
coeffs = rnd(1 << 20); N = 1 << 25
evals = torch.empty(N, dtype=torch.int32, device=dev)
salts = torch.randint(0, 256, (N, 16), dtype=torch.uint8, device=dev, generator=g)
nodes = torch.empty((D.merkle_node_count(N), 32), dtype=torch.uint8, device=dev)
t_lde = timeit(lambda: D.coset_fft(coeffs, N, 7, out=evals))
t_commit = timeit(lambda: D.merkle_commit(evals, salts, nodes=nodes, want_root=False), reps=5, warm=2)
def both():
    D.coset_fft(coeffs, N, 7, out=evals); D.merkle_commit(evals, salts, nodes=nodes, want_root=False)
t_both = timeit(both, reps=5, warm=1)
hashes = N + 2 * (N - 1)
out["config3_lde_commit"] = {"lde": t_lde, "commit": t_commit, "lde_plus_commit": t_both,
    "lde_algorithmic_bytes": 132 * (1 << 20), "lde_frac_of_hbm_roofline": 132.0 * (1 << 20) / (t_lde["median_us"] * 1e-6) / 1e9 / HBM,
    "lde_g_output_elem_s": N / t_lde["median_us"] / 1e3,
    "commit_sha256_compressions": hashes, "commit_gcompressions_s": hashes / t_commit["median_us"] / 1e3}
del evals, nodes
# ---- config 4: Ext fold chain 2^25 -> 16 (21 folds), fold-only with betas up front, and the real commit loop
l0 = rnd(N, 4)
betas = np.random.default_rng(1).integers(0, P, 4 * 21).astype(np.uint64)
t_fold = timeit(lambda: D.fri_commit(l0, 7, 16, betas=betas, hash_layers=False), reps=5, warm=2)
chain_bytes = sum(24 * (N >> k) for k in range(21))
one = timeit(lambda: D.fri_fold(l0, 7, [1, 2, 3, 4]), reps=10)
nsalt = sum(N >> k for k in range(21))
fsalts = torch.randint(0, 256, (nsalt, 16), dtype=torch.uint8, device=dev, generator=g)
import hashlib
class T:
    def __init__(s): s.state = b"toyni-stark-v1"
    def ch(s, root, layer):
        s.state += root
        o = []
        for _ in range(4):
            h = hashlib.sha256(s.state).digest(); s.state = h; o.append(int.from_bytes(h[:8], "little") % P)
        return o
def real():
    t = T(); D.fri_commit(l0, 7, 16, fsalts, challenge=t.ch)
t_real = timeit(real, reps=3, warm=1)
out["config4_fri_chain_ext"] = {"fold_only_chain": t_fold, "chain_algorithmic_bytes": chain_bytes,
    "chain_frac_of_hbm_roofline": chain_bytes / (t_fold["median_us"] * 1e-6) / 1e9 / HBM,
    "first_fold_2^25": one, "first_fold_frac_of_hbm_roofline": 24.0 * N / (one["median_us"] * 1e-6) / 1e9 / HBM,
    "commit_loop_with_transcript": t_real, "folds": 21}
print(json.dumps(out, indent=1))
