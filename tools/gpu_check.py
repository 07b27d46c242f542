"""Ad-hoc GPU parity sweep over the whole device API (development aid; the real tests are tests/ -m gpu)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import oracle as O
from toyni_b200 import device as D
from toyni_b200.lib import lib as _lib

ok_all = True
def report(name, ok):
    global ok_all
    ok_all &= bool(ok)
    print(f"{name:60s} {'OK' if ok else 'FAIL'}", flush=True)

def check_ntt(log_n, inverse=False):
    n = 1 << log_n
    x = O.random_field(n, seed=log_n + 7)
    t = D.to_device(x); D.ntt_(t, inverse)
    ref = O.intt(x, threads=8) if inverse else O.ntt(x, threads=8)
    return np.array_equal(D.to_host(t), ref)

maxlog = int(sys.argv[1]) if len(sys.argv) > 1 else 22
for log_n in range(0, maxlog + 1):
    for inv in (False, True):
        report(f"ntt log_n={log_n} inv={int(inv)} passes={_lib().bb_ntt_launches(log_n)}", check_ntt(log_n, inv))
# batched
for log_n, batch in [(0, 5), (3, 7), (8, 100), (10, 33), (12, 64), (14, 9), (16, 16), (18, 3)]:
    n = 1 << log_n
    x = O.random_field(n * batch, seed=99 + log_n).reshape(batch, n)
    for inv in (False, True):
        t = D.to_device(x); D.ntt_batch_(t, inv)
        ref = np.stack([(O.intt if inv else O.ntt)(x[b]) for b in range(batch)])
        report(f"batch ntt log_n={log_n} batch={batch} inv={int(inv)}", np.array_equal(D.to_host(t), ref))
# ext
for log_n in [0, 1, 3, 6, 8, 9, 12, 15, 17, 20]:
    n = 1 << log_n
    x = O.random_field(4 * n, seed=5 + log_n).reshape(n, 4)
    t = D.to_device(x); D.ntt_ext_(t, False)
    ref = O.domain_fft_ext(x, n, 1)
    report(f"ext ntt log_n={log_n}", np.array_equal(D.to_host(t), ref))
    D.ntt_ext_(t, True)
    report(f"ext intt roundtrip log_n={log_n}", np.array_equal(D.to_host(t), x))
# coset fft / ifft
for n_c, log_size, shift in [(8, 3, 7), (3, 3, 7), (0, 4, 7), (8, 8, 7), (100, 7, 7), (300, 8, 5), (1 << 10, 15, 7), (1 << 12, 17, 7), (5000, 18, 3), (1 << 15, 20, 7), (1<<9, 9, 1)]:
    c = O.random_field(max(n_c, 1), seed=n_c + log_size)[:n_c]
    size = 1 << log_size
    out = D.coset_fft(D.to_device(c) if n_c else torch.zeros(0, dtype=torch.int32, device="cuda"), size, shift)
    ref = O.domain_fft(c, size, shift)
    report(f"coset fft n_coeffs={n_c} size=2^{log_size} shift={shift}", np.array_equal(D.to_host(out), ref))
    back = D.coset_ifft_(out.clone(), shift)
    report(f"coset ifft size=2^{log_size} shift={shift}", np.array_equal(D.to_host(back), O.domain_ifft(ref, shift)))
for n_c, log_size in [(3, 3), (64, 11), (1 << 10, 15)]:
    c = O.random_field(4 * n_c, seed=3).reshape(n_c, 4)
    size = 1 << log_size
    out = D.coset_fft(D.to_device(c), size, 7)
    ref = O.domain_fft_ext(c, size, 7)
    report(f"coset fft_ext n_coeffs={n_c} size=2^{log_size}", np.array_equal(D.to_host(out), ref))
    back = D.coset_ifft_(out.clone(), 7)
    report(f"coset ifft_ext size=2^{log_size}", np.array_equal(D.to_host(back), O.domain_ifft_ext(ref, 7)))
# folds
for log_m in [1, 2, 5, 10, 16]:
    m = 1 << log_m
    ev = O.random_field(m, seed=m); xs = O.domain_elements(m, 7)
    report(f"fri_fold m=2^{log_m}", np.array_equal(D.to_host(D.fri_fold(D.to_device(ev), 7, 123456789)), O.fri_fold(ev, xs, 123456789)))
    report(f"fri_fold_xs m=2^{log_m}", np.array_equal(D.to_host(D.fri_fold_xs(D.to_device(ev), D.to_device(xs[: m // 2]), 987654321)), O.fri_fold(ev, xs, 987654321)))
    ee = O.random_field(4 * m, seed=m + 1).reshape(m, 4); beta = [5, 6, 7, 2013265920]
    report(f"fri_fold_ext m=2^{log_m}", np.array_equal(D.to_host(D.fri_fold(D.to_device(ee), 7, beta)), O.fri_fold_ext(ee, xs, beta)))
    report(f"fri_fold_ext_xs m=2^{log_m}", np.array_equal(D.to_host(D.fri_fold_xs(D.to_device(ee), D.to_device(xs[: m // 2]), beta)), O.fri_fold_ext(ee, xs, beta)))
    if m >= 8:
        G = 4
        parts = [D.to_host(D.fri_fold_shard(D.to_device(np.ascontiguousarray(ee[r::G])), log_m, 7, beta, G, r)) for r in range(G)]
        full = O.fri_fold_ext(ee, xs, beta)
        report(f"fri_fold shard x{G} m=2^{log_m}", all(np.array_equal(parts[r], full[r::G]) for r in range(G)))
# merkle
for n in [1, 2, 3, 4, 5, 100, 1 << 10, (1 << 12) + 1]:
    v = O.random_field(n, seed=n); salts = O.random_bytes(16 * n, seed=n + 1).reshape(n, 16)
    nodes, root = D.merkle_commit(D.to_device(v), torch.from_numpy(salts).cuda())
    rn, rr = O.commit_values(v, salts)
    report(f"merkle salted n={n}", root == rr and np.array_equal(nodes.cpu().numpy(), rn))
    nodes, root = D.merkle_commit(D.to_device(v))
    rn, rr = O.commit_values(v)
    report(f"merkle unsalted n={n}", root == rr and np.array_equal(nodes.cpu().numpy(), rn))
    v4 = O.random_field(4 * n, seed=n + 2).reshape(n, 4)
    nodes, root = D.merkle_commit(D.to_device(v4), torch.from_numpy(salts).cuda())
    rn, rr = O.commit_values(v4, salts, limbs=4)
    report(f"merkle ext salted n={n}", root == rr and np.array_equal(nodes.cpu().numpy(), rn))
# fri commit loop with transcript
import hashlib
class T:
    def __init__(s): s.state = b"toyni-stark-v1"
    def absorb(s, d): s.state += d
    def squeeze(s):
        h = hashlib.sha256(s.state).digest(); s.state = h
        return int.from_bytes(h[:8], "little") % O.P
for ext in (False, True):
    n, final = 1 << 12, 16
    l0 = O.random_field(n * (4 if ext else 1), seed=77)
    if ext: l0 = l0.reshape(n, 4)
    nsalt = sum(n >> k for k in range(0, 8))
    salts = O.random_bytes(16 * nsalt, seed=5)
    layers_ref, roots_ref, betas_ref = O.fri_commit(l0, 7, final, salts, ext=ext)
    t = T()
    def challenge(root, layer):
        t.absorb(root)
        return [t.squeeze() for _ in range(4)] if ext else t.squeeze()
    layers, nodes, roots = D.fri_commit(D.to_device(l0), 7, final, torch.from_numpy(salts).cuda(), challenge=challenge)
    ok = roots == roots_ref and all(np.array_equal(D.to_host(a), b) for a, b in zip(layers, layers_ref))
    report(f"fri commit loop ext={ext} (layers, roots vs oracle)", ok)
print("ALL OK" if ok_all else "FAILURES")
sys.exit(0 if ok_all else 1)
