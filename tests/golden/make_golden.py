"""Generate tests/golden/golden.npz: small input/output vectors of the hot path.

The reference is Rust and cannot run in this image (no cargo/rustc), so the vectors come from the C oracle
(oracle/toyni_oracle.c, a restatement of the cited reference lines), cross-checked here against the independent
pure-Python restatement (oracle/pyref.py) before they are written.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O, pyref as R  # noqa: E402

out = {}
# NTT / INTT (src/ntt.rs:24-66): the reference's own test input 7i+3 (src/ntt.rs:270-272,326) and random vectors
for log_n in (0, 1, 3, 8, 10, 13):
    n = 1 << log_n
    x = (np.arange(n, dtype=np.uint64) * 7 + 3) % O.P
    y = O.ntt(x)
    if n <= 1024:
        assert list(y) == R.ntt([int(v) for v in x], R.root_of_unity(log_n))
    out[f"ntt_in_{log_n}"] = x
    out[f"ntt_out_{log_n}"] = y
    r = O.random_field(n, seed=1000 + log_n)
    out[f"rnd_in_{log_n}"] = r
    out[f"rnd_ntt_{log_n}"] = O.ntt(r)
    out[f"rnd_intt_{log_n}"] = O.intt(r)
# coset LDE at blowup 32 (src/math/domain.rs:107-123 with the prover's shift 7, src/fibonacci.rs:13-17)
for log_n in (3, 6, 9):
    n = 1 << log_n
    c = O.random_field(n, seed=2000 + log_n)
    e = O.domain_fft(c, 32 * n, 7)
    if n <= 64:
        assert list(e) == R.domain_fft([int(v) for v in c], 32 * n, 7)
    out[f"lde_in_{log_n}"] = c
    out[f"lde_out_{log_n}"] = e
    out[f"lde_back_{log_n}"] = O.domain_ifft(e, 7)
# Ext transforms (src/math/domain.rs:129-151)
c = O.random_field(4 * 16, seed=3000).reshape(16, 4)
out["ext_in"] = c
out["ext_fft"] = O.domain_fft_ext(c, 64, 7)
# FRI folds (src/math/fri.rs)
ev = O.random_field(512, seed=4000)
xs = O.domain_elements(512, 7)
out["fold_in"] = ev
out["fold_out"] = O.fri_fold(ev, xs, 123456789)
assert list(out["fold_out"]) == R.fri_fold([int(v) for v in ev], [int(v) for v in xs], 123456789)
ee = O.random_field(4 * 256, seed=4001).reshape(256, 4)
out["fold_ext_in"] = ee
out["fold_ext_beta"] = np.array([5, 6, 7, 8], dtype=np.uint64)
out["fold_ext_out"] = O.fri_fold_ext(ee, O.domain_elements(256, 7), [5, 6, 7, 8])
assert [list(map(int, r)) for r in out["fold_ext_out"]] == R.fri_fold_ext([[int(v) for v in r] for r in ee],
                                                                          [int(v) for v in O.domain_elements(256, 7)], [5, 6, 7, 8])
# Merkle commits (src/merkle.rs + src/fibonacci.rs:340-363)
v = O.random_field(300, seed=5000)
s = O.random_bytes(16 * 300, seed=5001).reshape(300, 16)
out["merkle_vals"] = v
out["merkle_salts"] = s
out["merkle_root_salted"] = np.frombuffer(O.commit_values(v, s)[1], np.uint8)
out["merkle_root_unsalted"] = np.frombuffer(O.commit_values(v)[1], np.uint8)
leaves = [bytes(s[i]) + int(v[i]).to_bytes(8, "little") for i in range(300)]
assert R.merkle_root(leaves) == O.commit_values(v, s)[1]
# FRI commit loop (src/fibonacci.rs:200-247): 2^10 -> 16
l0 = O.random_field(1 << 10, seed=6000)
nsalt = sum((1 << 10) >> k for k in range(0, 6))
salts = O.random_bytes(16 * nsalt, seed=6001)
layers, roots, betas = O.fri_commit(l0, 7, 16, salts)
out["fri_layer0"] = l0
out["fri_salts"] = salts
out["fri_roots"] = np.frombuffer(b"".join(roots), np.uint8).reshape(-1, 32)
out["fri_betas"] = betas
out["fri_final"] = layers[-1]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz"), **out)
print("wrote golden.npz with", len(out), "arrays")
