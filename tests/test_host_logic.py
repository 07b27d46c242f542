"""Host-side logic that needs no GPU: Merkle openings over GPU-format node arrays, the sharding index maps of the
multi-GPU layer (checked with a 2-rank gloo group on CPU), and the product's refusal to run without a device."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from toyni_b200 import merkle as M
from toyni_b200 import multigpu as MG
from toyni_b200.lib import ToyniCudaError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_salted_tree_openings_verify_against_oracle_nodes():
    n = 37
    vals = O.random_field(n, seed=1)
    salts = O.random_bytes(16 * n, seed=2).reshape(n, 16)
    nodes, root = O.commit_values(vals, salts)
    tree = M.SaltedTree(n, nodes, root, salts)  # same node layout the GPU commit produces
    for i in (0, 1, 17, 35, 36):
        proof = tree.get_proof(i)
        leaf = salts[i].tobytes() + int(vals[i]).to_bytes(8, "little")
        assert M.verify_merkle_proof(leaf, proof, root)
        path, pos = O.merkle_open(nodes, n, i)
        assert [p.tobytes() for p in path] == proof.path and [bool(b) for b in pos] == proof.position
        assert not M.verify_merkle_proof(leaf[:-1] + b"\x00", proof, root) or leaf[-1] == 0
    assert tree.get_proof(n) is None


def test_fourstep_index_maps_cpu_simulation():
    """The sharded 2^k four-step NTT as pure index arithmetic (numpy, oracle transforms per shard)."""
    for log_n, G in ((6, 2), (8, 4), (10, 8)):
        n = 1 << log_n
        x = O.random_field(n, seed=log_n)
        got = MG.fourstep_reference_simulation(x, G, ntt=O.ntt, mul=lambda a, b: (a.astype(object) * b.astype(object) % O.P).astype(np.uint64))
        assert np.array_equal(got, O.ntt(x)), (log_n, G)


def test_cyclic_fold_layout_is_closed():
    for G in (1, 2, 4, 8):
        m = 1 << 10
        ev = O.random_field(m, seed=G)
        xs = O.domain_elements(m, 7)
        full = O.fri_fold(ev, xs, 31337)
        for r in range(G):
            local = ev[r::G]
            half = local.size // 2
            # pair (t, t+half) of the shard is the global pair (i, i + m/2)
            idx = r + G * np.arange(half)
            assert np.array_equal(local[:half], ev[idx]) and np.array_equal(local[half:], ev[idx + m // 2])
            assert np.array_equal(full[r::G], O.fri_fold(np.concatenate([ev[idx], ev[idx + m // 2]]), xs[idx], 31337))


_GLOO_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as O
from toyni_b200 import multigpu as MG
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
log_n = 10
n = 1 << log_n
x = O.random_field(n, seed=7)
# every rank owns its column block of the n1 x n2 matrix; local transforms by the oracle (no GPU here)
out = MG.fourstep_ntt_distributed(x, rank, world, backend_ntt=lambda a: O.ntt(a), device="cpu")
ref = O.ntt(x)
n1 = MG.fourstep_split(log_n, world)[0]
k1 = np.arange(rank * n1 // world, (rank + 1) * n1 // world)
mine = ref.reshape(n // n1, n1)[:, k1].T   # out[k1_local][k2] = X[k1 + n1*k2]
assert np.array_equal(out, mine), "rank %d mismatch" % rank
dist.barrier()
if rank == 0: print("gloo fourstep ok")
dist.destroy_process_group()
"""


def test_fourstep_all_to_all_with_gloo_world_size_2(tmp_path):
    script = tmp_path / "gloo_fourstep.py"
    script.write_text(_GLOO_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29511")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gloo fourstep ok" in out.stdout


_GLOO_COMMIT_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from oracle import oracle as O
from toyni_b200 import multigpu as MG
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
log_m, final_size, shift = 12, 16, 7
m = 1 << log_m
layer0 = O.random_field(m, seed=11)
nsalt, mm = 0, m
while mm > final_size:
    nsalt += mm; mm //= 2
salts = O.random_bytes(16 * nsalt, seed=12).reshape(-1, 16)
ref_layers, ref_roots, ref_betas = O.fri_commit(layer0, shift, final_size, salts.reshape(-1))

class OracleBackend:                       # the same loop with the CPU oracle's primitives instead of the CUDA kernels
    def __init__(self): self.t = O.FiatShamirTranscript()
    def fold(self, local, log_mk, x0, beta, world, rank):
        a = local.numpy().astype(np.uint64); half = a.size // 2
        idx = rank + world * np.arange(half)
        w = O.root_of_unity(log_mk)
        xs = np.array([x0 * pow(w, int(i), O.P) % O.P for i in idx], np.uint64)
        return torch.from_numpy(O.fri_fold(a, xs, beta).astype(np.int64))
    def commit(self, vals, s):
        nodes, root = O.commit_values(vals.numpy().astype(np.uint64), None if s is None else s.numpy())
        return nodes, root
    def finish(self, full, x0, final_size, s, challenge):
        layers, roots, _ = O.fri_commit(full.numpy().astype(np.uint64), x0, final_size,
                                        np.zeros(0, np.uint8) if s is None else s.numpy(), transcript=self.t)
        return layers, None, roots
    def bytes_tensor(self, b): return torch.frombuffer(bytearray(b), dtype=torch.uint8)

B = OracleBackend()
offs, o, mm = [], 0, m
while mm > final_size:
    offs.append(o); o += mm; mm //= 2
def salts_for(k, lo, hi): return torch.from_numpy(salts[offs[k] + lo: offs[k] + hi].copy())
def challenge(root, k):
    B.t.absorb(root); return B.t.squeeze_challenge()
local0 = torch.from_numpy(layer0[rank::world].astype(np.int64))
roots, layers, nodes, tail = MG.fri_commit_sharded(local0, log_m, shift, final_size, salts_for, challenge, rank, world, B,
                                                   gather_below=1 << 8)
assert roots == ref_roots, "rank %d: roots differ" % rank
assert np.array_equal(np.asarray(tail[-1], np.uint64), ref_layers[-1])
for k, lay in enumerate(layers):           # every sharded layer is the cyclic shard of the reference layer
    assert np.array_equal(lay.numpy().astype(np.uint64), ref_layers[k][rank::world])
dist.barrier()
if rank == 0: print("gloo sharded commit ok", len(roots), len(layers))
dist.destroy_process_group()
"""


def test_sharded_fri_commit_with_gloo_world_size_2(tmp_path):
    """The multi-GPU FRI commit loop (cyclic fold shards, cyclic -> block exchange per layer, per-rank subtrees, top of
    the tree on every rank) with a 2-rank gloo group and the oracle's primitives: same roots and layers as the
    single-process loop of src/fibonacci.rs:200-247."""
    script = tmp_path / "gloo_commit.py"
    script.write_text(_GLOO_COMMIT_SCRIPT.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gloo sharded commit ok" in out.stdout


def test_product_path_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from toyni_b200 import domain, ntt
    with pytest.raises(ToyniCudaError):
        ntt.ntt_cuda(np.arange(8, dtype=np.uint64))
    with pytest.raises(ToyniCudaError):
        domain.BabyBearDomain(8).fft(np.arange(8, dtype=np.uint64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "toyni_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "__init__.py" or "import oracle" not in text
                assert "from oracle" not in text and "import oracle" not in text and "toyni_oracle" not in text, f


def test_fast_shoup_companion_is_exact(tmp_path):
    """bb_field.cuh::shoup_companion_fast (device-side twiddle generation of the warp-private pass kernels, no 64-bit
    division) must equal floor(w 2^32 / p) for every canonical w: both ends of the range and 20 million random values."""
    import subprocess
    src = tmp_path / "t.cpp"
    src.write_text(r"""
#include "bb_field.cuh"
#include <cstdio>
#include <random>
int main() {
    std::mt19937_64 g(1);
    unsigned long long bad = 0;
    auto chk = [&](uint32_t w) { if (bb::shoup_companion_fast(w) != bb::shoup_companion(w)) bad++; };
    for (uint32_t w = 0; w < 1000000; w++) chk(w);
    for (uint32_t w = bb::P - 1000000; w < bb::P; w++) chk(w);
    for (int i = 0; i < 20000000; i++) chk((uint32_t)(g() % bb::P));
    printf("%llu\n", bad);
    return bad != 0;
}
""")
    exe = tmp_path / "t.bin"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "toyni_b200", "csrc"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)]).decode().strip() == "0"


def test_merkle_top_is_the_reference_node_hash():
    """multigpu.merkle_top finishes a sharded tree from its subtree roots exactly like MerkleTree::build_tree
    (src/merkle.rs:25-48): for 2^k leaves split into G blocks, root == top(roots of the blocks)."""
    vals = O.random_field(64, seed=4)
    salts = O.random_bytes(16 * 64, seed=5).reshape(64, 16)
    _, root = O.commit_values(vals, salts)
    for G in (1, 2, 4, 8):
        c = 64 // G
        subs = [O.commit_values(vals[r * c:(r + 1) * c], salts[r * c:(r + 1) * c])[1] for r in range(G)]
        assert MG.merkle_top(subs) == root


def test_pass_planner_without_a_device():
    """The pass planner is host code behind bb_ntt_get_plan: two TMA-staged 4096-point passes at 2^24, the measured uneven splits
    at 2^25..2^27, two passes up to 2^16, every plan multiplying out to n."""
    import ctypes as C
    from toyni_b200.lib import lib
    L = lib()
    rows, cols = (C.c_int * 3)(), (C.c_int * 3)()
    want = {24: [12, 12], 25: [8, 8, 9], 26: [10, 8, 8], 27: [8, 10, 9], 16: [8, 8], 12: [6, 6]}
    for log_n in range(1, 28):
        npass = L.bb_ntt_get_plan(log_n, rows, cols)
        assert 1 <= npass <= 3 and sum(rows[i] for i in range(npass)) == log_n
        if log_n in want:
            assert [rows[i] for i in range(npass)] == want[log_n]


def test_tma_kernel_index_model_is_a_dft():
    """tools/v7_model.py restates which shared-memory row, twiddle-table entry and inter-pass factor every lane of the
    TMA-staged two-pass kernel (ntt_pass_v7.cuh) uses, with the radix as a parameter; at radix 4 (n = 2^12) the model must
    be the DFT of src/ntt.rs:24-53 and its inverse, which pins the chunk layout, the swizzle and the A * beta split."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("v7_model", os.path.join(ROOT, "tools", "v7_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    g, n = 2, 4096
    x = O.random_field(n, seed=77)
    got = np.array([int(v) for v in m.two_pass_ntt([int(v) for v in x], g)], dtype=np.uint64)
    assert np.array_equal(got, O.ntt(x))
    back = np.array([int(v) for v in m.two_pass_ntt([int(v) for v in got], g, inverse=True)], dtype=np.uint64)
    assert np.array_equal(back, x)


def test_lde_expansion_pass_index_model():
    """tools/lde_model.py restates the index logic of lde_expand_kernel (two radix-16 DIT rounds whose butterfly twiddles
    carry the coset factors, table of w_8192 powers below 4096 only, rows >= 256 folded in with w_32^c) on the CPU: one
    column of Z[j][k1] against the direct sum, for 2^20 coefficients (256 rows) and the masked trace polynomial (257)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("lde_model", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "lde_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(7)
    for nrows in (256, 257, 3):
        m.check(nrows, rng, extra=6)
