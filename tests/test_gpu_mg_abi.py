"""Section 4 of the C ABI (bb_mg_*): the sharded paths driven from one process over the GPUs of the box, called through
ctypes the way a Rust host would bind them — no torch.distributed, no NCCL.  Runs with G = 1 on a single-GPU box and
with every power of two up to the device count otherwise; outputs are compared with the CPU oracle element by element
(src/ntt.rs:24-66, src/math/fri.rs:7-25)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

P = O.P
CORES = os.cpu_count() or 1


def _gs():
    import torch
    n = torch.cuda.device_count()
    return [g for g in (1, 2, 4, 8) if g <= n]


@pytest.fixture(scope="module")
def L():
    from toyni_b200.lib import lib
    return lib()


def _mg(L, G):
    h = C.c_void_p()
    rc = L.bb_mg_init(G, C.byref(h))
    assert rc == 0, L.cuda_get_error_string(rc)
    return h


@pytest.mark.parametrize("log_n", [12, 16, 21, 24])
def test_mg_ntt_host_matches_oracle(L, log_n):
    """bb_mg_ntt_host: natural order in, natural order out, like ntt_run_inplace (src/ntt.rs:108) but over G devices."""
    x = O.random_field(1 << log_n, seed=900 + log_n)
    for G in _gs():
        if (1 << (log_n // 2)) // G < 16:
            continue
        mg = _mg(L, G)
        try:
            v = x.copy()
            assert L.bb_mg_ntt_host(mg, v.ctypes.data, log_n, 0) == 0
            assert np.array_equal(v, O.ntt(x, threads=CORES)), f"forward, G={G}"
            assert L.bb_mg_ntt_host(mg, v.ctypes.data, log_n, 1) == 0
            assert np.array_equal(v, x), f"inverse, G={G}"
        finally:
            L.bb_mg_destroy(mg)


def test_mg_device_resident_fourstep_batch_and_fold_chain(L):
    """Device-resident forms with torch only as the allocator: four-step blocks / result slabs, a column batch, and the Ext
    fold chain on cyclic shards, against the oracle."""
    import torch
    from toyni_b200 import multigpu as MG
    for G in _gs():
        mg = _mg(L, G)
        try:
            # ---- four-step, 2^20
            log_n = 20
            x = O.random_field(1 << log_n, seed=31 + G)
            ref = O.ntt(x, threads=CORES)
            n1, n2 = MG.fourstep_split(log_n, G)
            blocks = [torch.from_numpy(MG.fourstep_scatter(x, r, G).astype(np.int32)).to(f"cuda:{r}") for r in range(G)]
            outs = [torch.empty((n1 // G, n2), dtype=torch.int32, device=f"cuda:{r}") for r in range(G)]
            for r in range(G):
                torch.cuda.synchronize(r)
            pb = (C.c_void_p * G)(*[b.data_ptr() for b in blocks])
            po = (C.c_void_p * G)(*[o.data_ptr() for o in outs])
            assert L.bb_mg_ntt_fourstep(mg, log_n, 0, pb, po) == 0
            assert L.bb_mg_sync(mg) == 0
            got = MG.fourstep_gather([o.cpu().numpy().astype(np.uint64) for o in outs], log_n)
            assert np.array_equal(got, ref), f"four-step G={G}"
            # ---- column batch: 8 columns of 2^14, column j on device j % G
            cols = O.random_field(8 << 14, seed=41).reshape(8, 1 << 14)
            dcols = [torch.from_numpy(np.ascontiguousarray(cols[r::G]).astype(np.int32)).to(f"cuda:{r}") for r in range(G)]
            for r in range(G):
                torch.cuda.synchronize(r)
            pc = (C.c_void_p * G)(*[d.data_ptr() for d in dcols])
            nc = (C.c_size_t * G)(*[d.shape[0] for d in dcols])
            assert L.bb_mg_ntt_batch(mg, 14, 0, pc, nc) == 0
            assert L.bb_mg_sync(mg) == 0
            for r in range(G):
                for i, j in enumerate(range(r, 8, G)):
                    assert np.array_equal(dcols[r][i].cpu().numpy().astype(np.uint64), O.ntt(cols[j]))
            # ---- Ext fold chain 2^16 -> 16 on cyclic shards
            log_m, shift, final = 16, 7, 16
            m = 1 << log_m
            full = O.random_field(4 * m, seed=51).reshape(m, 4)
            betas = np.array([[(5 * k + j + 2) % P for j in range(4)] for k in range(log_m)], dtype=np.uint32)
            shards = [torch.from_numpy(np.ascontiguousarray(full[r::G]).astype(np.int32)).to(f"cuda:{r}") for r in range(G)]
            louts = [torch.zeros((m // G, 4), dtype=torch.int32, device=f"cuda:{r}") for r in range(G)]
            for r in range(G):
                torch.cuda.synchronize(r)
            ps = (C.c_void_p * G)(*[s.data_ptr() for s in shards])
            pl = (C.c_void_p * G)(*[o.data_ptr() for o in louts])
            folds = C.c_size_t(0)
            assert L.bb_mg_fri_chain(mg, log_m, shift, 4, final, betas.ctypes.data, ps, pl, C.byref(folds)) == 0
            assert L.bb_mg_sync(mg) == 0
            xs = O.domain_elements(m, shift)
            cur, off = full, 0
            for k in range(folds.value):
                cur = O.fri_fold_ext(cur, xs, betas[k].astype(np.uint64))
                xs = (xs[: cur.shape[0]] * xs[: cur.shape[0]]) % np.uint64(P)
                loc = cur.shape[0] // G
                for r in range(G):
                    assert np.array_equal(louts[r][off:off + loc].cpu().numpy().astype(np.uint64), cur[r::G]), f"fold {k} G={G} r={r}"
                off += loc
            assert folds.value == log_m - 4 and cur.shape[0] == final
        finally:
            L.bb_mg_destroy(mg)


def test_mg_argument_errors(L):
    h = C.c_void_p()
    assert L.bb_mg_init(3, C.byref(h)) != 0      # not a power of two
    assert L.bb_mg_init(64, C.byref(h)) != 0
    assert L.bb_mg_ngpus(None) == 0
    L.bb_clear_error()
