"""Parity of the CUDA NTT / coset LDE path against the CPU oracle, through the C ABI.  Bit-exact: the test is
array equality of canonical values, no tolerance."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

P = O.P
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def D():
    import torch
    from toyni_b200 import device
    torch.cuda.set_device(0)
    return device


# ------------------------------------------------------------------ the drop-in boundary (src/ntt.rs:96-110)
def test_cuda_available_and_native_library_loaded():
    from toyni_b200 import ntt
    assert ntt.cuda_available()
    assert any("libntt_cuda.so" in line for line in open("/proc/self/maps")), "native CUDA library not loaded"


def test_cuda_ntt_vs_cpu_reference_test():
    """src/ntt.rs:263-287 (test_cuda_ntt_vs_cpu): n = 256, x[i] = 7i + 3, GPU == CPU element-wise."""
    from toyni_b200 import ntt
    x = (np.arange(256, dtype=np.uint64) * 7 + 3) % P
    gpu = x.copy()
    ntt.ntt_cuda(gpu)
    cpu = O.ntt(x)
    for i in range(256):
        assert gpu[i] == cpu[i], f"Mismatch at index {i}: CPU={cpu[i]}, GPU={gpu[i]}"


def test_cuda_intt_roundtrip_reference_test():
    """src/ntt.rs:289-310 (test_cuda_intt_roundtrip)."""
    from toyni_b200 import ntt
    x = (np.arange(256, dtype=np.uint64) * 7 + 3) % P
    v = x.copy()
    ntt.ntt_cuda(v)
    ntt.intt_cuda(v)
    assert np.array_equal(v, x)


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8, 9, 12, 13, 16, 17, 20])
def test_host_pointer_ntt_and_intt_match_oracle(log_n):
    from toyni_b200 import ntt
    x = O.random_field(1 << log_n, seed=log_n)
    v = x.copy()
    ntt.ntt_cuda(v)
    assert np.array_equal(v, O.ntt(x, threads=4))
    w = x.copy()
    ntt.intt_cuda(w)
    assert np.array_equal(w, O.intt(x, threads=4))


def test_bad_sizes_assert_like_the_reference():
    from toyni_b200 import ntt
    with pytest.raises(AssertionError):  # src/ntt.rs:229
        ntt.ntt_cuda(np.zeros(12, dtype=np.uint64))
    with pytest.raises(TypeError):
        ntt.ntt_cuda(np.zeros(8, dtype=np.uint32))


def test_cuda_buffer_roundtrip():
    """CudaBuffer (src/ntt.rs:153-212) over cuda_malloc / cuda_copy_* / cuda_free."""
    from toyni_b200 import ntt
    buf = ntt.CudaBuffer(1000)
    src = O.random_field(1000)
    dst = np.zeros(1000, dtype=np.uint64)
    buf.copy_from_host(src)
    buf.copy_to_host(dst)
    assert np.array_equal(src, dst)
    with pytest.raises(AssertionError):
        buf.copy_from_host(np.zeros(5, dtype=np.uint64))


def test_against_the_reference_cuda_implementation():
    """The reference's own cuda/ntt_kernel.cu (compiled unmodified into oracle/_ref) on the same inputs: pins the
    oracle AND the new kernels against real reference output."""
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libntt_cuda_ref.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    from toyni_b200 import ntt
    ref = ctypes.CDLL(ref_so, mode=ctypes.RTLD_LOCAL)
    ref.ntt_ctx_create.restype = ctypes.c_void_p
    ref.ntt_ctx_create.argtypes = [ctypes.c_uint32]
    ref.ntt_run_inplace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    ref.intt_run_inplace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    for log_n in (1, 8, 12, 16, 20):
        n = 1 << log_n
        x = O.random_field(n, seed=40 + log_n)
        ctx = ref.ntt_ctx_create(n)
        a = x.copy()
        ref.ntt_run_inplace(ctx, a.ctypes.data)
        assert np.array_equal(a, O.ntt(x, threads=4)), "oracle != reference CUDA"
        b = x.copy()
        ntt.ntt_cuda(b)
        assert np.array_equal(a, b), "new kernels != reference CUDA"
        ref.intt_run_inplace(ctx, a.ctypes.data)
        ntt.intt_cuda(b)
        assert np.array_equal(a, x) and np.array_equal(b, x)


# ------------------------------------------------------------------ device-resident transforms
@pytest.mark.parametrize("log_n", list(range(0, 19)) + [21, 22])
@pytest.mark.parametrize("inverse", [False, True])
def test_device_ntt_every_size(D, log_n, inverse):
    x = O.random_field(1 << log_n, seed=100 + log_n)
    t = D.to_device(x)
    D.ntt_(t, inverse)
    ref = O.intt(x, threads=4) if inverse else O.ntt(x, threads=4)
    assert np.array_equal(D.to_host(t), ref)


def test_golden_vectors(D, golden):
    for log_n in (0, 1, 3, 8, 10, 13):
        t = D.to_device(golden[f"ntt_in_{log_n}"])
        assert np.array_equal(D.to_host(D.ntt_(t)), golden[f"ntt_out_{log_n}"])
        t = D.to_device(golden[f"rnd_in_{log_n}"])
        assert np.array_equal(D.to_host(D.ntt_(t, True)), golden[f"rnd_intt_{log_n}"])
    for log_n in (3, 6, 9):
        ev = D.coset_fft(D.to_device(golden[f"lde_in_{log_n}"]), 32 << log_n, 7)
        assert np.array_equal(D.to_host(ev), golden[f"lde_out_{log_n}"])
        assert np.array_equal(D.to_host(D.coset_ifft_(ev, 7)), golden[f"lde_back_{log_n}"])
    ev = D.coset_fft(D.to_device(golden["ext_in"]), 64, 7)
    assert np.array_equal(D.to_host(ev), golden["ext_fft"])


def test_scalar_and_vector_kernels_agree(D):
    """The same transform forced through the scalar pass kernel must equal the vectorised one."""
    from toyni_b200.lib import lib
    L = lib()
    x = O.random_field(1 << 15, seed=9)
    a = D.to_host(D.ntt_(D.to_device(x)))
    lr, lc = (ctypes.c_int * 3)(8, 7, 0), (ctypes.c_int * 3)(0, 0, 0)  # one column per tile: scalar kernel
    assert L.bb_ntt_set_plan(15, 2, lr, lc) == 0
    try:
        b = D.to_host(D.ntt_(D.to_device(x)))
    finally:
        L.bb_ntt_set_plan(15, 0, lr, lc)
    assert np.array_equal(a, b) and np.array_equal(a, O.ntt(x))


@pytest.mark.parametrize("log_n,batch", [(0, 5), (3, 7), (8, 100), (10, 33), (12, 64), (14, 9), (16, 16)])
def test_batched_ntt(D, log_n, batch):
    n = 1 << log_n
    x = O.random_field(n * batch, seed=log_n).reshape(batch, n)
    for inv in (False, True):
        t = D.to_device(x)
        D.ntt_batch_(t, inv)
        ref = np.stack([(O.intt if inv else O.ntt)(x[b]) for b in range(batch)])
        assert np.array_equal(D.to_host(t), ref)


@pytest.mark.parametrize("log_n", [0, 1, 3, 6, 8, 9, 12, 15, 17])
def test_ext_ntt_is_four_base_ntts(D, log_n):
    """transform_ext, src/math/domain.rs:140-151."""
    n = 1 << log_n
    x = O.random_field(4 * n, seed=log_n).reshape(n, 4)
    t = D.to_device(x)
    D.ntt_ext_(t)
    assert np.array_equal(D.to_host(t), O.domain_fft_ext(x, n, 1))
    D.ntt_ext_(t, True)
    assert np.array_equal(D.to_host(t), x)


# ------------------------------------------------------------------ coset LDE (BabyBearDomain::fft / ifft)
@pytest.mark.parametrize("n_c,log_size,shift", [(8, 3, 7), (3, 3, 7), (0, 4, 7), (8, 8, 7), (100, 7, 7), (300, 8, 5),
                                                (1 << 10, 15, 7), (1 << 12, 17, 7), (5000, 18, 3), (1 << 9, 9, 1)])
def test_coset_fft_zero_pads_truncates_and_shifts(D, n_c, log_size, shift):
    """src/math/domain.rs:107-123: resize() zero-pads AND truncates (n_c > size); shift 1 skips the coset."""
    import torch
    size = 1 << log_size
    c = O.random_field(max(n_c, 1), seed=n_c + log_size)[:n_c]
    dev_c = D.to_device(c) if n_c else torch.zeros(0, dtype=torch.int32, device="cuda")
    ev = D.coset_fft(dev_c, size, shift)
    ref = O.domain_fft(c, size, shift)
    assert np.array_equal(D.to_host(ev), ref)
    assert np.array_equal(D.to_host(D.coset_ifft_(ev, shift)), O.domain_ifft(ref, shift))


def test_coset_fft_equals_horner_at_every_point(D):
    """src/math/domain.rs:221-242 (test_coset_evaluations_correct), on the GPU path."""
    from oracle import pyref as R
    ev = D.to_host(D.coset_fft(D.to_device(np.array([1, 2, 3], dtype=np.uint64)), 8, 7))
    for i, x in enumerate(O.domain_elements(8, 7)):
        assert int(ev[i]) == R.horner([1, 2, 3], int(x))


def test_domain_mirror_matches_reference_call_sites():
    """BabyBearDomain mirror (host u64 in/out): fft, ifft, fft_ext, ifft_ext, elements."""
    from toyni_b200.domain import BabyBearDomain
    dom = BabyBearDomain.new(1 << 11).with_gpu(True).get_coset(7)
    c = O.random_field(64, seed=1)
    ev = dom.fft(c)
    assert np.array_equal(ev, O.domain_fft(c, 1 << 11, 7))
    assert np.array_equal(dom.ifft(ev), O.domain_ifft(ev, 7))
    assert np.array_equal(dom.elements(), O.domain_elements(1 << 11, 7))
    ce = O.random_field(4 * 64, seed=2).reshape(64, 4)
    ee = dom.fft_ext(ce)
    assert np.array_equal(ee, O.domain_fft_ext(ce, 1 << 11, 7))
    assert np.array_equal(dom.ifft_ext(ee), O.domain_ifft_ext(ee, 7))
    with pytest.raises(AssertionError):  # src/math/domain.rs:86
        dom.ifft(ev[:-1])


# ------------------------------------------------------------------ full-size properties (no oracle run needed)
def test_roundtrip_and_linearity_at_2_24(D):
    import torch
    n = 1 << 24
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    x = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g)
    y = torch.randint(0, P, (n,), dtype=torch.int32, device="cuda", generator=g)
    fx, fy = D.ntt_(x.clone()), D.ntt_(y.clone())
    s = ((x.to(torch.int64) + y.to(torch.int64)) % P).to(torch.int32)
    fs = D.ntt_(s)
    assert torch.equal(fs, ((fx.to(torch.int64) + fy.to(torch.int64)) % P).to(torch.int32))  # linearity
    assert torch.equal(D.ntt_(fx, True), x)                                                     # inverse
    assert int(fx.max()) < P and int(fx.min()) >= 0                                             # canonical
    # X[0] = sum of inputs
    fx2 = D.ntt_(x.clone())
    assert int(fx2[0]) == int(x.to(torch.int64).sum() % P)


def test_lde_2_20_to_2_25_roundtrip(D):
    """Config 3's transform at full size: coset LDE, then coset IFFT gives the zero-padded coefficients back."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    c = torch.randint(0, P, (1 << 20,), dtype=torch.int32, device="cuda", generator=g)
    ev = D.coset_fft(c, 1 << 25, 7)
    back = D.coset_ifft_(ev, 7)
    assert torch.equal(back[: 1 << 20], c)
    assert int(back[1 << 20:].abs().max()) == 0


def test_ntt_2_24_against_oracle_checksum(D):
    """One full-size oracle run (multi-threaded port, identical arithmetic): whole-vector equality at 2^24."""
    x = O.random_field(1 << 24, seed=24)
    ref = O.ntt(x, threads=os.cpu_count() or 1)
    got = D.to_host(D.ntt_(D.to_device(x)))
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("log_n", [17, 18, 19, 20])
def test_large_batches_of_mid_size_vectors_use_two_long_passes(D, log_n):
    """Batches of 2^17..2^20-point vectors with at least 2^23 words in total are planned as two passes of 256..1024 rows
    (single vectors of the same length as three passes): both plans must give the oracle's transform."""
    from toyni_b200.lib import lib
    n = 1 << log_n
    batch = (1 << 23) // n
    x = O.random_field(batch * n, seed=300 + log_n).reshape(batch, n)
    for inv in (False, True):
        got = D.to_host(D.ntt_batch_(D.to_device(x), inv))
        single = D.to_host(D.ntt_(D.to_device(x[batch - 1]), inv))
        for r in (0, batch - 1):
            ref = O.intt(x[r], threads=8) if inv else O.ntt(x[r], threads=8)
            assert np.array_equal(got[r], ref)
        assert np.array_equal(single, got[batch - 1])


def test_kernel_selection_gives_identical_bits(D):
    """bb_ntt_set_kernel: 1 (default) runs the TMA-staged two-pass kernel where it applies (plain 2^24-point vectors, also
    batched), 0 the tile kernel everywhere.  Same transform, same bits, forward and inverse; shapes the TMA-staged kernel
    does not take (2^22, AoS Ext) go through the tile kernel either way and must not be disturbed by the switch."""
    import torch
    from toyni_b200.lib import lib
    L = lib()
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    try:
        for shape, fn in (((1 << 22,), D.ntt_), ((1 << 24,), D.ntt_), ((1 << 24, 4), D.ntt_ext_), ((3, 1 << 24), D.ntt_batch_)):
            t = torch.randint(0, P, shape, dtype=torch.int32, device="cuda", generator=g)
            for inv in (False, True):
                L.bb_ntt_set_kernel(0)
                a = fn(t.clone(), inv)
                L.bb_ntt_set_kernel(1)
                b = fn(t.clone(), inv)
                assert torch.equal(a, b)
                assert int(b.max()) < P and int(b.min()) >= 0
    finally:
        L.bb_ntt_set_kernel(1)
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ four-step building blocks (multi-GPU layer)
@pytest.mark.parametrize("log_n", [10, 16, 21])
def test_fourstep_paths_on_one_gpu(D, log_n):
    """The sharded four-step NTT with world = 1: both the NCCL-shaped path (twiddle kernel + row transforms) and
    the fused path (twiddle + transpose inside the last column pass, stores into the receive buffer) must equal
    the 1-D transform.  With world > 1 the same code runs under torchrun (tools/mg_check.py)."""
    from toyni_b200 import multigpu as MG
    n = 1 << log_n
    x = O.random_field(n, seed=log_n)
    for inv in (False, True):
        ref = O.intt(x, threads=4) if inv else O.ntt(x, threads=4)
        out = MG.fourstep_ntt_cuda(D.to_device(MG.fourstep_scatter(x, 0, 1)), log_n, 0, 1, inverse=inv)
        assert np.array_equal(MG.fourstep_gather([D.to_host(out)], log_n), ref)
        fs = MG.FourStepFused(log_n, 0, 1)
        out = fs.run(D.to_device(MG.fourstep_scatter(x, 0, 1)), inverse=inv)
        assert np.array_equal(MG.fourstep_gather([D.to_host(out)], log_n), ref)
        fs.close()


def test_fourstep_two_transforms_in_flight(D):
    """FourStepFused.run_async: consecutive transforms alternate over two internal streams and three receive buffers (the
    exchange of one beside the row passes of the previous one).  Five different vectors, outputs consumed one call late,
    then the serial form again on the same object; world = 1 here, world > 1 in tools/mg_check.py / bench.py."""
    import torch
    from toyni_b200 import multigpu as MG
    log_n = 18
    xs = [O.random_field(1 << log_n, seed=900 + i) for i in range(5)]
    refs = [O.ntt(x, threads=4) for x in xs]
    fs = MG.FourStepFused(log_n, 0, 1, nbuf=3)
    pend, got = [], []
    for x in xs:
        pend.append(fs.run_async(D.to_device(MG.fourstep_scatter(x, 0, 1))))
        if len(pend) == 2:
            o, ev = pend.pop(0)
            torch.cuda.current_stream().wait_event(ev)
            got.append(MG.fourstep_gather([D.to_host(o)], log_n))
    for o, ev in pend:
        torch.cuda.current_stream().wait_event(ev)
        got.append(MG.fourstep_gather([D.to_host(o)], log_n))
    fs.join()
    for g, r in zip(got, refs):
        assert np.array_equal(g, r)
    out = fs.run(D.to_device(MG.fourstep_scatter(xs[0], 0, 1)))
    assert np.array_equal(MG.fourstep_gather([D.to_host(out)], log_n), refs[0])
    fs.check_peers()
    fs.close()


def test_two_contexts_on_two_threads_do_not_share_scratch():
    """The reference's context is not re-entrant and its callers are single threaded (src/ntt.rs:118-120); here two
    sizes driven from two host threads at once must both come out right (per-context stream and scratch)."""
    import threading
    from toyni_b200 import ntt
    results = {}

    def work(log_n, reps):
        x = O.random_field(1 << log_n, seed=log_n)
        ref = O.ntt(x, threads=2)
        ok = True
        for _ in range(reps):
            v = x.copy()
            ntt.ntt_cuda(v)
            ok &= bool(np.array_equal(v, ref))
        results[log_n] = ok

    ts = [threading.Thread(target=work, args=(18, 6)), threading.Thread(target=work, args=(19, 4))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert results == {18: True, 19: True}


def test_roots_of_unity_domain_and_u32_host_entry(D):
    """a6: roots_of_unity_domain (src/ntt.rs:69-81, its test :360-379: 16 distinct roots, w^16 = 1) from the device twiddle
    cache; the device-resident domain elements; and the u32 host entry point (half the PCIe bytes of the u64 drop-in)."""
    import ctypes
    import torch
    from toyni_b200 import ntt
    from toyni_b200.lib import lib
    dom = ntt.roots_of_unity_domain(16)
    assert len(set(int(v) for v in dom)) == 16 and dom[0] == 1
    assert pow(int(dom[1]), 16, P) == 1
    for log_n in (0, 1, 9, 20):
        assert np.array_equal(ntt.roots_of_unity_domain(1 << log_n), O.roots_of_unity_domain(1 << log_n))
    d = torch.empty(1 << 18, dtype=torch.int32, device="cuda")
    assert lib().bb_domain_elements_device(18, 7, ctypes.c_void_p(d.data_ptr())) == 0
    assert np.array_equal(D.to_host(d), O.domain_elements(1 << 18, 7))
    L = lib()
    x = O.random_field(1 << 20, seed=5)
    ctx = L.ntt_ctx_create(1 << 20)
    v = x.astype(np.uint32)
    assert L.bb_ntt_host_u32(ctypes.c_void_p(ctx), v.ctypes.data, 0) == 0
    assert np.array_equal(v.astype(np.uint64), O.ntt(x, threads=4))
    assert L.bb_ntt_host_u32(ctypes.c_void_p(ctx), v.ctypes.data, 1) == 0
    assert np.array_equal(v.astype(np.uint64), x)
    L.ntt_ctx_destroy(ctypes.c_void_p(ctx))
