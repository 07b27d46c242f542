"""Full-size parity against the CPU oracle (BASELINE.json configs 2-4 at their own sizes), through the C ABI.

Every output element is compared: the oracle's loops run over all host threads (oracle.set_threads / threads=...),
which cuts the reference's serial loops into chunks without changing a single operation (src/ntt.rs:24-66,
src/math/domain.rs:107-123, src/math/fri.rs:7-25, src/merkle.rs:16-48), so a full 2^27 vector costs seconds.
The template is the reference's own GPU test: GPU == CPU, element by element (src/ntt.rs:264-287)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

P = O.P
CORES = os.cpu_count() or 1


@pytest.fixture(scope="module")
def D():
    import torch
    from toyni_b200 import device
    torch.cuda.set_device(0)
    return device


@pytest.fixture()
def oracle_threads():
    O.set_threads(CORES)
    yield
    O.set_threads(1)


@pytest.mark.parametrize("log_n", [24, 25, 26, 27])
def test_ntt_full_vector_matches_oracle(D, log_n):
    """Forward transform of a whole 2^24..2^27-point vector (2^27 is the two-adicity limit, src/babybear.rs:119)."""
    import torch
    x = O.random_field(1 << log_n, seed=100 + log_n)
    ref = O.ntt(x, threads=CORES)
    got = D.to_host(D.ntt_(D.to_device(x)))
    assert np.array_equal(got, ref)
    del got, ref
    torch.cuda.empty_cache()


@pytest.mark.parametrize("log_n", [24, 27])
def test_intt_full_vector_matches_oracle(D, log_n):
    import torch
    x = O.random_field(1 << log_n, seed=200 + log_n)
    ref = O.intt(x, threads=CORES)
    got = D.to_host(D.ntt_(D.to_device(x), True))
    assert np.array_equal(got, ref)
    torch.cuda.empty_cache()


def test_ntt_2_24_both_kernels_and_batches(D):
    """The TMA-staged two-pass kernel (default) and the tile kernel give the oracle's bits; batches of 2^24-point
    vectors (one launch over several vectors) as well."""
    import torch
    from toyni_b200.lib import lib
    L = lib()
    x = O.random_field(3 << 24, seed=77).reshape(3, 1 << 24)
    refs = [O.ntt(x[i], threads=CORES) for i in range(3)]
    try:
        for kernel in (1, 0):
            L.bb_ntt_set_kernel(kernel)
            got = D.to_host(D.ntt_batch_(D.to_device(x), False))
            for i in range(3):
                assert np.array_equal(got[i], refs[i]), (kernel, i)
            back = D.to_host(D.ntt_batch_(D.to_device(got), True))
            assert np.array_equal(back, x)
    finally:
        L.bb_ntt_set_kernel(1)
    torch.cuda.empty_cache()


def test_lde_2_20_to_2_25_every_evaluation_and_salted_root(D, oracle_threads):
    """Config 3: BabyBearDomain::fft on the coset 7 * <w_2^25> of 2^20 coefficients (src/math/domain.rs:107-123), then the
    salted Merkle commit of the 2^25 evaluations (src/fibonacci.rs:340-353): every evaluation and the root."""
    import torch
    c = O.random_field(1 << 20, seed=320)
    ref = O.domain_fft(c, 1 << 25, 7)
    ev = D.coset_fft(D.to_device(c), 1 << 25, 7)
    assert np.array_equal(D.to_host(ev), ref)
    salts = O.random_bytes(16 << 25, seed=321).reshape(-1, 16)
    _, ref_root = O.commit_values(ref, salts)
    nodes, root = D.merkle_commit(ev, torch.from_numpy(salts).cuda())
    assert bytes(root) == bytes(ref_root)
    del nodes, ev
    torch.cuda.empty_cache()


def test_ext_fold_chain_2_25_every_layer(D, oracle_threads):
    """Config 4 (i): fri_fold_ext (src/math/fri.rs:7-25) from a 2^25 Ext codeword down to 16 values, x_i squared from layer
    to layer as the prover does (src/fibonacci.rs:228-231): every value of every layer."""
    import torch
    from toyni_b200 import multigpu as MG
    log_m, shift = 25, 7
    m = 1 << log_m
    full = O.random_field(4 * m, seed=425).reshape(m, 4)
    betas = [[(11 * k + j + 3) % P for j in range(4)] for k in range(log_m)]
    layers = MG.fold_chain_cuda(D.to_device(full), log_m, shift, betas, 0, 1, until=16)
    xs = O.domain_elements(m, shift)
    cur = full
    for k in range(1, len(layers)):
        cur = O.fri_fold_ext(cur, xs, betas[k - 1])
        xs = (xs[: cur.shape[0]] * xs[: cur.shape[0]]) % np.uint64(P)
        assert np.array_equal(D.to_host(layers[k]), cur), f"layer {k}"
    assert len(layers) == 22 and layers[-1].shape[0] == 16
    torch.cuda.empty_cache()


def test_fri_commit_loop_2_25_ext_roots_match_oracle(D, oracle_threads):
    """Config 4 (ii): the prover's commit loop (src/fibonacci.rs:200-247) on a 2^25 Ext codeword — fold, salted tree per
    layer, unsalted final layer, transcript — against the oracle's loop with the same salts: all 22 roots, the betas the
    transcript produced and the final layer."""
    import torch
    from toyni_b200.prover import FiatShamirTranscript
    m, shift, final = 1 << 25, 7, 16
    full = O.random_field(4 * m, seed=426).reshape(m, 4)
    sizes, mm = [], m
    while mm > final:
        sizes.append(mm)
        mm //= 2
    salts = O.random_bytes(16 * sum(sizes), seed=427)
    ref_layers, ref_roots, ref_betas = O.fri_commit(full.reshape(-1), shift, final, salts, ext=True)
    tr = FiatShamirTranscript()
    betas = []

    def challenge(root, layer):
        tr.absorb(root)
        b = [tr.squeeze_challenge() for _ in range(4)]
        betas.append(b)
        return b
    layers, nodes, roots = D.fri_commit(D.to_device(full), shift, final, torch.from_numpy(salts).cuda(), challenge=challenge)
    assert [bytes(r) for r in roots] == [bytes(r) for r in ref_roots]
    assert np.array_equal(np.array(betas, dtype=np.uint64), ref_betas)
    assert np.array_equal(D.to_host(layers[-1]), ref_layers[-1])
    assert np.array_equal(D.to_host(layers[5]), ref_layers[5])
    del layers, nodes
    torch.cuda.empty_cache()


def test_columns_64_x_2_22(D):
    """Config 5: 64 batched 2^22-point column NTTs; four of the columns against the oracle, all 64 by round trip."""
    import torch
    x = O.random_field(64 << 22, seed=522).reshape(64, 1 << 22)
    dx = D.to_device(x)
    got = D.ntt_batch_(dx.clone(), False)
    for i in (0, 21, 42, 63):
        assert np.array_equal(D.to_host(got[i]), O.ntt(x[i], threads=CORES))
    assert torch.equal(D.ntt_batch_(got, True), dx)
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n_coeffs,shift", [((1 << 20) + 140, 7), (12345, 7), (1 << 21, 7), (1 << 20, 1), ((1 << 21) - 4097, 3), (1, 7), (4097, 5)])
def test_lde_to_2_25_ragged_inputs(D, oracle_threads, n_coeffs, shift):
    """The blowup-32 plan (expansion pass + TMA-staged pass 2) on the inputs it must cope with: the masked trace polynomial
    of a 2^20-row proof (2^20 + 140 coefficients, src/fibonacci.rs:117-121), short and ragged coefficient vectors, twice the
    trace length, and the unshifted domain (src/math/domain.rs:107-123): every evaluation against the oracle."""
    import torch
    c = O.random_field(n_coeffs, seed=330 + n_coeffs % 97)
    ref = O.domain_fft(c, 1 << 25, shift)
    ev = D.coset_fft(D.to_device(c), 1 << 25, shift)
    assert np.array_equal(D.to_host(ev), ref)
    del ev
    torch.cuda.empty_cache()


def test_lde_to_2_25_matches_the_tile_kernel_plan(D):
    """Same LDE through the general three-pass plan (bb_ntt_set_kernel(0)): identical bits."""
    import torch
    from toyni_b200.lib import lib
    c = D.to_device(O.random_field((1 << 20) + 140, seed=77))
    a = D.coset_fft(c, 1 << 25, 7)
    lib().bb_ntt_set_kernel(0)
    try:
        b = D.coset_fft(c, 1 << 25, 7)
    finally:
        lib().bb_ntt_set_kernel(1)
    assert torch.equal(a, b)
    torch.cuda.empty_cache()
