"""Parity of FRI folding, the SHA-256 Merkle commit and the prover's FRI commit loop against the CPU oracle."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

P = O.P


@pytest.fixture(scope="module")
def D():
    import torch
    from toyni_b200 import device
    torch.cuda.set_device(0)
    return device


@pytest.mark.parametrize("log_m", [1, 2, 5, 10, 16, 20])
def test_fri_fold_base_and_ext(D, log_m):
    """src/math/fri.rs:27-48 and :7-25 on the coset the prover uses (src/fibonacci.rs:214)."""
    m = 1 << log_m
    ev = O.random_field(m, seed=m)
    xs = O.domain_elements(m, 7)
    if log_m <= 16:
        assert np.array_equal(D.to_host(D.fri_fold(D.to_device(ev), 7, 123456789)), O.fri_fold(ev, xs, 123456789))
        assert np.array_equal(D.to_host(D.fri_fold_xs(D.to_device(ev), D.to_device(xs[: m // 2]), 5)), O.fri_fold(ev, xs, 5))
        ee = O.random_field(4 * m, seed=m + 1).reshape(m, 4)
        beta = [5, 6, 7, P - 1]
        assert np.array_equal(D.to_host(D.fri_fold(D.to_device(ee), 7, beta)), O.fri_fold_ext(ee, xs, beta))
        assert np.array_equal(D.to_host(D.fri_fold_xs(D.to_device(ee), D.to_device(xs[: m // 2]), beta)),
                              O.fri_fold_ext(ee, xs, beta))
    else:
        # size-independent property: folding evaluations of f gives evaluations of f_even + beta f_odd on x^2
        c = O.random_field(1 << 10, seed=3)
        evd = D.coset_fft(D.to_device(c), m, 7)
        beta = 987654321
        folded = D.fri_fold(evd, 7, beta)
        g = (c[0::2].astype(object) + beta * c[1::2].astype(object)) % P
        want = D.coset_fft(D.to_device(np.array(g, dtype=np.uint64)), m // 2, 49)
        assert np.array_equal(D.to_host(folded), D.to_host(want))


def test_fold_later_layers_use_squared_points(D):
    """xs are squared in place each round (src/fibonacci.rs:228-231): layer k has x0 = shift^(2^k)."""
    m = 1 << 12
    cur = O.random_field(m, seed=8)
    xs = O.domain_elements(m, 7)
    dev = D.to_device(cur)
    x0 = 7
    for k in range(6):
        beta = 1000 + k
        ref = O.fri_fold(cur, xs, beta)
        dev = D.fri_fold(dev, x0, beta)
        assert np.array_equal(D.to_host(dev), ref)
        xs = (xs[: ref.size].astype(object) ** 2 % P).astype(np.uint64)
        x0 = x0 * x0 % P
        cur = ref


def test_cyclic_shards_fold_without_exchange(D):
    m, G = 1 << 12, 8
    ee = O.random_field(4 * m, seed=1).reshape(m, 4)
    beta = [9, 8, 7, 6]
    full = O.fri_fold_ext(ee, O.domain_elements(m, 7), beta)
    for r in range(G):
        part = D.fri_fold_shard(D.to_device(np.ascontiguousarray(ee[r::G])), 12, 7, beta, G, r)
        assert np.array_equal(D.to_host(part), full[r::G])


@pytest.mark.parametrize("limbs", [1, 4])
def test_fold_chain_in_one_call(D, limbs):
    """bb_fri_fold_chain_shard_device: every fold of the prover's chain (src/fibonacci.rs:213-231 without the commits)
    launched by one C call, on cyclic shards of 1, 2 and 8 ranks; every layer against the oracle."""
    from toyni_b200 import multigpu as MG
    log_m, shift = 13, 7
    m = 1 << log_m
    ee = O.random_field(limbs * m, seed=40 + limbs).reshape(m, 4) if limbs == 4 else O.random_field(m, seed=41)
    betas = [[(3 * k + j + 5) % P for j in range(4)] if limbs == 4 else (3 * k + 5) % P for k in range(log_m)]
    ref, xs, cur = [], O.domain_elements(m, shift), ee
    while cur.shape[0] > 16:
        cur = O.fri_fold_ext(cur, xs, betas[len(ref)]) if limbs == 4 else O.fri_fold(cur, xs, betas[len(ref)])
        xs = (xs[: cur.shape[0]] * xs[: cur.shape[0]]) % np.uint64(P)
        ref.append(cur)
    for G in (1, 2, 8):
        for r in range(G):
            layers = MG.fold_chain_cuda(D.to_device(np.ascontiguousarray(ee[r::G])), log_m, shift, betas, r, G, until=16)
            assert len(layers) == len(ref) + 1
            for k, want in enumerate(ref):
                assert np.array_equal(D.to_host(layers[k + 1]), want[r::G]), (G, r, k)


def test_pool_allocations_and_device_copy(D):
    """bb_pool_alloc / bb_pool_free / bb_pool_trim (stream-ordered, kept between uses) and bb_d2d, as the C++ prover uses
    them: allocate, copy device to device both ways, free (NULL included), trim; the copies are exact."""
    import ctypes as C
    import torch
    from toyni_b200.lib import check, lib
    L = lib()
    D._bind_stream()
    x = D.to_device(O.random_field(1 << 16, seed=8))
    p1, p2 = C.c_void_p(), C.c_void_p()
    check(L.bb_pool_alloc(C.byref(p1), 4 << 16), "bb_pool_alloc")
    check(L.bb_d2d(p1, C.c_void_p(x.data_ptr()), 4 << 16), "bb_d2d")
    back = torch.empty_like(x)
    check(L.bb_d2d(C.c_void_p(back.data_ptr()), p1, 4 << 16), "bb_d2d")
    check(L.bb_sync(), "bb_sync")
    assert torch.equal(back, x)
    check(L.bb_pool_free(p1), "bb_pool_free")
    check(L.bb_pool_alloc(C.byref(p2), 4 << 16), "bb_pool_alloc")
    assert p2.value
    check(L.bb_pool_free(p2), "bb_pool_free")
    check(L.bb_pool_free(None), "bb_pool_free(NULL)")
    check(L.bb_pool_trim(), "bb_pool_trim")


def test_fold_host_mirror():
    from toyni_b200 import fri
    m = 1 << 10
    ev = O.random_field(m, seed=2)
    xs = O.domain_elements(m, 7)
    assert np.array_equal(fri.fri_fold(ev, xs, 77), O.fri_fold(ev, xs, 77))
    ee = O.random_field(4 * m, seed=3).reshape(m, 4)
    assert np.array_equal(fri.fri_fold_ext(ee, xs, [1, 2, 3, 4]), O.fri_fold_ext(ee, xs, [1, 2, 3, 4]))
    with pytest.raises(AssertionError):  # src/math/fri.rs:8
        fri.fri_fold(ev[:-1], xs, 1)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 100, 255, 256, 257, 512, 1 << 10, (1 << 12) + 1, 1 << 14, 259 * 256, (1 << 18) + 256, 1 << 19])
def test_merkle_commit_every_level(D, n):
    """src/merkle.rs:25-48 with the prover's leaves (src/fibonacci.rs:340-363): every node of every level equal.  The sizes
    walk every launch shape of merkle_upper_levels: single-CTA tail only (<= 256 nodes), one level per launch (sizes that
    are not multiples of 256, odd levels duplicating their last node), eight levels per launch (multiples of 256, also a
    non-power-of-two one), and the persistent large-level kernel (parents >= 2^18)."""
    import torch
    v = O.random_field(n, seed=n)
    salts = O.random_bytes(16 * n, seed=n + 1).reshape(n, 16)
    nodes, root = D.merkle_commit(D.to_device(v), torch.from_numpy(salts).cuda())
    rn, rr = O.commit_values(v, salts)
    assert root == rr and np.array_equal(nodes.cpu().numpy(), rn)
    nodes, root = D.merkle_commit(D.to_device(v))
    rn, rr = O.commit_values(v)
    assert root == rr and np.array_equal(nodes.cpu().numpy(), rn)
    v4 = O.random_field(4 * n, seed=n + 2).reshape(n, 4)
    nodes, root = D.merkle_commit(D.to_device(v4), torch.from_numpy(salts).cuda())
    assert root == O.commit_values(v4, salts, limbs=4)[1]
    nodes, root = D.merkle_commit(D.to_device(v4))
    assert root == O.commit_values(v4, None, limbs=4)[1]


def test_merkle_kat_roots(D, golden):
    """SURVEY KAT7 (leaves LE-u64(1..k), the inputs of src/merkle.rs:129-163) and the golden fixture."""
    want = {4: "082e8e29b028ef12e81530323943dc08834f103e41a73e41c4cbd14b115f85c9",
            3: "3c391efe69e4a5a3e6212efeb617a669e6b73fa053c9064b70c9ade3810a2a93",
            1: "51b09ceccfbec44595dd4241e6e2a693d279b72c899c8f60ec63524fe58b1d4f"}
    for k, h in want.items():
        assert D.merkle_commit(D.to_device(np.arange(1, k + 1, dtype=np.uint64)))[1].hex() == h
    import torch
    _, root = D.merkle_commit(D.to_device(golden["merkle_vals"]), torch.from_numpy(golden["merkle_salts"]).cuda())
    assert root == golden["merkle_root_salted"].tobytes()


def test_merkle_host_mirror_openings_verify():
    from toyni_b200 import merkle as M
    n = 300
    v = O.random_field(n, seed=5)
    salts = O.random_bytes(16 * n, seed=6).reshape(n, 16)
    tree = M.build_merkle_tree(v, salts)
    assert tree.root() == O.commit_values(v, salts)[1]
    for i in (0, 1, 150, 298, 299):
        leaf = salts[i].tobytes() + int(v[i]).to_bytes(8, "little")
        assert M.verify_merkle_proof(leaf, tree.get_proof(i), tree.root())
    assert M.build_unsalted_tree(v).root() == O.commit_values(v)[1]


def test_generic_byte_leaves_and_device_openings(D):
    """MerkleTree::new over raw byte leaves of any length, and get_proof gathered on the device."""
    import ctypes as C
    import torch
    from toyni_b200.lib import check, lib
    L = lib()
    for leaf_len in (1, 8, 24, 55, 56, 63, 64, 100):
        n = 37
        raw = O.random_bytes(n * leaf_len, seed=leaf_len)
        leaves = [raw[i * leaf_len:(i + 1) * leaf_len].tobytes() for i in range(n)]
        rn, rr = O.merkle_build(leaves)
        d_leaves = torch.from_numpy(raw).cuda()
        nodes = torch.empty((L.bb_merkle_node_count(n), 32), dtype=torch.uint8, device="cuda")
        root = np.zeros(32, np.uint8)
        L.bb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        check(L.bb_merkle_build_bytes_device(d_leaves.data_ptr(), n, leaf_len, nodes.data_ptr(), root.ctypes.data))
        assert root.tobytes() == rr and np.array_equal(nodes.cpu().numpy(), rn)
        path, pos, depth = np.zeros((64, 32), np.uint8), np.zeros(64, np.uint8), C.c_size_t(0)
        check(L.bb_merkle_open_device(nodes.data_ptr(), n, 36, path.ctypes.data, pos.ctypes.data, C.byref(depth)))
        rp, rpos = O.merkle_open(rn, n, 36)
        assert depth.value == len(rpos) and np.array_equal(path[: depth.value], rp) and np.array_equal(pos[: depth.value], rpos)
        assert O.merkle_verify(leaves[36], path[: depth.value], pos[: depth.value], rr)


class _Transcript:  # src/transcript.rs, host side of the commit loop
    def __init__(self):
        self.state = b"toyni-stark-v1"

    def absorb(self, d):
        self.state += d

    def squeeze(self):
        h = hashlib.sha256(self.state).digest()
        self.state = h
        return int.from_bytes(h[:8], "little") % P


@pytest.mark.parametrize("ext", [False, True])
def test_fri_commit_loop_matches_oracle(D, ext):
    """src/fibonacci.rs:200-247: layers, roots and transcript-derived betas all equal, final layer unsalted."""
    import torch
    n, final = 1 << 12, 16
    l0 = O.random_field(n * (4 if ext else 1), seed=77)
    if ext:
        l0 = l0.reshape(n, 4)
    salts = O.random_bytes(16 * sum(n >> k for k in range(8)), seed=5)
    layers_ref, roots_ref, betas_ref = O.fri_commit(l0, 7, final, salts, ext=ext)
    t, seen = _Transcript(), []

    def challenge(root, layer):
        t.absorb(root)
        b = [t.squeeze() for _ in range(4)] if ext else t.squeeze()
        seen.append(b)
        return b

    layers, nodes, roots = D.fri_commit(D.to_device(l0), 7, final, torch.from_numpy(salts).cuda(), challenge=challenge)
    assert roots == roots_ref
    assert len(layers) == len(layers_ref) == 9
    for a, b in zip(layers, layers_ref):
        assert np.array_equal(D.to_host(a), b)
    assert np.array_equal(np.array(seen, dtype=np.uint64).reshape(betas_ref.shape), betas_ref)
    assert roots[-1] == O.commit_values(layers_ref[-1], None, limbs=4 if ext else 1)[1]


def test_fri_commit_golden(D, golden):
    import torch
    t = _Transcript()

    def challenge(root, layer):
        t.absorb(root)
        return t.squeeze()

    layers, _, roots = D.fri_commit(D.to_device(golden["fri_layer0"]), 7, 16, torch.from_numpy(golden["fri_salts"]).cuda(),
                                    challenge=challenge)
    assert np.array_equal(np.frombuffer(b"".join(roots), np.uint8).reshape(-1, 32), golden["fri_roots"])
    assert np.array_equal(D.to_host(layers[-1]), golden["fri_final"])


def test_fold_chain_2_25_to_16_constant_final_layer(D):
    """Config 4 at full size, property form: the chain of a degree < 2^21 codeword over the extension field ends
    in a constant layer of 16 values (what the verifier checks, src/verifier.rs:69-75)."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    c = torch.randint(0, P, (1 << 21, 4), dtype=torch.int32, device="cuda", generator=g)
    l0 = D.coset_fft(c, 1 << 25, 7)
    betas = O.random_field(4 * 21, seed=9)
    layers, _, _ = D.fri_commit(l0, 7, 16, betas=betas, hash_layers=False)
    assert len(layers) == 22 and layers[-1].shape[0] == 16
    last = layers[-1]
    assert bool((last == last[0]).all())


def test_openings_of_several_trees_in_one_call(D):
    """bb_merkle_open_multi_device (the openings of a whole proof in one launch) against the per-tree calls
    (bb_merkle_open_batch_device + bb_gather_device, themselves checked against src/merkle.rs:50-80 in the oracle): an odd
    salted tree, a power-of-two unsalted one and a single-leaf tree; paths, position flags, values and salts."""
    import ctypes as C
    import torch
    from toyni_b200.lib import check, lib
    L = lib()

    class Req(C.Structure):
        _fields_ = [("d_nodes", C.c_void_p), ("nleaves", C.c_size_t), ("d_vals", C.c_void_p), ("d_salts", C.c_void_p),
                    ("first", C.c_size_t), ("count", C.c_size_t)]

    specs = [(1000, True, [0, 999, 998, 501, 7]), (64, False, [63, 0, 31]), (1, False, [0])]
    trees, reqs, indices = [], [], []
    for n, salted, idx in specs:
        v = D.to_device(O.random_field(n, seed=n + 5))
        s = torch.from_numpy(O.random_bytes(16 * n, seed=n + 6).reshape(n, 16)).cuda() if salted else None
        nodes, _ = D.merkle_commit(v, s)
        trees.append((v, s, nodes))
        reqs.append(Req(nodes.data_ptr(), n, v.data_ptr(), s.data_ptr() if salted else None, len(indices), len(idx)))
        indices += idx
    depths = [max(n - 1, 0).bit_length() for n, _, _ in specs]
    nq = len(indices)
    pbytes = sum(32 * d * len(idx) for d, (_, _, idx) in zip(depths, specs))
    paths = np.zeros(max(pbytes, 1), np.uint8)
    pos = np.zeros(max(pbytes // 32, 1), np.uint8)
    vals = np.zeros(nq, np.uint32)
    salts = np.zeros((nq, 16), np.uint8)
    ia = np.asarray(indices, np.uint64)
    D._bind_stream()
    check(L.bb_merkle_open_multi_device((Req * len(reqs))(*reqs), len(reqs), ia.ctypes.data, nq, 4, paths.ctypes.data, pbytes,
                                        pos.ctypes.data, vals.ctypes.data, salts.ctypes.data), "bb_merkle_open_multi_device")
    po = q = 0
    for (n, salted, idx), d, (v, s, nodes) in zip(specs, depths, trees):
        want_paths, want_pos = D.merkle_open_batch(nodes, n, idx)
        want_vals = D.gather(v, idx).view(np.uint32).reshape(-1)
        for k in range(len(idx)):
            assert np.array_equal(paths[32 * po:32 * (po + d)].reshape(d, 32), want_paths[k])
            assert np.array_equal(pos[po:po + d], want_pos[k])
            assert vals[q] == want_vals[k]
            assert np.array_equal(salts[q], D.gather(s, [idx[k]])[0] if salted else np.zeros(16, np.uint8))
            po += d
            q += 1
    # a request list that does not cover the indices, or a wrong paths size, is refused
    assert L.bb_merkle_open_multi_device((Req * 1)(reqs[0]), 1, ia.ctypes.data, nq, 4, paths.ctypes.data, pbytes, pos.ctypes.data,
                                         vals.ctypes.data, salts.ctypes.data) != 0
    L.bb_clear_error()
