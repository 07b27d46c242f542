"""Parity of the device-resident prover stages (prover_ew.cu) against the CPU oracle's restatement
of src/fibonacci.rs:133-198, src/math/polynomial.rs:134-144 and src/merkle.rs:50-80, through the C ABI.  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import fibonacci as F  # noqa: E402
from oracle import oracle as O  # noqa: E402

P = O.P


@pytest.fixture(scope="module")
def D():
    import torch
    from toyni_b200 import device
    torch.cuda.set_device(0)
    return device


@pytest.mark.parametrize("log_n", [5, 10, 15])
def test_constraint_quotient_and_deep_match_the_reference_formulas(D, log_n):
    """One random 'trace LDE' pushed through the three element-wise stages on the device and through the oracle
    prover's numpy formulas (same lines of src/fibonacci.rs): every value equal."""
    n = 1 << log_n
    trace_len = n // 32
    t = O.random_field(n, seed=log_n)
    xs = O.domain_elements(n, 7)
    g = O.root_of_unity(max(trace_len.bit_length() - 1, 0))
    b1v, b2v = pow(g, max(trace_len - 1, 0), P), pow(g, max(trace_len - 2, 0) if trace_len >= 2 else 0, P)
    t_gx, t_ggx = np.roll(t, -32), np.roll(t, -64)
    c_ref = F.mulmod(F.mulmod(F.submod(t_ggx, F.addmod(t_gx, t)), F.submod(xs, np.uint64(b1v))), F.submod(xs, np.uint64(b2v)))
    td = D.to_device(t)
    c_dev = D.fib_constraint(td, 32, 7, b1v, b2v)
    assert np.array_equal(D.to_host(c_dev), c_ref)
    # quotient: c / (x^trace_len - 1); Z_H takes 32 values on the coset (trace_len >= 1)
    z_h = F.submod(F.powmod_vec(xs, trace_len), np.uint64(1))
    q_ref = F.mulmod(c_ref, F.invmod_vec(z_h))
    g_ext = O.root_of_unity(log_n)
    sn, wn = pow(7, trace_len, P), pow(g_ext, trace_len, P)
    q_dev = D.scale_periodic_(c_dev.clone(), [pow((sn * pow(wn, i, P) - 1) % P, P - 2, P) for i in range(32)])
    assert np.array_equal(D.to_host(q_dev), q_ref)
    # DEEP composition at a point outside the coset
    z, q_z, t_z, t_gz, t_ggz = 123456789, 11, 22, 33, 44
    assert pow(z * pow(7, P - 2, P) % P, n, P) != 1
    inv_xz = F.invmod_vec(F.submod(xs, np.uint64(z)))
    d_ref = F.mulmod(F.submod(q_ref, np.uint64(q_z)), inv_xz)
    d_ref = F.addmod(d_ref, F.mulmod(F.submod(t_ggx, np.uint64(t_ggz)), inv_xz))
    d_ref = F.addmod(d_ref, F.mulmod(F.submod(t_gx, np.uint64(t_gz)), inv_xz))
    d_ref = F.addmod(d_ref, F.mulmod(F.submod(t, np.uint64(t_z)), inv_xz))
    d_dev = D.fib_deep(q_dev, td, 32, 7, z, q_z, t_z, t_gz, t_ggz)
    assert np.array_equal(D.to_host(d_dev), d_ref)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, (1 << 16) + 7])
def test_poly_eval_matches_horner(D, n):
    """Polynomial::evaluate (src/math/polynomial.rs:134-144): ragged lengths around the 64-coefficient chunk."""
    c = O.random_field(n, seed=n)
    for z in (0, 1, 2, P - 1, 987654321):
        assert D.poly_eval(D.to_device(c), z) == F.poly_eval(c, z)


@pytest.mark.parametrize("n,salted", [(1, False), (3, True), (4, True), (1000, True), (1 << 12, False)])
def test_batched_openings_match_the_reference_openings(D, n, salted):
    """MerkleTree::get_proof for a whole query set (src/merkle.rs:50-80), odd level sizes included, plus the gather of
    the opened values and salts; each opening verifies against the root (src/merkle.rs:86-101)."""
    import torch
    vals = O.random_field(n, seed=n + 1)
    salts = O.random_bytes(16 * n, seed=n + 2).reshape(n, 16) if salted else None
    ref_nodes, ref_root = O.commit_values(vals, salts)
    vd = D.to_device(vals)
    sd = torch.from_numpy(salts).cuda() if salted else None
    nodes, root = D.merkle_commit(vd, sd)
    assert root == ref_root
    idx = sorted(set([0, n - 1, n // 2, (n * 7) // 11] + list(range(min(n, 5)))))
    paths, pos = D.merkle_open_batch(nodes, n, idx)
    got_vals = D.gather(vd, idx).view(np.uint32).reshape(-1)
    got_salts = D.gather(sd, idx) if salted else None
    for k, i in enumerate(idx):
        rp, rpos = O.merkle_open(ref_nodes, n, i)
        assert [p.tobytes() for p in rp] == [paths[k, d].tobytes() for d in range(paths.shape[1])]
        assert [bool(b) for b in rpos] == [bool(b) for b in pos[k]]
        assert int(got_vals[k]) == int(vals[i])
        leaf = (got_salts[k].tobytes() if salted else b"") + int(vals[i]).to_bytes(8, "little")
        assert O.merkle_verify(leaf, paths[k], pos[k], root)


def test_interleave_is_the_cyclic_to_block_relayout(D):
    import torch
    for groups, chunk, limbs in ((2, 8, 1), (4, 16, 4), (8, 1, 1), (1, 5, 4)):
        shape = (groups * chunk, 4) if limbs == 4 else (groups * chunk,)
        src = torch.arange(groups * chunk * limbs, dtype=torch.int32, device="cuda").reshape(shape)
        dst = D.interleave(src, groups)
        ref = src.view((groups, chunk) + tuple(shape[1:])).transpose(0, 1).contiguous().view(shape)
        assert torch.equal(dst, ref)
