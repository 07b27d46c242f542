"""The reference's own identity tests for this path, re-expressed against the oracle.
Each test names the reference test it mirrors."""
import numpy as np

from oracle import oracle as O
from oracle import pyref as R

P = O.P


def test_ntt_intt_roundtrip():  # src/ntt.rs:322-336
    x = (np.arange(256, dtype=np.uint64) * 7 + 3) % P
    assert np.array_equal(O.intt(O.ntt(x)), x)


def test_roots_of_unity():  # src/ntt.rs:360-379
    d = O.roots_of_unity_domain(16)
    assert d[0] == 1 and O.bb_pow(int(d[1]), 16) == 1 and len(set(d.tolist())) == 16


def test_domain_elements_and_nesting():  # src/math/domain.rs:181-191,292-304; tests/fri.rs:10-25
    e8 = O.domain_elements(8)
    assert e8[0] == 1 and O.bb_pow(O.root_of_unity(3), 8) == 1
    small, big = O.domain_elements(4), O.domain_elements(32)
    assert all(small[i] == big[8 * i] for i in range(4))


def test_fft_ifft_roundtrips():  # src/math/domain.rs:194-218
    c = np.array([3 * i + 1 for i in range(8)], np.uint64)
    assert np.array_equal(O.domain_ifft(O.domain_fft(c, 8, 1), 1), c)
    assert np.array_equal(O.domain_ifft(O.domain_fft(c, 8, 7), 7), c)


def test_coset_evaluations_match_horner():  # src/math/domain.rs:221-242
    c = [1, 2, 3]
    ev = O.domain_fft(c, 8, 7)
    for i, x in enumerate(O.domain_elements(8, 7)):
        assert int(ev[i]) == R.horner(c, int(x))


def test_ext_fft_roundtrip_and_horner():  # src/math/domain.rs:245-278
    c = np.array([[3 * i + 1, i + 2, 7 * i, i + 5] for i in range(8)], np.uint64)
    ev = O.domain_fft_ext(c, 8, 1)
    assert np.array_equal(O.domain_ifft_ext(ev, 1), c)
    c3 = [[i + 1, 2 * i, i + 4, 7] for i in range(3)]
    ev3 = O.domain_fft_ext(np.array(c3, np.uint64), 8, 1)
    for i, x in enumerate(O.domain_elements(8)):
        for k in range(4):
            assert int(ev3[i][k]) == R.horner([row[k] for row in c3], int(x))


def test_ext_field_identities():  # src/ext.rs:210-275
    rng = np.random.default_rng(0xC0FFEE)
    x4 = [0, 1, 0, 0]
    x2 = O.ext_mul(x4, x4)
    assert list(O.ext_mul(x2, x2)) == [11, 0, 0, 0]  # X^4 = 11
    for _ in range(50):
        a = [int(v) for v in rng.integers(1, P, 4)]
        b = [int(v) for v in rng.integers(0, P, 4)]
        c = [int(v) for v in rng.integers(0, P, 4)]
        assert list(O.ext_mul(a, O.ext_inverse(a))) == [1, 0, 0, 0]
        lhs = O.ext_mul(a, [(b[k] + c[k]) % P for k in range(4)])
        rhs = [(int(u) + int(v)) % P for u, v in zip(O.ext_mul(a, b), O.ext_mul(a, c))]
        assert list(lhs) == rhs
        s = int(rng.integers(0, P))
        assert list(O.ext_mul(a, [s, 0, 0, 0])) == [x * s % P for x in a]  # mul_base == full multiply


def test_merkle_proofs():  # src/merkle.rs:129-189
    for k in (4, 3, 1):
        leaves = [int(i).to_bytes(8, "little") for i in range(1, k + 1)]
        nodes, root = O.merkle_build(leaves)
        for i, leaf in enumerate(leaves):
            path, pos = O.merkle_open(nodes, k, i)
            assert O.merkle_verify(leaf, path, pos, root)
            assert not O.merkle_verify(b"\xff" * 8, path, pos, root) or k == 0
    # leaf / node domain separation (src/merkle.rs:166-189)
    l, r = b"\x11" * 32, b"\x22" * 32
    assert O.hash_leaf(l + r) != O.hash_node(l, r)


def test_fri_fold_identity():  # what tests/fri.rs:101-133 means to check: fold = f_even(x^2) + beta f_odd(x^2)
    c = [int(v) for v in O.random_field(16, seed=5)]
    ev = O.domain_fft(c, 64, 7)
    xs = O.domain_elements(64, 7)
    beta = 424242
    folded = O.fri_fold(ev, xs, beta)
    g = [(c[2 * i] + beta * c[2 * i + 1]) % P for i in range(8)]
    for i in range(32):
        assert int(folded[i]) == R.horner(g, int(xs[i]) * int(xs[i]) % P)


def test_transcript_behaviour():  # src/transcript.rs:12-72
    t, r = O.FiatShamirTranscript(), R.Transcript()
    for blob in (b"root-one" * 4, b"\x00" * 32):
        t.absorb(blob)
        r.absorb(blob)
        assert t.squeeze_challenge() == r.squeeze_challenge()
    assert t.squeeze_indices(44, 1 << 14) == r.squeeze_indices(44, 1 << 14)


def test_fri_commit_loop_shape():  # src/fibonacci.rs:200-247 and the proof shapes of SURVEY Appendix A
    n, final = 1 << 11, 8  # trace_len 64: lde 2048, final layer 8, 8 folds, 9 commitments
    l0 = O.random_field(n, seed=3)
    salts = O.random_bytes(16 * sum(n >> k for k in range(8)), seed=4)
    layers, roots, betas = O.fri_commit(l0, 7, final, salts)
    assert len(layers) == 9 and len(roots) == 9 and len(betas) == 8
    assert len(layers[-1]) == 8
    # the final layer is committed unsalted and every beta comes from the transcript after the previous root
    assert O.commit_values(layers[-1])[1] == roots[-1]
    t = O.FiatShamirTranscript()
    t.absorb(roots[0])
    for k in range(8):
        assert t.squeeze_challenge() == int(betas[k])
        t.absorb(roots[k + 1])
