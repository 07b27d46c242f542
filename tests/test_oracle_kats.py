"""Pin the CPU oracle: known answers held by the reference's own tests, SURVEY Appendix A, an independent
pure-Python restatement, Python hashlib, and the committed golden fixtures."""
import hashlib

import numpy as np

from oracle import oracle as O
from oracle import pyref as R

P = O.P


def sha_u64(a):
    return hashlib.sha256(np.asarray(a, dtype="<u8").tobytes()).hexdigest()


# ---- values the reference's own tests hold (file:line in the reference tree)
def test_babybear_kats_from_reference_tests():
    assert O.bb_mul(100, 200) == 20000                      # src/babybear.rs:221-226
    assert O.lib().to_bb_new(P + 5) == 5                      # :228-232
    assert O.bb_pow(3, 4) == 81                             # :234-238
    for log_n in range(1, 11):                              # :240-247
        assert O.bb_pow(O.root_of_unity(log_n), 1 << log_n) == 1
    a = 12345
    assert O.bb_mul(a, O.bb_inverse(a)) == 1                # :249-254
    assert O.bb_add(a, O.lib().to_bb_neg(a)) == 0


def test_ntt_n8_kat_from_reference_test():
    """src/ntt.rs:339-357: coefficients [1,2,3], evals[0] = 6 and evals[1] = 1 + 2w + 3w^2."""
    ev = O.ntt([1, 2, 3, 0, 0, 0, 0, 0])
    w = int(O.roots_of_unity_domain(8)[1])
    assert int(ev[0]) == 6
    assert int(ev[1]) == (1 + 2 * w + 3 * w * w) % P


# ---- SURVEY Appendix A (survey-derived cross-checks)
def test_appendix_a_constants():
    assert O.bb_inverse(2) == 1006632961
    roots = {1: 2013265920, 2: 1728404513, 3: 1592366214, 8: 1732600167, 10: 341742893, 12: 1282623253, 15: 2009781145,
             20: 195061667, 22: 570250684, 24: 1003846038, 25: 1149491290, 27: 440564289}
    for k, v in roots.items():
        assert O.root_of_unity(k) == v


def test_kat1_ntt_256():
    x = (np.arange(256, dtype=np.uint64) * 7 + 3) % P
    y = O.ntt(x)
    assert list(y[:8]) == [229248, 1589689095, 848020359, 934041065, 308451887, 41376074, 1713463301, 1337956362]
    assert sha_u64(y) == "bf3697bc7d70c8f86b15faf504724ff11ecaefe903d4b3d4ccc4ef0f15e8d468"
    for log_n, head in ((12, [58718208, 1854648990, 621142519]), (16, [939491321, 721035978, 970451650])):
        n = 1 << log_n
        z = O.ntt((np.arange(n, dtype=np.uint64) * 7 + 3) % P)
        assert list(z[:3]) == head


def test_kat2_to_kat6():
    assert list(O.ntt([1, 2, 3, 0, 0, 0, 0, 0])) == [6, 316882284, 1443543103, 1278030613, 2, 2000481112, 569722814, 431137837]
    c = [3 * i + 1 for i in range(8)]
    assert list(O.domain_fft(c, 8, 7)) == [20657204, 514007405, 601857848, 847117793, 1997142493, 899025193, 1406999153,
                                           1766256603]
    k4 = O.domain_fft(c, 256, 7)
    assert list(k4[:4]) == [20657204, 1655697034, 1139485828, 1866165607]
    assert sha_u64(k4) == "3ce22ed967421a32766eeac91fc480f521e503ddc759362d1bbf1870b429119a"
    k5 = O.fri_fold(k4, O.domain_elements(256, 7), 123456789)
    assert list(k5[:4]) == [1017175357, 155561472, 894059586, 780182327]
    assert sha_u64(k5) == "d4911179c616d7f75036d494e07fd429f5120a2fb251598276bbb20ee58f5248"
    ev = np.array([[((4 * i + k) * 1000003) % P for k in range(4)] for i in range(16)], np.uint64)
    k6 = O.fri_fold_ext(ev, O.domain_elements(16, 7), [5, 6, 7, 8])
    assert list(k6[0]) == [1202226362, 765150799, 926151282, 1685227811]
    assert list(k6[1]) == [1729845515, 1398508500, 1011781982, 569665961]
    assert sha_u64(k6) == "1c331c473814738ba67b7ee197328e66813c31767662c4b7b2f482a8c669208e"


def test_kat7_kat8_merkle_and_transcript():
    want = {4: "082e8e29b028ef12e81530323943dc08834f103e41a73e41c4cbd14b115f85c9",
            3: "3c391efe69e4a5a3e6212efeb617a669e6b73fa053c9064b70c9ade3810a2a93",
            1: "51b09ceccfbec44595dd4241e6e2a693d279b72c899c8f60ec63524fe58b1d4f"}
    for k, h in want.items():
        assert O.commit_values(np.arange(1, k + 1, dtype=np.uint64))[1].hex() == h
    assert O.FiatShamirTranscript().squeeze_challenge() == 837208778


# ---- the two restatements agree (C vs pure Python), and NTT agrees with a naive DFT
def test_c_oracle_matches_python_restatement():
    rng = np.random.default_rng(1)
    for log_n in (1, 2, 5, 8):
        n = 1 << log_n
        x = [int(v) for v in rng.integers(0, P, n)]
        w = R.root_of_unity(log_n)
        assert list(O.ntt(x)) == R.ntt(x, w)
        assert list(O.intt(x)) == R.intt(x, w)
        if n <= 32:
            assert R.ntt(x, w) == R.naive_dft(x, w)
        assert list(O.domain_fft(x[: n // 2 + 1], n, 7)) == R.domain_fft(x[: n // 2 + 1], n, 7)
        assert list(O.domain_ifft(x, 7)) == R.domain_ifft(x, 7)
        xs = R.domain_elements(n, 7)
        assert list(O.fri_fold(x, xs, 99)) == R.fri_fold(x, xs, 99)
    a, b = [1, 2, 3, 4], [5, 6, 7, 8]
    assert list(O.ext_mul(a, b)) == R.ext_mul(a, b)


def test_multithreaded_port_is_bit_identical():
    x = O.random_field(1 << 14)
    assert np.array_equal(O.ntt(x, threads=4), O.ntt(x))
    assert np.array_equal(O.intt(x, threads=3), O.intt(x))


def test_sha256_against_hashlib():
    for n in (0, 1, 9, 25, 33, 49, 55, 56, 63, 64, 65, 119, 120, 1000):
        data = bytes((i * 7 + 1) & 0xFF for i in range(n))
        assert O.sha256(data) == hashlib.sha256(data).digest()
    assert O.hash_leaf(b"abc") == hashlib.sha256(b"\x00abc").digest()      # src/merkle.rs:109-114
    l, r = bytes(range(32)), bytes(range(32, 64))
    assert O.hash_node(l, r) == hashlib.sha256(b"\x01" + l + r).digest()    # src/merkle.rs:117-123


def test_golden_fixtures_reproduce(golden):
    for log_n in (0, 1, 3, 8, 10, 13):
        assert np.array_equal(O.ntt(golden[f"ntt_in_{log_n}"]), golden[f"ntt_out_{log_n}"])
        assert np.array_equal(O.intt(golden[f"rnd_in_{log_n}"]), golden[f"rnd_intt_{log_n}"])
    for log_n in (3, 6, 9):
        assert np.array_equal(O.domain_fft(golden[f"lde_in_{log_n}"], 32 << log_n, 7), golden[f"lde_out_{log_n}"])
    assert np.array_equal(O.fri_fold(golden["fold_in"], O.domain_elements(512, 7), 123456789), golden["fold_out"])
    assert O.commit_values(golden["merkle_vals"], golden["merkle_salts"])[1] == golden["merkle_root_salted"].tobytes()
