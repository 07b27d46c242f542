"""CPU model of the index logic of the blowup-32 LDE expansion pass (toyni_b200/csrc/lde_expand.cuh): one column j of the
4096 x 8192 intermediate array Z[j][k1] = sum_i1 a'[i1 * 4096 + j] w_8192^(i1 k1), computed the way the kernel does
(two radix-16 DIT rounds whose butterfly twiddles carry the coset factors, table w_8192^e for e < 4096 only), against
the direct sum.  Then the whole 2^25 LDE for a small analogue is NOT modelled here: pass 2 is the TMA-staged kernel.
  python tools/lde_model.py"""
import numpy as np

P = 2013265921
GEN27 = 440564289


def root(log_n):
    return pow(GEN27, 1 << (27 - log_n), P)


def brev4(k):
    return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3)


W = root(13)
TW = [pow(W, e, P) for e in range(4096)]  # the kernel's table: first half only


def dit16(x, H):
    """x[k] holds input digit brev4(k); returns x[k] = output digit k of sum_b in_b (w^H w_16^k)^b, twiddle exponents
    (H + 512 kp) << (3 - t) — all below 4096 for H < 512."""
    x = list(x)
    for t in range(4):
        for k in range(16):
            if k & (1 << t):
                continue
            kp = k & ((1 << t) - 1)
            e = (H + 512 * kp) << (3 - t)
            assert e < 4096
            v = x[k + (1 << t)] * TW[e] % P
            x[k], x[k + (1 << t)] = (x[k] + v) % P, (x[k] - v) % P
    return x


def column(xp, nrows):
    """xp: the (shifted) coefficients of one column, rows i1 < nrows <= 512.  Returns Z[k1], k1 < 8192."""
    xin = list(xp) + [0] * (512 - len(xp))
    tile = {}
    for c in range(32):          # lane
        w32 = pow(W, 256 * c, P)
        for b in range(16):      # warp of round A
            x = []
            for k in range(16):
                a = brev4(k)
                v = xin[16 * a + b]
                if 256 + 16 * a + b < nrows:
                    v = (v + xin[256 + 16 * a + b] * w32) % P
                x.append(v)
            x = dit16(x, 16 * c)
            for k in range(16):
                tile[(k, b, c)] = x[k]
    Z = [0] * 8192
    for c in range(32):
        for a1 in range(16):     # warp of round B
            x = dit16([tile[(a1, brev4(k), c)] for k in range(16)], 32 * a1 + c)
            for k in range(16):
                Z[512 * k + 32 * a1 + c] = x[k]
    return Z


def check(nrows, rng, extra=20):
    xp = [int(v) for v in rng.integers(0, P, nrows)]
    got = column(xp, nrows)
    ks = [0, 1, 31, 32, 33, 511, 512, 4095, 4096, 8191] + [int(v) for v in rng.integers(0, 8192, extra)]
    for k1 in ks:
        want = sum(xp[i1] * pow(W, i1 * k1, P) for i1 in range(nrows)) % P
        assert got[k1] == want, (nrows, k1)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for nrows in (256, 257, 300, 512, 100):
        check(nrows, rng)
        print("nrows", nrows, "ok")
