"""CPU model of the index logic of the TMA-staged two-pass NTT kernel (toyni_b200/csrc/ntt_pass_v7.cuh).

Not a product path: it restates, in numpy, exactly which shared-memory row every lane of the kernel reads and
writes in each of its three radix-G rounds, which twiddle-table entry each butterfly uses, and how the two passes
chain (transposing store of pass 1, inter-pass twiddle as A[d2] * beta(d1, d0) in pass 2).  It is checked against a
naive DFT at G = 4 (n = 2^12) by tests/test_host_logic.py and was used to derive the constants of the kernel, where
G = 16 (R = 4096 rows per tile).

Reference transform: src/ntt.rs:24-53 (natural order in and out).
"""
import numpy as np

P = 2013265921
GEN27 = 440564289


def root(log_n):
    return pow(GEN27, 1 << (27 - log_n), P)


def brev(v, bits):
    r = 0
    for i in range(bits):
        r |= ((v >> i) & 1) << (bits - 1 - i)
    return r


def mulmod(a, b):
    return (a.astype(np.uint64) * np.uint64(b) if np.isscalar(b) else a.astype(np.uint64) * b.astype(np.uint64)) % np.uint64(P)


class TileModel:
    """One tile: R = G^3 rows x C columns, three radix-G DIT rounds, positions as in the kernel."""

    def __init__(self, g, omega_R):
        self.g = g
        self.G = 1 << g
        self.R = self.G ** 3
        self.wR = omega_R
        # twiddle table: w_R^i, i < R/2 (the kernel's shared-memory Shoup table)
        self.tw = [pow(omega_R, i, P) for i in range(self.R // 2)]

    def phys(self, q):
        """physical row of logical position q = (q2, q1, q0): chunks (q1) are contiguous, and the low two bits of q0 are
        XOR-ed with the low two bits of q2 (bank spreading of the round-1 stores)"""
        G = self.G
        q2, q1, q0 = q // (G * G), (q // G) % G, q % G
        return q1 * G * G + q2 * G + (q0 ^ (q2 & 3))

    def dit(self, x, tw_of):
        """radix-G DIT on register arrays x[kk] (each an array over lanes/columns); tw_of(t, kp) -> per-lane factor"""
        g, G = self.g, self.G
        for t in range(g):
            for kk in range(G):
                if kk & (1 << t):
                    continue
                kp = kk & ((1 << t) - 1)
                w = tw_of(t, kp)
                v = (x[kk + (1 << t)] * w) % P
                u = x[kk]
                x[kk] = (u + v) % P
                x[kk + (1 << t)] = (u + P - v) % P

    def run(self, tile, pre=None):
        """tile[d][c] (natural rows) -> X[f][c]; pre = None or (A[d2][c], beta(k, r)[c]) factors of the pre-twiddle"""
        g, G, R = self.g, self.G, self.R
        C = tile.shape[1]
        tile = tile.astype(object)
        main = np.zeros((R, C), dtype=object)
        lanes = np.arange(G)  # r
        # ---- round 1, chunk k = d1; landing[d2*G + d0] = tile[d2*G*G + k*G + d0] (the TMA box order)
        for k in range(G):
            landing = np.zeros((G * G, C), dtype=object)
            for d2 in range(G):
                for d0 in range(G):
                    landing[d2 * G + d0] = tile[d2 * G * G + k * G + d0]
            x = []
            for kk in range(G):
                d2 = brev(kk, g)
                v = landing[d2 * G + lanes]  # lane r reads row d2*G + r
                if pre is not None:
                    v = (v * pre[0][d2][None, :]) % P
                x.append(v)
            # constant twiddles w_G^(kp * 2^(g-1-t))
            wG = pow(self.wR, R // G, P)
            self.dit(x, lambda t, kp: pow(wG, kp << (g - 1 - t), P))
            for kk in range(G):
                v = x[kk]
                if pre is not None:
                    v = (v * pre[1](k)) % P  # beta[r][c]
                q = lanes * G * G + k * G + kk
                main[[self.phys(int(qq)) for qq in q]] = v
        # ---- round 2, per q2; lane r = q0
        for q2 in range(G):
            x = []
            for kk in range(G):
                q = q2 * G * G + brev(kk, g) * G + lanes
                x.append(main[[self.phys(int(qq)) for qq in q]])
            # w_{2^(g+t+1)}^(r + G*kp) = tw[(r + G*kp) << (3g - (g+t+1))]
            def tw2(t, kp):
                return np.array([self.tw[(int(r) + G * kp) << (2 * g - t - 1)] for r in lanes], dtype=object)[:, None]
            self.dit(x, tw2)
            for kk in range(G):
                q = q2 * G * G + kk * G + lanes
                main[[self.phys(int(qq)) for qq in q]] = x[kk]
        # ---- round 3, chunk k = q1; lane r = q0; outputs f = kk*G*G + k*G + r
        out = np.zeros((R, C), dtype=object)
        for k in range(G):
            x = []
            for kk in range(G):
                q = brev(kk, g) * G * G + k * G + lanes
                x.append(main[[self.phys(int(qq)) for qq in q]])
            def tw3(t, kp):
                return np.array([self.tw[(int(r) + G * k + G * G * kp) << (g - t - 1)] for r in lanes], dtype=object)[:, None]
            self.dit(x, tw3)
            for kk in range(G):
                out[kk * G * G + k * G + lanes] = x[kk]
        return out


def two_pass_ntt(x, g, C=8, inverse=False):
    """n = R*R transform exactly as the kernel chains its two passes.  x: list/array of n canonical values."""
    G = 1 << g
    R = G ** 3
    n = R * R
    log_n = 6 * g
    wn = root(log_n)
    if inverse:
        wn = pow(wn, P - 2, P)
    wR = pow(wn, n // R, P)
    tm = TileModel(g, wR)
    x = np.array(x, dtype=object).reshape(R, R)  # [d][col]
    # pass 1: column tiles, transposing store out1[col*R + e]
    out1 = np.zeros((R, R), dtype=object)  # [col][e]
    for col0 in range(0, R, C):
        X = tm.run(x[:, col0:col0 + C])
        out1[col0:col0 + C, :] = X.T
    # pass 2: rows d = col of pass 1, tile columns e1; pre-twiddle t_c^d, t_c = wn^e1 (times n^-1 for the inverse)
    scale = pow(n, P - 2, P) if inverse else 1
    out = np.zeros((R, R), dtype=object)  # [f][e1]
    for col0 in range(0, R, C):
        e1 = np.arange(col0, col0 + C)
        A = [np.array([pow(wn, (G * G * d2 * int(e)) % n, P) for e in e1], dtype=object) for d2 in range(G)]
        U = [np.array([pow(wn, (G * k * int(e)) % n, P) for e in e1], dtype=object) for k in range(G)]
        V = [np.array([pow(wn, (r * int(e)) % n, P) * scale % P for e in e1], dtype=object) for r in range(G)]
        def beta(k):
            return np.array([(U[k] * V[r]) % P for r in range(G)], dtype=object)  # [r][c]
        out[:, col0:col0 + C] = tm.run(out1[:, col0:col0 + C], pre=(A, beta))
    return out.reshape(-1)


def naive_dft(x, w):
    n = len(x)
    return [sum(int(x[j]) * pow(w, j * k, P) for j in range(n)) % P for k in range(n)]


if __name__ == "__main__":
    g = 2
    n = (1 << g) ** 6
    rng = np.random.default_rng(1)
    x = [int(v) for v in rng.integers(0, P, n)]
    got = two_pass_ntt(x, g)
    # check a sample of outputs against the definition
    w = root(6 * g)
    for k in list(range(8)) + [n // 2, n - 1, 1234 % n]:
        want = sum(x[j] * pow(w, j * k, P) for j in range(n)) % P
        assert int(got[k]) == want, (k, got[k], want)
    inv = two_pass_ntt([int(v) for v in got], g, inverse=True)
    assert [int(v) for v in inv] == x
    print("v7 model ok, n =", n)
