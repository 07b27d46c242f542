"""Independent pure-Python restatement of the same reference lines, for small sizes only.

TEST INFRASTRUCTURE ONLY.  Written separately from toyni_oracle.c (Python ints, hashlib) so that
the two restatements can be checked against each other (SURVEY 8c "how to trust the restatement").
"""
import hashlib

P = 2013265921  # src/babybear.rs:8
W = 11  # src/ext.rs:20


def root_of_unity(log_n):  # src/babybear.rs:118-126
    assert log_n <= 27
    return pow(440564289, 1 << (27 - log_n), P)


def bit_reverse(x, log_n):  # src/ntt.rs:14-21
    r = 0
    for _ in range(log_n):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def ntt(values, omega):  # src/ntt.rs:24-53
    v = list(values)
    n = len(v)
    log_n = n.bit_length() - 1
    for i in range(n):
        j = bit_reverse(i, log_n)
        if i < j:
            v[i], v[j] = v[j], v[i]
    ln = 2
    while ln <= n:
        w_len = pow(omega, n // ln, P)
        for i in range(0, n, ln):
            w = 1
            for j in range(ln // 2):
                u = v[i + j]
                t = v[i + j + ln // 2] * w % P
                v[i + j] = (u + t) % P
                v[i + j + ln // 2] = (u - t) % P
                w = w * w_len % P
        ln *= 2
    return v


def intt(values, omega):  # src/ntt.rs:56-66
    n = len(values)
    v = ntt(values, pow(omega, n - 1, P))
    inv_n = pow(n % P, P - 2, P)
    return [x * inv_n % P for x in v]


def naive_dft(values, omega):
    n = len(values)
    return [sum(values[j] * pow(omega, j * k, P) for j in range(n)) % P for k in range(n)]


def domain_elements(size, shift=1):  # src/math/domain.rs:61-69
    omega = root_of_unity(size.bit_length() - 1)
    out, cur = [], shift
    for _ in range(size):
        out.append(cur)
        cur = cur * omega % P
    return out


def domain_fft(coeffs, size, shift=1):  # src/math/domain.rs:107-123,154-162
    v = list(coeffs)[:size] + [0] * max(0, size - len(coeffs))
    if shift != 1:
        sp = 1
        for i in range(size):
            v[i] = v[i] * sp % P
            sp = sp * shift % P
    return ntt(v, root_of_unity(size.bit_length() - 1))


def domain_ifft(evals, shift=1):  # src/math/domain.rs:85-102,165-174
    size = len(evals)
    v = intt(evals, root_of_unity(size.bit_length() - 1))
    if shift != 1:
        si = pow(shift, P - 2, P)
        sp = 1
        for i in range(size):
            v[i] = v[i] * sp % P
            sp = sp * si % P
    return v


def horner(coeffs, x):  # src/math/polynomial.rs:134-144
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


def ext_mul(a, b):  # src/ext.rs:178-192
    return [
        (a[0] * b[0] + W * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1])) % P,
        (a[0] * b[1] + a[1] * b[0] + W * (a[2] * b[3] + a[3] * b[2])) % P,
        (a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + W * (a[3] * b[3])) % P,
        (a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]) % P,
    ]


def fri_fold(evals, xs, beta):  # src/math/fri.rs:27-48
    half = len(evals) // 2
    hi = pow(2, P - 2, P)
    out = []
    for i in range(half):
        a, b, x = evals[i], evals[i + half], xs[i]
        avg = (a + b) * hi % P
        diff = (a - b) * hi % P
        out.append((avg + diff * beta % P * pow(x, P - 2, P)) % P)
    return out


def fri_fold_ext(evals, xs, beta):  # src/math/fri.rs:7-25
    half = len(evals) // 2
    hi = pow(2, P - 2, P)
    out = []
    for i in range(half):
        a, b = evals[i], evals[i + half]
        x_inv = pow(xs[i], P - 2, P)
        avg = [(a[k] + b[k]) * hi % P for k in range(4)]
        diff = [(a[k] - b[k]) * hi % P for k in range(4)]
        t = ext_mul(ext_mul(diff, beta), [x_inv, 0, 0, 0])
        out.append([(avg[k] + t[k]) % P for k in range(4)])
    return out


def hash_leaf(data: bytes) -> bytes:  # src/merkle.rs:109-114
    return hashlib.sha256(b"\x00" + data).digest()


def hash_node(l: bytes, r: bytes) -> bytes:  # src/merkle.rs:117-123
    return hashlib.sha256(b"\x01" + l + r).digest()


def merkle_levels(leaves):  # src/merkle.rs:25-48
    cur = [hash_leaf(x) for x in leaves]
    levels = [cur]
    while len(cur) > 1:
        nxt = []
        for i in range(0, len(cur), 2):
            r = cur[i + 1] if i + 1 < len(cur) else cur[i]
            nxt.append(hash_node(cur[i], r))
        cur = nxt
        levels.append(cur)
    return levels


def merkle_root(leaves):
    return merkle_levels(leaves)[-1][0]


class Transcript:  # src/transcript.rs
    def __init__(self):
        self.state = b"toyni-stark-v1"

    def absorb(self, data):
        self.state += data

    def squeeze_challenge(self):
        h = hashlib.sha256(self.state).digest()
        self.state = h
        return int.from_bytes(h[:8], "little") % P

    def squeeze_indices(self, count, mx):
        out = []
        while len(out) < count:
            h = hashlib.sha256(self.state).digest()
            self.state = h
            idx = int.from_bytes(h[:8], "little") % mx
            if idx not in out:
                out.append(idx)
        return out
