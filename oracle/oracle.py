"""ctypes binding of oracle/liboracle.so (the C restatement in toyni_oracle.c).

TEST INFRASTRUCTURE ONLY.  Arrays are numpy uint64, one canonical BabyBear value per entry
(the reference's storage, src/babybear.rs:10-14); Ext arrays have shape (n, 4).
"""
import ctypes as C
import os
import subprocess

import numpy as np

P = 2013265921
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "toyni_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/cuda/ntt_kernel.cu"):
        ref = os.path.join(_HERE, "_ref", "libntt_cuda_ref.so")
        if force or not os.path.exists(ref):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


class Transcript(C.Structure):
    _fields_ = [("state", u8p), ("len", C.c_size_t), ("cap", C.c_size_t)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        sz = C.c_size_t
        u64 = C.c_uint64
        sig = {
            "to_bb_new": ([u64], u64), "to_bb_add": ([u64, u64], u64), "to_bb_sub": ([u64, u64], u64),
            "to_bb_mul": ([u64, u64], u64), "to_bb_neg": ([u64], u64), "to_bb_pow": ([u64, u64], u64),
            "to_bb_inverse": ([u64], u64), "to_bb_root_of_unity": ([C.c_uint32], u64),
            "to_ext_add": ([u64p, u64p, u64p], None), "to_ext_sub": ([u64p, u64p, u64p], None),
            "to_ext_mul": ([u64p, u64p, u64p], None), "to_ext_mul_base": ([u64p, u64, u64p], None),
            "to_ext_inverse": ([u64p, u64p], None),
            "to_ntt": ([u64p, sz, u64], None), "to_intt": ([u64p, sz, u64], None),
            "to_ntt_mt": ([u64p, sz, u64, C.c_int], None), "to_intt_mt": ([u64p, sz, u64, C.c_int], None),
            "to_roots_of_unity_domain": ([u64p, sz], None),
            "to_domain_elements": ([u64p, sz, u64], None),
            "to_domain_fft": ([u64p, sz, sz, u64, u64p], None), "to_domain_ifft": ([u64p, sz, u64, u64p], None),
            "to_domain_fft_ext": ([u64p, sz, sz, u64, u64p], None), "to_domain_ifft_ext": ([u64p, sz, u64, u64p], None),
            "to_fri_fold": ([u64p, sz, u64p, u64, u64p], None), "to_fri_fold_ext": ([u64p, sz, u64p, u64p, u64p], None),
            "to_sha256": ([u8p, sz, u8p], None), "to_hash_leaf": ([u8p, sz, u8p], None),
            "to_hash_node": ([u8p, u8p, u8p], None), "to_merkle_node_count": ([sz], sz),
            "to_merkle_build": ([u8p, sz, sz, u8p, u8p], None),
            "to_commit_values": ([u64p, sz, C.c_int, u8p, u8p, u8p], None),
            "to_merkle_open": ([u8p, sz, sz, u8p, u8p], sz),
            "to_merkle_verify": ([u8p, sz, u8p, u8p, sz, u8p], C.c_int),
            "to_transcript_init": ([C.POINTER(Transcript)], None), "to_transcript_free": ([C.POINTER(Transcript)], None),
            "to_transcript_absorb": ([C.POINTER(Transcript), u8p, sz], None),
            "to_transcript_squeeze": ([C.POINTER(Transcript)], u64),
            "to_transcript_squeeze_ext": ([C.POINTER(Transcript), u64p], None),
            "to_transcript_squeeze_indices": ([C.POINTER(Transcript), sz, sz, u64p], None),
            "to_fri_commit": ([u64p, sz, u64, sz, u8p, C.POINTER(Transcript), u64p, u8p, u64p], sz),
            "to_fri_commit_ext": ([u64p, sz, u64, sz, u8p, C.POINTER(Transcript), u64p, u8p, u64p], sz),
            "to_set_threads": ([C.c_int], None), "to_get_threads": ([], C.c_int),
            "to_fill_random": ([u64p, sz, u64], None), "to_fill_random_bytes": ([u8p, sz, u64], None),
        }
        for name, (args, res) in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _LIB = L
    return _LIB


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _p8(a):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def set_threads(threads):
    """Thread count of the element-wise oracle loops (coset shift, folds, hashing) and of the transforms inside
    domain_fft / domain_ifft; 1 = the reference's serial loops verbatim.  Same bits either way."""
    lib().to_set_threads(int(threads))


# ---- field -------------------------------------------------------------------------------
def bb_mul(a, b): return lib().to_bb_mul(a, b)
def bb_add(a, b): return lib().to_bb_add(a, b)
def bb_sub(a, b): return lib().to_bb_sub(a, b)
def bb_pow(a, e): return lib().to_bb_pow(a, e)
def bb_inverse(a): return lib().to_bb_inverse(a)
def root_of_unity(log_n): return lib().to_bb_root_of_unity(log_n)


def ext_mul(a, b):
    a, b = _arr(a), _arr(b)
    r = np.zeros(4, np.uint64)
    lib().to_ext_mul(_p64(a), _p64(b), _p64(r))
    return r


def ext_inverse(a):
    a = _arr(a)
    r = np.zeros(4, np.uint64)
    lib().to_ext_inverse(_p64(a), _p64(r))
    return r


# ---- transforms --------------------------------------------------------------------------
def ntt(values, omega=None, threads=1):
    v = _arr(values).copy()
    n = v.size
    if omega is None:
        omega = root_of_unity(n.bit_length() - 1)
    if threads > 1:
        lib().to_ntt_mt(_p64(v), n, omega, threads)
    else:
        lib().to_ntt(_p64(v), n, omega)
    return v


def intt(values, omega=None, threads=1):
    v = _arr(values).copy()
    n = v.size
    if omega is None:
        omega = root_of_unity(n.bit_length() - 1)
    if threads > 1:
        lib().to_intt_mt(_p64(v), n, omega, threads)
    else:
        lib().to_intt(_p64(v), n, omega)
    return v


def ntt_inplace(v, threads=1):
    """In-place forward NTT with the canonical root (timing helper: no copies)."""
    n = v.size
    omega = root_of_unity(n.bit_length() - 1)
    if threads > 1:
        lib().to_ntt_mt(_p64(v), n, omega, threads)
    else:
        lib().to_ntt(_p64(v), n, omega)


def roots_of_unity_domain(n):
    out = np.zeros(n, np.uint64)
    lib().to_roots_of_unity_domain(_p64(out), n)
    return out


def domain_elements(size, shift=1):
    out = np.zeros(size, np.uint64)
    lib().to_domain_elements(_p64(out), size, shift)
    return out


def domain_fft(coeffs, size, shift=1):
    c = _arr(coeffs)
    out = np.zeros(size, np.uint64)
    lib().to_domain_fft(_p64(c), c.size, size, shift, _p64(out))
    return out


def domain_ifft(evals, shift=1):
    e = _arr(evals)
    out = np.zeros(e.size, np.uint64)
    lib().to_domain_ifft(_p64(e), e.size, shift, _p64(out))
    return out


def domain_fft_ext(coeffs, size, shift=1):
    c = _arr(coeffs).reshape(-1, 4)
    out = np.zeros((size, 4), np.uint64)
    lib().to_domain_fft_ext(_p64(c), c.shape[0], size, shift, _p64(out))
    return out


def domain_ifft_ext(evals, shift=1):
    e = _arr(evals).reshape(-1, 4)
    out = np.zeros_like(e)
    lib().to_domain_ifft_ext(_p64(e), e.shape[0], shift, _p64(out))
    return out


# ---- FRI ---------------------------------------------------------------------------------
def fri_fold(evals, xs, beta):
    e, x = _arr(evals), _arr(xs)
    out = np.zeros(e.size // 2, np.uint64)
    lib().to_fri_fold(_p64(e), e.size, _p64(x), beta, _p64(out))
    return out


def fri_fold_ext(evals, xs, beta):
    e, x, b = _arr(evals).reshape(-1, 4), _arr(xs), _arr(beta)
    out = np.zeros((e.shape[0] // 2, 4), np.uint64)
    lib().to_fri_fold_ext(_p64(e), e.shape[0], _p64(x), _p64(b), _p64(out))
    return out


# ---- hashing -----------------------------------------------------------------------------
def sha256(data: bytes) -> bytes:
    d = np.frombuffer(data, np.uint8) if len(data) else np.zeros(0, np.uint8)
    d = np.ascontiguousarray(d)
    out = np.zeros(32, np.uint8)
    lib().to_sha256(_p8(d) if len(data) else None, len(data), _p8(out))
    return out.tobytes()


def hash_leaf(data: bytes) -> bytes:
    d = np.ascontiguousarray(np.frombuffer(data, np.uint8))
    out = np.zeros(32, np.uint8)
    lib().to_hash_leaf(_p8(d), len(data), _p8(out))
    return out.tobytes()


def hash_node(l: bytes, r: bytes) -> bytes:
    a = np.ascontiguousarray(np.frombuffer(l, np.uint8))
    b = np.ascontiguousarray(np.frombuffer(r, np.uint8))
    out = np.zeros(32, np.uint8)
    lib().to_hash_node(_p8(a), _p8(b), _p8(out))
    return out.tobytes()


def merkle_node_count(nleaves):
    return lib().to_merkle_node_count(nleaves)


def merkle_build(leaves):
    """leaves: list of equal-length byte strings. Returns (nodes uint8[count,32], root bytes)."""
    n, ll = len(leaves), len(leaves[0])
    flat = np.ascontiguousarray(np.frombuffer(b"".join(leaves), np.uint8))
    nodes = np.zeros((merkle_node_count(n), 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().to_merkle_build(_p8(flat), n, ll, _p8(nodes), _p8(root))
    return nodes, root.tobytes()


def commit_values(values, salts=None, limbs=1):
    """Prover-style commit (src/fibonacci.rs:340-363). salts: uint8[n,16] or None (unsalted)."""
    v = _arr(values)
    n = v.size // limbs
    nodes = np.zeros((merkle_node_count(n), 32), np.uint8)
    root = np.zeros(32, np.uint8)
    s = None if salts is None else np.ascontiguousarray(salts, dtype=np.uint8)
    lib().to_commit_values(_p64(v), n, limbs, _p8(s), _p8(nodes), _p8(root))
    return nodes, root.tobytes()


def merkle_open(nodes, nleaves, index):
    path = np.zeros((64, 32), np.uint8)
    pos = np.zeros(64, np.uint8)
    d = lib().to_merkle_open(_p8(np.ascontiguousarray(nodes)), nleaves, index, _p8(path), _p8(pos))
    return path[:d].copy(), pos[:d].copy()


def merkle_verify(leaf: bytes, path, pos, root: bytes):
    lf = np.ascontiguousarray(np.frombuffer(leaf, np.uint8))
    rt = np.ascontiguousarray(np.frombuffer(root, np.uint8))
    path = np.ascontiguousarray(path, dtype=np.uint8)
    pos = np.ascontiguousarray(pos, dtype=np.uint8)
    return bool(lib().to_merkle_verify(_p8(lf), len(leaf), _p8(path), _p8(pos), len(pos), _p8(rt)))


# ---- transcript --------------------------------------------------------------------------
class FiatShamirTranscript:
    """src/transcript.rs"""

    def __init__(self):
        self._t = Transcript()
        lib().to_transcript_init(C.byref(self._t))

    def __del__(self):
        try:
            lib().to_transcript_free(C.byref(self._t))
        except Exception:
            pass

    def absorb(self, data: bytes):
        d = np.ascontiguousarray(np.frombuffer(data, np.uint8))
        lib().to_transcript_absorb(C.byref(self._t), _p8(d), len(data))

    def absorb_field(self, v):
        self.absorb(int(v).to_bytes(8, "little"))

    def squeeze_challenge(self):
        return lib().to_transcript_squeeze(C.byref(self._t))

    def squeeze_ext_challenge(self):
        out = np.zeros(4, np.uint64)
        lib().to_transcript_squeeze_ext(C.byref(self._t), _p64(out))
        return out

    def squeeze_indices(self, count, mx):
        out = np.zeros(count, np.uint64)
        lib().to_transcript_squeeze_indices(C.byref(self._t), count, mx, _p64(out))
        return [int(x) for x in out]

    def state(self) -> bytes:
        return bytes(self._t.state[: self._t.len])


def fri_commit(layer0, shift, final_size, salts, transcript=None, ext=False):
    """src/fibonacci.rs:200-247 with explicit salts. Returns (layers list, roots list, betas)."""
    limbs = 4 if ext else 1
    l0 = _arr(layer0)
    n = l0.size // limbs
    t = transcript or FiatShamirTranscript()
    total = 0
    m = n
    sizes = []
    while True:
        sizes.append(m)
        total += m
        if m <= final_size:
            break
        m //= 2
    layers = np.zeros(total * limbs, np.uint64)
    roots = np.zeros((len(sizes), 32), np.uint8)
    betas = np.zeros(max(1, (len(sizes) - 1)) * limbs, np.uint64)
    s = np.ascontiguousarray(salts, dtype=np.uint8)
    fn = lib().to_fri_commit_ext if ext else lib().to_fri_commit
    folds = fn(_p64(l0), n, shift, final_size, _p8(s), C.byref(t._t), _p64(layers), _p8(roots), _p64(betas))
    assert folds == len(sizes) - 1
    out, off = [], 0
    for m in sizes:
        a = layers[off * limbs:(off + m) * limbs]
        out.append(a.reshape(m, 4) if ext else a)
        off += m
    b = betas[: folds * limbs]
    return out, [r.tobytes() for r in roots], (b.reshape(folds, 4) if ext else b)


# ---- synthetic data ----------------------------------------------------------------------
SEED = 0x70796E69  # "toyni" 


def random_field(n, seed=SEED):
    out = np.zeros(n, np.uint64)
    lib().to_fill_random(_p64(out), n, seed)
    return out


def random_bytes(n, seed=SEED + 2):
    out = np.zeros(n, np.uint8)
    lib().to_fill_random_bytes(_p8(out), n, seed)
    return out
