"""Host-side mirror of StarkProver::generate_proof (src/fibonacci.rs:99-310) with the hot path on the GPU.

What runs on the device (through the C ABI): the trace interpolation (coset-free INTT), the
blowup-32 coset LDE of the masked trace polynomial (NTT instead of the reference's per-point Horner,
src/fibonacci.rs:124-128 — same values, the arithmetic is exact), both coset IFFTs (:145,151), every salted /
unsalted Merkle tree (:129,153,206,234-238) and the FRI commit loop (:200-247) with the transcript as its callback.
What stays on the host: the Fiat-Shamir transcript, the constraint / quotient / DEEP element-wise formulas
(numpy), out-of-domain evaluations and the query openings — serial protocol code, out of scope.

The reference draws blinding from thread_rng(); here mask coefficients and salts are explicit inputs so that a proof
is reproducible (and comparable byte for byte with the CPU oracle's)."""
import hashlib

import numpy as np
import torch

from . import device as D
from .domain import get_root_of_unity
from .lib import P

NUM_QUERIES, BLOWUP, COSET_SHIFT = 44, 32, 7  # src/fibonacci.rs:11-16
MASK_DEGREE = 3 * NUM_QUERIES + 8             # :19
_P = np.uint64(P)


class FiatShamirTranscript:
    """src/transcript.rs"""

    def __init__(self):
        self.state = b"toyni-stark-v1"

    def absorb(self, data):
        self.state += bytes(data)

    def absorb_field(self, v):
        self.absorb(int(v).to_bytes(8, "little"))

    def squeeze_challenge(self):
        h = hashlib.sha256(self.state).digest()
        self.state = h
        return int.from_bytes(h[:8], "little") % P

    def squeeze_indices(self, count, mx):
        out, seen = [], set()
        while len(out) < count:
            h = hashlib.sha256(self.state).digest()
            self.state = h
            idx = int.from_bytes(h[:8], "little") % mx
            if idx not in seen:
                seen.add(idx)
                out.append(idx)
        return out


def _mul(a, b):
    return (np.asarray(a, np.uint64) * np.asarray(b, np.uint64)) % _P


def _add(a, b):
    return (np.asarray(a, np.uint64) + np.asarray(b, np.uint64)) % _P


def _sub(a, b):
    return (np.asarray(a, np.uint64) + _P - np.asarray(b, np.uint64)) % _P


def _pow(a, e):
    a = np.asarray(a, np.uint64)
    r = np.ones_like(a)
    while e:
        if e & 1:
            r = _mul(r, a)
        a = _mul(a, a)
        e >>= 1
    return r


def _trim(c):  # Polynomial::new, src/math/polynomial.rs:11-16
    n = c.size
    while n > 0 and c[n - 1] == 0:
        n -= 1
    return c[:n]


def _horner(c, x):  # src/math/polynomial.rs:134-144
    acc = 0
    for v in reversed([int(t) for t in c]):
        acc = (acc * x + v) % P
    return acc


class _Tree:
    """A committed layer: device values + device nodes, host copies made lazily for the openings."""

    def __init__(self, vals_dev, nodes_dev, root, salts_host):
        self.vals_dev, self.nodes_dev, self.root, self.salts = vals_dev, nodes_dev, root, salts_host
        self.n = vals_dev.shape[0]
        self._vals = self._nodes = None

    def open(self, index):  # open_merkle, src/fibonacci.rs:366-374 + MerkleTree::get_proof, src/merkle.rs:50-80
        if self._vals is None:
            self._vals = D.to_host(self.vals_dev)
            self._nodes = self.nodes_dev.cpu().numpy()
        path, position, cur, off, n = [], [], index, 0, self.n
        while n > 1:
            sib = cur + 1 if cur % 2 == 0 else cur - 1
            if sib >= n:
                path.append(self._nodes[off + cur].tobytes())
                position.append(True)
            else:
                path.append(self._nodes[off + sib].tobytes())
                position.append(cur % 2 == 1)
            cur //= 2
            off += n
            n = (n + 1) // 2
        return {"index": int(index), "value": int(self._vals[index]), "path": path, "position": position,
                "salt": b"" if self.salts is None else bytes(self.salts[index])}


def generate_proof(trace_column, mask, salts_trace, salts_quot, salts_fri, device="cuda"):
    trace_len = len(trace_column)
    lde = trace_len * BLOWUP
    g = get_root_of_unity(trace_len.bit_length() - 1)
    g_ext = get_root_of_unity(lde.bit_length() - 1)
    salts_trace = np.ascontiguousarray(salts_trace, np.uint8).reshape(lde, 16)
    salts_quot = np.ascontiguousarray(salts_quot, np.uint8).reshape(lde, 16)
    salts_fri = np.ascontiguousarray(salts_fri, np.uint8).reshape(-1)

    # 1. trace polynomial (INTT on the GPU) + masking T + Z_H * R (:110-121)
    coeffs = _trim(D.to_host(D.coset_ifft_(D.to_device(np.asarray(trace_column, np.uint64), device), 1)))
    tp = np.zeros(trace_len + MASK_DEGREE, np.uint64)
    tp[:coeffs.size] = coeffs
    zr = np.zeros(trace_len + MASK_DEGREE, np.uint64)
    zr[trace_len:] = mask
    zr[:MASK_DEGREE] = _sub(zr[:MASK_DEGREE], mask)
    trace_poly = _trim(_add(tp, zr))
    # LDE on the shifted domain + commit (:124-130), all on the device
    shifted_dev = D.coset_fft(D.to_device(np.array([0, 1], np.uint64), device), lde, COSET_SHIFT)  # the coset itself
    trace_lde_dev = D.coset_fft(D.to_device(trace_poly, device), lde, COSET_SHIFT)
    nodes, root = D.merkle_commit(trace_lde_dev, torch.from_numpy(salts_trace).to(device))
    trace_tree = _Tree(trace_lde_dev, nodes, root, salts_trace)
    # 2. constraint and quotient (:133-153): element-wise on the host, both IFFTs on the device
    xs = D.to_host(shifted_dev)
    t_x = D.to_host(trace_lde_dev)
    t_gx, t_ggx = np.roll(t_x, -BLOWUP), np.roll(t_x, -2 * BLOWUP)  # T(g x_i) = trace_lde[(i + 32) % N]
    b1 = _sub(xs, np.uint64(pow(g, trace_len - 1, P)))
    b2 = _sub(xs, np.uint64(pow(g, trace_len - 2, P)))
    c_evals = _mul(_mul(_sub(t_ggx, _add(t_gx, t_x)), b1), b2)
    c_poly = _trim(D.to_host(D.coset_ifft_(D.to_device(c_evals, device), COSET_SHIFT)))
    c_on_domain = D.to_host(D.coset_fft(D.to_device(c_poly, device), lde, COSET_SHIFT))  # c_poly.evaluate(x)
    q_evals = _mul(c_on_domain, _pow(_sub(_pow(xs, trace_len), np.uint64(1)), P - 2))
    q_dev = D.to_device(q_evals, device)
    q_poly = _trim(D.to_host(D.coset_ifft_(q_dev.clone(), COSET_SHIFT)))
    nodes, root = D.merkle_commit(q_dev, torch.from_numpy(salts_quot).to(device))
    quot_tree = _Tree(q_dev, nodes, root, salts_quot)
    # 3. Fiat-Shamir: z outside both domains (:156-161, :378-399)
    tr = FiatShamirTranscript()
    tr.absorb(trace_tree.root)
    tr.absorb(quot_tree.root)
    ext_set = set(int(v) for v in D.to_host(D.coset_fft(D.to_device(np.array([0, 1], np.uint64), device), lde, 1)))
    shift_set = set(int(v) for v in xs)
    while True:
        z = tr.squeeze_challenge()
        if (z not in ext_set and z not in shift_set and g_ext * z % P not in shift_set
                and g_ext * g_ext % P * z % P not in shift_set):
            break
    # 4. OOD evaluations (:164-183)
    t_z, t_gz, t_ggz = _horner(trace_poly, z), _horner(trace_poly, g * z % P), _horner(trace_poly, g * g % P * z % P)
    q_z = _horner(q_poly, z)
    c_z = (t_ggz - (t_gz + t_z)) % P * ((z - pow(g, trace_len - 1, P)) % P) % P * ((z - pow(g, trace_len - 2, P)) % P) % P
    assert c_z == q_z * ((pow(z, trace_len, P) - 1) % P) % P, "Constraint check at z failed"  # :173-177
    for v in (t_z, t_gz, t_ggz, q_z):
        tr.absorb_field(v)
    # 5. DEEP polynomial (:186-198)
    inv_xz = _pow(_sub(xs, np.uint64(z)), P - 2)
    d_evals = _mul(_sub(q_evals, np.uint64(q_z)), inv_xz)
    d_evals = _add(d_evals, _mul(_sub(t_ggx, np.uint64(t_ggz)), inv_xz))
    d_evals = _add(d_evals, _mul(_sub(t_gx, np.uint64(t_gz)), inv_xz))
    d_evals = _add(d_evals, _mul(_sub(t_x, np.uint64(t_z)), inv_xz))
    # 6. FRI commit loop on the device, transcript as the callback (:200-247)
    bound = 1 << (trace_len + MASK_DEGREE - 1).bit_length()
    final_size = lde // bound

    def challenge(root, layer):
        tr.absorb(root)
        return tr.squeeze_challenge()

    layers, nodes_l, roots = D.fri_commit(D.to_device(d_evals, device), COSET_SHIFT, final_size,
                                          torch.from_numpy(salts_fri).to(device), challenge=challenge)
    tr.absorb(roots[-1])  # the callback absorbed every root but the last (no fold follows it), :239-242
    trees, off = [], 0
    for k, (lay, nd) in enumerate(zip(layers, nodes_l)):
        s = None
        if k < len(layers) - 1:
            s = salts_fri[16 * off:16 * (off + lay.shape[0])].reshape(-1, 16)
            off += lay.shape[0]
        trees.append(_Tree(lay, nd, roots[k], s))
    # 7. query phase (:250-295)
    queries = tr.squeeze_indices(NUM_QUERIES, lde // 2)
    qps = []
    for qi in queries:
        qp = {"index": qi,
              "deep_opening": trees[0].open(qi), "deep_opening_pair": trees[0].open(qi + lde // 2),
              "trace_opening": trace_tree.open(qi), "trace_opening_g": trace_tree.open((qi + BLOWUP) % lde),
              "trace_opening_gg": trace_tree.open((qi + 2 * BLOWUP) % lde), "quotient_opening": quot_tree.open(qi),
              "fri_openings": []}
        idx = qi
        for k in range(1, len(trees) - 1):
            half = trees[k].n // 2
            idx %= half
            qp["fri_openings"].append((trees[k].open(idx), trees[k].open(idx + half)))
        qps.append(qp)
    return {"trace_len": trace_len, "lde_size": lde, "trace_commitment": trace_tree.root,
            "quotient_commitment": quot_tree.root, "t_z": t_z, "t_gz": t_gz, "t_ggz": t_ggz, "q_z": q_z,
            "fri_commitments": roots, "fri_final_layer": [int(v) for v in D.to_host(layers[-1])], "query_proofs": qps}
