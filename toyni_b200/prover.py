"""Host-side mirror of StarkProver::generate_proof (src/fibonacci.rs:99-310) with the hot path on the GPU.

What runs on the device (through the C ABI): the trace interpolation (coset-free INTT), the
blowup-32 coset LDE of the masked trace polynomial (NTT instead of the reference's per-point Horner,
src/fibonacci.rs:124-128 — same values, the arithmetic is exact), both coset IFFTs (:145,151), every salted /
unsalted Merkle tree (:129,153,206,234-238) and the FRI commit loop (:200-247) with the transcript as its callback.
the constraint / quotient / DEEP element-wise formulas (closed forms over the coset, batched
inversion), the out-of-domain evaluations and the openings of the whole query set (rank 3).
What stays on the host: the Fiat-Shamir transcript, the trace_len + 140 coefficients of the masked trace polynomial
and the assembly of the proof object — serial protocol code, out of scope.

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

    def squeeze_ext_challenge(self):  # src/transcript.rs:43-50: four independent base squeezes
        return [self.squeeze_challenge() for _ in range(4)]

    def absorb_ext(self, limbs):  # src/transcript.rs:53-55: the 32 bytes of Ext::to_bytes (src/ext.rs:83-89)
        for v in limbs:
            self.absorb_field(v)

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




def _add(a, b):  # canonical in, canonical out: a conditional subtraction instead of numpy's slow 64-bit modulo
    s = np.asarray(a, np.uint64) + np.asarray(b, np.uint64)
    return s - _P * (s >= _P).astype(np.uint64)


def _sub(a, b):
    s = np.asarray(a, np.uint64) + _P - np.asarray(b, np.uint64)
    return s - _P * (s >= _P).astype(np.uint64)




def _trim(c):  # Polynomial::new, src/math/polynomial.rs:11-16
    n = c.size
    while n > 0 and c[n - 1] == 0:
        n -= 1
    return c[:n]




class _Tree:
    """A committed layer: device values + device nodes; openings are gathered on the device (one launch per query
    set), nothing but the opened paths, values and salts ever crosses PCIe."""

    def __init__(self, vals_dev, nodes_dev, root, salts_dev):
        self.vals_dev, self.nodes_dev, self.root, self.salts_dev = vals_dev, nodes_dev, root, salts_dev
        self.n = vals_dev.shape[0]

    def open_many(self, indices):  # open_merkle, src/fibonacci.rs:366-374 + MerkleTree::get_proof, src/merkle.rs:50-80
        indices = [int(i) for i in indices]
        paths, pos = D.merkle_open_batch(self.nodes_dev, self.n, indices)
        vals = D.gather(self.vals_dev, indices).view(np.uint32).reshape(-1)
        salts = None if self.salts_dev is None else D.gather(self.salts_dev, indices)
        return [{"index": idx, "value": int(vals[k]), "path": [paths[k, d].tobytes() for d in range(paths.shape[1])],
                 "position": [bool(b) for b in pos[k]], "salt": b"" if salts is None else salts[k].tobytes()}
                for k, idx in enumerate(indices)]


def _as_salts(s, device):
    if isinstance(s, torch.Tensor):
        return s.to(device).contiguous().view(-1, 16)
    return torch.from_numpy(np.ascontiguousarray(s, np.uint8).reshape(-1, 16)).to(device)


def generate_proof_native(trace_column, mask, salts_trace, salts_quot, salts_fri, device="cuda", as_bytes=False):
    """The same proof from ONE call into the library (toyni_prove_fibonacci: the loop of toyni_prover.hpp, compiled host
    code, transcript included) — no Python between the stages.  Salts may be numpy arrays (copied by the library) or CUDA
    uint8 tensors (used in place).  Returns the proof dict of generate_proof, or its canonical bytes."""
    import ctypes as C

    from .lib import check, lib
    from .proof import deserialize_proof
    D._bind_stream()
    trace = np.ascontiguousarray(np.asarray(trace_column, dtype=np.uint64))
    mask = np.ascontiguousarray(np.asarray(mask, dtype=np.uint64).reshape(-1))
    assert mask.size == MASK_DEGREE, f"mask: {mask.size} coefficients given, {MASK_DEGREE} required"
    L = lib()
    need = L.toyni_fri_salt_bytes(trace.size)
    on_dev = all(isinstance(s, torch.Tensor) and s.is_cuda for s in (salts_trace, salts_quot, salts_fri))
    if on_dev:
        keep = [s.contiguous().view(torch.uint8).reshape(-1) for s in (salts_trace, salts_quot, salts_fri)]
        ptrs = [C.c_void_p(k.data_ptr()) for k in keep]
        nfri = keep[2].numel()
    else:
        keep = [np.ascontiguousarray(s.cpu().numpy() if isinstance(s, torch.Tensor) else s, dtype=np.uint8).reshape(-1)
                for s in (salts_trace, salts_quot, salts_fri)]
        ptrs = [C.c_void_p(k.ctypes.data) for k in keep]
        nfri = keep[2].size
    lde = trace.size * BLOWUP
    assert (keep[0].numel() if on_dev else keep[0].size) == 16 * lde and (keep[1].numel() if on_dev else keep[1].size) == 16 * lde
    assert nfri >= need, f"salts_fri: {nfri} bytes given, {need} needed"
    cap = 4 << 20
    while True:
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = L.toyni_prove_fibonacci(trace.ctypes.data, trace.size, mask.ctypes.data, ptrs[0], ptrs[1], ptrs[2], nfri, 1 if on_dev else 0,
                                     out.ctypes.data, cap, C.byref(n))
        if rc and n.value > cap:  # cannot happen below 2^27 LDE points; kept for safety
            cap = n.value
            continue
        if rc:
            raise RuntimeError(f"toyni_prove_fibonacci: {L.toyni_prover_error().decode()} (code {rc})")
        break
    data = out[:n.value].tobytes()
    return data if as_bytes else deserialize_proof(data)


def generate_proof(trace_column, mask, salts_trace, salts_quot, salts_fri, device="cuda", timings=None):
    """Every array of LDE size stays on the device: LDE, constraint / quotient / DEEP formulas, commits, FRI commit
    loop and openings.  Salts may be numpy arrays or CUDA uint8 tensors.  `timings` (a dict) receives the wall time
    of each stage in seconds, measured with a device synchronisation at every mark."""
    import time
    t_last = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_last[0]
            t_last[0] = now

    trace_len = len(trace_column)
    lde = trace_len * BLOWUP
    mask = np.asarray(mask, dtype=np.uint64).reshape(-1)
    # the reference draws exactly MASK_DEGREE blinding coefficients (src/fibonacci.rs:117-120); a scalar would broadcast
    assert mask.size == MASK_DEGREE, f"mask: {mask.size} coefficients given, {MASK_DEGREE} required"
    g = get_root_of_unity(trace_len.bit_length() - 1)
    salts_trace, salts_quot = _as_salts(salts_trace, device), _as_salts(salts_quot, device)
    salts_fri = _as_salts(salts_fri, device)

    # 1. trace polynomial (INTT on the GPU) + masking T + Z_H * R (:110-121): trace_len + 140 coefficients, host
    coeffs = _trim(D.to_host(D.coset_ifft_(D.to_device(np.asarray(trace_column, np.uint64), device), 1)))
    tp = np.zeros(trace_len + MASK_DEGREE, np.uint64)
    tp[:coeffs.size] = coeffs
    zr = np.zeros(trace_len + MASK_DEGREE, np.uint64)
    zr[trace_len:] = mask
    zr[:MASK_DEGREE] = _sub(zr[:MASK_DEGREE], mask)
    trace_poly_dev = D.to_device(_trim(_add(tp, zr)), device)
    # LDE on the shifted domain + commit (:124-130)
    trace_lde_dev = D.coset_fft(trace_poly_dev, lde, COSET_SHIFT)
    nodes, root = D.merkle_commit(trace_lde_dev, salts_trace)
    trace_tree = _Tree(trace_lde_dev, nodes, root, salts_trace)
    mark("interpolate + LDE + trace commit")
    # 2. constraint and quotient (:133-153).  c_poly.evaluate(x) over the coset is c_evals itself (interpolate, then
    #    evaluate at the same points), and Z_H(x_i) = 7^n (w_N^n)^i - 1 takes BLOWUP values.
    c_dev = D.fib_constraint(trace_lde_dev, BLOWUP, COSET_SHIFT, pow(g, trace_len - 1, P), pow(g, trace_len - 2, P))
    g_ext = get_root_of_unity(lde.bit_length() - 1)
    sn, wn = pow(COSET_SHIFT, trace_len, P), pow(g_ext, trace_len, P)
    q_dev = D.scale_periodic_(c_dev, [pow((sn * pow(wn, i, P) - 1) % P, P - 2, P) for i in range(BLOWUP)])
    q_coeffs_dev = D.coset_ifft_(q_dev.clone(), COSET_SHIFT)
    nodes, root = D.merkle_commit(q_dev, salts_quot)
    quot_tree = _Tree(q_dev, nodes, root, salts_quot)
    mark("constraint + quotient + quotient commit")
    # 3. Fiat-Shamir: z outside both domains (:156-161, :378-399).  Membership without building the 2 x 32n-element
    #    sets: z is in the extended domain iff z^N = 1, in the shifted domain iff (z / 7)^N = 1, and g_ext^k z is in the
    #    shifted domain iff z is (g_ext generates the extended domain).
    tr = FiatShamirTranscript()
    tr.absorb(trace_tree.root)
    tr.absorb(quot_tree.root)
    inv_shift = pow(COSET_SHIFT, P - 2, P)
    while True:
        z = tr.squeeze_challenge()
        if pow(z, lde, P) != 1 and pow(z * inv_shift % P, lde, P) != 1:
            break
    # 4. OOD evaluations (:164-183)
    t_z, t_gz = D.poly_eval(trace_poly_dev, z), D.poly_eval(trace_poly_dev, g * z % P)
    t_ggz, q_z = D.poly_eval(trace_poly_dev, g * g % P * z % P), D.poly_eval(q_coeffs_dev, z)
    del q_coeffs_dev
    c_z = (t_ggz - (t_gz + t_z)) % P * ((z - pow(g, trace_len - 1, P)) % P) % P * ((z - pow(g, trace_len - 2, P)) % P) % P
    assert c_z == q_z * ((pow(z, trace_len, P) - 1) % P) % P, "Constraint check at z failed"  # :173-177
    for v in (t_z, t_gz, t_ggz, q_z):
        tr.absorb_field(v)
    mark("z + out-of-domain evaluations")
    # 5. DEEP polynomial (:186-198)
    d_dev = D.fib_deep(q_dev, trace_lde_dev, BLOWUP, COSET_SHIFT, z, q_z, t_z, t_gz, t_ggz)
    # 6. FRI commit loop on the device, transcript as the callback (:200-247)
    bound = 1 << (trace_len + MASK_DEGREE - 1).bit_length()
    final_size = lde // bound

    def challenge(root, layer):
        tr.absorb(root)
        return tr.squeeze_challenge()

    layers, nodes_l, roots = D.fri_commit(d_dev, COSET_SHIFT, final_size, salts_fri.view(-1), challenge=challenge)
    mark("DEEP + FRI commit loop")
    tr.absorb(roots[-1])  # the callback absorbed every root but the last (no fold follows it), :239-242
    trees, off = [], 0
    for k, (lay, nd) in enumerate(zip(layers, nodes_l)):
        s = None
        if k < len(layers) - 1:
            s = salts_fri[off:off + lay.shape[0]]
            off += lay.shape[0]
        trees.append(_Tree(lay, nd, roots[k], s))
    # 7. query phase (:250-295): one batched opening per tree
    queries = tr.squeeze_indices(NUM_QUERIES, lde // 2)
    deep = trees[0].open_many([i for qi in queries for i in (qi, qi + lde // 2)])
    trc = trace_tree.open_many([i for qi in queries for i in (qi, (qi + BLOWUP) % lde, (qi + 2 * BLOWUP) % lde)])
    quo = quot_tree.open_many(queries)
    fri, idxs = [], list(queries)
    for k in range(1, len(trees) - 1):
        half = trees[k].n // 2
        idxs = [i % half for i in idxs]
        fri.append(trees[k].open_many([j for i in idxs for j in (i, i + half)]))
    qps = []
    for n_q, qi in enumerate(queries):
        qps.append({"index": qi, "deep_opening": deep[2 * n_q], "deep_opening_pair": deep[2 * n_q + 1],
                    "trace_opening": trc[3 * n_q], "trace_opening_g": trc[3 * n_q + 1], "trace_opening_gg": trc[3 * n_q + 2],
                    "quotient_opening": quo[n_q],
                    "fri_openings": [(f[2 * n_q], f[2 * n_q + 1]) for f in fri]})
    final_layer = [int(v) for v in D.to_host(layers[-1])]
    mark("queries: openings + proof object")
    return {"trace_len": trace_len, "lde_size": lde, "trace_commitment": trace_tree.root,
            "quotient_commitment": quot_tree.root, "t_z": t_z, "t_gz": t_gz, "t_ggz": t_ggz, "q_z": q_z,
            "fri_commitments": roots, "fri_final_layer": final_layer, "query_proofs": qps}
