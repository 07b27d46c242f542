"""CPU restatement of the Fibonacci AIR prover and verifier (src/fibonacci.rs, src/verifier.rs) with the
randomness made explicit.  TEST INFRASTRUCTURE ONLY (config 1 of BASELINE.json: the acceptance oracle for proofs
assembled from GPU-computed pieces).

The reference draws its blinding from rand::thread_rng() (src/fibonacci.rs:117, :342), which is OS-seeded and
unpinnable, so the mask coefficients and the per-leaf salts are inputs here:
    mask        : MASK_DEGREE field elements            (src/fibonacci.rs:118-120)
    salts_trace : (lde, 16) bytes                         (:129 -> :340-353)
    salts_quot  : (lde, 16) bytes                         (:153)
    salts_fri   : 16 bytes per leaf of FRI layers 0..folds-1, back to back  (:206, :237)
Field vectors are numpy uint64 arrays of canonical values; products of two values fit in 62 bits, so
(a * b) % P is exact in uint64.
"""
import numpy as np

from . import oracle as O

P = O.P
NUM_QUERIES, BLOWUP, COSET_SHIFT = 44, 32, 7  # src/fibonacci.rs:11-16
MASK_DEGREE = 3 * NUM_QUERIES + 8             # :19

_P = np.uint64(P)


def mulmod(a, b):
    return (np.asarray(a, np.uint64) * np.asarray(b, np.uint64)) % _P


def addmod(a, b):
    return (np.asarray(a, np.uint64) + np.asarray(b, np.uint64)) % _P


def submod(a, b):
    return (np.asarray(a, np.uint64) + _P - np.asarray(b, np.uint64)) % _P


def powmod_vec(a, e):
    a = np.asarray(a, np.uint64)
    r = np.ones_like(a)
    while e:
        if e & 1:
            r = mulmod(r, a)
        a = mulmod(a, a)
        e >>= 1
    return r


def invmod_vec(a):  # Fermat, like BabyBear::inverse (src/babybear.rs:111-114)
    return powmod_vec(a, P - 2)


def poly_trim(c):  # Polynomial::new, src/math/polynomial.rs:11-16
    c = np.asarray(c, np.uint64)
    n = c.size
    while n > 0 and c[n - 1] == 0:
        n -= 1
    return c[:n].copy()


def poly_eval(c, x):  # Horner, src/math/polynomial.rs:134-144, at one point (Python ints)
    acc = 0
    for v in reversed([int(t) for t in c]):
        acc = (acc * x + v) % P
    return acc


def poly_eval_vec(c, xs):  # Horner at many points
    xs = np.asarray(xs, np.uint64)
    if len(c) == 0:
        return np.zeros_like(xs)
    acc = np.full(xs.shape, c[-1], np.uint64)
    for v in c[-2::-1]:
        acc = (acc * xs + v) % _P
    return acc


def lagrange_interpolate_column(domain, ys):
    """ExecutionTrace::interpolate_column, src/program/trace.rs:28-56 (O(n^3), literal)."""
    n = len(domain)
    xs = [int(v) for v in domain]
    poly = np.zeros(n, np.uint64)
    for i in range(n):
        num = np.zeros(n, np.uint64)
        num[0] = 1
        deg = 0
        denom = 1
        for j in range(n):
            if i == j:
                continue
            shifted = np.zeros(n, np.uint64)            # numerator * (x - xj)
            shifted[1:deg + 2] = num[:deg + 1]
            num = submod(shifted, mulmod(num, np.uint64(xs[j])))
            deg += 1
            denom = denom * ((xs[i] - xs[j]) % P) % P
        scale = pow(denom, P - 2, P) * int(ys[i]) % P
        poly = addmod(poly, mulmod(num, np.uint64(scale)))
    return poly_trim(poly)


def fibonacci_trace(n):  # the test trace of src/fibonacci.rs:421-431
    t = [1, 1]
    while len(t) < n:
        t.append((t[-1] + t[-2]) % P)
    return np.array(t[:n], np.uint64)


def derive_z(transcript, ext_elements, shifted_elements, g_ext):  # src/fibonacci.rs:378-399
    ext_set = set(int(v) for v in ext_elements)
    shift_set = set(int(v) for v in shifted_elements)
    while True:
        z = transcript.squeeze_challenge()
        if z not in ext_set and z not in shift_set and (g_ext * z) % P not in shift_set and (g_ext * g_ext % P * z) % P not in shift_set:
            return z


def open_merkle(nodes, n, values, salts, index):  # src/fibonacci.rs:366-374
    path, pos = O.merkle_open(nodes, n, index)
    return {"index": int(index), "value": int(values[index]), "path": [p.tobytes() for p in path], "position": [bool(b) for b in pos],
            "salt": b"" if salts is None else bytes(salts[index])}


def generate_proof(trace_column, mask, salts_trace, salts_quot, salts_fri, interpolate="lagrange", backend=None):
    """StarkProver::generate_proof, src/fibonacci.rs:99-310.  `backend` swaps the hot-path primitives (NTT-based
    coset fft / ifft, fold + commit loop, Merkle commit); None = the CPU oracle's."""
    B = backend or CpuBackend()
    trace_len = len(trace_column)
    lde = trace_len * BLOWUP
    log_n = trace_len.bit_length() - 1
    g = O.root_of_unity(log_n)                       # domain.group_gen(), :108
    g_ext = O.root_of_unity(lde.bit_length() - 1)
    # 1. trace polynomial + masking (:110-121)
    domain_elements = O.domain_elements(trace_len, 1)
    if interpolate == "lagrange":
        trace_poly = lagrange_interpolate_column(domain_elements, trace_column)
    else:                                            # the unique interpolant, by INTT
        trace_poly = poly_trim(B.ifft(np.asarray(trace_column, np.uint64), 1))
    zr = np.zeros(trace_len + MASK_DEGREE, np.uint64)  # Z_H * R = R x^n - R
    zr[trace_len:] = mask
    zr[:MASK_DEGREE] = submod(zr[:MASK_DEGREE], mask)
    tp = np.zeros(trace_len + MASK_DEGREE, np.uint64)
    tp[:trace_poly.size] = trace_poly
    trace_poly = poly_trim(addmod(tp, zr))
    shifted_elements = O.domain_elements(lde, COSET_SHIFT)
    trace_lde = B.lde(trace_poly, lde)               # :124-128 (Horner at every coset point == coset FFT)
    trace_nodes, trace_commitment = B.commit(trace_lde, salts_trace)
    # 2. constraint & quotient (:133-151)
    t_x, t_gx, t_ggx = trace_lde, np.roll(trace_lde, -BLOWUP), np.roll(trace_lde, -2 * BLOWUP)  # g = w_N^32
    b1 = submod(shifted_elements, np.uint64(pow(g, trace_len - 1, P)))
    b2 = submod(shifted_elements, np.uint64(pow(g, trace_len - 2, P)))
    c_evals = mulmod(mulmod(submod(t_ggx, addmod(t_gx, t_x)), b1), b2)
    c_poly = poly_trim(B.ifft(c_evals, COSET_SHIFT))
    z_h = submod(powmod_vec(shifted_elements, trace_len), np.uint64(1))       # z_poly.evaluate(x) = x^n - 1
    q_evals = mulmod(B.lde(c_poly, lde), invmod_vec(z_h))                      # c_poly.evaluate(x) / z_poly.evaluate(x)
    q_poly = poly_trim(B.ifft(q_evals, COSET_SHIFT))
    quot_nodes, quotient_commitment = B.commit(q_evals, salts_quot)
    # 3. Fiat-Shamir (:156-161)
    tr = O.FiatShamirTranscript()
    tr.absorb(trace_commitment)
    tr.absorb(quotient_commitment)
    z = derive_z(tr, O.domain_elements(lde, 1), shifted_elements, g_ext)
    # 4. OOD evaluations (:164-183)
    t_z, t_gz, t_ggz = poly_eval(trace_poly, z), poly_eval(trace_poly, g * z % P), poly_eval(trace_poly, g * g % P * z % P)
    q_z = poly_eval(q_poly, z)
    c_z = (t_ggz - (t_gz + t_z)) % P * ((z - pow(g, trace_len - 1, P)) % P) % P * ((z - pow(g, trace_len - 2, P)) % P) % P
    assert c_z == q_z * ((pow(z, trace_len, P) - 1) % P) % P, "Constraint check at z failed"
    for v in (t_z, t_gz, t_ggz, q_z):
        tr.absorb_field(v)
    # 5. DEEP polynomial (:186-198)
    inv_xz = invmod_vec(submod(shifted_elements, np.uint64(z)))
    d_evals = mulmod(submod(q_evals, np.uint64(q_z)), inv_xz)
    d_evals = addmod(d_evals, mulmod(submod(t_ggx, np.uint64(t_ggz)), inv_xz))
    d_evals = addmod(d_evals, mulmod(submod(t_gx, np.uint64(t_gz)), inv_xz))
    d_evals = addmod(d_evals, mulmod(submod(t_x, np.uint64(t_z)), inv_xz))
    # 6. FRI commit loop (:200-247)
    bound = 1 << (trace_len + MASK_DEGREE - 1).bit_length()          # next_power_of_two, :220
    final_size = lde // bound
    layers, fri_nodes, roots, layer_salts = B.fri_commit(d_evals, final_size, salts_fri, tr)
    # 7. query phase (:250-295)
    queries = tr.squeeze_indices(NUM_QUERIES, lde // 2)
    qps = []
    for qi in queries:
        half0 = lde // 2
        qp = {
            "index": qi,
            "deep_opening": open_merkle(fri_nodes[0], lde, layers[0], layer_salts[0], qi),
            "deep_opening_pair": open_merkle(fri_nodes[0], lde, layers[0], layer_salts[0], qi + half0),
            "trace_opening": open_merkle(trace_nodes, lde, trace_lde, salts_trace, qi),
            "trace_opening_g": open_merkle(trace_nodes, lde, trace_lde, salts_trace, (qi + BLOWUP) % lde),
            "trace_opening_gg": open_merkle(trace_nodes, lde, trace_lde, salts_trace, (qi + 2 * BLOWUP) % lde),
            "quotient_opening": open_merkle(quot_nodes, lde, q_evals, salts_quot, qi),
            "fri_openings": [],
        }
        idx = qi
        for k in range(1, len(layers) - 1):
            half = len(layers[k]) // 2
            idx %= half
            qp["fri_openings"].append((open_merkle(fri_nodes[k], len(layers[k]), layers[k], layer_salts[k], idx),
                                       open_merkle(fri_nodes[k], len(layers[k]), layers[k], layer_salts[k], idx + half)))
        qps.append(qp)
    return {"trace_len": trace_len, "lde_size": lde, "trace_commitment": trace_commitment,
            "quotient_commitment": quotient_commitment, "t_z": t_z, "t_gz": t_gz, "t_ggz": t_ggz, "q_z": q_z,
            "fri_commitments": roots, "fri_final_layer": [int(v) for v in layers[-1]], "query_proofs": qps}


class CpuBackend:
    """Hot-path primitives from the CPU oracle."""

    def lde(self, coeffs, size):
        return O.domain_fft(coeffs, size, COSET_SHIFT)

    def ifft(self, evals, shift):
        return O.domain_ifft(evals, shift)

    def commit(self, values, salts):
        return O.commit_values(values, salts)

    def fri_commit(self, d_evals, final_size, salts_fri, transcript):
        layers, roots, _ = O.fri_commit(d_evals, COSET_SHIFT, final_size, salts_fri, transcript=transcript)
        nodes, layer_salts, off = [], [], 0
        for k, layer in enumerate(layers):
            if k == len(layers) - 1:
                s = None
            else:
                s = np.ascontiguousarray(np.asarray(salts_fri, np.uint8).reshape(-1)[16 * off:16 * (off + len(layer))]).reshape(-1, 16)
                off += len(layer)
            nodes.append(O.commit_values(layer, s)[0])
            layer_salts.append(s)
        return layers, nodes, roots, layer_salts


# ------------------------------------------------------------------------------------------------ serialization
def serialize_proof(p):
    """Canonical byte form of StarkProof (the reference has none: `#[derive(Debug)]` only, src/fibonacci.rs:62-86).
    Little-endian u64 for integers and field values, raw 32-byte digests, length-prefixed lists, in field order."""
    out = bytearray()
    u64 = lambda v: out.extend(int(v).to_bytes(8, "little"))

    def opening(o):
        u64(o["index"]); u64(o["value"])
        out.append(len(o["salt"])); out.extend(o["salt"])
        u64(len(o["path"]))
        for d, r in zip(o["path"], o["position"]):
            out.extend(d); out.append(1 if r else 0)

    u64(p["trace_len"]); u64(p["lde_size"])
    out.extend(p["trace_commitment"]); out.extend(p["quotient_commitment"])
    for k in ("t_z", "t_gz", "t_ggz", "q_z"):
        u64(p[k])
    u64(len(p["fri_commitments"]))
    for r in p["fri_commitments"]:
        out.extend(r)
    u64(len(p["fri_final_layer"]))
    for v in p["fri_final_layer"]:
        u64(v)
    u64(len(p["query_proofs"]))
    for q in p["query_proofs"]:
        u64(q["index"])
        for k in ("deep_opening", "deep_opening_pair", "trace_opening", "trace_opening_g", "trace_opening_gg", "quotient_opening"):
            opening(q[k])
        u64(len(q["fri_openings"]))
        for a, b in q["fri_openings"]:
            opening(a); opening(b)
    return bytes(out)


# ------------------------------------------------------------------------------------------------ verifier
def _verify_opening(o, root):  # src/verifier.rs:235-238
    leaf = o["salt"] + int(o["value"]).to_bytes(8, "little")
    path = np.frombuffer(b"".join(o["path"]), np.uint8).reshape(-1, 32) if o["path"] else np.zeros((0, 32), np.uint8)
    return O.merkle_verify(leaf, path, np.array(o["position"], np.uint8), root)


class _CosetPoints:
    """shifted_elements[i] = 7 w^i on demand (large proofs: the verifier reads 44 * O(log N) of the 32n points)."""

    def __init__(self, lde):
        self.w = O.root_of_unity(lde.bit_length() - 1)

    def __getitem__(self, i):
        return COSET_SHIFT * pow(self.w, int(i), P) % P


def derive_z_algebraic(transcript, lde):
    """derive_z without the two 32n-element sets: z is in the extended domain iff z^N = 1 and in the shifted domain
    iff (z / 7)^N = 1; g_ext^k z is in the shifted domain iff z is.  Same z as src/fibonacci.rs:378-399."""
    inv_shift = pow(COSET_SHIFT, P - 2, P)
    while True:
        z = transcript.squeeze_challenge()
        if pow(z, lde, P) != 1 and pow(z * inv_shift % P, lde, P) != 1:
            return z


def verify(p, algebraic=None):
    """StarkVerifier::verify, src/verifier.rs:14-232.  `algebraic` (default: for lde >= 2^21) swaps the materialised
    domains for their closed forms; both ways give the same verdict (tests/test_oracle_reference_tests.py)."""
    n, lde = p["trace_len"], p["lde_size"]
    if lde != n * BLOWUP:
        return False
    if algebraic is None:
        algebraic = lde >= (1 << 21)
    g = O.root_of_unity(n.bit_length() - 1)
    g_ext = O.root_of_unity(lde.bit_length() - 1)
    tr = O.FiatShamirTranscript()
    tr.absorb(p["trace_commitment"]); tr.absorb(p["quotient_commitment"])
    if algebraic:
        shifted = _CosetPoints(lde)
        z = derive_z_algebraic(tr, lde)
    else:
        shifted = O.domain_elements(lde, COSET_SHIFT)
        z = derive_z(tr, O.domain_elements(lde, 1), shifted, g_ext)
    for k in ("t_z", "t_gz", "t_ggz", "q_z"):
        tr.absorb_field(p[k])
    c_z = (p["t_ggz"] - (p["t_gz"] + p["t_z"])) % P * ((z - pow(g, n - 1, P)) % P) % P * ((z - pow(g, n - 2, P)) % P) % P
    if c_z != p["q_z"] * ((pow(z, n, P) - 1) % P) % P:
        return False
    if not p["fri_commitments"]:
        return False
    bound = 1 << (n + MASK_DEGREE - 1).bit_length()
    final_size = lde // bound
    folds = (lde // final_size).bit_length() - 1
    if len(p["fri_commitments"]) != folds + 1 or len(p["fri_final_layer"]) != final_size:
        return False
    if any(v != p["fri_final_layer"][0] for v in p["fri_final_layer"]):
        return False
    if O.commit_values(np.array(p["fri_final_layer"], np.uint64))[1] != p["fri_commitments"][-1]:
        return False
    tr.absorb(p["fri_commitments"][0])
    betas = []
    for i in range(1, len(p["fri_commitments"])):
        betas.append(tr.squeeze_challenge())
        tr.absorb(p["fri_commitments"][i])
    queries = tr.squeeze_indices(NUM_QUERIES, lde // 2)
    if len(p["query_proofs"]) != NUM_QUERIES:
        return False
    half_inv = pow(2, P - 2, P)
    for qi, qp in zip(queries, p["query_proofs"]):
        if qp["index"] != qi or len(qp["fri_openings"]) != folds - 1:
            return False
        for k in ("trace_opening", "trace_opening_g", "trace_opening_gg"):
            if not _verify_opening(qp[k], p["trace_commitment"]):
                return False
        if (qp["trace_opening"]["index"] != qi or qp["trace_opening_g"]["index"] != (qi + BLOWUP) % lde
                or qp["trace_opening_gg"]["index"] != (qi + 2 * BLOWUP) % lde):
            return False
        if not _verify_opening(qp["quotient_opening"], p["quotient_commitment"]):
            return False
        if not _verify_opening(qp["deep_opening"], p["fri_commitments"][0]) or not _verify_opening(qp["deep_opening_pair"], p["fri_commitments"][0]):
            return False
        x_i = int(shifted[qi])
        inv = pow((x_i - z) % P, P - 2, P)
        exp = ((qp["quotient_opening"]["value"] - p["q_z"]) * inv + (qp["trace_opening_gg"]["value"] - p["t_ggz"]) * inv
               + (qp["trace_opening_g"]["value"] - p["t_gz"]) * inv + (qp["trace_opening"]["value"] - p["t_z"]) * inv) % P
        if qp["deep_opening"]["value"] != exp:
            return False
        a0, b0 = qp["deep_opening"]["value"], qp["deep_opening_pair"]["value"]
        prev = ((a0 + b0) * half_inv + (a0 - b0) * half_inv % P * betas[0] % P * pow(x_i, P - 2, P)) % P
        pos = qi
        for layer, (op, op_pair) in enumerate(qp["fri_openings"]):
            k = layer + 1
            half = (lde >> k) // 2
            lo = pos % half
            if not _verify_opening(op, p["fri_commitments"][k]) or not _verify_opening(op_pair, p["fri_commitments"][k]):
                return False
            if (op["value"] if pos == lo else op_pair["value"]) != prev:
                return False
            x = pow(int(shifted[lo]), 1 << k, P)
            a, b = op["value"], op_pair["value"]
            prev = ((a + b) * half_inv + (a - b) * half_inv % P * betas[k] % P * pow(x, P - 2, P)) % P
            pos = lo
        if p["fri_final_layer"][pos] != prev:
            return False
    return True


def proof_randomness(trace_len, seed=0x70796E69):
    """Deterministic stand-in for thread_rng: mask coefficients and salts from the SplitMix64 stream."""
    lde = trace_len * BLOWUP
    bound = 1 << (trace_len + MASK_DEGREE - 1).bit_length()
    final = lde // bound
    nfri, m = 0, lde
    while m > final:
        nfri += m
        m //= 2
    return (O.random_field(MASK_DEGREE, seed), O.random_bytes(16 * lde, seed + 1).reshape(lde, 16),
            O.random_bytes(16 * lde, seed + 2).reshape(lde, 16), O.random_bytes(16 * nfri, seed + 3))
