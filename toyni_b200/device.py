"""Device-resident forms of the hot path over torch tensors (torch is plumbing: memory and streams).

Field arrays are torch.int32 CUDA tensors holding canonical values (p < 2^31, so the sign bit is
never set); Ext arrays have shape (n, 4).  Every call is stream-ordered on torch's current stream
and launches the library's own kernels through the C ABI."""
import ctypes as C

import numpy as np
import torch

from .lib import P, check, lib, u32p


def _bind_stream():
    lib().bb_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


def _chk(t, what="tensor"):
    assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous(), f"{what}: contiguous int32 CUDA tensor required"
    return C.c_void_p(t.data_ptr())


def _beta(beta, limbs):
    b = (C.c_uint32 * 4)(0, 0, 0, 0)
    vals = [int(beta)] if limbs == 1 else [int(x) for x in beta]
    for k, v in enumerate(vals):
        b[k] = v % P
    return b


def to_device(host_u64, device="cuda"):
    """Upload a numpy uint64 field array (reference storage) as an int32 device tensor."""
    a = np.ascontiguousarray(np.asarray(host_u64, dtype=np.uint64))
    if a.size and int(a.max()) >= P:  # canonical inputs (the contract, src/babybear.rs:26-30) skip the slow 64-bit modulo
        a = a % P
    return torch.from_numpy(a.astype(np.int32)).to(device)


def to_host(t):
    """Download to the reference's u64 storage."""
    return t.detach().cpu().numpy().astype(np.uint64)  # canonical values: the sign bit is never set


def ntt_(t, inverse=False):
    """In-place NTT of a 1-D tensor (src/ntt.rs:24-66)."""
    _bind_stream()
    n = t.numel()
    assert n & (n - 1) == 0
    check(lib().bb_ntt_device(_chk(t), n.bit_length() - 1, 1 if inverse else 0), "bb_ntt_device")
    return t


def ntt_batch_(t, inverse=False):
    """In-place NTT of every row of a (batch, n) tensor."""
    _bind_stream()
    batch, n = t.shape
    assert n & (n - 1) == 0
    check(lib().bb_ntt_batch_device(_chk(t), n.bit_length() - 1, batch, 1 if inverse else 0), "bb_ntt_batch_device")
    return t


def ntt_ext_(t, inverse=False):
    """In-place NTT of an (n, 4) Ext tensor: the four coordinate transforms in one launch set."""
    _bind_stream()
    n = t.shape[0]
    assert t.shape[1] == 4 and n & (n - 1) == 0
    check(lib().bb_ntt_ext_device(_chk(t), n.bit_length() - 1, 1 if inverse else 0), "bb_ntt_ext_device")
    return t


def coset_fft(coeffs, size, shift=1, out=None):
    """BabyBearDomain::fft / fft_ext (src/math/domain.rs:107-137) on the device."""
    _bind_stream()
    limbs = 4 if coeffs.dim() == 2 else 1
    n_coeffs = coeffs.shape[0]
    if out is None:
        out = torch.empty((size, 4) if limbs == 4 else (size,), dtype=torch.int32, device=coeffs.device)
    check(lib().bb_coset_fft_device(_chk(coeffs), n_coeffs, size.bit_length() - 1, shift % P, limbs, _chk(out)),
          "bb_coset_fft_device")
    return out


def coset_ifft_(evals, shift=1):
    """BabyBearDomain::ifft / ifft_ext (src/math/domain.rs:85-102,130-132), in place."""
    _bind_stream()
    limbs = 4 if evals.dim() == 2 else 1
    size = evals.shape[0]
    check(lib().bb_coset_ifft_device(_chk(evals), size.bit_length() - 1, shift % P, limbs), "bb_coset_ifft_device")
    return evals


def fri_fold(evals, x0, beta, out=None):
    """fri_fold / fri_fold_ext (src/math/fri.rs) for xs[i] = x0 * w_m^i."""
    _bind_stream()
    limbs = 4 if evals.dim() == 2 else 1
    m = evals.shape[0]
    if out is None:
        out = torch.empty((m // 2, 4) if limbs == 4 else (m // 2,), dtype=torch.int32, device=evals.device)
    check(lib().bb_fri_fold_device(_chk(evals), m, x0 % P, _beta(beta, limbs), limbs, _chk(out)), "bb_fri_fold_device")
    return out


def fri_fold_shard(evals, log_m, x0, beta, nranks, rank, out=None):
    """One cyclic shard (global indices rank, rank+nranks, ...) of a fold; no communication."""
    _bind_stream()
    limbs = 4 if evals.dim() == 2 else 1
    m_local = evals.shape[0]
    if out is None:
        out = torch.empty((m_local // 2, 4) if limbs == 4 else (m_local // 2,), dtype=torch.int32, device=evals.device)
    check(lib().bb_fri_fold_shard_device(_chk(evals), m_local, log_m, x0 % P, _beta(beta, limbs), limbs, nranks, rank,
                                         _chk(out)), "bb_fri_fold_shard_device")
    return out


def fri_fold_xs(evals, xs, beta, out=None):
    _bind_stream()
    limbs = 4 if evals.dim() == 2 else 1
    m = evals.shape[0]
    if out is None:
        out = torch.empty((m // 2, 4) if limbs == 4 else (m // 2,), dtype=torch.int32, device=evals.device)
    check(lib().bb_fri_fold_xs_device(_chk(evals), m, _chk(xs), _beta(beta, limbs), limbs, _chk(out)), "bb_fri_fold_xs_device")
    return out


def interleave(src, groups):
    """dst[j*G + r] = src[r*c + j] (G = groups runs of c rows each): the re-layout after the cyclic -> block exchange
    of the sharded FRI commit."""
    _bind_stream()
    rows = src.shape[0]
    limbs = 4 if src.dim() == 2 else 1
    dst = torch.empty_like(src)
    check(lib().bb_interleave_device(_chk(src), groups, rows // groups, limbs, _chk(dst)), "bb_interleave_device")
    return dst


def merkle_node_count(n):
    return lib().bb_merkle_node_count(n)


def merkle_commit(vals, salts=None, nodes=None, want_root=True):
    """Salted / unsalted commit (src/fibonacci.rs:340-363). vals: int32 (n,) or (n,4); salts: uint8 (n,16).
    Returns (nodes uint8 tensor (count,32), root bytes or None)."""
    _bind_stream()
    limbs = 4 if vals.dim() == 2 else 1
    n = vals.shape[0]
    if nodes is None:
        nodes = torch.empty((merkle_node_count(n), 32), dtype=torch.uint8, device=vals.device)
    sp = None
    if salts is not None:
        assert salts.is_cuda and salts.dtype == torch.uint8 and salts.is_contiguous() and salts.numel() == 16 * n
        sp = C.c_void_p(salts.data_ptr())
    root = np.empty(32, dtype=np.uint8) if want_root else None
    check(lib().bb_merkle_commit_device(_chk(vals), limbs, n, sp, C.c_void_p(nodes.data_ptr()),
                                        None if root is None else root.ctypes.data), "bb_merkle_commit_device")
    return nodes, (root.tobytes() if want_root else None)


def merkle_open_batch(nodes, nleaves, indices):
    """Authentication paths of a whole query set in one launch (src/merkle.rs:50-80 for every index).
    Returns (paths uint8 [nq, depth, 32], positions uint8 [nq, depth])."""
    _bind_stream()
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
    depth = C.c_size_t(0)
    d = max(int(nleaves) - 1, 0).bit_length()
    paths = np.zeros((idx.size, d, 32), dtype=np.uint8)
    pos = np.zeros((idx.size, d), dtype=np.uint8)
    check(lib().bb_merkle_open_batch_device(C.c_void_p(nodes.data_ptr()), nleaves, idx.ctypes.data, idx.size, paths.ctypes.data,
                                            pos.ctypes.data, C.byref(depth)), "bb_merkle_open_batch_device")
    assert depth.value == d
    return paths, pos


def gather(t, indices):
    """Rows t[indices] of a contiguous device tensor, fetched with one kernel and one copy."""
    _bind_stream()
    assert t.is_cuda and t.is_contiguous()
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64))
    row = t[0].numel() * t.element_size() if t.dim() > 1 else t.element_size()
    out = np.zeros((idx.size, row), dtype=np.uint8)
    check(lib().bb_gather_device(C.c_void_p(t.data_ptr()), row, idx.ctypes.data, idx.size, out.ctypes.data), "bb_gather_device")
    return out


def fib_constraint(trace_lde, step, shift, b1, b2, out=None):
    """src/fibonacci.rs:133-143 over the whole shifted domain: (T(g^2 x) - T(g x) - T(x)) (x - b1) (x - b2)."""
    _bind_stream()
    n = trace_lde.numel()
    if out is None:
        out = torch.empty_like(trace_lde)
    check(lib().bb_fib_constraint_device(_chk(trace_lde), n.bit_length() - 1, step, shift % P, b1 % P, b2 % P, _chk(out)),
          "bb_fib_constraint_device")
    return out


def scale_periodic_(vals, table):
    """vals[i] *= table[i mod len(table)] in place (len a power of two <= 64)."""
    _bind_stream()
    tab = np.ascontiguousarray(np.asarray([int(v) % P for v in table], dtype=np.uint32))
    check(lib().bb_scale_periodic_device(_chk(vals), vals.numel(), tab.ctypes.data, tab.size), "bb_scale_periodic_device")
    return vals


def fib_deep(quotient, trace_lde, step, shift, z, q_z, t_z, t_gz, t_ggz, out=None):
    """DEEP composition polynomial over the shifted domain (src/fibonacci.rs:186-198), 1/(x - z) by batched inversion."""
    _bind_stream()
    n = trace_lde.numel()
    if out is None:
        out = torch.empty_like(trace_lde)
    check(lib().bb_fib_deep_device(_chk(quotient), _chk(trace_lde), n.bit_length() - 1, step, shift % P, z % P, q_z % P, t_z % P,
                                   t_gz % P, t_ggz % P, _chk(out)), "bb_fib_deep_device")
    return out


def poly_eval(coeffs, z):
    """Polynomial::evaluate (src/math/polynomial.rs:134-144) of device-resident coefficients at one point."""
    _bind_stream()
    v = C.c_uint32(0)
    check(lib().bb_poly_eval_device(_chk(coeffs), coeffs.numel(), int(z) % P, C.byref(v)), "bb_poly_eval_device")
    return int(v.value)


def fri_commit(layer0, shift, final_size, salts=None, challenge=None, betas=None, hash_layers=True):
    """The prover's FRI commit loop (src/fibonacci.rs:200-247) on device-resident data.
    challenge(root: bytes, layer: int) -> beta (int or 4 limbs) plays the transcript; alternatively
    `betas` supplies them up front.  Returns (layers tensor list, nodes tensor list or None, roots list)."""
    from .lib import CHALLENGE_FN

    _bind_stream()
    limbs = 4 if layer0.dim() == 2 else 1
    n = layer0.shape[0]
    sizes, m = [], n
    while True:
        sizes.append(m)
        if m <= final_size:
            break
        m //= 2
    total = max(1, sum(sizes[1:]))  # folded layers only; layer 0 stays in the caller's tensor
    dev = layer0.device
    layers = torch.empty((total, 4) if limbs == 4 else (total,), dtype=torch.int32, device=dev)
    nodes = roots = None
    if hash_layers:
        nodes = torch.empty((sum(merkle_node_count(s) for s in sizes), 32), dtype=torch.uint8, device=dev)
        roots = np.zeros((len(sizes), 32), dtype=np.uint8)
    if salts is not None and hash_layers:
        # 16 bytes per leaf of every salted (= non-final) layer, back to back: the fused fold + leaf-hash kernels read
        # straight through this pointer
        need = 16 * sum(sizes[:-1])
        assert salts.is_cuda and salts.dtype == torch.uint8 and salts.is_contiguous(), "salts: contiguous uint8 CUDA tensor"
        assert salts.numel() >= need, f"salts: {salts.numel()} bytes given, {need} needed for layers {sizes[:-1]}"
    cb = None
    cb_error = []
    if challenge is not None:
        def _cb(_user, root_ptr, layer, beta_out):
            # ctypes swallows exceptions raised inside a callback: keep the first one and re-raise it after the C call
            # (beta stays 0 for the remaining layers; their results are discarded)
            if cb_error:
                return
            try:
                b = challenge(bytes(root_ptr[:32]), int(layer))
                vals = [int(b)] if limbs == 1 else [int(x) for x in b]
                assert len(vals) == limbs, f"challenge returned {len(vals)} limbs, {limbs} expected"
                for k, v in enumerate(vals):
                    beta_out[k] = v % P
            except BaseException as e:  # noqa: BLE001 - re-raised below
                cb_error.append(e)
        cb = CHALLENGE_FN(_cb)
    bt = None
    if betas is not None:
        bt = np.ascontiguousarray(np.asarray(betas, dtype=np.uint64) % P).astype(np.uint32)
    folds = C.c_size_t(0)
    sp = None if salts is None else C.c_void_p(salts.data_ptr())
    check(lib().bb_fri_commit_device(_chk(layer0), n, shift % P, final_size, limbs, sp,
                                     C.cast(cb, C.c_void_p) if cb is not None else None, None,
                                     None if bt is None else bt.ctypes.data, _chk(layers),
                                     None if nodes is None else C.c_void_p(nodes.data_ptr()),
                                     None if roots is None else roots.ctypes.data, C.byref(folds)), "bb_fri_commit_device")
    if cb_error:
        raise cb_error[0]
    out_layers, out_nodes, off, noff = [], [], 0, 0
    for k, s in enumerate(sizes):
        if k == 0:
            out_layers.append(layer0)
        else:
            out_layers.append(layers[off:off + s])
            off += s
        if nodes is not None:
            c = merkle_node_count(s)
            out_nodes.append(nodes[noff:noff + c])
            noff += c
    return out_layers, (out_nodes if nodes is not None else None), ([r.tobytes() for r in roots] if roots is not None else None)
