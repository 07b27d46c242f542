"""Sharding of the hot path over the GPUs of one box (one process per GPU, torch.distributed).

Nothing like this exists in the reference (single device, default stream).  Three layouts:

* independent column NTTs (batched traces): column c lives on rank c % G — no communication at all;
* one large NTT (2^25..2^27): four-step.  n = n1*n2, input index j = j1*n2 + j2, output index k = k1 + n1*k2.
  Rank r owns the column block j2 in [r*n2/G, (r+1)*n2/G) of the n1 x n2 input matrix (all j1) and ends up with
  the row block k1 in [r*n1/G, (r+1)*n1/G) of the output matrix out[k1][k2] = X[k1 + n1*k2] (all k2):
      1. n1-point NTTs down the local columns            (bb_ntt_columns_device)
      2. multiply (k1, j2) by w_n^(j2*k1)                (bb_fourstep_twiddle_device)
      3. all-to-all: row block s of every rank goes to rank s  (the only exchange: n/G^2 values per peer)
      4. n2-point NTTs along the local rows              (bb_ntt_batch_device)
* FRI folding: cyclic layout, rank r owns indices i = r (mod G).  The fold partner i + m/2 is on the same rank
  while m/2 >= G, so a 2^25 -> 2^4 chain needs no exchange for G <= 8 (bb_fri_fold_shard_device).
* per-layer Merkle commits of the FRI loop (`fri_commit_sharded`): the tree pairs ADJACENT leaves
  (src/merkle.rs:36-43), which the cyclic layout spreads over the ranks, so every layer is exchanged once from the
  cyclic to the block layout (rank t gets leaves [t*m/G, (t+1)*m/G): one all-to-all of m/G^2 values per peer, 4 or
  16 bytes per leaf instead of 32-byte digests), each rank hashes its leaves and builds its subtree, the G subtree
  roots are all-gathered and the top log2(G) levels are finished redundantly on every rank's host, so that the
  transcript (absorb root, squeeze beta) runs replicated and no broadcast is needed.

The index arithmetic is written once (`fourstep_*`) and runs either on CUDA tensors with NCCL or on CPU tensors
with gloo, which is how the N > 1 path is tested without GPUs.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .lib import P


def fourstep_split(log_n, world):
    """(n1, n2) with n1*n2 = 2^log_n, both divisible by `world`; n1 >= n2."""
    l1 = (log_n + 1) // 2
    if os.environ.get("TOYNI_FOURSTEP_LOG_N2"):  # tuning hook (tools/nvlink_fourstep.py)
        l1 = log_n - int(os.environ["TOYNI_FOURSTEP_LOG_N2"])
    n1, n2 = 1 << l1, 1 << (log_n - l1)
    assert n1 % world == 0 and n2 % world == 0, "transform too small for this many ranks"
    return n1, n2


def fourstep_scatter(x, rank, world):
    """Local input block of `rank`: columns j2 in its range, all rows j1, row-major (n1, n2/G)."""
    n = x.size
    n1, n2 = fourstep_split(n.bit_length() - 1, world)
    w = n2 // world
    return np.ascontiguousarray(x.reshape(n1, n2)[:, rank * w:(rank + 1) * w])


def fourstep_gather(blocks, log_n):
    """Natural-order X from the per-rank output blocks out_r[k1_local][k2] = X[k1 + n1*k2]."""
    world = len(blocks)
    n1, n2 = fourstep_split(log_n, world)
    full = np.concatenate(blocks, axis=0)  # (n1, n2) indexed [k1][k2]
    return np.ascontiguousarray(full.T).reshape(-1)  # index k1 + n1*k2


def fourstep_reference_simulation(x, world, ntt, mul):
    """All ranks simulated in one process with plain numpy: pins the index maps against a 1-D transform."""
    n = x.size
    log_n = n.bit_length() - 1
    n1, n2 = fourstep_split(log_n, world)
    w_n = pow(440564289, 1 << (27 - log_n), P)
    cw, rw = n2 // world, n1 // world
    stage = []
    for r in range(world):
        a = fourstep_scatter(x, r, world)
        a = np.stack([ntt(np.ascontiguousarray(a[:, c])) for c in range(cw)], axis=1)        # step 1
        tw = np.array([[pow(w_n, (r * cw + c) * k1, P) for c in range(cw)] for k1 in range(n1)], dtype=np.uint64)
        stage.append(mul(a, tw))                                                             # step 2
    outs = []
    for r in range(world):                                                                   # step 3 (exchange)
        b = np.concatenate([stage[s][r * rw:(r + 1) * rw, :] for s in range(world)], axis=1)  # (n1/G, n2)
        outs.append(np.stack([ntt(np.ascontiguousarray(b[k])) for k in range(rw)], axis=0))  # step 4
    return fourstep_gather(outs, log_n)


def _all_to_all_rows(block, world):
    """block: (n1, cols) tensor; row block s goes to rank s.  Returns (n1/G, G*cols): my rows, all columns."""
    n1, cols = block.shape
    rw = n1 // world
    if world == 1:
        return block
    recv = torch.empty_like(block)
    dist.all_to_all_single(recv.view(-1), block.contiguous().view(-1))
    # recv is [source s][k1_local][c]; the row transforms want [k1_local][s][c]
    return recv.view(world, rw, cols).permute(1, 0, 2).reshape(rw, world * cols).contiguous()


def fourstep_ntt_distributed(x_full, rank, world, backend_ntt, device="cpu"):
    """CPU/gloo form used by the tests: local transforms through `backend_ntt` (a 1-D natural-order NTT),
    exchange through torch.distributed.  Returns this rank's output block as numpy (n1/G, n2)."""
    n = x_full.size
    log_n = n.bit_length() - 1
    n1, n2 = fourstep_split(log_n, world)
    cw = n2 // world
    a = fourstep_scatter(x_full, rank, world)
    a = np.stack([backend_ntt(np.ascontiguousarray(a[:, c])) for c in range(cw)], axis=1)
    w_n = pow(440564289, 1 << (27 - log_n), P)
    tw = np.array([[pow(w_n, (rank * cw + c) * k1, P) for c in range(cw)] for k1 in range(n1)], dtype=object)
    a = (a.astype(object) * tw % P).astype(np.int64)
    b = _all_to_all_rows(torch.from_numpy(a).to(device), world).cpu().numpy().astype(np.uint64)
    return np.stack([backend_ntt(np.ascontiguousarray(b[k])) for k in range(b.shape[0])], axis=0)


# ------------------------------------------------------------------------------------------------ CUDA / NCCL
def fourstep_ntt_cuda(block, log_n, rank, world, inverse=False):
    """One 2^log_n NTT sharded over `world` GPUs.  block: int32 CUDA tensor (n1, n2/G), this rank's columns.
    Returns (n1/G, n2): out[k1_local][k2] = X[k1 + n1*k2].  Device kernels through the C ABI; the exchange is one
    NCCL all-to-all (world == 1 runs the same steps without it)."""
    import ctypes as C

    from .device import _bind_stream, _chk, ntt_batch_
    from .lib import check, lib

    n1, n2 = fourstep_split(log_n, world)
    cw = n2 // world
    assert tuple(block.shape) == (n1, cw)
    _bind_stream()
    d = 1 if inverse else 0
    check(lib().bb_ntt_columns_device(_chk(block), n1.bit_length() - 1, cw, d), "bb_ntt_columns_device")
    check(lib().bb_fourstep_twiddle_device(_chk(block), log_n, n1.bit_length() - 1, cw, rank * cw, d),
          "bb_fourstep_twiddle_device")
    rows = _all_to_all_rows(block, world)
    return ntt_batch_(rows, inverse)


class _RawCudaBuffer:
    """A cudaMalloc'ed buffer (not from torch's caching allocator, so that CUDA IPC maps exactly this allocation)
    exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, nwords):
        import ctypes as C

        from .lib import check, lib
        self.nwords = nwords
        self.ptr = C.c_void_p()
        check(lib().bb_dev_alloc(C.byref(self.ptr), nwords * 4), "bb_dev_alloc")
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i4", "data": (self.ptr.value, False), "version": 2}

    def tensor(self, shape):
        return torch.as_tensor(self, device="cuda").view(*shape)

    def free(self):
        from .lib import lib
        if self.ptr:
            lib().bb_dev_free(self.ptr)
            self.ptr = None


class FourStepFused:
    """One 2^log_n NTT over `world` GPUs with the exchange fused into the compute: the last pass of the column
    transforms multiplies by the inter-half twiddle and stores every row directly into the receive buffer of the
    rank that owns it (CUDA-IPC peer mappings: the stores travel over NVLink / NVSwitch), already in the layout the
    row transforms read.  Compared with `fourstep_ntt_cuda` this removes the twiddle pass, the NCCL all-to-all and
    the re-layout copy.  The ranks rendezvous on the device (flag words in peer memory, one signal + one wait kernel
    per transform, two or three receive buffers in rotation), so a transform involves no host synchronisation and no NCCL
    call; torch.distributed is used once, to exchange the IPC handles.

    `run` executes one transform on the caller's stream.  `run_async` (objects made with nbuf=3) keeps TWO transforms
    in flight on two internal streams, so that the column passes of transform e could overlap the row passes of e-1;
    `join` makes the caller's stream wait for everything issued.  Measured at 8 GPUs it gains nothing: the persistent
    CTAs of the scattering pass hold every SM while they wait on their NVLink stores, so the other stream's row passes
    only start when they drain."""

    FLAG_WORDS = 512  # 8 lines of 32 words for the flags, then the error word

    def __init__(self, log_n, rank, world, nbuf=None):
        import ctypes as C

        from .lib import check, lib
        self.log_n, self.rank, self.world = log_n, rank, world
        # receive buffers in rotation: 2 for `run` (default), `run_async` needs 3 (see its contract).  Measured on 8 B200s
        # at 2^27: 0.226 ms per transform with two buffers, 0.296 ms with three (the NVLink-store pass slows from 0.149 to
        # 0.216 ms, profiles/fourstep_nbuf_ab_8gpu_r2.json) — so three are allocated only on request.
        self.NBUF = int(nbuf or os.environ.get("TOYNI_FOURSTEP_NBUF", 2))
        assert self.NBUF in (2, 3)
        self.n1, self.n2 = fourstep_split(log_n, world)
        self.rw, self.cw = self.n1 // world, self.n2 // world
        self.buf_words = self.rw * self.n2
        # one allocation per rank: [receive buffers 0 .. NBUF-1][flags + error word]
        nb = self.NBUF
        self.mem = _RawCudaBuffer(nb * self.buf_words + self.FLAG_WORDS)
        whole = self.mem.tensor((nb * self.buf_words + self.FLAG_WORDS,))
        whole[nb * self.buf_words:].zero_()
        torch.cuda.synchronize()
        handle = (C.c_uint8 * 64)()
        check(lib().bb_ipc_get_handle(self.mem.ptr, handle), "bb_ipc_get_handle")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        if world > 1:
            allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
        else:
            allh = [mine]
        base = []
        self._opened = []
        for r in range(world):
            if r == rank:
                base.append(self.mem.ptr.value)
            else:
                hb = (C.c_uint8 * 64)(*allh[r].cpu().tolist())
                p = C.c_void_p()
                check(lib().bb_ipc_open_handle(hb, C.byref(p)), "bb_ipc_open_handle")
                base.append(p.value)
                self._opened.append(p)
        self.peers = [(C.c_void_p * world)(*[b + 4 * k * self.buf_words for b in base]) for k in range(nb)]
        flag_ptrs = [b + 4 * nb * self.buf_words for b in base]
        self.d_peer_flags = torch.tensor(flag_ptrs, dtype=torch.int64, device="cuda")  # device array of pointers
        self.flags_ptr = C.c_void_p(base[rank] + 4 * nb * self.buf_words)
        self.err_ptr = C.c_void_p(base[rank] + 4 * (nb * self.buf_words + 256))
        self.err = whole[nb * self.buf_words + 256:nb * self.buf_words + 257]
        self.out = [whole[k * self.buf_words:(k + 1) * self.buf_words].view(self.rw, self.n2) for k in range(nb)]
        self.epoch = 0
        self._streams = None   # run_async: two internal streams, events "peers waited" / "rows done" per epoch parity
        self._ev_waited = [None, None]
        self._ev_rows = [None, None]
        if world > 1:
            dist.barrier()  # every rank has mapped every buffer and zeroed its flags

    def run(self, block, inverse=False):
        """block: int32 CUDA tensor (n1, n2/G), this rank's columns (destroyed).  Returns the (n1/G, n2) receive
        buffer holding out[k1_local][k2] = X[k1 + n1*k2]; it stays valid until the transform after next is started."""
        import ctypes as C

        from .device import _bind_stream, _chk, ntt_batch_
        from .lib import check, lib
        assert tuple(block.shape) == (self.n1, self.cw)
        _bind_stream()
        if self._ev_rows[0] is not None:  # transforms issued by run_async may still be running on the internal streams
            self.join()
        self.epoch += 1
        k = self.epoch % self.NBUF
        L = lib()
        # buffer k was last read by the row transforms of epoch-3; every rank finished those before it signalled
        # epoch-1, and this rank has already waited for all epoch-1 signals: the peers' buffers are free
        check(L.bb_ntt_columns_scatter_device(_chk(block), self.log_n, self.n1.bit_length() - 1, self.cw,
                                              1 if inverse else 0, self.peers[k], self.world, self.rank),
              "bb_ntt_columns_scatter_device")
        check(L.bb_peer_signal_device(C.c_void_p(self.d_peer_flags.data_ptr()), self.world, self.rank, self.epoch),
              "bb_peer_signal_device")
        check(L.bb_peer_wait_device(self.flags_ptr, self.world, self.epoch, self.err_ptr), "bb_peer_wait_device")
        return ntt_batch_(self.out[k], inverse)

    def run_async(self, block, inverse=False):
        """Same transform, issued on one of two internal streams (alternating), so that consecutive calls overlap: the
        column passes + NVLink stores of this transform run beside the row passes of the previous one.  Returns
        (out, done): `out` as for `run`, `done` a CUDA event recorded after the row passes.  Contract: `block` was
        produced on the caller's current stream; `out` must have been consumed (work enqueued on the caller's stream
        after waiting for `done`) before run_async is called for the second time after this call."""
        import ctypes as C

        from .device import _bind_stream, _chk, ntt_batch_
        from .lib import check, lib
        assert tuple(block.shape) == (self.n1, self.cw)
        assert self.NBUF >= 3, "run_async needs three receive buffers (FourStepFused(..., nbuf=3))"
        if self._streams is None:
            self._streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        self.epoch += 1
        e = self.epoch
        k, par = e % self.NBUF, e & 1
        st = self._streams[par]
        L = lib()
        ready = torch.cuda.Event()
        ready.record()  # the caller's stream: the input block exists, and what the caller consumed so far is ordered
        with torch.cuda.stream(st):
            st.wait_event(ready)
            # all peers have signalled e-1 (they are done reading buffer k, last used by epoch e-3) once the OTHER
            # stream's wait kernel of e-1 has returned
            if self._ev_waited[par ^ 1] is not None:
                st.wait_event(self._ev_waited[par ^ 1])
            _bind_stream()
            check(L.bb_ntt_columns_scatter_device(_chk(block), self.log_n, self.n1.bit_length() - 1, self.cw,
                                                  1 if inverse else 0, self.peers[k], self.world, self.rank),
                  "bb_ntt_columns_scatter_device")
            # our signal of e tells the peers that our row passes of e-1 are finished
            if self._ev_rows[par ^ 1] is not None:
                st.wait_event(self._ev_rows[par ^ 1])
            check(L.bb_peer_signal_device(C.c_void_p(self.d_peer_flags.data_ptr()), self.world, self.rank, e), "bb_peer_signal_device")
            check(L.bb_peer_wait_device(self.flags_ptr, self.world, e, self.err_ptr), "bb_peer_wait_device")
            self._ev_waited[par] = torch.cuda.Event()
            self._ev_waited[par].record(st)
            out = ntt_batch_(self.out[k], inverse)
            self._ev_rows[par] = torch.cuda.Event()
            self._ev_rows[par].record(st)
        _bind_stream()
        return out, self._ev_rows[par]

    def join(self):
        """The caller's current stream waits for every transform issued by run_async."""
        cur = torch.cuda.current_stream()
        for ev in self._ev_rows:
            if ev is not None:
                cur.wait_event(ev)
        self._ev_rows = [None, None]
        self._ev_waited = [None, None]

    def check_peers(self):
        """Host-side check (synchronises): raises if a wait kernel gave up on a peer."""
        e = int(self.err.item())
        if e:
            raise RuntimeError(f"four-step rendezvous timed out waiting for rank {e - 1}")

    def close(self):
        from .lib import lib
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        for p in self._opened:
            lib().bb_ipc_close_handle(p)
        self._opened = []
        self.mem.free()


class FoldChain:
    """A prepared FRI fold chain on one cyclic shard (global indices rank, rank+G, ...): layer sizes, the betas as a
    C array and the output arena are set up once; `run(local)` is then ONE C call (bb_fri_fold_chain_shard_device)
    that launches every fold back to back.  At eight ranks a 2^25 chain is ~60 us of kernels, so per-call Python
    set-up (and, before that, a trip through Python per layer) was most of what the caller saw.
    No communication while the layer has at least 2*world values."""

    def __init__(self, log_m, shift, betas, rank, world, limbs, until=16, device="cuda"):
        self.log_m, self.shift, self.rank, self.world, self.limbs, self.until = log_m, shift % P, rank, world, limbs, until
        self.sizes, m = [], 1 << log_m
        while m > until and (m // 2) >= world:
            m //= 2
            self.sizes.append(m // world)
        self.m_local = (1 << log_m) // world
        self.bt = np.ascontiguousarray(np.asarray([[int(v) % P for v in (b if limbs == 4 else [b])] for b in betas[:len(self.sizes)]],
                                                  dtype=np.uint32).reshape(-1))
        total = max(1, sum(self.sizes))
        self.flat = torch.empty((total, 4) if limbs == 4 else (total,), dtype=torch.int32, device=device)

    def run(self, local):
        """local: this shard of layer 0.  Returns [local, layer 1, layer 2, ...] (views of the arena: valid until the
        next run)."""
        import ctypes as C

        from .device import _bind_stream, _chk
        from .lib import check, lib
        assert local.shape[0] == self.m_local and (4 if local.dim() == 2 else 1) == self.limbs
        if not self.sizes:
            return [local]
        _bind_stream()
        folds = C.c_size_t(0)
        check(lib().bb_fri_fold_chain_shard_device(_chk(local), self.m_local, self.log_m, self.shift, self.bt.ctypes.data, len(self.sizes),
                                                   self.limbs, self.world, self.rank, self.until, _chk(self.flat), C.byref(folds)),
              "bb_fri_fold_chain_shard_device")
        assert folds.value == len(self.sizes)
        layers, off = [local], 0
        for s in self.sizes:
            layers.append(self.flat[off:off + s])
            off += s
        return layers


def fold_chain_cuda(local, log_m, shift, betas, rank, world, until=16):
    """FRI fold chain on one cyclic shard, betas supplied up front: FoldChain prepared and run once.  Returns the list
    of local layers (owning their arena)."""
    return FoldChain(log_m, shift, betas, rank, world, 4 if local.dim() == 2 else 1, until, local.device).run(local)


# ------------------------------------------------------------------ sharded FRI commit loop
def merkle_top(roots):
    """Root of the tree whose leaves are the G subtree roots (node = SHA256(0x01 || L || R), src/merkle.rs:117-123)."""
    import hashlib
    level = [bytes(r) for r in roots]
    while len(level) > 1:
        level = [hashlib.sha256(b"\x01" + level[i] + level[i + 1]).digest() for i in range(0, len(level), 2)]
    return level[0]


def cyclic_to_block(local, world):
    """local[j] = layer[rank + G*j]  ->  block[l] = layer[rank*m/G + l].  One all-to-all: the values of rank t's block
    that this rank holds are the contiguous run local[t*c : (t+1)*c], c = m/G^2; on arrival the run from rank r
    supplies block positions r, r+G, r+2G, ... (an interleave of G runs)."""
    if world == 1:
        return local
    rows = local.shape[0]
    assert rows % world == 0, "layer too small to exchange: gather it instead"
    recv = torch.empty_like(local)
    dist.all_to_all_single(recv, local.contiguous())
    c = rows // world
    if local.is_cuda:
        from . import device as D
        return D.interleave(recv, world)
    tail = tuple(local.shape[1:])
    return recv.view((world, c) + tail).transpose(0, 1).contiguous().view((rows,) + tail)


def gather_cyclic(local, world):
    """The whole layer in natural order on every rank (small layers at the end of the chain)."""
    if world == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    full = torch.stack(parts, dim=1)  # full[j][r] = layer[r + G*j]
    return full.reshape((local.shape[0] * world,) + tuple(local.shape[1:])).contiguous()


class CudaFriBackend:
    """The device primitives the sharded loop is made of (C ABI kernels)."""

    def __init__(self, device):
        self.device = device

    def fold(self, local, log_m, x0, beta, world, rank):
        from .device import fri_fold_shard
        return fri_fold_shard(local, log_m, x0, beta, world, rank)

    def commit(self, vals, salts):
        from .device import merkle_commit
        nodes, _ = merkle_commit(vals, salts, want_root=False)
        return nodes, nodes[-1]  # the subtree root stays on the device: it goes straight into the all-gather

    def finish(self, layer_full, x0, final_size, salts, challenge):
        from .device import fri_commit
        return fri_commit(layer_full, x0, final_size, salts, challenge=challenge)

    def bytes_tensor(self, b):
        return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(self.device)


def fri_commit_sharded(local0, log_m, shift, final_size, salts_for, challenge, rank, world, backend, gather_below=1 << 16):
    """The prover's FRI commit loop (src/fibonacci.rs:200-247) over cyclic shards.
      local0      : this rank's shard of layer 0, local0[j] = layer0[rank + G*j]  ((m/G,) or (m/G, 4))
      salts_for   : salts_for(layer_index, lo, hi) -> (hi-lo, 16) uint8 tensor of the salts of leaves lo..hi-1 of that
                    layer (None for the unsalted final layer), in the backend's memory
      challenge   : challenge(root bytes, layer index) -> beta, called identically on every rank
    Layers of at least `gather_below` values (and G^2) are committed sharded; below that the layer is all-gathered and
    the rest of the loop runs replicated.  Returns (roots, local layers, per-layer local nodes, final layer)."""
    m = 1 << log_m
    x0 = shift % P
    roots, layers, nodes = [], [local0], []
    k = 0
    # (one rank: nothing to shard — the whole loop is the single-device one below, with its fused fold + leaf-hash kernels)
    while world > 1 and m > final_size and m >= gather_below and m >= world * world and (m // 2) >= world:
        block = cyclic_to_block(layers[-1], world)
        c = m // world
        nd, local_root = backend.commit(block, salts_for(k, rank * c, (rank + 1) * c))
        nodes.append(nd)
        mine = local_root if isinstance(local_root, torch.Tensor) else backend.bytes_tensor(local_root)
        if world > 1:
            allr = torch.empty((world, 32), dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(allr, mine.reshape(1, 32).contiguous())
            allr = allr.cpu().numpy()
            root = merkle_top([allr[r].tobytes() for r in range(world)])
        else:
            root = mine.cpu().numpy().tobytes()
        roots.append(root)
        beta = challenge(root, k)
        layers.append(backend.fold(layers[-1], log_m - k, x0, beta, world, rank))
        x0 = x0 * x0 % P
        m //= 2
        k += 1
    # tail: replicated single-device loop on the gathered layer (commits it, then folds down to final_size)
    full = gather_cyclic(layers[-1], world)
    off = [0]

    def tail_salts():
        parts, mm, kk = [], m, k
        while mm > final_size:
            parts.append(salts_for(kk, 0, mm))
            mm //= 2
            kk += 1
        if not parts:
            return None
        flat = [q.reshape(-1) for q in parts]
        # slices of one buffer that already lie back to back (the usual case: one salt arena, layer after layer) are used in
        # place; only scattered pieces are concatenated (1 GB at 2^25 Ext: 0.4 ms and the memory)
        if all(q.is_contiguous() for q in parts) and all(
                flat[i].untyped_storage().data_ptr() == flat[0].untyped_storage().data_ptr() and
                flat[i].storage_offset() == flat[i - 1].storage_offset() + flat[i - 1].numel() for i in range(1, len(flat))):
            total = sum(q.numel() for q in flat)
            return torch.empty(0, dtype=flat[0].dtype, device=flat[0].device).set_(flat[0].untyped_storage(), flat[0].storage_offset(), (total,))
        return torch.cat(flat)

    t_layers, t_nodes, t_roots = backend.finish(full, x0, final_size, tail_salts(), lambda root, layer: challenge(root, k + layer))
    roots.extend(t_roots)
    return roots, layers, nodes, t_layers
