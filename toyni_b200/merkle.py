"""Host-side mirror of the prover's Merkle commits (src/merkle.rs:10-84, src/fibonacci.rs:325-374) over
the GPU library: the tree is hashed on the device; openings and verification are host logic."""
import hashlib

import numpy as np

from .lib import check, lib


class MerkleProof:
    """src/merkle.rs:4-8"""

    def __init__(self, path, position):
        self.path = path          # list of 32-byte digests
        self.position = position  # list of bools


def _hash_leaf(data):  # src/merkle.rs:109-114 (host side of verify)
    return hashlib.sha256(b"\x00" + data).digest()


def _hash_node(l, r):  # src/merkle.rs:117-123
    return hashlib.sha256(b"\x01" + l + r).digest()


def verify_merkle_proof(leaf, proof, root):
    """src/merkle.rs:86-101"""
    cur = _hash_leaf(leaf)
    for sib, is_right in zip(proof.path, proof.position):
        cur = _hash_node(sib, cur) if is_right else _hash_node(cur, sib)
    return cur == root


class SaltedTree:
    """src/fibonacci.rs:325-337: a Merkle tree over field values plus the per-leaf salts.
    `nodes` holds every level (leaf level first) as uint8[count, 32], as hashed on the GPU."""

    def __init__(self, nleaves, nodes, root, salts):
        self.nleaves = nleaves
        self.nodes = nodes
        self._root = root
        self.salts = salts  # uint8[n,16] or None

    def root(self):
        return self._root

    def level_offsets(self):
        offs, n, o = [], self.nleaves, 0
        while True:
            offs.append((o, n))
            if n <= 1:
                break
            o += n
            n = (n + 1) // 2
        return offs

    def get_proof(self, index):
        """src/merkle.rs:50-80"""
        if index >= self.nleaves:
            return None
        path, position, cur = [], [], index
        for off, n in self.level_offsets()[:-1]:
            sib = cur + 1 if cur % 2 == 0 else cur - 1
            if sib >= n:
                path.append(self.nodes[off + cur].tobytes())
                position.append(True)
            else:
                path.append(self.nodes[off + sib].tobytes())
                position.append(cur % 2 == 1)
            cur //= 2
        return MerkleProof(path, position)


def _commit(evals, salts, limbs):
    v = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64))
    n = v.size // limbs
    L = lib()
    nodes = np.empty((L.bb_merkle_node_count(n), 32), dtype=np.uint8)
    root = np.empty(32, dtype=np.uint8)
    s = None
    if salts is not None:
        s = np.ascontiguousarray(np.asarray(salts, dtype=np.uint8)).reshape(n, 16)
    check(L.toyni_merkle_commit(v.ctypes.data, n, limbs, None if s is None else s.ctypes.data, nodes.ctypes.data,
                                root.ctypes.data), "merkle commit")
    return SaltedTree(n, nodes, root.tobytes(), s)


def build_merkle_tree(evals, salts, limbs=1):
    """src/fibonacci.rs:340-353 with the salts passed in (the reference draws them from thread_rng)."""
    return _commit(evals, salts, limbs)


def build_unsalted_tree(evals, limbs=1):
    """src/fibonacci.rs:357-363"""
    return _commit(evals, None, limbs)
