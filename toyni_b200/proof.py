"""Canonical byte form of a StarkProof (product side).

The reference has no serialization: `StarkProof` only derives `Debug` (src/fibonacci.rs:62-86).  This module defines
the format the B200 prover emits, so that "the final serialized Fibonacci proof" of two implementations can be
compared byte for byte:

  * integers and field values: 8 bytes little-endian (the reference's `BabyBear::to_bytes`, src/babybear.rs:53-55)
  * digests: 32 raw bytes; lists: u64 length prefix, then the items; structs: fields in declaration order
  * MerkleOpening (src/fibonacci.rs:44-51): index, value, salt length (1 byte) + salt, path length,
    then per level the 32-byte sibling followed by one byte (1 = sibling on the right)
  * QueryProof (:53-60): index, the six trace / quotient / DEEP openings in the order below, then the FRI pairs
  * StarkProof (:62-86): trace_len, lde_size, the two commitments, t_z, t_gz, t_ggz, q_z, FRI roots, final layer, queries

`deserialize_proof` is the inverse; `serialize_proof(deserialize_proof(b)) == b`.
"""
import struct

OPENINGS = ("deep_opening", "deep_opening_pair", "trace_opening", "trace_opening_g", "trace_opening_gg", "quotient_opening")
EVALS = ("t_z", "t_gz", "t_ggz", "q_z")


def _opening(out, o):
    out += struct.pack("<QQ", int(o["index"]), int(o["value"]))
    salt = bytes(o["salt"])
    assert len(salt) < 256
    out.append(len(salt))
    out += salt
    out += struct.pack("<Q", len(o["path"]))
    for digest, right in zip(o["path"], o["position"]):
        d = bytes(digest)
        assert len(d) == 32
        out += d
        out.append(1 if right else 0)


def serialize_proof(p):
    out = bytearray()
    out += struct.pack("<QQ", int(p["trace_len"]), int(p["lde_size"]))
    out += bytes(p["trace_commitment"]) + bytes(p["quotient_commitment"])
    for k in EVALS:
        out += struct.pack("<Q", int(p[k]))
    out += struct.pack("<Q", len(p["fri_commitments"]))
    for r in p["fri_commitments"]:
        out += bytes(r)
    out += struct.pack("<Q", len(p["fri_final_layer"]))
    for v in p["fri_final_layer"]:
        out += struct.pack("<Q", int(v))
    out += struct.pack("<Q", len(p["query_proofs"]))
    for q in p["query_proofs"]:
        out += struct.pack("<Q", int(q["index"]))
        for k in OPENINGS:
            _opening(out, q[k])
        out += struct.pack("<Q", len(q["fri_openings"]))
        for a, b in q["fri_openings"]:
            _opening(out, a)
            _opening(out, b)
    return bytes(out)


class _Reader:
    def __init__(self, data):
        self.b, self.o = memoryview(data), 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated proof")
        v = bytes(self.b[self.o:self.o + n])
        self.o += n
        return v

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def opening(self):
        index, value = self.u64(), self.u64()
        salt = self.take(self.take(1)[0])
        n = self.u64()
        if n > 64:
            raise ValueError("implausible path length")
        path, pos = [], []
        for _ in range(n):
            path.append(self.take(32))
            pos.append(self.take(1)[0] != 0)
        return {"index": index, "value": value, "salt": salt, "path": path, "position": pos}


def deserialize_proof(data):
    r = _Reader(data)
    p = {"trace_len": r.u64(), "lde_size": r.u64(), "trace_commitment": r.take(32), "quotient_commitment": r.take(32)}
    for k in EVALS:
        p[k] = r.u64()
    p["fri_commitments"] = [r.take(32) for _ in range(r.u64())]
    p["fri_final_layer"] = [r.u64() for _ in range(r.u64())]
    qs = []
    for _ in range(r.u64()):
        q = {"index": r.u64()}
        for k in OPENINGS:
            q[k] = r.opening()
        q["fri_openings"] = [(r.opening(), r.opening()) for _ in range(r.u64())]
        qs.append(q)
    p["query_proofs"] = qs
    if r.o != len(r.b):
        raise ValueError("trailing bytes after the proof")
    return p
