"""Config 1: Fibonacci AIR proofs assembled from GPU-computed pieces are byte-identical to the CPU oracle's, the
oracle verifier (src/verifier.rs restated) accepts them, and the reference's tamper cases are rejected."""
import copy

import numpy as np
import pytest

from oracle import fibonacci as F


def test_oracle_prover_and_verifier_cpu():
    """test_fibonacci / test_verify_valid_proof at the reference's size (trace_len 64, src/fibonacci.rs:423,
    src/verifier.rs:290), literal Lagrange interpolation vs the INTT shortcut, and the proof shape of Appendix A."""
    tr = F.fibonacci_trace(64)
    rnd = F.proof_randomness(64)
    p = F.generate_proof(tr, *rnd)
    assert F.verify(p)
    assert len(p["fri_commitments"]) == 9 and len(p["fri_final_layer"]) == 8 and len(p["query_proofs"]) == 44
    assert F.serialize_proof(p) == F.serialize_proof(F.generate_proof(tr, *rnd, interpolate="intt"))
    other = F.generate_proof(tr, *F.proof_randomness(64, seed=99))
    assert F.serialize_proof(other) != F.serialize_proof(p)  # src/verifier.rs:304-312: blinding changes the proof
    bad_trace = tr.copy()
    bad_trace[10] = (bad_trace[10] + 1) % F.P
    with pytest.raises(AssertionError):  # src/fibonacci.rs:433-455: a corrupted trace fails the OOD check
        F.generate_proof(bad_trace, *rnd)


def test_product_serializer_matches_the_oracle_format_cpu():
    """toyni_b200/proof.py (product) against oracle/fibonacci.py::serialize_proof (checker) on an oracle-made proof:
    same bytes, lossless round trip, truncation and trailing bytes rejected."""
    from toyni_b200 import proof as product_proof
    p = F.generate_proof(F.fibonacci_trace(64), *F.proof_randomness(64), interpolate="intt")
    blob = product_proof.serialize_proof(p)
    assert blob == F.serialize_proof(p)
    back = product_proof.deserialize_proof(blob)
    assert product_proof.serialize_proof(back) == blob and F.verify(back)
    with pytest.raises(ValueError):
        product_proof.deserialize_proof(blob[:-1])
    with pytest.raises(ValueError):
        product_proof.deserialize_proof(blob + b"\x00")


def _tamper_cases(p):  # src/verifier.rs:314-380
    a = copy.deepcopy(p); a["t_z"] = (a["t_z"] + 1) % F.P; yield a
    a = copy.deepcopy(p); a["q_z"] = (a["q_z"] + 1) % F.P; yield a
    a = copy.deepcopy(p); a["trace_commitment"] = bytes(32); yield a
    a = copy.deepcopy(p); a["fri_final_layer"][0] = (a["fri_final_layer"][0] + 1) % F.P; yield a
    a = copy.deepcopy(p); a["query_proofs"][0]["deep_opening"]["value"] = (a["query_proofs"][0]["deep_opening"]["value"] + 1) % F.P; yield a
    a = copy.deepcopy(p); a["fri_commitments"] = a["fri_commitments"][:-1]; yield a
    a = copy.deepcopy(p); a["query_proofs"][3]["fri_openings"][0][0]["value"] ^= 1; yield a


def test_verifier_rejects_tampering_cpu():
    p = F.generate_proof(F.fibonacci_trace(64), *F.proof_randomness(64))
    for bad in _tamper_cases(p):
        assert not F.verify(bad)


def test_algebraic_verifier_agrees_with_the_literal_one_cpu():
    """The closed-form domain membership / coset points used for large proofs give the same verdicts and the same z."""
    from oracle import oracle as O
    p = F.generate_proof(F.fibonacci_trace(64), *F.proof_randomness(64), interpolate="intt")
    assert F.verify(p, algebraic=True) and F.verify(p, algebraic=False)
    for bad in _tamper_cases(p):
        assert not F.verify(bad, algebraic=True)
    for seed in range(5):
        t1, t2 = O.FiatShamirTranscript(), O.FiatShamirTranscript()
        t1.absorb(bytes([seed]) * 32); t2.absorb(bytes([seed]) * 32)
        lde = 2048
        z1 = F.derive_z(t1, O.domain_elements(lde, 1), O.domain_elements(lde, F.COSET_SHIFT), O.root_of_unity(11))
        assert z1 == F.derive_z_algebraic(t2, lde)


@pytest.mark.gpu
@pytest.mark.parametrize("trace_len", [64, 1 << 10])
def test_gpu_proof_is_byte_identical_and_verifies(trace_len):
    from toyni_b200 import prover
    tr = F.fibonacci_trace(trace_len)
    rnd = F.proof_randomness(trace_len)
    ref = F.generate_proof(tr, *rnd, interpolate="lagrange" if trace_len == 64 else "intt")
    got = prover.generate_proof(tr, *rnd)
    assert F.serialize_proof(got) == F.serialize_proof(ref)
    # the product's own serializer (toyni_b200/proof.py) emits the same bytes as the oracle's, and reads them back
    from toyni_b200 import proof as product_proof
    blob = product_proof.serialize_proof(got)
    assert blob == F.serialize_proof(ref)
    assert product_proof.serialize_proof(product_proof.deserialize_proof(blob)) == blob
    assert F.verify(product_proof.deserialize_proof(blob))
    assert F.verify(got)
    if trace_len == 1 << 10:  # shape of BASELINE config 1
        assert got["lde_size"] == 32768 and len(got["fri_commitments"]) == 12 and len(got["fri_final_layer"]) == 16
    for bad in _tamper_cases(got):
        assert not F.verify(bad)


@pytest.mark.gpu
def test_gpu_proof_at_2_14_stays_on_the_device_and_verifies():
    """Beyond the size the O(n^3) reference prover can reach: trace 2^14 (LDE 2^19), every LDE-sized array on the device
    (constraint / quotient / DEEP kernels, batched openings), salts drawn on the device; the restated verifier accepts
    and the tamper cases are rejected."""
    import torch
    from oracle import oracle as O
    from toyni_b200 import prover
    trace_len = 1 << 14
    lde = trace_len * 32
    tr = F.fibonacci_trace(trace_len)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    salts = [torch.randint(0, 256, (m, 16), dtype=torch.uint8, device="cuda", generator=g) for m in (lde, lde, 2 * lde)]
    p = prover.generate_proof(tr, O.random_field(prover.MASK_DEGREE, 3), *salts)
    assert p["lde_size"] == lde and len(p["query_proofs"]) == 44
    assert F.verify(p, algebraic=False)
    assert F.verify(p, algebraic=True)
    for bad in _tamper_cases(p):
        assert not F.verify(bad, algebraic=True)


@pytest.mark.gpu
@pytest.mark.parametrize("trace_len", [64, 1 << 10])
def test_one_call_prover_gives_the_oracle_proof_bytes(trace_len):
    """toyni_prove_fibonacci (C ABI section 5: the whole loop of src/fibonacci.rs:99-310 behind one call, host salts) against
    the oracle prover on the same randomness: identical canonical bytes; a wrong mask length is refused like the
    reference's own length assumption."""
    from toyni_b200 import prover
    tr = F.fibonacci_trace(trace_len)
    rnd = F.proof_randomness(trace_len)
    ref = F.generate_proof(tr, *rnd, interpolate="intt")
    blob = prover.generate_proof_native(tr, *rnd, as_bytes=True)
    assert blob == F.serialize_proof(ref)
    assert F.verify(prover.generate_proof_native(tr, *rnd))
    with pytest.raises(AssertionError):
        prover.generate_proof_native(tr, rnd[0][:-1], *rnd[1:])


@pytest.mark.gpu
def test_one_call_prover_with_device_salts_equals_the_python_loop():
    """Same call with the salts resident on the device (trace 2^14): the bytes of the Python-driven loop over the same
    device primitives, and a proof the restated verifier accepts."""
    import torch
    from oracle import oracle as O
    from toyni_b200 import proof as product_proof
    from toyni_b200 import prover
    trace_len = 1 << 14
    lde = trace_len * 32
    tr = F.fibonacci_trace(trace_len)
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    salts = [torch.randint(0, 256, (m, 16), dtype=torch.uint8, device="cuda", generator=g) for m in (lde, lde, 2 * lde)]
    mask = O.random_field(prover.MASK_DEGREE, 4)
    a = prover.generate_proof_native(tr, mask, *salts, as_bytes=True)
    b = product_proof.serialize_proof(prover.generate_proof(tr, mask, *salts))
    assert a == b
    assert F.verify(product_proof.deserialize_proof(a), algebraic=True)
