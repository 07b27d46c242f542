"""Host-side mirror of src/math/fri.rs (fri_fold, fri_fold_ext) over the GPU library."""
import numpy as np

from .lib import P, check, lib


def fri_fold(evals, xs, beta):
    """src/math/fri.rs:27-48. evals: uint64[m]; xs: uint64[>= m/2]; beta: int."""
    e = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64))
    x = np.ascontiguousarray(np.asarray(xs, dtype=np.uint64))
    assert e.size % 2 == 0, "Evaluations length must be even"
    assert x.size >= e.size // 2
    out = np.empty(e.size // 2, dtype=np.uint64)
    check(lib().toyni_fri_fold(e.ctypes.data, e.size, x.ctypes.data, int(beta) % P, out.ctypes.data), "fri_fold")
    return out


def fri_fold_ext(evals, xs, beta):
    """src/math/fri.rs:7-25. evals: uint64[m,4]; xs: uint64[>= m/2] (base field); beta: 4 limbs."""
    e = np.ascontiguousarray(np.asarray(evals, dtype=np.uint64)).reshape(-1, 4)
    x = np.ascontiguousarray(np.asarray(xs, dtype=np.uint64))
    b = np.ascontiguousarray(np.asarray(beta, dtype=np.uint64))
    assert e.shape[0] % 2 == 0, "Evaluations length must be even"
    assert x.size >= e.shape[0] // 2 and b.size == 4
    out = np.empty((e.shape[0] // 2, 4), dtype=np.uint64)
    check(lib().toyni_fri_fold_ext(e.ctypes.data, e.shape[0], x.ctypes.data, b.ctypes.data, out.ctypes.data), "fri_fold_ext")
    return out
