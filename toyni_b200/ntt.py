"""Host-side mirror of src/ntt.rs `mod cuda` (lines 85-315): same names, argument meaning and
error behaviour, over the same C ABI the Rust side binds (src/ntt.rs:96-110)."""
import ctypes as C
import threading

import numpy as np

from .lib import ToyniCudaError, check, cuda_available, lib  # noqa: F401

_CTX_CACHE = {}
_CTX_LOCK = threading.Lock()


def _get_or_create_ctx(n):
    """src/ntt.rs:128-141: per-size context cache, kept for the life of the process."""
    with _CTX_LOCK:
        ctx = _CTX_CACHE.get(n)
        if ctx is None:
            ctx = lib().ntt_ctx_create(n)
            if not ctx:
                raise ToyniCudaError(f"ntt_ctx_create({n}) failed: {lib().bb_last_error_string().decode()}")
            _CTX_CACHE[n] = ctx
        return ctx


def _as_u64(values):
    if not (isinstance(values, np.ndarray) and values.dtype == np.uint64 and values.flags["C_CONTIGUOUS"]):
        raise TypeError("values must be a C-contiguous numpy uint64 array (the reference's BabyBear storage)")
    return values


class CudaBuffer:
    """src/ntt.rs:153-212: RAII wrapper of `size` u64 elements of device memory."""

    def __init__(self, size):
        self.size = size
        self._ptr = C.c_void_p()
        err = lib().cuda_malloc(C.byref(self._ptr), size)
        if err != 0:
            raise ToyniCudaError("CUDA malloc failed: " + lib().cuda_get_error_string(err).decode())

    def copy_from_host(self, data):
        data = _as_u64(data)
        assert data.size == self.size, "Size mismatch"
        err = lib().cuda_copy_to_device(self._ptr, data.ctypes.data, self.size)
        if err != 0:
            raise ToyniCudaError("CUDA copy to device failed: " + lib().cuda_get_error_string(err).decode())

    def copy_to_host(self, data):
        data = _as_u64(data)
        assert data.size == self.size, "Size mismatch"
        err = lib().cuda_copy_from_device(data.ctypes.data, self._ptr, self.size)
        if err != 0:
            raise ToyniCudaError("CUDA copy from device failed: " + lib().cuda_get_error_string(err).decode())

    def as_ptr(self):
        return self._ptr.value

    def __del__(self):
        try:
            if self._ptr:
                lib().cuda_free(self._ptr)
                self._ptr = C.c_void_p()
        except Exception:
            pass


def _run(values, inverse):
    if not cuda_available():
        raise ToyniCudaError("CUDA not available")  # src/ntt.rs:225-227
    values = _as_u64(values)
    n = values.size
    assert n > 0 and (n & (n - 1)) == 0, "NTT size must be power of 2"  # :229
    assert n.bit_length() - 1 <= 27, "BabyBear only supports NTT up to 2^27"  # :230
    ctx = _get_or_create_ctx(n)
    L = lib()
    # the reference's void entry points (src/ntt.rs:108-109) cannot report a failed copy or launch; the _rc forms can
    err = (L.intt_run_inplace_rc if inverse else L.ntt_run_inplace_rc)(ctx, values.ctypes.data)
    if err:
        raise ToyniCudaError("CUDA NTT failed: " + L.cuda_get_error_string(err).decode())


def ntt_cuda(values):
    """Forward NTT on the GPU, in place (src/ntt.rs:224-236)."""
    _run(values, False)


def intt_cuda(values):
    """Inverse NTT on the GPU, in place (src/ntt.rs:239-251)."""
    _run(values, True)


def roots_of_unity_domain(n):
    """[omega^i for i < n] (src/ntt.rs:69-81), from the device twiddle cache instead of n sequential multiplications."""
    if not cuda_available():
        raise ToyniCudaError("CUDA not available")
    assert n > 0 and (n & (n - 1)) == 0, "Domain size must be power of 2"  # :70
    out = np.empty(n, dtype=np.uint64)
    err = lib().toyni_roots_of_unity_domain(n, out.ctypes.data)
    if err:
        raise ToyniCudaError("CUDA roots_of_unity_domain failed: " + lib().cuda_get_error_string(err).decode())
    return out
