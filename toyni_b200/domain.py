"""Host-side mirror of BabyBearDomain (src/math/domain.rs:9-175) over the GPU library.

The reference picks GPU or CPU per call (`use_gpu && cuda_available()`, :90-97,113-119) and keeps the
coset shift on the CPU.  This mirror IS the GPU branch: the shift is fused into the NTT passes on the
device, and without a device it raises instead of falling back."""
import numpy as np

from .lib import P, ToyniCudaError, check, cuda_available, lib


def get_root_of_unity(log_n):
    """src/babybear.rs:118-126 (host scalar, Python ints)."""
    assert log_n <= 27, "BabyBear only supports NTT up to 2^27"
    return pow(440564289, 1 << (27 - log_n), P)


def _u64(a, width=1):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
    if width > 1:
        a = a.reshape(-1, width)
    return a


class BabyBearDomain:
    def __init__(self, size, shift=1, use_gpu=True):
        assert size > 0 and (size & (size - 1)) == 0, "Domain size must be power of 2"  # :21
        self.size = size
        self.log_size = size.bit_length() - 1
        self.omega = get_root_of_unity(self.log_size)
        self.shift = shift % P
        self.use_gpu = use_gpu

    @classmethod
    def new(cls, size):
        return cls(size)

    def get_coset(self, shift):  # :34-42
        return BabyBearDomain(self.size, shift, self.use_gpu)

    def with_gpu(self, use_gpu):  # :45-48
        self.use_gpu = use_gpu
        return self

    def group_gen(self):
        return self.omega

    def _require_gpu(self):
        if not self.use_gpu:
            raise ToyniCudaError("this mirror only implements the GPU branch (use_gpu=False is the reference's CPU path)")
        if not cuda_available():
            raise ToyniCudaError("CUDA not available")

    def elements(self):
        """{shift * omega^i} (:61-69) = the coset evaluation of the polynomial X."""
        self._require_gpu()
        out = np.empty(self.size, dtype=np.uint64)
        check(lib().toyni_domain_elements(self.size, int(self.shift) % P, out.ctypes.data), "toyni_domain_elements")
        return out

    def fft(self, coeffs):
        """:107-123: zero-pad / truncate to `size`, coset shift, forward NTT."""
        self._require_gpu()
        c = _u64(coeffs)
        out = np.empty(self.size, dtype=np.uint64)
        check(lib().toyni_domain_fft(c.ctypes.data, c.size, self.size, self.shift, out.ctypes.data), "CUDA NTT")
        return out

    def ifft(self, evals):
        """:85-102: INTT, then divide coefficient i by shift^i."""
        self._require_gpu()
        e = _u64(evals)
        assert e.size == self.size, "Evaluation count must match domain size"  # :86
        out = np.empty(self.size, dtype=np.uint64)
        check(lib().toyni_domain_ifft(e.ctypes.data, self.size, self.shift, out.ctypes.data), "CUDA INTT")
        return out

    def fft_ext(self, coeffs):
        """:135-137: Ext arrays have shape (n, 4)."""
        self._require_gpu()
        c = _u64(coeffs, 4)
        out = np.empty((self.size, 4), dtype=np.uint64)
        check(lib().toyni_domain_fft_ext(c.ctypes.data, c.shape[0], self.size, self.shift, out.ctypes.data), "CUDA NTT")
        return out

    def ifft_ext(self, evals):
        """:130-132"""
        self._require_gpu()
        e = _u64(evals, 4)
        assert e.shape[0] == self.size, "Evaluation count must match domain size"
        out = np.empty((self.size, 4), dtype=np.uint64)
        check(lib().toyni_domain_ifft_ext(e.ctypes.data, self.size, self.shift, out.ctypes.data), "CUDA INTT")
        return out

    def vanishing_poly_coeffs(self):  # :74-80 (host scalar work)
        h_n = pow(self.shift, self.size, P)
        coeffs = np.zeros(self.size + 1, dtype=np.uint64)
        coeffs[0] = (P - h_n) % P
        coeffs[self.size] = 1
        return coeffs
