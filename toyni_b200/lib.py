"""ctypes loader for libntt_cuda.so (C ABI in include/toyni_ntt_cuda.h).  Fails loudly if the
extension is missing: the product path never falls back to a CPU implementation."""
import ctypes as C
import os

P = 2013265921  # src/babybear.rs:8
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
CHALLENGE_FN = C.CFUNCTYPE(None, C.c_void_p, u8p, C.c_uint32, u32p)

_SIGNATURES = {
    # 1. reference symbols (src/ntt.rs:96-110)
    "cuda_copy_to_device": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "cuda_copy_from_device": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "cuda_malloc": ([C.POINTER(C.c_void_p), C.c_size_t], C.c_int),
    "cuda_free": ([C.c_void_p], C.c_int),
    "cuda_get_error_string": ([C.c_int], C.c_char_p),
    "ntt_ctx_create": ([C.c_uint32], C.c_void_p),
    "ntt_ctx_destroy": ([C.c_void_p], None),
    "ntt_run_inplace": ([C.c_void_p, C.c_void_p], None),
    "intt_run_inplace": ([C.c_void_p, C.c_void_p], None),
    "ntt_run_inplace_rc": ([C.c_void_p, C.c_void_p], C.c_int),
    "intt_run_inplace_rc": ([C.c_void_p, C.c_void_p], C.c_int),
    "bb_ntt_host_u32": ([C.c_void_p, C.c_void_p, C.c_int], C.c_int),
    "bb_domain_elements_device": ([C.c_uint32, C.c_uint32, C.c_void_p], C.c_int),
    "toyni_roots_of_unity_domain": ([C.c_size_t, C.c_void_p], C.c_int),
    "toyni_domain_elements": ([C.c_size_t, C.c_uint64, C.c_void_p], C.c_int),
    # 2. device-resident API
    "bb_last_error": ([], C.c_int),
    "bb_last_error_string": ([], C.c_char_p),
    "bb_clear_error": ([], None),
    "bb_device_ok": ([], C.c_int),
    "bb_set_stream": ([C.c_void_p], None),
    "bb_sync": ([], C.c_int),
    "bb_dev_alloc": ([C.POINTER(C.c_void_p), C.c_size_t], C.c_int),
    "bb_dev_free": ([C.c_void_p], C.c_int),
    "toyni_fri_salt_bytes": ([C.c_size_t], C.c_size_t),
    "toyni_prove_fibonacci": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                               C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
    "toyni_prover_error": ([], C.c_char_p),
    "bb_merkle_open_multi_device": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_void_p], C.c_int),
    "bb_pool_alloc": ([C.POINTER(C.c_void_p), C.c_size_t], C.c_int),
    "bb_pool_free": ([C.c_void_p], C.c_int),
    "bb_pool_trim": ([], C.c_int),
    "bb_h2d": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "bb_d2h": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "bb_d2d": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "bb_narrow_u64_to_u32": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "bb_widen_u32_to_u64": ([C.c_void_p, C.c_void_p, C.c_size_t], C.c_int),
    "bb_ntt_device": ([C.c_void_p, C.c_uint32, C.c_int], C.c_int),
    "bb_ntt_batch_device": ([C.c_void_p, C.c_uint32, C.c_size_t, C.c_int], C.c_int),
    "bb_ntt_ext_device": ([C.c_void_p, C.c_uint32, C.c_int], C.c_int),
    "bb_ntt_columns_device": ([C.c_void_p, C.c_uint32, C.c_size_t, C.c_int], C.c_int),
    "bb_fourstep_twiddle_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t, C.c_int], C.c_int),
    "bb_ntt_columns_scatter_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.c_uint32,
                                       C.c_uint32], C.c_int),
    "bb_peer_signal_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32], C.c_int),
    "bb_peer_wait_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p], C.c_int),
    "bb_ipc_get_handle": ([C.c_void_p, C.c_void_p], C.c_int),
    "bb_ipc_open_handle": ([C.c_void_p, C.POINTER(C.c_void_p)], C.c_int),
    "bb_ipc_close_handle": ([C.c_void_p], C.c_int),
    "bb_coset_fft_device": ([C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p], C.c_int),
    "bb_coset_ifft_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_int], C.c_int),
    "bb_fri_fold_device": ([C.c_void_p, C.c_size_t, C.c_uint32, u32p, C.c_int, C.c_void_p], C.c_int),
    "bb_fri_fold_shard_device": ([C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, u32p, C.c_int, C.c_uint32, C.c_uint32,
                                  C.c_void_p], C.c_int),
    "bb_fri_fold_chain_shard_device": ([C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32,
                                        C.c_uint32, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)], C.c_int),
    "bb_fri_fold_xs_device": ([C.c_void_p, C.c_size_t, C.c_void_p, u32p, C.c_int, C.c_void_p], C.c_int),
    "bb_merkle_node_count": ([C.c_size_t], C.c_size_t),
    "bb_merkle_commit_device": ([C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "bb_merkle_build_bytes_device": ([C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p], C.c_int),
    "bb_merkle_open_device": ([C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)], C.c_int),
    "bb_merkle_open_batch_device": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)], C.c_int),
    "bb_gather_device": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int),
    "bb_interleave_device": ([C.c_void_p, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p], C.c_int),
    "bb_fib_constraint_device": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p], C.c_int),
    "bb_scale_periodic_device": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32], C.c_int),
    "bb_fib_deep_device": ([C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                            C.c_uint32, C.c_void_p], C.c_int),
    "bb_poly_eval_device": ([C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)], C.c_int),
    "bb_fri_commit_device": ([C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)], C.c_int),
    "bb_ntt_set_plan": ([C.c_uint32, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "bb_ntt_get_plan": ([C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "bb_ntt_set_kernel": ([C.c_int], None),
    "bb_ntt_diag": ([C.POINTER(C.c_uint32)], C.c_int),
    "bb_ntt_launches": ([C.c_uint32], C.c_int),
    "bb_kernel_launch_count": ([], C.c_ulonglong),
    "bb_warmup": ([C.c_uint32], C.c_int),
    "bb_release": ([], None),
    # 3. host-pointer forms of the reference call sites
    "toyni_domain_fft": ([C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_void_p], C.c_int),
    "toyni_domain_ifft": ([C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p], C.c_int),
    "toyni_domain_fft_ext": ([C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_void_p], C.c_int),
    "toyni_domain_ifft_ext": ([C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p], C.c_int),
    "toyni_fri_fold": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p], C.c_int),
    "toyni_fri_fold_ext": ([C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    "toyni_merkle_commit": ([C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
    # section 4: one process, G devices
    "bb_mg_init": ([C.c_int, C.POINTER(C.c_void_p)], C.c_int),
    "bb_mg_destroy": ([C.c_void_p], None),
    "bb_mg_ngpus": ([C.c_void_p], C.c_int),
    "bb_mg_sync": ([C.c_void_p], C.c_int),
    "bb_mg_stream": ([C.c_void_p, C.c_int], C.c_void_p),
    "bb_mg_ntt_fourstep": ([C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)], C.c_int),
    "bb_mg_ntt_batch": ([C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)], C.c_int),
    "bb_mg_fri_chain": ([C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                         C.POINTER(C.c_size_t)], C.c_int),
    "bb_mg_ntt_host": ([C.c_void_p, C.c_void_p, C.c_uint32, C.c_int], C.c_int),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path():
    return os.environ.get("TOYNI_NTT_LIB") or os.path.join(_HERE, "libntt_cuda.so")  # env: tuning builds only


class ToyniCudaError(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (python toyni_b200/build.py)."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ToyniCudaError(f"{path} is missing: build it with `python toyni_b200/build.py` "
                                 "(there is no CPU fallback)")
        L = C.CDLL(path)
        for name, (args, res) in _SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _LIB = L
    return _LIB


def check(rc, what=""):
    if rc != 0:
        L = lib()
        msg = L.cuda_get_error_string(rc)
        raise ToyniCudaError(f"{what} failed: CUDA error {rc} ({msg.decode() if msg else '?'})")


def cuda_available():
    """src/ntt.rs:144-150: cudaGetDeviceCount succeeds and reports a device (resolved from libcudart)."""
    try:
        L = lib()
    except ToyniCudaError:
        raise
    rt = C.CDLL("libcudart.so", mode=C.RTLD_GLOBAL) if not hasattr(L, "cudaGetDeviceCount") else L
    cnt = C.c_int(0)
    try:
        err = rt.cudaGetDeviceCount(C.byref(cnt))
    except AttributeError:
        return False
    return err == 0 and cnt.value > 0
