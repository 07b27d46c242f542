"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess
import tempfile

from toyni_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "toyni_ntt_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:bb_|toyni_|cuda_|ntt_|intt_)[a-z0-9_]+)\s*\(", text))
    names.discard("bb_challenge_fn")
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_symbols()
    assert len(names) >= 45
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/toyni_ntt_cuda.h but not exported"
    assert names == set(L.EXPORTED_SYMBOLS), names ^ set(L.EXPORTED_SYMBOLS)


def test_reference_symbols_keep_their_names():
    """src/ntt.rs:96-110 imports exactly these from the static library `ntt_cuda`."""
    want = ["cuda_copy_to_device", "cuda_copy_from_device", "cuda_malloc", "cuda_free", "cuda_get_error_string",
            "ntt_ctx_create", "ntt_run_inplace", "intt_run_inplace"]
    out = subprocess.check_output(["nm", "-g", os.path.join(ROOT, "toyni_b200", "libntt_cuda.a")], text=True)
    for w in want:
        assert re.search(rf" T {w}$", out, flags=re.M), w


def test_build_lists_agree():
    """toyni_b200/build.py names its translation units explicitly; rust/build.rs compiles every .cu under cuda/ (a copy
    of toyni_b200/csrc).  Both build the same library iff the explicit list is exactly the .cu files that exist."""
    from toyni_b200 import build as B
    on_disk = sorted(f for f in os.listdir(os.path.join(ROOT, "toyni_b200", "csrc")) if f.endswith(".cu"))
    assert sorted(B.SOURCES) == on_disk
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    assert 'ends_with(".cu")' in rs and "read_dir" in rs and "arch=compute_100a,code=sm_100a" in rs
    assert not re.search(r"sm_(7|8|9|12)\d", rs), "sm_100a only"


def test_tma_kernel_is_in_the_library():
    """The hot NTT pass stages its tiles with TMA: the SASS of the shipped library carries UTMALDG and mbarrier (SYNCS) ops."""
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "ntt_v7.sm_100a.cubin", L.library_path()], cwd=td, stdout=subprocess.DEVNULL)
        sass = subprocess.check_output(["cuobjdump", "-sass", os.path.join(td, "ntt_v7.sm_100a.cubin")], text=True)
    assert "ntt_pass_v7_kernel" in sass and "UTMALDG" in sass and "SYNCS" in sass


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", L.library_path()], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    ptx = subprocess.run(["cuobjdump", "-lptx", L.library_path()], capture_output=True, text=True).stdout
    assert "sm_" not in ptx, "no PTX fallback is shipped"


def test_pure_host_entry_points_without_gpu():
    lib = L.lib()
    assert lib.bb_merkle_node_count(1) == 1
    assert lib.bb_merkle_node_count(4) == 7
    assert lib.bb_merkle_node_count(3) == 6     # 3 + 2 + 1 (odd level duplicates its last node)
    assert lib.bb_merkle_node_count(5) == 5 + 3 + 2 + 1
    assert lib.cuda_get_error_string(0) == b"no error"
    lr, lc = (ctypes.c_int * 3)(), (ctypes.c_int * 3)()
    assert lib.bb_ntt_get_plan(24, lr, lc) == 2 and list(lr)[:2] == [12, 12]   # the TMA-staged two-pass plan
    lib.bb_ntt_set_kernel(0)
    assert lib.bb_ntt_get_plan(24, lr, lc) == 3 and sum(lr) == 24               # tile kernel: three passes
    lib.bb_ntt_set_kernel(1)
    assert lib.bb_ntt_get_plan(27, lr, lc) == 3 and sum(lr) == 27
    assert lib.bb_ntt_get_plan(12, lr, lc) == 2 and sum(lr[:2]) == 12
    assert lib.ntt_ctx_create(3) is None          # not a power of two
    assert lib.ntt_ctx_create(1 << 28) is None     # beyond the two-adicity, cuda/ntt_kernel.cu:220
    lib.bb_clear_error()
