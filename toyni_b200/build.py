"""Build the CUDA library in-tree for sm_100a only (the north star's build.rs contract: no other arch,
no PTX fallback).  Produces toyni_b200/libntt_cuda.so (loaded by the Python host mirror and the tests) and
toyni_b200/libntt_cuda.a (the static library name the reference's Rust side links, src/ntt.rs:95)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "nvcc")
EXTRA = os.environ.get("BB_NVCC_EXTRA", "").split()
FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-diag-suppress", "186,128", "-I", os.path.join(HERE, "..", "include")]
SOURCES = ["ntt_v7.cu", "ntt_v4_inst_d.cu", "ntt_v4_inst_c.cu", "ntt_v4_inst_b.cu", "ntt_v4_inst_a.cu", "ntt_inst_a.cu", "ntt_inst_b.cu", "ntt_inst_c.cu", "ntt_inst_d.cu", "ntt_dispatch.cu", "ntt_engine.cu",
           "fri_fold.cu", "elementwise.cu", "prover_ew.cu", "merkle.cu", "c_abi.cu", "mg.cu", "prover_abi.cu"]
SO = os.path.join(HERE, "libntt_cuda.so")
AR = os.path.join(HERE, "libntt_cuda.a")


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "toyni_ntt_cuda.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _newest_header()):
        return obj
    cmd = [NVCC] + FLAGS + ["-c", spath, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call([NVCC, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs)
        if os.path.exists(AR):
            os.remove(AR)
        subprocess.check_call(["ar", "rcs", AR] + objs)
    return SO


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
