"""Take the warp-private pass kernel (ntt_pass_v5.cuh) apart: build variants of the library with the global loads, the
global stores and / or the butterflies compiled out (-DV5_NO_LOAD / -DV5_NO_STORE / -DV5_NO_COMPUTE) and time 2^24
transforms with each (results are garbage, only the time matters).  This is where DESIGN.md 3.2's "arithmetic only
81 us, memory only 105 us, both 115 us" comes from.

  python tools/pass_decompose.py build     # here (nvcc cross-compiles): toyni_b200/build/variants/lib_*.so
  python tools/pass_decompose.py run       # on the B200: one line per variant
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "toyni_b200", "build", "variants")
VARIANTS = {"full": [], "arith_only": ["-DV5_NO_LOAD", "-DV5_NO_STORE"], "loads_only": ["-DV5_NO_COMPUTE", "-DV5_NO_STORE"],
            "stores_only": ["-DV5_NO_COMPUTE", "-DV5_NO_LOAD"], "memory_only": ["-DV5_NO_COMPUTE"]}
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "186,128"]

if sys.argv[1:] == ["build"]:
    sys.path.insert(0, ROOT)
    from toyni_b200 import build as B
    B.build()
    os.makedirs(VAR, exist_ok=True)
    objs = [os.path.join(B.OBJ, f) for f in os.listdir(B.OBJ) if f.endswith(".o") and f not in ("ntt_v5_inst.o", "ntt_engine.o")]
    for name, defs in VARIANTS.items():
        mine = []
        for src in ("ntt_v5_inst.cu", "ntt_engine.cu"):
            o = os.path.join(VAR, f"{name}_{src[:-3]}.o")
            subprocess.check_call(["nvcc"] + FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", o])
            mine.append(o)
        subprocess.check_call(["nvcc", "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o",
                               os.path.join(VAR, f"lib_{name}.so")] + mine + objs)
        print("built", name)
else:
    for name in VARIANTS:
        env = dict(os.environ, TOYNI_NTT_LIB=os.path.join(VAR, f"lib_{name}.so"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "v5_time.py"), "1"], env=env, capture_output=True, text=True).stdout
        print(name, "|", " | ".join(l.strip() for l in out.splitlines()[:2]), flush=True)
