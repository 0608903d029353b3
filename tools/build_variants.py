#!/usr/bin/env python3
"""Build alternative libcodexcommit.so variants (extra -D flags) into build/variants/ for same-box A/B sweeps with
tools/sweep_lib.py.  build/ is git-ignored but travels with the gpurun snapshot.
usage: build_variants.py name1:-DFOO=1,-DBAR=2 name2: ..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "codex-storage-proofs-circuits_b200", "csrc", "capi.cu")
OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)

def build(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join(OUT, f"lib_{name}.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v", "-o", out, SRC] + [f for f in flags.split(",") if f]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        return name, "FAILED\n" + r.stderr[-2000:]
    lines = r.stderr.splitlines()
    info = ""
    for i, l in enumerate(lines):
        if "k_hash_cells_tma" in l and "Compiling" in l:
            info = " | ".join(x.strip() for x in lines[i + 1:i + 4])
    return name, info

with ThreadPoolExecutor(4) as ex:
    for name, info in ex.map(build, sys.argv[1:]):
        print(name, "::", info)
