#!/usr/bin/env python
"""Builds variants of libcustma_b200.so that differ in -D switches of ONE source file (kernel experiments):
    python tools/build_variants.py sliding_backward.cu name1:-DFOO name2:-DFOO,-DBAR ...
The variants land in custereomatching_b200/_variants/libcustma_<name>.so (git-ignored, shipped by gpurun); select one
with CUSTMA_LIB=<path> (custereomatching_b200/binding.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custereomatching_b200 import build as B  # noqa: E402

src_name, specs = sys.argv[1], sys.argv[2:]
B.build()
obj_dir = os.path.join(B.PKG_DIR, "_build")
out_dir = os.path.join(B.PKG_DIR, "_variants")
os.makedirs(out_dir, exist_ok=True)
nvcc = B.find_nvcc()
procs = []
for spec in specs:
    name, _, flags = spec.partition(":")
    obj = os.path.join(out_dir, f"{src_name}.{name}.o")
    cmd = [nvcc, *B.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-Xptxas=-v", "-I", B.INCLUDE, "-I", B.CSRC, "-c",
           os.path.join(B.CSRC, src_name), "-o", obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, pr in procs:
    out, _ = pr.communicate()
    if pr.returncode:
        print(out)
        raise SystemExit(f"variant {name} failed")
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and ("ILi5ELi3ELi5E" in l or "ILi5ELi3ELi4E" in l) and "finalize" not in l:
            print(name, "|", " ".join(x.strip() for x in lines[i + 1:i + 4] if "spill" in x or "Used" in x))
    objs = [os.path.join(obj_dir, os.path.basename(s) + ".o") for s in B.sources() if os.path.basename(s) != src_name]
    lib = os.path.join(out_dir, f"libcustma_{name}.so")
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj, *objs])
    print("built", lib)
