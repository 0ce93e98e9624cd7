"""Builds libcustma_b200.so (the C-ABI library, include/custma_b200.h) in-tree with nvcc for sm_100a only.

Used by setup.py (build_ext), `__graft_entry__.build()` and `python -m custereomatching_b200.build`.
The reference's build (setup.py:28-57) passes no arch flags and compiles torch headers into the .cu (about 3 min);
here the .cu files see only cuda_runtime.h, so a full rebuild takes seconds and the library has no torch dependency.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_NAME = "libcustma_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = ["custma_api.cu", "direct.cu", "sliding_prep.cu", "sliding_forward.cu", "sliding_backward.cu", "sliding_fallback.cu", "tc_forward.cu", "tc_backward.cu", "host_pipeline.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(INCLUDE, "custma_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = os.path.join(PKG_DIR, "_build")
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(extra_flags or []), "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            print(out, file=sys.stderr if pr.returncode else sys.stdout, flush=True)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB_PATH + ".tmp"
    # static cudart (nvcc default): the library depends on libcuda/libc only and needs no LD_LIBRARY_PATH
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", tmp, *objs])
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def torch_module_path() -> str:
    import sysconfig
    return os.path.join(ROOT, "custma", "src" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_torch_module(force: bool = False, verbose: bool = False) -> str:
    """Builds custma/src.<abi>.so - the reference's native module name (setup.py:30-38 builds `custma.src`) - from
    csrc/torch_binding.cpp with g++ against the installed torch, linked to libcustma_b200.so.  No CUDA code here."""
    import sysconfig
    out = torch_module_path()
    src = os.path.join(CSRC, "torch_binding.cpp")
    hdr = os.path.join(INCLUDE, "custma_b200.h")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return out
    build()   # libcustma_b200.so first
    import torch
    from torch.utils import cpp_extension as ce
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           "-DTORCH_EXTENSION_NAME=src", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
    cmd += [f"-I{i}" for i in ce.include_paths()] + ["-I/usr/local/cuda/include", f"-I{sysconfig.get_paths()['include']}",
                                                     f"-I{INCLUDE}"]
    cmd += [src, "-o", out + ".tmp", f"-L{PKG_DIR}", "-lcustma_b200", f"-L{torch_lib}", "-lc10", "-lc10_cuda",
            "-ltorch_cpu", "-ltorch", "-ltorch_python",
            "-Wl,-rpath,$ORIGIN/../custereomatching_b200", f"-Wl,-rpath,{torch_lib}"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout, file=sys.stderr, flush=True)
        raise RuntimeError("building custma.src failed")
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_module(force="--force" in sys.argv, verbose="-v" in sys.argv))
