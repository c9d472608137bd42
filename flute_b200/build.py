"""Builds flute_b200/libflute_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree.

    python flute_b200/build.py [--force] [--verbose]

No torch headers are involved: the library is plain CUDA C++ behind include/flute_b200.h.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libflute_b200.so")
SOURCES = ["qgemm_sm100.cu", "qgemm_decode_sm100.cu", "qgemm_prefill_sm100.cu", "aux_kernels.cu", "api.cu"]
HEADERS = ["ptx.cuh", "qgemm_sm100.h", "aux_kernels.h", os.path.join("..", "..", "include", "flute_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    if profile:
        extra = ["-DFB_PROFILE=1"] + os.environ.get("FB_EXTRA_DEFINES", "").split()
        return _build(os.path.join(HERE, "libflute_b200_prof.so"), verbose, extra, "_obj_prof")
    if not force and not _stale():
        return LIB
    return _build(LIB, verbose, [], "_obj")


def _build(lib_path: str, verbose: bool, extra: list, objname: str) -> str:
    objdir = os.path.join(HERE, "csrc", objname)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-Xlinker", "--exclude-libs,ALL", "-o", lib_path, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    return lib_path


TORCH_LIB = os.path.join(HERE, "_flute_b200_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch_binding.cpp")


def build_torch_binding(force: bool = False, verbose: bool = False) -> str:
    """The compiled torch-operator binding (TORCH_LIBRARY(flute) over the C ABI): one host-only C++ file, g++."""
    build(False, verbose)
    deps = [TORCH_SRC, os.path.join(HERE, "..", "include", "flute_b200.h")]   # (only links the C ABI by name)
    if not force and os.path.exists(TORCH_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_LIB) for d in deps):
        return TORCH_LIB
    import torch
    from torch.utils import cpp_extension
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for inc in cpp_extension.include_paths("cuda"):
        cmd += ["-isystem", inc]
    cmd += [TORCH_SRC, "-o", TORCH_LIB, f"-L{libdir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
            f"-L{HERE}", "-lflute_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"torch binding build failed:\n{res.stdout}")
    return TORCH_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile="--profile" in sys.argv)
    print(path)
    if "--torch" in sys.argv:
        print(build_torch_binding(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
