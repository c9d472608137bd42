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


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile="--profile" in sys.argv)
    print(path)
