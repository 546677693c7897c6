#!/usr/bin/env python
"""Builds lib/libstencils_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "lib", "libstencils_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG, "..", "include", "stencils_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    sp = os.path.join(HERE, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), headers_mtime()):
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", sp, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: compile_one(s, verbose), srcs))
    objs = [o for o, _ in res]
    stale = [f for f in os.listdir(OBJ) if os.path.join(OBJ, f) not in objs]
    for f in stale:
        os.remove(os.path.join(OBJ, f))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs) or stale:
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
               "-Xcompiler", "-fPIC", "-lcuda" if False else "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
