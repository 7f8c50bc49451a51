"""Builds csrc/*.cu with nvcc for sm_100a, in-tree (so the .so files travel to the GPU box with the repo snapshot):

  simple_image_compression_network_b200/libfinnconv_b200.so   the product library (no experiment switches, no v1 kernel)
  tools/libfinnconv_exp.so                                    the same sources with -DFCB_EXPERIMENT -DFCB_U2_PROF: environment
                                                              switches that bend plans, the first-generation kernel, cross-check
                                                              instantiations and in-kernel clock accounting (tools/, tests/test_experimental.py)

`python -m simple_image_compression_network_b200.build [--force] [--exp] [-v]`.  Translation units compile in parallel.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libfinnconv_b200.so")
EXP_LIB = os.path.join(PKG, "..", "tools", "libfinnconv_exp.so")
OBJ_DIR = os.path.join(PKG, "csrc", "_obj")
CC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
            "--expt-relaxed-constexpr"]
EXP_DEFS = ["-DFCB_EXPERIMENT", "-DFCB_U2_PROF"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(PKG, "..", "include", "finnconv_b200.h")]


def stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in _deps())


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _compile(args):
    src, obj, defs, verbose = args
    cmd = [_nvcc()] + CC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    subprocess.check_call(cmd)
    return obj


def _build_one(lib: str, tag: str, defs, verbose: bool) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs = [(s, os.path.join(OBJ_DIR, f"{os.path.basename(s)[:-3]}.{tag}.o"), defs, verbose) for s in sources()]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, jobs))
    subprocess.check_call([_nvcc(), "-shared", "-o", lib] + objs + ["-lcudart"])
    return lib


def build(force: bool = False, verbose: bool = False, exp: bool = False) -> str:
    extra = os.environ.get("FCB_NVCC_EXTRA", "").split()
    if force or stale(LIB):
        _build_one(LIB, "prod", extra, verbose)
    if exp and (force or stale(EXP_LIB)):
        _build_one(EXP_LIB, "exp", EXP_DEFS + extra, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, exp="--exp" in sys.argv))
