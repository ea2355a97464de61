"""Builds fastrank_b200/libfastrank_b200.so in-tree: nvcc for the sm_100a kernels, g++ for the
host, one shared object with cudart linked statically (so it loads on a machine without a
GPU and reports the missing device at call time).

    python fastrank_b200/build.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfastrank_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off"]

CU_SOURCES = ["device.cu", "sweep_fast.cu", "trees.cu", "long_queries.cu", "rf_induction.cu"]
CXX_SOURCES = ["dataset.cpp", "model.cpp", "evaluator.cpp", "coordinate_ascent.cpp",
               "random_forest.cpp", "training.cpp", "capi.cpp", "io_helper.cpp"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".h", ".cuh"))]
    hs.append(os.path.join(INCLUDE, "fastrank_b200.h"))
    return hs


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    path = os.path.join(CSRC, src)
    if not (force or _newer(obj, [path] + _headers())):
        return obj
    if src.endswith(".cu"):
        cmd = [NVCC] + CU_FLAGS + ["-c", path, "-o", obj]
    else:
        cmd = [CXX] + CXX_FLAGS + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("compile failed: " + " ".join(cmd))
    if src.endswith(".cu"):
        with open(os.path.join(OBJ, os.path.basename(src) + ".ptxas.txt"), "w") as fp:
            fp.write(res.stderr)
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(lambda s: _compile(s, force), CU_SOURCES + CXX_SOURCES))
    if force or _newer(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
