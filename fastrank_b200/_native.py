"""Loads libfastrank_b200.so through cffi (ABI mode) -- the stand-in for the
maturin-generated `fastrank.fastrank` sub-package the reference imports as
`from .fastrank import lib, ffi` (reference fastrank/clib.py:2).

The declarations come from include/fastrank_b200.h itself, so the Python binding cannot
drift from the C ABI.  There is no fallback: if the shared library is missing this import
fails, and every compute call fails on a machine without a CUDA device.
"""
from __future__ import annotations

import os
import re

import cffi

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "fastrank_b200.h")
LIB_PATH = os.environ.get("FASTRANK_B200_LIB", os.path.join(_HERE, "libfastrank_b200.so"))


def header_cdef(path: str = _HEADER) -> str:
    """The header minus what cffi's cdef() cannot digest (include guards, extern "C")."""
    text = open(path, "r").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("#"):
            if re.match(r"#define\s+\w+\s+-?\d+\s*$", s):
                out.append(s)
            continue
        if s.startswith('extern "C"') or s == "}":
            continue
        out.append(line)
    return "\n".join(out)


def exported_symbols(path: str = _HEADER):
    """Function names the header declares (used by the CPU-side ABI test)."""
    text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(\w+)\s*\([^;{}]*\)\s*;", text)))


ffi = cffi.FFI()
ffi.cdef(header_cdef())
if not os.path.exists(LIB_PATH):
    raise ImportError(
        "fastrank_b200: %s is missing -- build it with `python fastrank_b200/build.py` "
        "(there is no CPU fallback)" % LIB_PATH
    )
lib = ffi.dlopen(LIB_PATH)
