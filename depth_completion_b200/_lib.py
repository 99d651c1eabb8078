"""ctypes loader for libmdc_b200.so, the C-ABI boundary of the hot path (include/mdc.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU or
PyTorch fallback: if the shared object is missing or fails to load, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmdc_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile csrc/mdc.cu into lib/libmdc_b200.so for sm_100a (cross-compiles without a GPU)."""
    src = os.path.join(CSRC, "mdc.cu")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, src]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


class MdcError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MdcError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mdc_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MdcError(lib().mdc_last_error().decode("utf-8", "replace"))


def ptr(t) -> ctypes.c_void_p:
    """Device (or host) pointer of a torch tensor, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())
