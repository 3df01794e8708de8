"""ctypes loader for libcniic_b200.so (the C ABI declared in include/cniic_b200.h).

The library is built in-tree by ``cniic_b200/csrc/Makefile`` (nvcc, sm_100a).  There is no CPU fallback: if the
shared object is missing the import fails loudly, and every entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcniic_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "cniic_b200.h")

OK, ERR_BAD_ARG, ERR_TOO_FEW_POINTS, ERR_TOO_FEW_ACTIVE, ERR_CUDA, ERR_NCCL, ERR_DECODE, ERR_BUFFER_TOO_SMALL, \
    ERR_UNSUPPORTED = range(9)
TIE_KEEP_CURRENT, TIE_LOWEST_INDEX = 0, 1
POINTS_RGB, POINTS_XYRGB = 0, 1
KMEANS_NO_CULL = 1
KMEANS_FORCE_CULL = 2
MAX_K, MAX_DIM = 4096, 16384


class KMeansStats(C.Structure):
    _fields_ = [("iterations", C.c_uint32), ("empty_events", C.c_uint32), ("moved_last", C.c_uint64),
                ("moved_total", C.c_uint64), ("converged", C.c_uint32), ("gpu_launches", C.c_uint32),
                ("device_ms", C.c_float), ("assign_ms_avg", C.c_float), ("pairs_scored", C.c_uint64)]


class KMeansDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("k", C.c_uint32), ("tie_rule", C.c_int), ("n_local", C.c_uint64),
                ("n_total", C.c_uint64), ("first_index", C.c_uint64), ("w", C.c_uint32), ("h_local", C.c_uint32),
                ("y0", C.c_uint32), ("rgb", C.c_void_p), ("weights", C.c_void_p), ("points_on_device", C.c_int), ("flags", C.c_int)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (sm_100a).  Used by __graft_entry__.build()."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j8"] + (["-B"] if force else [])
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `make -C cniic_b200/csrc` "
                              "(__graft_entry__.build()); there is no CPU fallback")
        _lib = _declare(C.CDLL(SO_PATH))
    return _lib


def _declare(L: C.CDLL) -> C.CDLL:
    """ctypes signatures of the entry points whose defaults (int in / int out) are wrong."""
    L.cniic_last_error.restype = C.c_char_p
    L.cniic_ctx_stream.restype = C.c_void_p
    L.cniic_device_alloc.restype = C.c_void_p
    L.cniic_device_alloc.argtypes = [C.c_void_p, C.c_size_t]
    L.cniic_device_free.argtypes = [C.c_void_p, C.c_void_p]
    L.cniic_device_free.restype = None
    L.cniic_ctx_destroy.argtypes = [C.c_void_p]
    L.cniic_ctx_destroy.restype = None
    L.cniic_kmeans_close.argtypes = [C.c_void_p]
    L.cniic_kmeans_close.restype = None
    L.cniic_kmeans_device_assign.restype = C.c_void_p
    L.cniic_kmeans_device_assign.argtypes = [C.c_void_p]
    L.cniic_ctx_launches.restype = C.c_uint32
    L.cniic_ctx_launches.argtypes = [C.c_void_p]
    L.cniic_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.cniic_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    return L


def declared_symbols() -> list[str]:
    """Every function name declared in include/cniic_b200.h (used by the symbol-export test)."""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cniic_[a-z0-9_]+)\s*\(", text)))
