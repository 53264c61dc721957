"""ctypes wrapper of the native BVH builder (clive2_b200/csrc/clv_bvh_build.cpp -> _lib/libclive_bvh.so).

Host-only code (g++, OpenMP): builds in any container.  Emits exactly the arrays `bvh.np_flatten_bvh(construct_BVH(soup))`
emits with the stable sort (tests/test_native_bvh.py compares them byte for byte).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .constants import MAX_DEPTH, MAX_MEMBERS
from .struct_types import Box

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "clv_bvh_build.cpp")
LIB = os.path.join(HERE, "_lib", "libclive_bvh.so")
_lib = None


def build_library(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O3", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                    "-fvisibility=hidden", "-shared", "-o", LIB, SRC], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_library())
        lib.clv_bvh_build.restype = C.c_int64
        lib.clv_bvh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int64, C.c_void_p]
        lib.clv_emit_triangles.restype = None
        lib.clv_emit_triangles.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 7
        _lib = lib
    return _lib


def build_order(mins, maxes, max_members=MAX_MEMBERS, max_depth=MAX_DEPTH):
    """(Box[], leaf-ordered triangle permutation) for per-triangle boxes `mins`, `maxes` ((n,3) float32 or float64)."""
    lib = _load()
    if mins.dtype != maxes.dtype or mins.dtype not in (np.float32, np.float64):
        raise TypeError("mins/maxes must both be float32 or both float64")
    mins, maxes = np.ascontiguousarray(mins), np.ascontiguousarray(maxes)
    n = len(mins)
    cap = max(2 * n, 1)
    boxes = np.zeros(cap, dtype=Box)
    order = np.empty(n, dtype=np.int64)
    nb = lib.clv_bvh_build(mins.ctypes.data_as(C.c_void_p), maxes.ctypes.data_as(C.c_void_p),
                           int(mins.dtype == np.float64), n, int(max_members), int(max_depth),
                           boxes.ctypes.data_as(C.c_void_p), cap, order.ctypes.data_as(C.c_void_p))
    if nb < 0:
        raise RuntimeError("box capacity too small")
    return boxes[:nb].copy(), order


def emit_triangles(soup, order):
    """`bvh.emit_triangles` as one native gather (numpy's fancy indexing of the ten soup arrays is as slow as the build
    itself at a million triangles); falls back to the numpy version for soups whose arrays it cannot take as they are."""
    from .bvh import emit_triangles as numpy_emit
    from .struct_types import Triangle
    f = soup.triangles.dtype
    same = all(a.dtype == f for a in (soup.smoothed_normals, soup.face_normals))
    if f not in (np.float32, np.float64) or not same or soup.material.dtype != np.int32:
        return numpy_emit(soup, order)
    lib = _load()
    order = np.ascontiguousarray(order, dtype=np.int64)
    arrays = [np.ascontiguousarray(a) for a in (soup.triangles, soup.smoothed_normals, soup.face_normals, soup.material,
                                                soup.emitter.astype(np.uint8), soup.camera.astype(np.uint8))]
    out = np.empty(len(order), dtype=Triangle)
    lib.clv_emit_triangles(order.ctypes.data_as(C.c_void_p), len(order), int(f == np.float64),
                           *[a.ctypes.data_as(C.c_void_p) for a in arrays], out.ctypes.data_as(C.c_void_p))
    return out


def build(soup):
    """Soup (bvh.FastTreeBox) -> (Box[], Triangle[]) like bvh.np_flatten_bvh(bvh.construct_BVH(soup))."""
    boxes, order = build_order(soup.mins, soup.maxes)
    return boxes, emit_triangles(soup, order)
