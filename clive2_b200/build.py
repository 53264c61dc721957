"""Builds libclive_b200.so (the sm_100a CUDA library behind include/clive_b200.h) in-tree with nvcc.

The library is the product's only execution path; there is no CPU build of it.  nvcc cross-compiles for sm_100a
without a GPU, so this runs in the build container; the resulting .so travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libclive_b200.so")
SOURCES = ["clive_b200.cu"]
HEADERS = ["clv_device.cuh", "clv_math.h", "clv_accel_build.h", os.path.join("..", "..", "include", "clive_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # arithmetic contract: no FMA contraction (bit-exact against the oracle)
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libclive_b200.so cannot be built")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        from . import native_bvh
        native_bvh.build_library()
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("CLV_NVCC_EXTRA", "").split()  # e.g. -DCLV_WALK_MIN_BLOCKS=8 for tuning experiments
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    from . import native_bvh
    native_bvh.build_library(force=force)  # host-only BVH builder (g++), next to the CUDA library
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
