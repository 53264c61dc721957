"""Numeric constants of the scene/camera/BVH host layer.

Values follow the reference (src/constants.py:1-37, src/trace.metal:4-5, src/renderer.py:8,299) because they
define the images: colours are in cv2 (BGR) channel order, the BVH limits shape the flattened tree, and the
kernel constants are baked into both the oracle and the CUDA kernels.
"""
import time

import numpy as np

# --- camera (constants.py:5) -------------------------------------------------------------------------------
H_FOV = np.deg2rad(110.0)

# --- axes --------------------------------------------------------------------------------------------------
UNIT_X, UNIT_Y, UNIT_Z = (np.eye(3, dtype=np.float64)[i] for i in range(3))
ZERO_VECTOR = np.zeros(3, dtype=np.float64)

# --- colours, BGR, [0,1] (constants.py:16-23) ----------------------------------------------------------------
def _c(b, g, r):
    return np.array([b, g, r], dtype=np.float64)


BLACK = _c(0.0, 0.0, 0.0)
WHITE = _c(0.7, 0.7, 0.7)
FULL_WHITE = _c(1.0, 1.0, 1.0)
GRAY = _c(0.5, 0.5, 0.5)
RED = _c(0.3, 0.3, 0.8)
GREEN = _c(0.541, 0.807, 0.0)
BLUE = _c(0.8, 0.3, 0.3)
CYAN = _c(0.8, 0.8, 0.3)

# --- BVH limits (constants.py:28-29) -------------------------------------------------------------------------
MAX_MEMBERS = 8   # a node with <= MAX_MEMBERS triangles becomes a leaf
MAX_DEPTH = 32    # compared against the build stack size, not the tree depth (bvh.py:294)

# --- Cornell enclosure (constants.py:33-36) --------------------------------------------------------------------
DEFAULT_BOX_MIN_CORNER = np.array([-10, -2, -10])
DEFAULT_BOX_MAX_CORNER = np.array([10, 10, 10])
DEFAULT_LIGHT_HEIGHT = 0.95
DEFAULT_LIGHT_SCALE = 0.25

# --- kernel constants (trace.metal:4-5,407; renderer.py:8,299) --------------------------------------------------
KERNEL_PI = np.float32(3.14159265359)
KERNEL_DELTA = np.float32(0.0001)
MAX_PATH_LENGTH = 8
TONE_MAP_EXPOSURE = 4.0


def timed(func):
    """Wall-clock print decorator with the reference's message format (constants.py:39-49)."""

    def wrapper(*args, **kwargs):
        t0 = time.time()
        out = func(*args, **kwargs)
        print(f"Function {func.__name__} took {time.time() - t0:.4f} seconds")
        return out

    wrapper.__name__ = getattr(func, "__name__", "wrapper")
    wrapper.__doc__ = func.__doc__
    return wrapper
