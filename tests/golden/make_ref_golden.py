"""Golden vectors of the KERNELS, produced by the reference itself: /root/reference/src/trace.metal compiled unmodified
as C++ (oracle/_ref, `make -C oracle ref`) and driven through the reference's stage order (oracle/ref_glue.c =
renderer.py:113-291).  Run in the build container, where /root/reference exists:

    python tests/golden/make_ref_golden.py

Writes tests/golden/ref_kernels.npz: for each case the SHA-256 of every stage buffer after two passes (RNG state, rays,
Path records, light entries, aggregators, finalized samples) and the raw float accumulators of the final image.
tests/test_ref_pin.py::test_restatement_matches_reference_golden checks oracle/trace_ref.c against it everywhere (also
where /root/reference is absent).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

CASES = {  # name: (preset, width, height, passes)
    "c1_empty_48x48": ("empty", 48, 48, 2),
    "c2_spheres_ggx_64x48": ("bench-spheres", 64, 48, 2),
}
HASHED = ("rand", "light_rays", "camera_rays", "camera_paths", "light_paths", "out_camera_image", "light_pixel_indices",
          "light_path_indices", "light_ray_indices", "light_weights", "light_shade", "aggregators", "out_samples",
          "finalized_samples", "sample_counts", "sample_weights", "out_light_image")
RAW = ("summed_image", "summed_sample_weights", "summed_sample_counts", "unidirectional")


def run_case(preset, width, height, passes, **oracle_kw):
    from oracle.oracle import OracleRenderer
    from tests import common
    arrays = common.arrays(preset, width, height)
    r = OracleRenderer(arrays, common.seeds_for(width * height), **oracle_kw)
    for _ in range(passes):
        r.run_sample()
    out = {}
    for name in HASHED:
        out["sha256_" + name] = np.frombuffer(hashlib.sha256(r.buffer(name).tobytes()).digest(), np.uint8).copy()
    for name in RAW:
        out[name] = r.buffer(name, np.uint32).copy()  # bit patterns
    r.close()
    return out


def main():
    from oracle import oracle as oracle_mod
    assert oracle_mod.build_ref(force=True), "needs /root/reference"
    blob = {}
    for case, (preset, w, h, passes) in CASES.items():
        for sm in (False, True):
            res = run_case(preset, w, h, passes, shared_math=sm, reference=True)
            for k, v in res.items():
                blob[f"{case}/{'shared_math' if sm else 'glibc'}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **blob)
    print("wrote ref_kernels.npz:", len(blob), "arrays")


if __name__ == "__main__":
    main()
