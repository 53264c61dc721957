"""Pins the oracle (oracle/trace_ref.c, a restatement) to the REFERENCE ITSELF.

`oracle/_ref` is /root/reference/src/trace.metal compiled unmodified, where it lies, as C++ behind a small
`metal_stdlib` shim (oracle/ref_build/, `make -C oracle ref`), with the same host glue (oracle/ref_glue.c =
renderer.py:113-291) around its eight kernels.  Two kinds of pin:

* live: restatement == compiled reference, byte for byte, on every buffer after every stage of every pass, for the
  Cornell scene (BASELINE C1), the GGX / transmission scene (C2's materials) and a mesh scene, in both arithmetic
  flavours (glibc libm; the shared clv_math.h the CUDA kernels use).  Runs wherever oracle/_ref is built or can be
  built (the build container has /root/reference; the GPU box gets the prebuilt .so with the snapshot);
* golden: restatement vs tests/golden/ref_kernels.npz, produced by the compiled reference with
  tests/golden/make_ref_golden.py — runs everywhere, also without the reference tree.

CPU only (no `gpu` marker): this is the oracle being checked, not the product.
"""
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as oracle_mod
from oracle.oracle import Oracle, OracleRenderer
from tests import common
from tests.golden import make_ref_golden  # the generator's case table and buffer lists are the test's too

HAVE_REF = bool(oracle_mod.build_ref())
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built and /root/reference absent")

STAGE_BUFFERS = make_ref_golden.HASHED + make_ref_golden.RAW


@needs_ref
def test_reference_source_compiles_unmodified_and_exports_the_kernels():
    import ctypes as C
    for path in oracle_mod.ref_libs():
        lib = C.CDLL(path)
        assert lib.ref_is_compiled_reference() == 1
        # the eight Metal kernels themselves (extern "C" through the shim's `kernel` macro) and their id-loop wrappers
        for k in ("generate_paths", "connect_paths", "light_sort", "light_image_gather", "reset_light_indices",
                  "adaptive_finalize_samples", "generate_camera_rays", "generate_light_rays"):
            assert hasattr(lib, k) and hasattr(lib, "ref_" + k), k
    s, f = Oracle(reference=True).xorshift32(1)
    assert s == 270369 and f == np.float32(270369) / np.float32(4294967296.0)  # SURVEY Appendix B KAT


@needs_ref
@pytest.mark.parametrize("shared_math", [False, True], ids=["glibc", "shared_math"])
@pytest.mark.parametrize("preset,width,height,passes", [
    ("empty", 64, 64, 3),            # BASELINE C1's scene (diffuse Cornell box)
    ("bench-spheres", 96, 72, 3),    # BASELINE C2's scene: GGX reflection, rough-dielectric transmission, coated diffuse
])
def test_restatement_equals_compiled_reference_after_every_stage(preset, width, height, passes, shared_math):
    arrays = common.arrays(preset, width, height)
    seeds = common.seeds_for(width * height)
    port = OracleRenderer(arrays, seeds, shared_math=shared_math)
    ref = OracleRenderer(arrays, seeds, shared_math=shared_math, reference=True)
    assert ref.oracle.reference and not port.oracle.reference
    for p in range(passes):
        for stage in range(len(oracle_mod.STAGES)):
            port.stage(stage)
            ref.stage(stage)
            for name in STAGE_BUFFERS:
                a, b = port.buffer(name), ref.buffer(name)
                assert np.array_equal(a, b), (f"pass {p}, after stage {oracle_mod.STAGES[stage]}: buffer {name} differs in "
                                              f"{int((a != b).sum())} of {a.size} bytes")
    # the scene exercises what it is there for
    mats = arrays.materials["type"][port.buffer("camera_paths", port.T.Path)["rays"]["material"]]
    if preset == "bench-spheres":
        assert (mats == 1).any() and (mats == 2).any()
    port.close()
    ref.close()


@needs_ref
def test_restatement_equals_compiled_reference_on_a_mesh_scene():
    """1 k-triangle terrain in the enclosure (the C3/C4 generator at low resolution): deeper tree, 15-triangle leaves."""
    from clive2_b200 import meshgen, scene
    arrays = scene.build_scene_arrays(48, 48, np.array([0, 1.5, 6]), np.array([0, 0, -1]),
                                      [{"mesh": meshgen.terrain(24, 24), "material": 3}], bvh_builder="numpy")
    seeds = common.seeds_for(48 * 48)
    port, ref = OracleRenderer(arrays, seeds), OracleRenderer(arrays, seeds, reference=True)
    for _ in range(2):
        port.run_sample()
        ref.run_sample()
    for name in STAGE_BUFFERS:
        assert np.array_equal(port.buffer(name), ref.buffer(name)), name
    port.close()
    ref.close()


@needs_ref
def test_c1_at_baseline_size_restatement_equals_compiled_reference():
    """BASELINE.json configs[0] at its own frame size (256x256), 2 of its 16 passes: accumulators bit-equal."""
    arrays = common.arrays("empty", 256, 256)
    seeds = common.seeds_for(256 * 256)
    port, ref = OracleRenderer(arrays, seeds), OracleRenderer(arrays, seeds, reference=True)
    for r in (port, ref):
        r.oracle.use_all_cores()
        for _ in range(2):
            r.run_sample()
    for name in make_ref_golden.RAW + ("rand", "aggregators", "light_weights", "light_shade"):
        assert np.array_equal(port.buffer(name), ref.buffer(name)), name
    port.close()
    ref.close()


@needs_ref
def test_traversal_and_visibility_hooks_equal_the_reference_helpers():
    """traverse_bvh / visibility_test (trace.metal:144-197) on caller-supplied rays: tri index and t, u, v bits."""
    arrays = common.arrays("bench-spheres", 32, 32)
    lo, hi = arrays.boxes["min"][0][:3], arrays.boxes["max"][0][:3]
    rays = common.random_rays(20000, lo, hi, seed=3)
    P, R = Oracle(), Oracle(reference=True)
    for x, y, what in zip(P.traverse(rays, arrays.boxes, arrays.triangles), R.traverse(rays, arrays.boxes, arrays.triangles),
                          ("tri", "t", "u", "v")):
        hit = P.traverse(rays, arrays.boxes, arrays.triangles)[0] >= 0
        assert np.array_equal(x[hit].view(np.uint32), y[hit].view(np.uint32)), what
    assert (hit.mean() > 0.9)
    a, b = rays[:10000].copy(), rays[10000:].copy()
    tri = P.traverse(b, arrays.boxes, arrays.triangles)[0]
    b["triangle"] = tri  # half of the targets are what a ray from b's origin sees: a mix of visible and occluded pairs
    va, vb = P.visibility(a, b, arrays.boxes, arrays.triangles), R.visibility(a, b, arrays.boxes, arrays.triangles)
    assert np.array_equal(va, vb)


@pytest.mark.parametrize("flavour", ["shared_math", "glibc"])
@pytest.mark.parametrize("case", sorted(make_ref_golden.CASES))
def test_restatement_matches_reference_golden(case, flavour):
    """oracle/trace_ref.c against vectors the compiled reference produced (tests/golden/make_ref_golden.py).  The
    shared_math flavour is portable; the glibc flavour holds for the libm of this image (the GPU box runs the same)."""
    gold = np.load(os.path.join(common.GOLDEN, "ref_kernels.npz"))
    preset, w, h, passes = make_ref_golden.CASES[case]
    got = make_ref_golden.run_case(preset, w, h, passes, shared_math=(flavour == "shared_math"))
    for key, value in got.items():
        want = gold[f"{case}/{flavour}/{key}"]
        assert np.array_equal(value, want), f"{case}/{flavour}: {key} differs from the compiled reference's output"


def test_golden_file_is_what_the_generator_writes():
    gold = np.load(os.path.join(common.GOLDEN, "ref_kernels.npz"))
    n_keys = len(make_ref_golden.HASHED) + len(make_ref_golden.RAW)
    assert len(gold.files) == 2 * len(make_ref_golden.CASES) * n_keys
    assert hashlib.sha256(b"").digest_size == gold[f"{sorted(make_ref_golden.CASES)[0]}/glibc/sha256_rand"].size
