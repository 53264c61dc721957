"""Native BVH builder against the numpy specification and against the REFERENCE's own builder
output on tie-free input (tests/golden/bvh_random_soup.npz)."""
import os

import numpy as np
import pytest

from clive2_b200 import bvh, meshgen, native_bvh, scene
from clive2_b200.load import fast_load
from tests.common import GOLDEN


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_matches_reference_builder_on_tie_free_soup():
    g = np.load(os.path.join(GOLDEN, "bvh_random_soup.npz"))
    soup = fast_load(g["mesh_vertices"], g["mesh_faces"], material=2)
    boxes, tris = native_bvh.build(soup)
    assert _same(boxes, g["boxes"]) and _same(tris, g["triangles"])


@pytest.mark.parametrize("quads,dtype", [(12, np.float64), (40, np.float64), (40, np.float32)])
def test_matches_numpy_specification(quads, dtype):
    v, f = meshgen.terrain(quads, quads, jitter=0.0)  # regular grid: full of centroid ties -> exercises the stable order
    soup = fast_load(v.astype(dtype), f, material=3)
    b0, t0 = bvh.np_flatten_bvh(bvh.construct_BVH(soup))
    b1, t1 = native_bvh.build(soup)
    assert _same(b0, b1) and _same(t0, t1)


@pytest.mark.parametrize("preset", ["empty", "bench-spheres"])
def test_presets_identical_with_either_builder(preset):
    a0 = scene.preset_arrays(preset, 32, 32, bvh_builder="numpy")
    a1 = scene.preset_arrays(preset, 32, 32, bvh_builder="native")
    assert _same(a0.boxes, a1.boxes) and _same(a0.triangles, a1.triangles)
    assert np.array_equal(a0.light_triangle_indices, a1.light_triangle_indices)


def test_tree_invariants_on_a_larger_mesh():
    v, f = meshgen.terrain(150, 150)
    soup = fast_load(v, f, material=3)
    boxes, order = native_bvh.build_order(soup.mins, soup.maxes)
    assert np.array_equal(np.sort(order), np.arange(len(f)))
    inner = boxes["right"] == 0
    leaves = boxes[~inner]
    assert (leaves["right"] - leaves["left"]).max() <= 8 and (leaves["right"] - leaves["left"]).min() >= 1
    assert np.array_equal(np.sort(leaves["left"])[1:], np.sort(leaves["right"])[:-1])  # ranges tile [0, n)
    kids = boxes["left"][inner]
    assert (kids > np.flatnonzero(inner)).all() and (kids + 1 < len(boxes)).all()
    # every child box lies inside its parent's
    for k in (0, 1):
        c = boxes[kids + k]
        assert (c["min"][:, :3] >= boxes["min"][inner][:, :3]).all() and (c["max"][:, :3] <= boxes["max"][inner][:, :3]).all()


def test_single_triangle_and_small_inputs():
    for n in (1, 5, 9):
        rng = np.random.default_rng(n)
        lo = rng.uniform(-1, 1, (n, 3))
        hi = lo + rng.uniform(0.01, 0.5, (n, 3))
        boxes, order = native_bvh.build_order(lo, hi)
        assert sorted(order.tolist()) == list(range(n))
        assert len(boxes) == (1 if n <= 8 else 3)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_awkward_soups_identical_with_either_builder(seed):
    """Exact duplicates (tied centroids and tied SAH costs), coplanar overlaps, zero-area triangles, flat boxes, one far
    outlier: the native builder still emits the numpy specification's bytes."""
    from tests.common import fuzz_triangle_soup
    soup = fuzz_triangle_soup(seed)
    b1, t1 = bvh.build_flat_bvh(soup, builder="numpy")
    b2, t2 = native_bvh.build(soup)
    assert _same(b1, b2) and _same(t1, t2)


def test_thread_count_does_not_change_the_tree():
    """The builder splits subtrees across OpenMP tasks; the emitted arrays must not depend on how many threads ran them."""
    import subprocess
    import sys
    code = ("import hashlib, sys; sys.path.insert(0, %r)\n"
            "from clive2_b200 import meshgen, native_bvh\n"
            "from clive2_b200.load import fast_load\n"
            "v, f = meshgen.terrain(120, 120)\n"
            "b, t = native_bvh.build(fast_load(v, f, material=3))\n"
            "print(hashlib.sha1(b.tobytes() + t.tobytes()).hexdigest())\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = set()
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-500:]
        digests.add(out.stdout.strip())
    assert len(digests) == 1


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("n,levels", [(300, 2), (1000, 3), (2000, 1)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_massive_centroid_ties(seed, n, levels, dtype):
    """Coordinates quantised to a handful of values: almost every centroid ties on every axis and many SAH costs tie too.
    The native builder sorts once at the root and then partitions, repairing the order inside runs of equal keys; the
    numpy specification re-sorts every node.  Same bytes."""
    rng = np.random.default_rng(seed)
    corners = (rng.integers(0, levels, (n, 1, 3)) + rng.integers(0, 2, (n, 3, 3)) * 0.5).astype(dtype)
    zeros = np.zeros((n, 3), dtype)
    soup = bvh.FastTreeBox(np.zeros((n, 3), np.uint32), corners, corners.min(axis=1), corners.max(axis=1), zeros,
                           np.repeat(zeros[:, None, :], 3, axis=1), np.zeros(n, dtype), np.zeros(n, np.int32),
                           np.zeros(n, np.bool_), np.zeros(n, np.bool_))
    b1, t1 = bvh.build_flat_bvh(soup, builder="numpy")
    b2, t2 = native_bvh.build(soup)
    assert _same(b1, b2) and _same(t1, t2)


def test_bench_trees_stay_far_below_the_reference_stack():
    """trace.metal:144-176 stops traversing when its 64-entry stack fills (one entry per level of the path being
    descended); the kernels do not reproduce that abort, so the trees in use must be much shallower than 62 levels."""
    def depth(boxes):
        d = np.zeros(len(boxes), np.int32)
        for b in range(len(boxes)):  # breadth-first layout: parents come before their children
            if boxes["right"][b] == 0:
                d[boxes["left"][b]] = d[boxes["left"][b] + 1] = d[b] + 1
        return int(d.max())
    assert depth(scene.preset_arrays("empty", 32, 32).boxes) <= 4
    assert depth(scene.preset_arrays("bench-spheres", 32, 32).boxes) <= 24
    v, f = meshgen.terrain(200, 200)
    assert depth(native_bvh.build(fast_load(v, f, material=3))[0]) <= 30
