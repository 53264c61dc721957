"""Shared helpers of the test suite: scenes, seeds, oracle drivers, bit-level comparisons."""
import functools
import os

import numpy as np

from clive2_b200 import scene, struct_types as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def arrays(preset, width, height):
    return scene.preset_arrays(preset, width, height, bvh_builder="numpy")


def seeds_for(batch, seed=20240928):
    s = np.random.RandomState(seed).randint(0, 2 ** 32, size=(batch, 2), dtype=np.uint32)
    s[s == 0] = 1
    return s


def n_light_entries(batch):
    return 1 << (8 * batch - 1).bit_length()


class StageBuffers:
    """The renderer-owned buffers of renderer.py:47-80 as host arrays."""

    def __init__(self, batch, seeds):
        n = n_light_entries(batch)
        self.batch, self.n = batch, n
        self.rand = np.array(seeds, dtype=np.uint32, copy=True).reshape(batch, 2)
        self.indices = np.arange(batch, dtype=np.uint32)
        self.bins_identity = np.arange(batch + 1, dtype=np.uint32)
        self.light_rays = np.zeros(batch, T.Ray)
        self.camera_rays = np.zeros(batch, T.Ray)
        self.light_paths = np.zeros(batch, T.Path)
        self.camera_paths = np.zeros(batch, T.Path)
        self.out_light = np.zeros((batch, 4), np.float32)
        self.out_camera = np.zeros((batch, 4), np.float32)
        self.debug = np.zeros((batch, 4), np.float32)
        self.pix = np.full(n, -1, np.int32)
        self.pth = np.zeros(n, np.int32)
        self.ry = np.zeros(n, np.int32)
        self.w = np.zeros(n, np.float32)
        self.sh = np.zeros(n, np.float32)
        self.aggs = np.zeros(batch, T.WeightAggregator)
        self.out_samples = np.zeros((batch, 4), np.float32)
        self.finalized = np.zeros((batch, 4), np.float32)
        self.sample_counts = np.zeros(batch, np.uint32)
        self.sample_weights = np.zeros(batch, np.float32)


def oracle_stages(O, a, B, upto="connect"):
    """Run the reference stage sequence on host buffers with the oracle (renderer.py:281-286)."""
    b = B.batch
    O.generate_light_rays(b, a.light_triangles, a.light_surface_areas, a.light_triangle_indices, a.materials, B.rand,
                          B.light_rays, a.light_counts)
    O.generate_camera_rays(b, a.camera, B.rand, B.indices, B.camera_rays)
    if upto == "rays":
        return B
    O.generate_paths(b, B.light_rays, a.boxes, a.triangles, a.materials, B.rand, B.out_light, B.light_paths, B.debug)
    O.generate_paths(b, B.camera_rays, a.boxes, a.triangles, a.materials, B.rand, B.out_camera, B.camera_paths, B.debug)
    if upto == "paths":
        return B
    O.reset_light_indices(B.n, B.pix, B.pth, B.ry, B.w, B.sh)
    O.connect_paths(b, B.camera_paths, B.light_paths, a.triangles, a.materials, a.boxes, a.camera, B.aggs,
                    B.out_samples, B.pix, B.pth, B.ry, B.w, B.sh)
    if upto == "connect":
        return B
    O.adaptive_finalize_samples(b, B.aggs, a.camera, B.finalized, B.sample_counts, B.bins_identity, B.sample_weights)
    return B


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def assert_bits_equal(x, y, what):
    bx, by = bits(x), bits(y)
    assert bx.shape == by.shape, (what, bx.shape, by.shape)
    if not np.array_equal(bx, by):
        rows = np.flatnonzero((bx.reshape(len(x), -1) != by.reshape(len(y), -1)).any(axis=1))
        raise AssertionError(f"{what}: {len(rows)} of {len(x)} records differ, first at {rows[:5]}")


def assert_paths_equal(p, q, what):
    """Path records agree on length/from_camera and on every stored vertex (slots >= length are unspecified)."""
    assert np.array_equal(p["length"], q["length"]), f"{what}: lengths differ at {np.flatnonzero(p['length'] != q['length'])[:5]}"
    assert np.array_equal(p["from_camera"], q["from_camera"]), f"{what}: from_camera"
    for k in range(8):
        m = p["length"] > k
        if m.any():
            assert_bits_equal(p["rays"][m, k], q["rays"][m, k], f"{what} vertex {k}")
    empty = p["length"] == 0
    if empty.any():  # rays[0] of an empty path is read by connect_paths' splat code: defined as zeros
        assert_bits_equal(p["rays"][empty, 0], q["rays"][empty, 0], f"{what} empty-path slot 0")


def rel_rmse(img, ref):
    """Relative RMSE: sqrt(mean((img-ref)^2)) / mean(ref)."""
    img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    return float(np.sqrt(np.mean((img - ref) ** 2)) / max(np.mean(ref), 1e-30))


def random_rays(n, lo, hi, seed=1):
    """Rays with origins uniform in the box [lo, hi] and uniform directions, as Ray records."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, T.Ray)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["origin"][:, :3] = o
    rays["direction"][:, :3] = d
    with np.errstate(divide="ignore"):
        rays["inv_direction"][:, :3] = np.float32(1.0) / d
    rays["triangle"] = -1
    return rays


@functools.lru_cache(maxsize=None)
def zoo_arrays(width, height):
    """Cornell box + two small spheres with a material table that exercises every branch of the material dispatch
    (trace.metal:474-487): walls mirror (type 3), glossy coated (type 2, alpha 0.3), rough/smooth dielectrics (type 1,
    alpha 0 and 0.1, different ior), a coloured emitter and a rough diffuse with alpha > 0 (GGX normal != n)."""
    import numpy as _np
    from clive2_b200 import meshgen
    from clive2_b200.load import make_materials
    mats = make_materials({
        0: dict(type=1, alpha=0.0, ior=1.33),             # smooth dielectric sphere
        5: dict(type=1, alpha=0.1, ior=1.8),              # rough dielectric sphere
        1: dict(type=3, alpha=0.0),                       # left wall: mirror
        2: dict(type=2, alpha=0.3),                       # right wall: coated
        3: dict(type=0, alpha=0.25),                      # front wall: diffuse, but its microfacet normal wanders
        6: dict(emission=[0.9, 1.0, 0.8]),                # tinted light
    })
    mesh = meshgen.icosphere(2, 1.2)
    specs = [{"mesh": mesh, "offset": _np.array([1.5, -0.5, 1.0]), "material": 5},
             {"mesh": mesh, "offset": _np.array([-1.8, 0.2, -1.0]), "material": 0}]
    return scene.build_scene_arrays(width, height, _np.array([0, 1.5, 6]), _np.array([0, 0, -1]), specs, mats,
                                    bvh_builder="numpy")
