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


def _nan_canonical(a):
    """32-bit words with every NaN bit pattern replaced by one quiet NaN: x86 and the GPU produce different payloads /
    signs for the same invalid operation (0/0 is 0xffc00000 on SSE, 0x7fffffff on CUDA); which NaN it is never matters
    to the path (comparisons with NaN are false, fminf/fmaxf drop it on both)."""
    w = np.ascontiguousarray(a).view(np.uint32).copy()
    w[(w & 0x7f800000 == 0x7f800000) & (w & 0x007fffff != 0)] = 0x7fc00000
    return w


def assert_bits_equal(x, y, what, nan_equal=False):
    bx, by = (_nan_canonical(x).view(np.uint8), _nan_canonical(y).view(np.uint8)) if nan_equal else (bits(x), bits(y))
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


def fuzz_soup(seed, n_random=300):
    """Box[]/Triangle[] (numpy builder) of `fuzz_triangle_soup`."""
    from clive2_b200.bvh import build_flat_bvh
    return build_flat_bvh(fuzz_triangle_soup(seed, n_random), builder="numpy")


def fuzz_triangle_soup(seed, n_random=300):
    """A deliberately awkward triangle soup (FastTreeBox): random triangles, EXACT duplicates (equal t on every
    ray: the reference keeps the first in its visiting order), coplanar overlapping pairs, zero-area and collinear
    triangles (Moeller-Trumbore divides by a = 0), large axis-aligned quads whose boxes are flat, and a far outlier."""
    from clive2_b200.bvh import FastTreeBox
    rng = np.random.default_rng(seed)
    f32 = np.float32
    base = rng.uniform(-2, 2, (n_random, 1, 3))
    tris = [(base + rng.normal(0, 0.4, (n_random, 3, 3))).astype(f32)]
    tris.append(tris[0][:40].copy())                                   # exact duplicates
    shifted = tris[0][40:70].copy()
    e = shifted[:, 1] - shifted[:, 0]
    tris.append((shifted + 0.25 * e[:, None, :]).astype(f32))          # coplanar, overlapping their originals
    deg = tris[0][70:80].copy(); deg[:, 1] = deg[:, 0]                 # zero area: v1 == v0
    col = tris[0][80:90].copy(); col[:, 2] = col[:, 0] + 2 * (col[:, 1] - col[:, 0])  # collinear
    tris += [deg, col]
    quads = []
    for axis in range(3):
        for c in (-2.5, 0.0, 2.5):
            p = np.zeros((4, 3), f32)
            u, v = [a for a in range(3) if a != axis]
            p[:, axis] = c
            p[:, u] = [-3, 3, 3, -3]
            p[:, v] = [-3, -3, 3, 3]
            quads += [p[[0, 1, 2]], p[[0, 2, 3]]]
    tris.append(np.array(quads, f32))
    tris.append(np.array([[[50, 50, 50], [51, 50, 50], [50, 51, 50]]], f32))
    corners = np.concatenate(tris).astype(f32)
    n = len(corners)
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = np.cross(corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0])
        area = np.linalg.norm(nrm, axis=1)
        nrm = np.nan_to_num(nrm / area[:, None]).astype(f32)
    return FastTreeBox(np.zeros((n, 3), np.uint32), corners, corners.min(axis=1), corners.max(axis=1), nrm,
                       np.repeat(nrm[:, None, :], 3, axis=1), (area / 2).astype(f32), np.zeros(n, np.int32),
                       np.zeros(n, np.bool_), np.zeros(n, np.bool_))


def fuzz_rays(triangles, n, seed):
    """Random rays plus the awkward ones: axis-parallel directions (zero components: inv_direction = +-inf, 0 * inf = NaN
    in the slab test), origins exactly on triangle vertices and inside triangle planes, rays along triangle edges."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    o = rng.uniform(-3.5, 3.5, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3))
    kind = rng.integers(0, 6, n)
    ax = rng.integers(0, 3, n)
    sign = rng.choice([-1.0, 1.0], n)
    for a in range(3):                       # kind 1: one exact axis direction; kind 2: one zero component
        m1 = (kind == 1) & (ax == a)
        d[m1] = 0.0
        d[m1, a] = sign[m1]
        d[(kind == 2) & (ax == a), a] = 0.0
    k = rng.integers(0, len(triangles), n)
    v0, v1, v2 = (triangles[name][k, :3] for name in ("v0", "v1", "v2"))
    m3 = kind == 3                           # origin exactly on a vertex
    o[m3] = v0[m3]
    m4 = kind == 4                           # origin inside a triangle's plane (barycentric point)
    w = rng.random((n, 2)).astype(f32)
    o[m4] = (v0 + (v1 - v0) * w[:, :1] * 0.5 + (v2 - v0) * w[:, 1:] * 0.5)[m4]
    m5 = kind == 5                           # along an edge, from outside
    d[m5] = (v1 - v0)[m5]
    o[m5] = (v0 - 3 * (v1 - v0))[m5]
    with np.errstate(invalid="ignore", divide="ignore"):
        nd = np.linalg.norm(d, axis=1, keepdims=True)
        d = np.where(nd > 0, d / nd, [[1.0, 0.0, 0.0]]).astype(f32)
        rays = np.zeros(n, T.Ray)
        rays["origin"][:, :3], rays["direction"][:, :3] = o, d
        rays["inv_direction"][:, :3] = f32(1.0) / d
    rays["triangle"] = -1
    return rays


def check_ordered_against_reference(rays, tris, tri, t, tri_o, t_o):
    """ORDERED may leave the reference only where the reference's own hit is a rounding artefact: Moeller-Trumbore has
    no parallel-ray guard (trace.metal:117-142), so a ray within ~1e-7 of a triangle's plane gets f = 1/a ~ 1e7 and
    garbage (u, v, t) that can pass the range checks although o + t d is nowhere near the triangle.  Whether such a
    'hit' is ever looked at depends on the visiting order; near-first traversal prunes its box.  Everything else — exact
    ties between duplicates included — must resolve to the reference's triangle."""
    mm = np.flatnonzero(tri != tri_o)
    assert len(mm) < 0.01 * len(tri), len(mm)
    o, d = rays["origin"][mm, :3], rays["direction"][mm, :3]

    def real_hit(k, tt):  # the hit point lies inside the triangle's own (slightly inflated) bounding box
        ok = (k >= 0) & np.isfinite(tt)
        kk = np.where(ok, k, 0)
        corners = np.stack([tris[name][kk, :3] for name in ("v0", "v1", "v2")], axis=1)
        with np.errstate(invalid="ignore", over="ignore"):
            p = o + np.where(ok, tt, 0)[:, None] * d
        lo, hi = corners.min(axis=1) - 1e-3, corners.max(axis=1) + 1e-3
        return ok & ((p >= lo) & (p <= hi)).all(axis=1)

    ref_real, ord_real = real_hit(tri[mm], t[mm]), real_hit(tri_o[mm], t_o[mm])
    assert not (tri_o[mm] >= 0)[~ord_real].any()               # ORDERED never reports an artefact or an infinite hit ...
    bad = ref_real & ~(ord_real & (np.abs(t[mm] - t_o[mm]) <= 1e-3 * np.abs(t[mm])) & (t[mm] != t_o[mm]))
    assert not bad.any(), (int(bad.sum()), mm[bad][:5])        # ... and only leaves the reference where IT does


AWKWARD_CAMERAS = [((0, 1.5, 30), (0, 0, -1)),      # far outside the enclosure: most camera paths have length 0 or 1
                   ((0, 1.5, 6), (0, 0, 1)),        # looking away from the scene, at the wall behind the camera
                   ((0, 8.5, 0), (0, -1, 0)),       # straight down from next to the light
                   ((4.9, 0, 0), (-1, 0.2, 0.1)),   # centimetres from a wall
                   ((40, 40, 40), (-1, -1, -1))]    # diagonal view from far away


@functools.lru_cache(maxsize=None)
def awkward_camera_arrays(which, width, height):
    """Cornell box + one glossy sphere seen from AWKWARD_CAMERAS[which]: empty and one-vertex camera paths, t = 1
    projections that land on the last pixel row/column and beyond (pixel index == W*H + ...)."""
    from clive2_b200 import meshgen
    from clive2_b200.load import make_materials
    mats = make_materials({0: dict(type=1, alpha=0.05, ior=1.5), 5: dict(type=2, alpha=0.2)})
    center, direction = AWKWARD_CAMERAS[which]
    specs = [{"mesh": meshgen.icosphere(1, 0.8), "offset": np.array([0.5, -0.5, 0.0]), "material": 0}]
    return scene.build_scene_arrays(width, height, np.array(center, float), np.array(direction, float), specs, mats,
                                    bvh_builder="numpy")


@functools.lru_cache(maxsize=None)
def random_material_arrays(trial, width, height):
    """Cornell box + two spheres with a random material table drawn from extreme values (alpha 0 ... 3, ior 0.7 ... 10,
    every type on every surface): total internal reflection everywhere, GGX with alpha > 1, ior = 1, and — for some
    tables — NaN radiance, which must then be the same NaN-ness on both sides."""
    from clive2_b200 import meshgen
    from clive2_b200.load import make_materials
    rng = np.random.default_rng(1000 + trial)
    alphas, iors = [0.0, 1e-4, 0.05, 0.5, 1.0, 3.0], [1.0, 1.0001, 1.33, 2.4, 0.7, 10.0]
    over = {m: dict(type=int(rng.integers(0, 4)), alpha=float(rng.choice(alphas)), ior=float(rng.choice(iors))) for m in range(6)}
    over[6] = dict(emission=[1.0, 1.0, 1.0])
    mesh = meshgen.icosphere(2, 1.0)
    specs = [{"mesh": mesh, "offset": np.array([1.2, -0.5, 0.5]), "material": 5},
             {"mesh": mesh, "offset": np.array([-1.5, 0.0, -0.8]), "material": 0}]
    return scene.build_scene_arrays(width, height, np.array([0, 1.5, 6.0]), np.array([0, 0, -1.0]), specs,
                                    make_materials(over), bvh_builder="numpy")


@functools.lru_cache(maxsize=None)
def many_lights_arrays(n_extra, width, height):
    """Cornell box + two spheres where `n_extra` faces of one sphere are emitters too (n_extra = 0: ONE light, the
    second ceiling triangle switched off): light pick, pdf 1 / (count * area) and hit_light on mesh lights."""
    import dataclasses
    from clive2_b200 import meshgen
    from clive2_b200.load import make_materials
    from clive2_b200.scene import surface_area
    mesh = meshgen.icosphere(1, 0.7)
    specs = [{"mesh": mesh, "offset": np.array([1.0, 0.5, 0.0]), "material": 5},
             {"mesh": mesh, "offset": np.array([-1.5, -0.8, 0.5]), "material": 0}]
    a = scene.build_scene_arrays(width, height, np.array([0, 1.5, 6.0]), np.array([0, 0, -1.0]), specs,
                                 make_materials({0: dict(type=2, alpha=0.2)}), bvh_builder="numpy")
    tris = a.triangles.copy()
    extra = np.flatnonzero(tris["material"] == 5)[:n_extra]
    tris["is_light"][extra] = 1
    tris["material"][extra] = 6
    if n_extra == 0:
        tris["is_light"][np.flatnonzero(tris["is_light"] != 0)[1]] = 0
    idx = np.flatnonzero(tris["is_light"] != 0).astype(np.int32)
    lights = np.ascontiguousarray(tris[idx])
    return dataclasses.replace(a, triangles=tris, light_triangles=lights, light_counts=np.array(len(lights), np.int32),
                               light_surface_areas=np.array([surface_area(t) for t in lights], np.float32),
                               light_triangle_indices=idx)
