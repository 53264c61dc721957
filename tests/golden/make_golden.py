#!/usr/bin/env python
"""Regenerate tests/golden/* by running the REFERENCE's own Python (read-only at /root/reference/src).

Only runs in the build container (the GPU box has no /root/reference).  The reference's host code imports three
modules that are not installed here and only matter for file parsing / device upload; they are stubbed:
  metalcompute -> Device().buffer(x) returns x unchanged (scene.py:29,74-82), release() is a no-op
  objloader    -> Obj.open(path) serves an in-memory mesh registered under that path (load.py:80-82)
  plyfile      -> PlyData.read(path) likewise (load.py:90-94)
Everything written is an output of unmodified reference code: struct_types, camera.py, load.py, bvh.py, scene.py.
The committed fixtures pin the host layer and the inputs of the oracle; the Metal kernels themselves cannot be
executed on Linux, so there are no reference-produced kernel outputs (parity of the kernel arithmetic is unpinned).

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"

_MESHES = {}


def _install_stubs():
    mc = types.ModuleType("metalcompute")

    class Device:
        def buffer(self, x):
            return x

    mc.Device = Device
    mc.release = lambda *_a, **_k: None
    mc.error = RuntimeError
    sys.modules["metalcompute"] = mc

    ol = types.ModuleType("objloader")

    class Obj:
        def __init__(self, v, f):
            self.vert = [tuple(r) for r in v]
            self.face = [(int(i) + 1, 0, 0) for i in np.asarray(f).ravel()]

        @classmethod
        def open(cls, path):
            return cls(*_MESHES[path])

    ol.Obj = Obj
    sys.modules["objloader"] = ol

    pf = types.ModuleType("plyfile")

    class PlyData(dict):
        @classmethod
        def read(cls, path):
            v, f = _MESHES[path]
            rec = np.zeros(len(v), dtype=[("x", "f4"), ("y", "f4"), ("z", "f4")])
            rec["x"], rec["y"], rec["z"] = v[:, 0], v[:, 1], v[:, 2]
            d = cls()
            d["vertex"] = types.SimpleNamespace(data=rec)
            d["face"] = {"vertex_indices": [np.asarray(r) for r in f]}
            return d

    pf.PlyData = PlyData
    sys.modules["plyfile"] = pf


def _scene_dump(scene):
    return dict(
        camera=np.asarray(scene.camera), boxes=np.asarray(scene.boxes), triangles=np.asarray(scene.triangles),
        materials=np.asarray(scene.materials), light_triangles=np.asarray(scene.light_triangles),
        light_counts=np.asarray(scene.light_counts), light_surface_areas=np.asarray(scene.light_surface_areas),
        light_triangle_indices=np.asarray(scene.light_triangle_indices),
        camera_triangle_indices=np.asarray(scene.camera_triangle_indices),
    )


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, REPO)
    import struct_types as ref_types
    import scene as ref_scene
    import camera as ref_camera
    from clive2_b200 import meshgen  # inputs only; saved alongside so the fixture is self-contained

    # 1. record layouts
    layouts = {}
    for name in ("Ray", "Path", "Box", "Triangle", "Material", "Camera"):
        dt = getattr(ref_types, name)
        layouts[name] = {"itemsize": dt.itemsize, "offsets": {k: int(v[1]) for k, v in dt.fields.items()}}
    with open(os.path.join(HERE, "layouts.json"), "w") as fh:
        json.dump(layouts, fh, indent=1, sort_keys=True)

    # 2. preset "empty" at 256x256 (BASELINE config 1) and its camera at a second resolution
    s = ref_scene.create_scene_from_preset("empty", 256, 256)
    np.savez_compressed(os.path.join(HERE, "scene_empty_256.npz"), **_scene_dump(s))

    cams = {}
    for tag, (center, direction, w, h) in {
        "empty_256": (np.array([0, 1.5, 6]), np.array([0, 0, -1]), 256, 256),
        "teapots_1280x720": (np.array([7, 0, 8]), np.array([-1, 0, -1]), 1280, 720),
        "orbit_f17_of_120": (None, None, 1280, 720),
    }.items():
        if center is None:
            theta = 2 * np.pi * 17 / 120
            center = np.array([np.sin(theta) * 7.5, 1.5, np.cos(theta) * 7.5])
            direction = np.array([-np.sin(theta), 0, -np.cos(theta)])
        c = ref_camera.Camera(center=center, direction=direction, pixel_width=w, pixel_height=h,
                              phys_width=w / h, phys_height=1)
        cams[tag] = c.to_struct()
    np.savez_compressed(os.path.join(HERE, "cameras.npz"), **cams)

    # 3. scenes with meshes: OBJ path (float64 vertices) and PLY path (float32 vertices), tie-free (jittered) inputs
    rng = np.random.default_rng(7)
    sv, sf = meshgen.icosphere(2, 1.5)
    sv = sv + rng.normal(0, 1e-3, sv.shape)
    _MESHES["sphere.obj"] = (sv, sf)
    tv, tf = meshgen.terrain(24, 24, jitter=0.01, seed=3)
    _MESHES["terrain.ply"] = (tv.astype(np.float32), tf)

    s = ref_scene.create_scene(
        pixel_width=64, pixel_height=48, cam_center=np.array([7, 0, 8]), cam_direction=np.array([-1, 0, -1]),
        file_specs=[{"file_path": "sphere.obj", "offset": np.array([0, 0, 2.5]), "material": 5},
                    {"file_path": "sphere.obj", "offset": np.array([0, 0, -2.5]), "material": 0}])
    np.savez_compressed(os.path.join(HERE, "scene_spheres_obj.npz"), mesh_vertices=sv, mesh_faces=sf, **_scene_dump(s))

    s = ref_scene.create_scene(
        pixel_width=32, pixel_height=32, cam_center=np.array([0, 1.5, 6]), cam_direction=np.array([0, 0, -1]),
        file_specs=[{"file_path": "terrain.ply", "material": 3, "scale": 1.0}])
    np.savez_compressed(os.path.join(HERE, "scene_terrain_ply.npz"), mesh_vertices=tv.astype(np.float32), mesh_faces=tf,
                        **_scene_dump(s))

    # 3b. tie-free soup straight through the reference builder (platform independent: no equal centroids anywhere)
    import bvh as ref_bvh
    import load as ref_load
    rv = np.random.default_rng(5).uniform(-4, 4, size=(3 * 2500, 3))
    rv[1::3] = rv[0::3] + np.random.default_rng(6).normal(0, 0.15, size=(2500, 3))
    rv[2::3] = rv[0::3] + np.random.default_rng(8).normal(0, 0.15, size=(2500, 3))
    rf = np.arange(3 * 2500, dtype=np.int32).reshape(-1, 3)
    soup = ref_load.fast_load(rv, rf, material=2)
    rb, rt = ref_bvh.np_flatten_bvh(ref_bvh.construct_BVH(soup))
    np.savez_compressed(os.path.join(HERE, "bvh_random_soup.npz"), mesh_vertices=rv, mesh_faces=rf, boxes=rb, triangles=rt)

    # 3c. how this platform's default argsort orders ties (lets tests tell whether byte-equality with 2./3. is expected)
    probe = np.repeat(np.random.default_rng(9).integers(0, 40, 64), 4).astype(np.float64)
    np.random.default_rng(10).shuffle(probe)
    np.savez_compressed(os.path.join(HERE, "argsort_tie_probe.npz"), probe=probe, order=np.argsort(probe))

    # 4. tone map of a fixed random image (camera.py:73-82)
    img = np.random.default_rng(11).random((12, 16, 3)).astype(np.float32) * 3.0
    np.savez_compressed(os.path.join(HERE, "tone_map.npz"), image=img, out=ref_camera.tone_map(img, exposure=4.0))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
