"""Host layer against arrays produced by the REFERENCE's own Python (tests/golden/make_golden.py):
scene assembly, flattened BVH, camera struct, material table, tone map.  (reference: scene.py, bvh.py, camera.py, load.py)"""
import os

import numpy as np
import pytest

from clive2_b200 import scene, struct_types as T
from clive2_b200.bvh import build_flat_bvh
from clive2_b200.camera import Camera, tone_map
from clive2_b200.load import fast_load
from tests.common import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _platform_sorts_ties_like_golden():
    g = _load("argsort_tie_probe.npz")
    return np.array_equal(np.argsort(g["probe"]), g["order"])


def _same_up_to_leaf_permutation(boxes, tris, g_boxes, g_tris):
    """Boxes equal byte for byte; triangles equal as multisets inside each leaf."""
    assert np.array_equal(boxes.view(np.uint8), g_boxes.view(np.uint8))
    for bx in boxes[boxes["right"] != 0]:
        a = np.sort(tris[bx["left"]:bx["right"]].view(np.uint8).reshape(-1, 128), axis=0)
        b = np.sort(g_tris[bx["left"]:bx["right"]].view(np.uint8).reshape(-1, 128), axis=0)
        assert np.array_equal(a, b)


def test_empty_preset_matches_reference():
    g = _load("scene_empty_256.npz")
    a = scene.preset_arrays("empty", 256, 256, bvh_builder="numpy")
    assert np.array_equal(a.camera.view(np.uint8), g["camera"].view(np.uint8).reshape(-1))
    assert np.array_equal(a.materials.view(np.uint8), g["materials"].view(np.uint8))
    assert len(a.boxes) == 5 and len(a.triangles) == 16
    _same_up_to_leaf_permutation(a.boxes, a.triangles, g["boxes"], g["triangles"])
    assert int(a.light_counts) == int(g["light_counts"]) == 2
    assert np.array_equal(a.light_surface_areas, g["light_surface_areas"])
    assert np.array_equal(np.sort(a.light_triangle_indices), np.sort(g["light_triangle_indices"]))
    assert np.array_equal(a.camera_triangle_indices, g["camera_triangle_indices"])


def test_empty_preset_bytes_when_sort_matches():
    if not _platform_sorts_ties_like_golden():
        pytest.skip("this platform's unstable argsort orders ties differently from the fixture's")
    g = _load("scene_empty_256.npz")
    soup_arrays = scene.preset_arrays("empty", 256, 256, bvh_builder="numpy")
    # the default (stable) order can differ inside leaves; the reference order is reproduced with sort_kind=None
    from clive2_b200.bvh import FastTreeBox, construct_BVH, np_flatten_bvh
    from clive2_b200.load import camera_geometry, triangles_for_box
    cam = Camera(center=np.array([0, 1.5, 6]), direction=np.array([0, 0, -1]), pixel_width=256, pixel_height=256,
                 phys_width=1.0, phys_height=1)
    soup = FastTreeBox.from_triangle_objects(camera_geometry(cam) + triangles_for_box())
    boxes, tris = np_flatten_bvh(construct_BVH(soup, sort_kind=None))
    assert np.array_equal(boxes.view(np.uint8), g["boxes"].view(np.uint8))
    assert np.array_equal(tris.view(np.uint8), g["triangles"].view(np.uint8))
    assert len(soup_arrays.boxes) == len(boxes)


def test_tie_free_soup_bvh_is_byte_identical():
    g = _load("bvh_random_soup.npz")
    soup = fast_load(g["mesh_vertices"], g["mesh_faces"], material=2)
    boxes, tris = build_flat_bvh(soup, builder="numpy")
    assert np.array_equal(boxes.view(np.uint8), g["boxes"].view(np.uint8))
    assert np.array_equal(tris.view(np.uint8), g["triangles"].view(np.uint8))


@pytest.mark.parametrize("fixture,kwargs", [
    ("scene_spheres_obj.npz", dict(pixel_width=64, pixel_height=48, cam_center=np.array([7, 0, 8]),
                                   cam_direction=np.array([-1, 0, -1]))),
    ("scene_terrain_ply.npz", dict(pixel_width=32, pixel_height=32, cam_center=np.array([0, 1.5, 6]),
                                   cam_direction=np.array([0, 0, -1]))),
])
def test_mesh_scenes_match_reference(fixture, kwargs):
    g = _load(fixture)
    mesh = (g["mesh_vertices"], g["mesh_faces"])
    if "spheres" in fixture:
        specs = [{"mesh": mesh, "offset": np.array([0, 0, 2.5]), "material": 5},
                 {"mesh": mesh, "offset": np.array([0, 0, -2.5]), "material": 0}]
    else:
        specs = [{"mesh": mesh, "material": 3, "scale": 1.0}]
    # grid/sphere neighbours share extreme vertices -> tied centroids -> the tree depends on how the platform's
    # unstable argsort orders ties: reproduce the reference's call (sort_kind=None) where it applies
    exact = _platform_sorts_ties_like_golden()
    a = scene.build_scene_arrays(file_specs=specs, bvh_builder="numpy", sort_kind=None if exact else "stable", **kwargs)
    assert np.array_equal(a.camera.view(np.uint8), g["camera"].view(np.uint8).reshape(-1))
    assert len(a.triangles) == len(g["triangles"])
    if exact:
        assert np.array_equal(a.boxes.view(np.uint8), g["boxes"].view(np.uint8))
        assert np.array_equal(a.triangles.view(np.uint8), g["triangles"].view(np.uint8))
    # whatever the tie order, the same triangle set must come out
    mine = np.sort(a.triangles.view(np.uint8).reshape(-1, 128), axis=0)
    ref = np.sort(g["triangles"].view(np.uint8).reshape(-1, 128), axis=0)
    assert np.array_equal(mine, ref)


def test_camera_structs():
    g = _load("cameras.npz")
    cases = {"empty_256": (np.array([0, 1.5, 6]), np.array([0, 0, -1]), 256, 256),
             "teapots_1280x720": (np.array([7, 0, 8]), np.array([-1, 0, -1]), 1280, 720)}
    c, d = scene.orbit_camera(17, 120)
    cases["orbit_f17_of_120"] = (c, d, 1280, 720)
    for tag, (center, direction, w, h) in cases.items():
        cam = Camera(center=center, direction=direction, pixel_width=w, pixel_height=h, phys_width=w / h, phys_height=1)
        assert np.array_equal(cam.to_struct().view(np.uint8), g[tag].view(np.uint8).reshape(-1)), tag


def test_tone_map():
    g = _load("tone_map.npz")
    assert np.array_equal(tone_map(g["image"], exposure=4.0), g["out"])


def test_unknown_preset_raises_like_reference():
    with pytest.raises(ValueError):
        scene.preset_arrays("no-such-preset", 8, 8)
