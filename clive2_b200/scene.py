"""Scene assembly: sensor quad + Cornell enclosure + meshes -> flattened BVH -> kernel argument arrays.

Surface mirror of the reference's src/scene.py:21-245.  `build_scene_arrays` does the host work and returns plain numpy
arrays (usable without a GPU: oracle, tests); `create_scene` uploads them through `runtime.Device` and returns the
`Scene` attribute bag the renderer reads (same attribute names as scene.py:107-134).
"""
from dataclasses import dataclass

import numpy as np

from . import meshgen
from .bvh import FastTreeBox, build_flat_bvh
from .camera import Camera
from .constants import UNIT_Z, ZERO_VECTOR
from .load import (camera_geometry, fast_load, fast_load_obj, fast_load_ply, get_materials, make_materials,
                   surface_area, triangles_for_box)
from .struct_types import Triangle as TriangleRecord


@dataclass
class SceneArrays:
    """Host-side kernel arguments, all in `struct_types` layouts."""
    pixel_width: int
    pixel_height: int
    camera: np.ndarray                  # Camera[1]
    boxes: np.ndarray                   # Box[n_boxes]
    triangles: np.ndarray               # Triangle[n_tris]
    materials: np.ndarray               # Material[n_mats]
    light_triangles: np.ndarray         # Triangle[n_lights]
    light_counts: np.ndarray            # int32 0-d
    light_surface_areas: np.ndarray     # float32[n_lights]
    light_triangle_indices: np.ndarray  # int32[n_lights]
    camera_triangle_indices: np.ndarray  # int32[n_camera_tris]


_SOUP_CACHE = {}


def _mesh_soup_cached(spec):
    """`_mesh_soup` memoised per (mesh source, scale, offset, material): the movie driver rebuilds the scene for every
    frame (movie.py:29-55) although only the camera moves; loading and smoothing the meshes once removes most of the
    per-frame host work (the BVH itself is rebuilt: the sensor triangles are part of it and shape it)."""
    mesh = spec.get("mesh")
    if "file_path" not in spec and not callable(mesh):
        return _mesh_soup(spec)  # in-memory arrays: no identity that is safe to key on
    if "file_path" in spec:
        import os
        path = spec["file_path"]
        src = (path, os.path.getmtime(path) if os.path.exists(path) else None)
    else:
        src = mesh  # a preset's generator function object (kept alive by the preset table)
    key = (src, float(spec.get("scale", 1.0)), tuple(np.asarray(spec.get("offset", ZERO_VECTOR), dtype=np.float64).tolist()),
           int(spec.get("material", 0)))
    soup = _SOUP_CACHE.get(key)
    if soup is None:
        if len(_SOUP_CACHE) > 16:
            _SOUP_CACHE.clear()
        soup = _SOUP_CACHE[key] = _mesh_soup(spec)
    return soup


def _mesh_soup(spec):
    """One entry of `file_specs` -> soup.  Besides the reference's `file_path` (.obj/.ply, scene.py:49-66) a spec may
    carry an in-memory mesh as `mesh=(vertices, faces)`; scale/offset/material mean the same in both cases."""
    material = spec.get("material", 0)
    scale = spec.get("scale", 1.0)
    offset = spec.get("offset", ZERO_VECTOR)
    if "mesh" in spec:
        vertices, faces = spec["mesh"]() if callable(spec["mesh"]) else spec["mesh"]
        return fast_load(np.asarray(vertices) * scale + offset, np.asarray(faces), material=material)
    path = spec["file_path"]
    if path.endswith(".ply"):
        return fast_load_ply(ply_path=path, material=material, scale=scale, offset=offset)
    if path.endswith(".obj"):
        return fast_load_obj(obj_path=path, material=material, scale=scale, offset=offset)
    raise NotImplementedError(path)


def build_scene_arrays(pixel_width=1280, pixel_height=720, cam_center=ZERO_VECTOR, cam_direction=UNIT_Z,
                       file_specs=None, materials=None, bvh_builder="auto", sort_kind="stable"):
    camera = Camera(center=cam_center, direction=cam_direction, pixel_width=pixel_width, pixel_height=pixel_height,
                    phys_width=pixel_width / pixel_height, phys_height=1)
    # sensor triangles first, then the enclosure: this order fixes the soup order the BVH build starts from
    soup = FastTreeBox.from_triangle_objects(camera_geometry(camera) + triangles_for_box())
    for spec in file_specs or ():
        soup = soup + _mesh_soup_cached(spec)
    boxes, triangles = build_flat_bvh(soup, builder=bvh_builder, sort_kind=sort_kind)

    is_light = triangles["is_light"] != 0
    light_idx = np.flatnonzero(is_light).astype(np.int32)
    lights = triangles[light_idx] if len(light_idx) else np.zeros(0, dtype=TriangleRecord)
    areas = np.array([surface_area(t) for t in lights], dtype=np.float32)
    return SceneArrays(
        pixel_width=pixel_width,
        pixel_height=pixel_height,
        camera=np.array([camera.to_struct()]).reshape(1),
        boxes=boxes,
        triangles=triangles,
        materials=get_materials() if materials is None else materials,
        light_triangles=np.ascontiguousarray(lights),
        light_counts=np.array(len(lights), dtype=np.int32),
        light_surface_areas=areas,
        light_triangle_indices=light_idx,
        camera_triangle_indices=np.flatnonzero(triangles["is_camera"] != 0).astype(np.int32),
    )


class Scene:
    """Attribute bag consumed by `Renderer` (scene.py:107-134).  Buffer-valued attributes are `runtime.Buffer`s;
    the two index lists stay bare numpy arrays as in the reference (scene.py:84-89).  `arrays` keeps the host copy."""

    def __init__(self, device, pixel_width, pixel_height, camera, triangles, boxes, materials, light_triangles,
                 light_counts, light_surface_areas, light_triangle_indices, camera_triangle_indices, arrays=None):
        self.device = device
        self.pixel_width = pixel_width
        self.pixel_height = pixel_height
        self.camera = camera
        self.triangles = triangles
        self.boxes = boxes
        self.materials = materials
        self.light_triangles = light_triangles
        self.light_counts = light_counts
        self.light_surface_areas = light_surface_areas
        self.light_triangle_indices = light_triangle_indices
        self.camera_triangle_indices = camera_triangle_indices
        self.arrays = arrays

    def release(self):
        for name in ("camera", "triangles", "boxes", "materials", "light_triangles", "light_counts",
                     "light_surface_areas"):
            buf = getattr(self, name, None)
            if buf is not None and hasattr(buf, "release"):
                buf.release()
            setattr(self, name, None)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def scene_from_arrays(arrays, device=None):
    """Upload `SceneArrays` (host -> device boundary, scene.py:74-89)."""
    from . import runtime
    dev = runtime.Device() if device is None else device
    up = dev.buffer
    return Scene(
        device=dev, pixel_width=arrays.pixel_width, pixel_height=arrays.pixel_height,
        # boxes / triangles: device copies are made only if a stage-level kernel asks for them (runtime.LazyBuffer); the
        # fused renderer traverses the acceleration structure built from the host arrays
        camera=up(arrays.camera), triangles=dev.lazy_buffer(arrays.triangles), boxes=dev.lazy_buffer(arrays.boxes),
        materials=up(arrays.materials),
        light_triangles=up(arrays.light_triangles), light_counts=up(arrays.light_counts),
        light_surface_areas=up(arrays.light_surface_areas),
        light_triangle_indices=arrays.light_triangle_indices,
        camera_triangle_indices=arrays.camera_triangle_indices,
        arrays=arrays,
    )


def create_scene(pixel_width=1280, pixel_height=720, cam_center=ZERO_VECTOR, cam_direction=UNIT_Z, file_specs=None,
                 materials=None, device=None, bvh_builder="auto"):
    arrays = build_scene_arrays(pixel_width, pixel_height, cam_center, cam_direction, file_specs, materials,
                                bvh_builder)
    return scene_from_arrays(arrays, device)


# ---------------------------------------------------------------------------------------------------------------
# presets.  The first five are the reference's (scene.py:149-206); their mesh files are not shipped with it.
# The `bench-*` presets are the procedural stand-ins named in BASELINE.md §5.
# ---------------------------------------------------------------------------------------------------------------
def _dragon(path):
    return {"cam_center": np.array([0, 1.5, 7.5]), "cam_direction": np.array([0, 0, -1]),
            "file_specs": [{"file_path": path, "offset": np.array([0, -4, 0]), "material": 5, "scale": 50}]}


def _two_of(mesh_or_path_key, value):
    return [{mesh_or_path_key: value, "offset": np.array([0, 0, 2.5]), "material": 5},
            {mesh_or_path_key: value, "offset": np.array([0, 0, -2.5]), "material": 0}]


#: rough dielectric on material 5, coated diffuse on material 0 (the stock table has alpha = 0)
GLOSSY_MATERIAL_OVERRIDES = {5: dict(type=1, alpha=0.05, ior=1.5), 0: dict(type=2, alpha=0.2)}

scene_presets = {
    "empty": {"cam_center": np.array([0, 1.5, 6]), "cam_direction": np.array([0, 0, -1])},
    "teapots": {"cam_center": np.array([7, 0, 8]), "cam_direction": np.array([-1, 0, -1]),
                "file_specs": _two_of("file_path", "../resources/teapot.obj")},
    "dragon": _dragon("../resources/dragon_vrip_res3.ply"),
    "medium-dragon": _dragon("../resources/dragon_vrip_res2.ply"),
    "big-dragon": _dragon("../resources/dragon_vrip.ply"),
    # --- procedural ---
    "bench-spheres": {"cam_center": np.array([7, 0, 8]), "cam_direction": np.array([-1, 0, -1]),
                      "file_specs": _two_of("mesh", lambda: meshgen.icosphere(5, 1.5)),
                      "materials": GLOSSY_MATERIAL_OVERRIDES},
    "bench-terrain-1m": {"cam_center": np.array([0, 1.5, 6]), "cam_direction": np.array([0, 0, -1]),
                         "file_specs": [{"mesh": lambda: meshgen.terrain(708, 708), "material": 3}]},
    "bench-terrain-10m": {"cam_center": np.array([0, 1.5, 6]), "cam_direction": np.array([0, 0, -1]),
                          "file_specs": [{"mesh": lambda: meshgen.terrain(2236, 2236), "material": 3}]},
}


def _preset(name):
    preset = scene_presets.get(name)
    if not preset:
        raise ValueError(f"Preset '{name}' not found.")
    return preset


def preset_arrays(name, pixel_width=1280, pixel_height=720, cam_center=None, cam_direction=None, bvh_builder="auto"):
    p = _preset(name)
    mats = make_materials(p["materials"]) if "materials" in p else None
    return build_scene_arrays(pixel_width, pixel_height,
                              p["cam_center"] if cam_center is None else cam_center,
                              p["cam_direction"] if cam_direction is None else cam_direction,
                              p.get("file_specs"), mats, bvh_builder)


def create_scene_from_preset(preset_name, pixel_width=1280, pixel_height=720, device=None):
    return scene_from_arrays(preset_arrays(preset_name, pixel_width, pixel_height), device)


def orbit_camera(frame_idx, total_frames):
    """Camera centre/direction of movie frame `frame_idx` (scene.py:232-237): radius 7.5 circle at y = 1.5."""
    theta = 2 * np.pi * frame_idx / total_frames
    return (np.array([np.sin(theta) * 7.5, 1.5, np.cos(theta) * 7.5]),
            np.array([-np.sin(theta), 0, -np.cos(theta)]))


def create_scene_from_preset_with_params(preset_name, pixel_width=1280, pixel_height=720, frame_idx=0, total_frames=1,
                                         device=None):
    center, direction = orbit_camera(frame_idx, total_frames)
    return scene_from_arrays(preset_arrays(preset_name, pixel_width, pixel_height, center, direction), device)
