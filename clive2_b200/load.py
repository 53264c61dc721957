"""Geometry sources: Cornell enclosure, sensor quad, material table, mesh ingestion (arrays, OBJ, PLY).

Host-side, one-time scene preparation with the reference's numerics (src/load.py:28-277) so that the flattened
arrays match the ones the reference uploads.  The reference parses files with the third-party `objloader` and
`plyfile` modules (load.py:76-95); neither exists here, so `read_obj` / `read_ply` below are small self-contained
parsers that yield the same `(vertices, faces)` arrays those call sites consume.
"""
import numpy as np

from .bvh import FastTreeBox
from .constants import (BLUE, DEFAULT_BOX_MAX_CORNER, DEFAULT_BOX_MIN_CORNER, DEFAULT_LIGHT_HEIGHT,
                        DEFAULT_LIGHT_SCALE, FULL_WHITE, GREEN, RED, UNIT_X, UNIT_Y, UNIT_Z, WHITE)
from .struct_types import Material


def unit(v):
    return v / np.linalg.norm(v)


class Triangle:
    """A single host triangle (Cornell walls, light, sensor).  Meshes never go through this class."""

    def __init__(self, v0, v1, v2, _min=None, _max=None, normal=None, surface_area=None, material=0, emitter=False,
                 camera=False):
        self.v0, self.v1, self.v2 = v0, v1, v2
        self.n0 = self.n1 = self.n2 = np.zeros(3)
        self.material, self.emitter, self.camera = material, emitter, camera
        self.min = np.minimum(v0, np.minimum(v1, v2)) if _min is None else _min
        self.max = np.maximum(v0, np.maximum(v1, v2)) if _max is None else _max
        self.n = unit(np.cross(v1 - v0, v2 - v0)) if normal is None else normal
        # The reference stores the constructor argument unconditionally (load.py:71-73), i.e. None -> NaN in the soup;
        # nothing downstream reads it (light areas are recomputed by `surface_area` below).
        self.surface_area = surface_area


# ---------------------------------------------------------------------------------------------------------------
# fixed geometry
# ---------------------------------------------------------------------------------------------------------------
def triangles_for_box(box_min=DEFAULT_BOX_MIN_CORNER, box_max=DEFAULT_BOX_MAX_CORNER,
                      light_height=DEFAULT_LIGHT_HEIGHT, light_scale=DEFAULT_LIGHT_SCALE):
    """12 wall triangles + 2 emitter triangles, in the reference's order and winding (load.py:203-258)."""
    span = box_max - box_min
    lbb = box_min
    rbb = box_min + span * UNIT_X
    ltb = box_min + span * UNIT_Y
    lbf = box_min + span * UNIT_Z
    rtf = box_max
    ltf = box_max - span * UNIT_X
    rbf = box_max - span * UNIT_Y
    rtb = box_max - span * UNIT_Z
    walls = [
        (lbb, rbb, rtb, 4), (lbb, rtb, ltb, 4),      # back
        (lbb, ltf, lbf, 1), (lbb, ltb, ltf, 1),      # left
        (rbb, rbf, rtf, 2), (rbb, rtf, rtb, 2),      # right
        (lbf, rtf, rbf, 3), (lbf, ltf, rtf, 3),      # front
        (lbb, rbf, rbb, 4), (lbb, lbf, rbf, 4),      # floor
        (ltb, rtb, rtf, 4), (ltb, rtf, ltf, 4),      # ceiling
    ]
    tris = [Triangle(a, b, c, material=m) for a, b, c, m in walls]
    # ceiling light: the ceiling corners scaled about the origin (float32 factors, as in the reference)
    k = np.array([light_scale, light_height, light_scale], dtype=np.float32)
    tris.append(Triangle(ltb * k, rtb * k, rtf * k, material=6, emitter=True))
    tris.append(Triangle(ltb * k, rtf * k, ltf * k, material=6, emitter=True))
    return tris


def camera_geometry(camera):
    """The two sensor triangles (material 7, is_camera) spanning the physical sensor (load.py:261-271)."""
    o = camera.origin
    right = o + camera.dx * camera.phys_width
    top_right = o + camera.dx * camera.phys_width + camera.dy * camera.phys_height
    top_left = o + camera.dy * camera.phys_height
    return [Triangle(o, right, top_right, material=7, camera=True),
            Triangle(o, top_right, top_left, material=7, camera=True)]


def surface_area(t):
    """Area of a `struct_types.Triangle` record (load.py:274-277)."""
    e1 = (t["v1"] - t["v0"])[0:3]
    e2 = (t["v2"] - t["v0"])[0:3]
    return np.linalg.norm(np.cross(e1, e2)) / 2


# ---------------------------------------------------------------------------------------------------------------
# materials
# ---------------------------------------------------------------------------------------------------------------
#: material `type` values understood by the kernels (trace.metal:474-487)
MATERIAL_DIFFUSE, MATERIAL_DIELECTRIC, MATERIAL_COATED, MATERIAL_MIRROR = 0, 1, 2, 3


def get_materials():
    """The reference's eight-entry table (load.py:179-200): alpha 0, ior 1.5, types 1 for entries 0 and 5."""
    table = np.zeros(8, dtype=Material)
    for i, colour in enumerate((RED, GREEN, BLUE, WHITE, WHITE, BLUE, FULL_WHITE, FULL_WHITE)):
        table["color"][i][:3] = colour
    table["emission"][6] = 1.0
    table["ior"] = 1.5
    table["alpha"] = 0.0
    table["type"] = MATERIAL_DIFFUSE
    table["type"][[0, 5]] = MATERIAL_DIELECTRIC
    return table


def make_materials(overrides=None):
    """Default table with per-entry field overrides, e.g. `{5: dict(type=1, alpha=0.05), 0: dict(type=2, alpha=0.2)}`."""
    table = get_materials()
    for idx, fields in (overrides or {}).items():
        for name, value in fields.items():
            if name in ("color", "emission"):
                table[name][idx][: len(value)] = value
            else:
                table[name][idx] = value
    return table


# ---------------------------------------------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------------------------------------------
def _cross(a, b):
    """np.cross for (..., 3) arrays, component by component: the same products and differences (same bits), several times
    faster than numpy's general routine on millions of rows."""
    out = np.empty(np.broadcast(a, b).shape, dtype=np.result_type(a, b))
    ax, ay, az = a[..., 0], a[..., 1], a[..., 2]
    bx, by, bz = b[..., 0], b[..., 1], b[..., 2]
    out[..., 0] = ay * bz - az * by
    out[..., 1] = az * bx - ax * bz
    out[..., 2] = ax * by - ay * bx
    return out


def _norm3(a):
    """np.linalg.norm(a, axis=-1) for (..., 3): sqrt((x*x + y*y) + z*z), the order numpy's reduction uses (same bits)."""
    x, y, z = a[..., 0], a[..., 1], a[..., 2]
    return np.sqrt((x * x + y * y) + z * z)


def smooth_vertex_normals(vertices, faces, face_n):
    """Angle-weighted vertex normals (load.py:137-176): each face normal is weighted by the corner's interior angle."""
    corners = vertices[faces]                                  # (M, 3, 3)
    to_next, to_prev = np.empty_like(corners), np.empty_like(corners)
    for k in range(3):  # np.roll(corners, -1 / +1, axis=1) - corners, without the rolled copies
        np.subtract(corners[:, (k + 1) % 3], corners[:, k], out=to_next[:, k])
        np.subtract(corners[:, (k + 2) % 3], corners[:, k], out=to_prev[:, k])
    sin_part = _norm3(_cross(to_next, to_prev))
    cos_part = np.einsum("ijk,ijk->ij", to_next, to_prev)
    angle = np.arctan2(sin_part, cos_part)                     # (M, 3)
    weighted = face_n[:, None, :] * angle[..., None]
    if weighted.dtype == np.float64:
        # np.bincount adds its float64 weights in input order, exactly like np.add.at below, and is twice as fast
        idx, w = faces.ravel(), weighted.reshape(-1, 3)
        acc = np.stack([np.bincount(idx, weights=w[:, k], minlength=len(vertices)) for k in range(3)], axis=1)
        acc = acc.astype(vertices.dtype, copy=False)
    else:
        acc = np.zeros_like(vertices, dtype=vertices.dtype)
        np.add.at(acc, faces.ravel(), weighted.reshape(-1, 3))
    length = _norm3(acc)[:, None]
    np.divide(acc, length, out=acc, where=length > 0)
    return acc


def fast_load(vertices, faces, emitter=False, material=None):
    """Indexed mesh -> soup (load.py:98-134).  Note the smoothing uses the *un-normalised* face normals, as there."""
    corners = vertices[faces]
    raw_n = _cross(corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0])
    smooth = smooth_vertex_normals(vertices, faces, raw_n)[faces]
    norm = _norm3(raw_n)
    n = len(corners)
    return FastTreeBox(
        faces=faces,
        triangles=corners,
        mins=np.minimum(np.minimum(corners[:, 0], corners[:, 1]), corners[:, 2]),
        maxes=np.maximum(np.maximum(corners[:, 0], corners[:, 1]), corners[:, 2]),
        face_normals=raw_n / norm[:, None],
        smoothed_normals=smooth,
        surface_areas=norm / 2,
        material=np.zeros(n, np.int32) if material is None else np.full(n, material, np.int32),
        emitter=np.full(n, bool(emitter), np.bool_),
        camera=np.zeros(n, np.int32),
    )


def fast_load_obj(obj_path, offset=None, material=None, emitter=False, scale=1.0):
    vertices, faces = read_obj(obj_path)
    vertices = vertices * scale + (np.zeros(3) if offset is None else offset)
    return fast_load(vertices, faces, material=material, emitter=emitter)


def fast_load_ply(ply_path, offset=None, material=None, scale=1.0, emitter=False):
    vertices, faces = read_ply(ply_path)
    vertices = vertices * scale + (np.zeros(3) if offset is None else offset)
    return fast_load(vertices, faces, material=material, emitter=emitter)


# ---------------------------------------------------------------------------------------------------------------
# file formats (triangle meshes only)
# ---------------------------------------------------------------------------------------------------------------
def read_obj(path):
    """Wavefront OBJ -> (float64 vertices (N,3), int32 faces (M,3)), 1-based `v[/vt[/vn]]` indices, triangles only."""
    vs, fs = [], []
    with open(path, "r") as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                vs.append(tok[1:4])
            elif tok[0] == "f":
                if len(tok) != 4:
                    raise ValueError(f"{path}: only triangular faces are supported, got {len(tok) - 1} corners")
                face = [int(t.split("/", 1)[0]) for t in tok[1:]]
                # a negative index counts back from the vertices defined SO FAR (relative indexing), not from the end of the file
                fs.append([i + len(vs) if i < 0 else i - 1 for i in face])
    vertices = np.array(vs, dtype=np.float64).reshape(-1, 3)
    faces = np.array(fs, dtype=np.int64).reshape(-1, 3)
    if len(faces) and (faces.min() < 0 or faces.max() >= len(vertices)):
        raise ValueError(f"{path}: face index out of range")
    faces = faces.astype(np.int32)
    return vertices, faces


def write_obj(path, vertices, faces):
    with open(path, "w") as fh:
        fh.write("# clive2_b200 procedural mesh\n")
        np.savetxt(fh, vertices, fmt="v %.9g %.9g %.9g")
        np.savetxt(fh, np.asarray(faces) + 1, fmt="f %d %d %d")


_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def read_ply(path):
    """PLY (ascii or binary_little_endian) -> (float32 xyz vertices (N,3), faces (M,3)); triangle lists only.

    Vertices come back as float32 like `plyfile` data viewed at load.py:90-93 (scale/offset then promote them)."""
    with open(path, "rb") as fh:
        if fh.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = fh.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                elements[-1][2].append(tok[1:])
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        vertices = faces = None
        for name, count, props in elements:
            is_list = any(p[0] == "list" for p in props)
            if not is_list:
                dt = np.dtype([(p[1], "<" + _PLY_TYPES[p[0]]) for p in props])
                if fmt == "ascii":
                    raw = np.loadtxt(fh, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
                    rec = np.zeros(count, dtype=dt)
                    for k, p in enumerate(props):
                        rec[p[1]] = raw[:, k]
                else:
                    rec = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt)
                if name == "vertex":
                    vertices = np.stack([rec["x"], rec["y"], rec["z"]], axis=1).astype(np.float32)
            else:
                if len(props) != 1:
                    raise ValueError(f"{path}: element {name} mixes list and scalar properties")
                _, cnt_t, idx_t, _pname = props[0]
                if fmt == "ascii":
                    raw = np.loadtxt(fh, dtype=np.int64, max_rows=count, ndmin=2)
                    if count and (raw[:, 0] != 3).any():
                        raise ValueError(f"{path}: only triangular faces are supported")
                    lists = raw[:, 1:4]
                else:
                    dt = np.dtype([("n", "<" + _PLY_TYPES[cnt_t]), ("idx", "<" + _PLY_TYPES[idx_t], 3)])
                    rec = np.frombuffer(fh.read(dt.itemsize * count), dtype=dt)
                    if count and (rec["n"] != 3).any():
                        raise ValueError(f"{path}: only triangular faces are supported")
                    lists = rec["idx"]
                if name == "face":
                    faces = np.ascontiguousarray(lists)
        if vertices is None or faces is None:
            raise ValueError(f"{path}: needs vertex and face elements")
        return vertices, faces


def write_ply(path, vertices, faces):
    """binary_little_endian PLY with float32 xyz and `uchar int` triangle lists."""
    v = np.ascontiguousarray(vertices, dtype="<f4")
    rec = np.zeros(len(faces), dtype=np.dtype([("n", "u1"), ("idx", "<i4", 3)]))
    rec["n"] = 3
    rec["idx"] = faces
    header = (f"ply\nformat binary_little_endian 1.0\ncomment clive2_b200 procedural mesh\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(rec)}\nproperty list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as fh:
        fh.write(header.encode("ascii"))
        fh.write(v.tobytes())
        fh.write(rec.tobytes())
