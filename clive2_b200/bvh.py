"""Host BVH: triangle soup container, greedy SAH object-split build and the breadth-first flat layout.

This is the readable numpy specification of the tree the reference emits (src/bvh.py:132-191 object_split,
:288-313 construct_BVH, :329-389 np_flatten_bvh).  It is index based (one permutation per node instead of copying
ten attribute arrays per split) and is pinned against arrays produced by the reference module itself
(tests/golden/bvh_*.npz).  `clive2_b200.native_bvh` is the fast C++ builder that must agree with this one bit for bit.

Reference behaviours that shape the output and are therefore kept:
  * centroid = (min + max) / 2 in the soup's own dtype (float32 for the Cornell-only scene, float64 once a mesh joins);
  * prefix/suffix box areas are formed in the soup dtype, then the SAH products with the integer counts are float64;
  * the left-count factor is `arange(n-1)` (0-based, i.e. one less than the true left size) (bvh.py:150-154);
  * first minimum wins inside an axis, strict improvement across axes x->y->z;
  * a node becomes a leaf when it has <= MAX_MEMBERS triangles or when the *work stack* holds more than MAX_DEPTH
    entries after it was popped (bvh.py:294);
  * children are visited left first while building, the flat array is breadth first, an inner node stores the index of
    its left child in `left` (right child is `left + 1`) and 0 in `right`; a leaf stores its triangle range
    [left, right) and triangles are emitted leaf by leaf in breadth-first leaf order.
The reference calls `np.argsort` with the default (unstable) kind, so tied centroids (every pair of coplanar wall
triangles, neighbouring grid triangles that share their extreme vertices) are ordered by whatever the platform's
SIMD quicksort does; a stable sort is the default here and `sort_kind=None` reproduces the reference on the same
platform.  On tie-free input all variants agree exactly.
"""
from collections import deque

import subprocess

import numpy as np

from .constants import MAX_DEPTH, MAX_MEMBERS
from .struct_types import Box, Triangle


class FastTreeBox:
    """Struct-of-arrays triangle soup (the reference's name is kept: scene code adds soups with `+`)."""

    FIELDS = ("faces", "triangles", "mins", "maxes", "face_normals", "smoothed_normals", "surface_areas",
              "material", "emitter", "camera")

    def __init__(self, faces, triangles, mins, maxes, face_normals, smoothed_normals, surface_areas, material,
                 emitter, camera):
        self.faces = faces
        self.triangles = triangles            # (n, 3 corners, 3 coords)
        self.mins = mins
        self.maxes = maxes
        self.face_normals = face_normals
        self.smoothed_normals = smoothed_normals  # (n, 3 corners, 3 coords)
        self.surface_areas = surface_areas
        self.material = material
        self.emitter = emitter
        self.camera = camera
        n = len(mins)
        self.min = mins.min(axis=0) if n else np.full(3, np.inf)
        self.max = maxes.max(axis=0) if n else np.full(3, -np.inf)
        # tree links, filled by construct_BVH
        self.left = None
        self.right = None
        self.parent = None

    def __len__(self):
        return len(self.mins)

    @classmethod
    def empty_box(cls):
        f32 = np.float32
        return cls(np.empty((0, 3), np.uint32), np.empty((0, 3, 3), f32), np.empty((0, 3), f32), np.empty((0, 3), f32),
                   np.empty((0, 3), f32), np.empty((0, 3, 3), f32), np.empty((0,), f32), np.empty((0,), np.int32),
                   np.empty((0,), np.bool_), np.empty((0,), np.bool_))

    @classmethod
    def from_triangle_objects(cls, tris):
        """Soup from `load.Triangle` objects; everything is narrowed to float32 here (bvh.py:54-83)."""
        f32 = np.float32
        corners = np.array([[t.v0, t.v1, t.v2] for t in tris], dtype=f32)
        normals = np.array([t.n for t in tris], dtype=f32)
        return cls(
            faces=np.zeros((len(tris), 3), np.uint32),
            triangles=corners,
            mins=corners.min(axis=1),
            maxes=corners.max(axis=1),
            face_normals=normals,
            smoothed_normals=np.repeat(normals[:, None, :], 3, axis=1),
            surface_areas=np.array([t.surface_area for t in tris], dtype=f32),
            material=np.array([t.material for t in tris], dtype=np.int32),
            emitter=np.array([t.emitter for t in tris], dtype=np.bool_),
            camera=np.array([t.camera for t in tris], dtype=np.bool_),
        )

    def __add__(self, other):
        if not isinstance(other, FastTreeBox):
            raise TypeError("Can only add another FastTreeBox")
        return FastTreeBox(*(np.concatenate((getattr(self, f), getattr(other, f)), axis=0) for f in self.FIELDS))

    def take(self, idx):
        return FastTreeBox(*(getattr(self, f)[idx] for f in self.FIELDS))


def _half_area_x2(lo, hi):
    s = hi - lo
    return 2 * (s[:, 0] * s[:, 1] + s[:, 1] * s[:, 2] + s[:, 2] * s[:, 0])


def sah_object_split(mins, maxes, centers, idx, sort_kind="stable"):
    """Best SAH object split of the triangles `idx` (in their current order).

    Returns (cost, ordered_idx, n_left).  Mirrors bvh.py:132-160 including the 0-based left count.
    `sort_kind=None` uses numpy's default unstable argsort exactly like the reference (same tie order on the same
    platform); "stable" is the portable definition that the native builder reproduces.
    """
    n = len(idx)
    counts = np.arange(n - 1)
    best = (np.inf, None, 0)
    for axis in range(3):
        order = idx[np.argsort(centers[idx, axis], kind=sort_kind)]
        hi_s, lo_s = maxes[order], mins[order]
        left_area = _half_area_x2(np.minimum.accumulate(lo_s), np.maximum.accumulate(hi_s))[:-1]
        right_area = _half_area_x2(np.minimum.accumulate(lo_s[::-1])[::-1], np.maximum.accumulate(hi_s[::-1])[::-1])[1:]
        cost = left_area * counts + right_area * ((n - 1) - counts)
        k = int(np.argmin(cost))
        if cost[k] < best[0]:
            best = (cost[k], order, k + 1)
    return best


class _Node:
    __slots__ = ("idx", "left", "right", "lo", "hi")

    def __init__(self, idx, lo, hi):
        self.idx, self.lo, self.hi = idx, lo, hi
        self.left = self.right = None


class BVHTree:
    """Result of `construct_BVH`: the soup plus the linked node tree."""

    def __init__(self, soup, root, max_stack):
        self.soup, self.root, self.max_stack = soup, root, max_stack


def construct_BVH(soup, verbose=False, sort_kind="stable"):
    mins, maxes = soup.mins, soup.maxes
    centers = (mins + maxes) / 2
    all_idx = np.arange(len(soup))
    root = _Node(all_idx, soup.min, soup.max)
    stack, max_stack = [root], 0
    while stack:
        node = stack.pop()
        max_stack = max(max_stack, len(stack))
        if len(node.idx) <= MAX_MEMBERS or len(stack) > MAX_DEPTH:
            continue
        _, order, n_left = sah_object_split(mins, maxes, centers, node.idx, sort_kind)
        li, ri = order[:n_left], order[n_left:]
        node.left = _Node(li, mins[li].min(axis=0), maxes[li].max(axis=0))
        node.right = _Node(ri, mins[ri].min(axis=0), maxes[ri].max(axis=0))
        stack.append(node.right)
        stack.append(node.left)
    if verbose:
        print(f"max BVH depth: {max_stack}")
    return BVHTree(soup, root, max_stack)


def count_boxes(tree):
    n, todo = 0, [tree.root]
    while todo:
        node = todo.pop()
        n += 1
        if node.left is not None:
            todo.extend((node.right, node.left))
    return n


def emit_triangles(soup, order):
    """`struct_types.Triangle` records for soup members `order` (narrowing to float32 like bvh.py:361-370)."""
    out = np.zeros(len(order), dtype=Triangle)
    corners, smooth = soup.triangles[order], soup.smoothed_normals[order]
    for k, name in enumerate(("v0", "v1", "v2")):
        out[name][:, :3] = corners[:, k, :]
    for k, name in enumerate(("n0", "n1", "n2")):
        out[name][:, :3] = smooth[:, k, :]
    out["normal"][:, :3] = soup.face_normals[order]
    out["material"] = soup.material[order]
    out["is_light"] = soup.emitter[order]
    out["is_camera"] = soup.camera[order]
    return out


def np_flatten_bvh(tree, verbose=False):
    """Breadth-first flat arrays `(Box[], Triangle[])` in the reference encoding (bvh.py:329-389)."""
    n_boxes = count_boxes(tree)
    boxes = np.zeros(n_boxes, dtype=Box)
    leaf_members = []
    queue = deque([tree.root])
    bi = ti = 0
    while queue:
        node = queue.popleft()
        boxes[bi]["min"][:3] = node.lo
        boxes[bi]["max"][:3] = node.hi
        if node.left is not None:
            boxes[bi]["left"] = bi + len(queue) + 1
            boxes[bi]["right"] = 0
            queue.append(node.left)
            queue.append(node.right)
        else:
            boxes[bi]["left"] = ti
            ti += len(node.idx)
            boxes[bi]["right"] = ti
            leaf_members.append(node.idx)
        bi += 1
    order = np.concatenate(leaf_members) if leaf_members else np.zeros(0, np.int64)
    assert bi == n_boxes and ti == len(tree.soup) == len(order)
    if verbose:
        print("flattened", n_boxes, "boxes and", len(order), "triangles.")
    return boxes, emit_triangles(tree.soup, order)


def build_flat_bvh(soup, builder="auto", sort_kind="stable"):
    """Soup -> `(Box[], Triangle[])`.  `builder`: "numpy" (this module), "native" (C++), "auto" (native when built).
    The native builder implements the stable tie order only."""
    if sort_kind != "stable":
        builder = "numpy"
    if builder in ("auto", "native"):
        try:
            from . import native_bvh
            return native_bvh.build(soup)
        except (ImportError, OSError, subprocess.CalledProcessError):
            if builder == "native":
                raise
    return np_flatten_bvh(construct_BVH(soup, sort_kind=sort_kind))
