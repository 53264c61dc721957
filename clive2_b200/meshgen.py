"""Procedural meshes standing in for the assets the reference's presets point at but does not ship.

The presets load `../resources/teapot.obj` and the Stanford dragon PLYs (scene.py:154-199); none of them is in the
repository and there is no network, so the benchmark configurations use generated `(vertices, faces)` pairs that go
through the same `load.fast_load` entry point.
"""
import numpy as np


def icosphere(subdivisions=5, radius=1.0):
    """Unit icosahedron subdivided `subdivisions` times: 20 * 4**s faces (s=5 -> 20 480), outward winding."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        # one new vertex per undirected edge, shared by the two faces that use it
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        uniq, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = inv.reshape(3, -1) + len(v)          # rows: midpoints of edges 01, 12, 20 of every face
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        ab, bc, ca = m[0], m[1], m[2]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1),
                            np.stack([ab, bc, ca], 1)])
        v = np.concatenate([v, mid])
    return v * radius, f.astype(np.int32)


def terrain(quads_x, quads_z, extent=5.0, base_y=-1.5, relief=0.6, jitter=0.01, seed=0):
    """Height-field over x,z in [-extent, extent] with 2 * quads_x * quads_z triangles facing +y.

    Heights are a sum of a few sinusoids; every vertex coordinate gets N(0, jitter) noise from `default_rng(seed)` so
    that no two triangle centroids tie (the reference's argsort is unstable on ties).
    708 x 708 quads -> 1 002 528 triangles (config 3); 2236 x 2236 -> 9 999 392 (config 4).
    """
    xs = np.linspace(-extent, extent, quads_x + 1)
    zs = np.linspace(-extent, extent, quads_z + 1)
    gx, gz = np.meshgrid(xs, zs, indexing="xy")
    h = (np.sin(gx * 1.3) * np.cos(gz * 0.9) + 0.5 * np.sin(gx * 3.1 + 1.0) * np.sin(gz * 2.7)
         + 0.25 * np.cos(gx * 6.3 - gz * 5.1))
    gy = base_y + relief * h
    v = np.stack([gx, gy, gz], axis=-1).reshape(-1, 3)
    if jitter:
        v = v + np.random.default_rng(seed).normal(0.0, jitter, size=v.shape)
    row = quads_x + 1
    i = (np.arange(quads_z)[:, None] * row + np.arange(quads_x)[None, :]).ravel()
    # counter-clockwise seen from above (+y normal): (i, i+row, i+1), (i+1, i+row, i+row+1)
    f = np.concatenate([np.stack([i, i + row, i + 1], 1), np.stack([i + 1, i + row, i + row + 1], 1)])
    return v, f.astype(np.int32)
