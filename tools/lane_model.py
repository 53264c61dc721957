#!/usr/bin/env python
"""A MODEL (host emulation, no GPU) of what limits the lanes of k_fused_walks: per-ray traversal work of every bounce
of a frame, grouped into warps the way the kernel groups them, against groupings a wavefront walk could choose.

    python tools/lane_model.py [--preset bench-spheres] [--width 128] [--height 128]

For each bounce the 32 rays of a warp cost the warp max(ray cost) (the kernel is bounce-synchronous); utilisation =
mean / max.  Cost of a ray = nodes * C_NODE + triangles * C_TRI (SASS instruction counts of the two loop bodies).
Groupings: (a) sample order = what the kernel does; (b) rays of a bounce sorted by origin cell and direction octant =
what a wavefront walk with a cheap sort could do; (c) sorted by the cost itself = the bound no ordering can beat.
This says how much of the measured 11 of 32 lanes is imbalance BETWEEN rays (orderable) and how much is divergence
inside the steps (not orderable).  Numbers are a planning aid, not a measurement.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clive2_b200 import scene, struct_types as T  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.common import StageBuffers, oracle_stages, seeds_for  # noqa: E402
from tests.hostemu import HostEmu  # noqa: E402

C_NODE, C_TRI = 92.0, 45.0  # SASS instructions of one inner-node step / one triangle test in k_fused_walks


def util(cost, order):
    c = cost[order]
    c = c[: len(c) // 32 * 32].reshape(-1, 32)
    return float(c.mean() / c.max(axis=1).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="bench-spheres")
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--height", type=int, default=128)
    args = ap.parse_args()
    w, h = args.width, args.height
    a = scene.preset_arrays(args.preset, w, h)
    O, E = Oracle(shared_math=True), HostEmu()
    R = oracle_stages(O, a, StageBuffers(w * h, seeds_for(w * h)), upto="paths")
    acc = E.accel(a.boxes, a.triangles)
    lo, hi = a.boxes["min"][0, :3], a.boxes["max"][0, :3]
    print(f"{args.preset} {w}x{h}: {len(a.triangles)} triangles")
    print("walk   bounce  rays   nodes/ray tris/ray | utilisation: sample order | cell+octant sort | cost sort (bound)")
    total = {"sample": [0.0, 0.0], "sorted": [0.0, 0.0], "bound": [0.0, 0.0]}
    for name, paths in (("light", R.light_paths), ("camera", R.camera_paths)):
        for k in range(T.MAX_BOUNCES):
            # the ray traced at bounce k leaves stored vertex k; a walk of length L traced min(L + 1, 6) rays, the last
            # of them from the vertex that was not kept — only the kept ones are available here (a lower bound on work)
            live = np.flatnonzero(paths["length"] > k)
            if len(live) < 64:
                continue
            rays = np.ascontiguousarray(paths["rays"][live, k])
            nodes, tris = E.traverse_counts(rays, acc)
            cost = nodes * C_NODE + tris * C_TRI
            o, d = rays["origin"][:, :3], rays["direction"][:, :3]
            cell = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * 8).astype(np.int64), 0, 7)
            octant = (d[:, 0] > 0) * 4 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0)
            key = ((cell[:, 0] * 8 + cell[:, 1]) * 8 + cell[:, 2]) * 8 + octant
            orders = {"sample": np.arange(len(live)), "sorted": np.argsort(key, kind="stable"), "bound": np.argsort(cost)}
            us = {n: util(cost, od) for n, od in orders.items()}
            for n in us:
                total[n][0] += cost.sum()
                total[n][1] += cost.sum() / us[n]
            print(f"{name:6s} {k:6d} {len(live):6d} {nodes.mean():9.1f} {tris.mean():8.1f} |"
                  f" {us['sample']:24.2f} | {us['sorted']:16.2f} | {us['bound']:17.2f}")
    E.accel_destroy(acc)
    print("all bounces, cost weighted:                      |"
          + " | ".join(f" {total[n][0] / total[n][1]:.2f} ({n})" for n in ("sample", "sorted", "bound")))


if __name__ == "__main__":
    main()
