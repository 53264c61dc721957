#!/usr/bin/env python
"""Lane utilisation of closest-hit traversal under different warp schedules, replayed on the real rays of a frame
(PLANNING AID, host only; see warp_sim.cpp).

    python tools/warp_sim/run.py [--preset bench-terrain-1m] [--width 96] [--height 96]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from clive2_b200 import scene, struct_types as T  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.common import StageBuffers, oracle_stages, seeds_for  # noqa: E402

LIB = os.path.join(HERE, "libwarpsim.so")
SRC = os.path.join(HERE, "warp_sim.cpp")
if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC], check=True)
lib = C.CDLL(LIB)
lib.sim_accel_create.restype = C.c_void_p
lib.sim_trace_rays.restype = C.c_void_p


class Costs(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("c_inner", "c_tri", "c_leaf_setup", "c_iter", "c_fetch", "c_pop")]


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="bench-terrain-1m")
    ap.add_argument("--width", type=int, default=96)
    ap.add_argument("--height", type=int, default=96)
    args = ap.parse_args()
    w, h = args.width, args.height
    a = scene.preset_arrays(args.preset, w, h)
    O = Oracle(shared_math=True)
    R = oracle_stages(O, a, StageBuffers(w * h, seeds_for(w * h)), upto="paths")
    boxes, tris = np.ascontiguousarray(a.boxes), np.ascontiguousarray(a.triangles)
    acc = C.c_void_p(lib.sim_accel_create(p(boxes), C.c_int32(len(boxes)), p(tris), C.c_int32(len(tris))))
    # all rays of both walks, grouped per bounce (a wavefront round traces the rays of one bounce of every live path)
    rounds = []
    for paths in (R.light_paths, R.camera_paths):
        for k in range(T.MAX_BOUNCES):
            live = np.flatnonzero(paths["length"] > k)
            if len(live):
                rounds.append(np.ascontiguousarray(paths["rays"][live, k]))
    rays = np.concatenate(rounds)
    bad = C.c_int()
    traces = C.c_void_p(lib.sim_trace_rays(acc, p(rays), C.c_uint32(len(rays)), C.byref(bad)))
    st = (C.c_double * 4)()
    lib.sim_trace_stats(traces, st)
    print(f"{args.preset} {w}x{h}: {len(tris)} triangles, {len(rays)} walk rays; per ray {st[0]:.1f} inner steps, "
          f"{st[1]:.1f} leaves, {st[2]:.1f} triangles; trace/closest_hit mismatches {bad.value}")
    cost = Costs(92.0, 45.0, 8.0, 6.0, 40.0, 0.0)
    n = len(rays)
    # queue orders: (a) megakernel = per bounce, 32 consecutive samples per warp; (b) wavefront = rounds concatenated
    order = np.arange(n, dtype=np.uint32)

    def run(schedule, refill=16, batch=0, weight=2.0, chunk=512, costs=cost):
        out = (C.c_double * 4)()
        lib.sim_run(traces, p(order), C.c_uint32(n), schedule, refill, batch, C.c_double(weight), C.byref(costs), chunk, out)
        return out[0], out[3]

    base_u, base_slots = run(0)
    print(f"{'schedule':58s} util   issue slots vs sync")
    rows = [("sync while-while, no refill (k_fused_walks today)", dict(schedule=0))]
    for r in (16, 24, 32):
        rows.append((f"refill<{r}, while-while rounds", dict(schedule=1, refill=r)))
    for r in (16, 32):
        for wgt in (1.0, 2.0):
            rows.append((f"refill<{r}, vote whole leaves, weight {wgt}", dict(schedule=2, refill=r, weight=wgt)))
    for r in (24, 32):
        for b in (1, 2, 3, 4, 0):
            rows.append((f"refill<{r}, inner block + <= {b or 'all'} tris per iteration", dict(schedule=3, refill=r, batch=b)))
    for r in (32,):
        for b in (1, 2, 4):
            for wgt in (1.0, 2.0):
                rows.append((f"refill<{r}, vote inner vs <= {b} tris, weight {wgt}", dict(schedule=4, refill=r, batch=b, weight=wgt)))
    for coop in (25, 40):
        for wgt in (1.0, 2.0):
            rows.append((f"refill<32, vote inner vs COOPERATIVE leaf (+{coop}), w {wgt}", dict(schedule=5, refill=32, batch=coop, weight=wgt)))
    for name, kw in rows:
        u, slots = run(**kw)
        print(f"{name:58s} {u:.3f}  {slots / base_slots:.3f}")
    # queue sorted inside every round by origin cell (8^3) + direction octant: does coherence help the vote schedule?
    lo, hi = boxes["min"][0, :3], boxes["max"][0, :3]
    o, d = rays["origin"][:, :3], rays["direction"][:, :3]
    cell = np.clip(((o - lo) / np.maximum(hi - lo, 1e-9) * 8).astype(np.int64), 0, 7)
    octant = (d[:, 0] > 0) * 4 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0)
    round_id = np.concatenate([np.full(len(r), i) for i, r in enumerate(rounds)])
    key = (((round_id * 8 + cell[:, 0]) * 8 + cell[:, 1]) * 8 + cell[:, 2]) * 8 + octant
    order = np.argsort(key, kind="stable").astype(np.uint32)
    for name, kw in (("SORTED queue: refill<32, vote whole leaves, weight 1.0", dict(schedule=2, refill=32, weight=1.0)),
                     ("SORTED queue: sync while-while", dict(schedule=0))):
        u, slots = run(**kw)
        print(f"{name:58s} {u:.3f}  {slots / base_slots:.3f}")
    # K rays per lane (costs as measured in SASS: inner step 70, triangle 85 instructions)
    real = Costs(70.0, 85.0, 8.0, 35.0, 60.0, 0.0)
    u1, s1 = run(schedule=2, refill=24, weight=2.0, costs=real)
    print(f"SASS costs, 1 ray per lane, refill<24, vote: util {u1:.3f}")
    for K, extra in ((2, 0.0), (2, 30.0), (2, 55.0), (3, 30.0), (4, 30.0)):
        out = (C.c_double * 4)()
        lib.sim_run_multi(traces, p(order), C.c_uint32(n), K, int(24 * K), C.c_double(2.0), C.byref(real), C.c_double(extra), 512 * K, out)
        print(f"{K} rays per lane, +{extra:.0f} slots per iteration for picking the ray: util {out[0]:.3f}, issue slots vs 1 ray per lane {out[3] / s1:.3f}")
    # leaves capped by chaining (synthetic same-box nodes): does a more uniform leaf length pay for the extra inner steps?
    lib.sim_traces_split_leaves.restype = C.c_void_p
    order = np.arange(n, dtype=np.uint32)
    keep = traces
    for cap in (15, 8, 6, 4, 3, 2):
        traces = C.c_void_p(lib.sim_traces_split_leaves(keep, cap))
        u, slots = run(schedule=2, refill=24, weight=2.0)
        u1, slots1 = run(schedule=2, refill=24, weight=1.0)
        print(f"leaves <= {cap:2d}: refill<24 vote, weight 2: issue slots vs sync {slots / base_slots:.3f}   weight 1: {slots1 / base_slots:.3f}")
        lib.sim_traces_destroy(traces)
    traces = keep
    lib.sim_traces_destroy(traces)
    lib.sim_accel_destroy(acc)


if __name__ == "__main__":
    main()
