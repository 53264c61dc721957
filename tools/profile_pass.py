#!/usr/bin/env python
"""Runs N plain fused passes of a bench configuration (no instrumentation, no timing) — the command to wrap in ncu.

    python tools/profile_pass.py [--config c2] [--passes 6] [--mode ordered|reference]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from clive2_b200 import runtime, scene  # noqa: E402
from clive2_b200.renderer import Renderer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--passes", type=int, default=6)
ap.add_argument("--mode", default="ordered")
ap.add_argument("--mesh-in-memory", action="store_true")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
arrays = bench.build_arrays(cfg, via_file=not args.mesh_in_memory)
dev = runtime.Device(0)
sc = scene.scene_from_arrays(arrays, dev)
r = Renderer(sc, seeds=bench.seeds_for(cfg["width"] * cfg["height"], 0),
             traverse_mode=runtime.TRAVERSE_ORDERED if args.mode == "ordered" else runtime.TRAVERSE_REFERENCE_ORDER)
for _ in range(args.passes):
    r.run_sample()
dev.synchronize()
print("ok", float(r.estimate().mean()))
