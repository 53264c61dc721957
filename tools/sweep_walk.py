#!/usr/bin/env python
"""Times the fused pass of one configuration under several walk / scheduling settings with ONE scene build.

    python tools/sweep_walk.py --config c4 [--passes 8] [--settings "mode=1;mode=2;mode=2,walk_refill_below=32"]

Per setting: ms per pass over `passes` passes (CUDA events on the launch stream, slices as in production) and the
per-kernel times of `passes` more passes run as one slice.  Results are identical under every setting (checked: RNG
state after the run).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from clive2_b200 import runtime, scene  # noqa: E402
from clive2_b200 import struct_types as T  # noqa: E402
from clive2_b200.renderer import Renderer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4")
ap.add_argument("--passes", type=int, default=8)
ap.add_argument("--settings", default="mode=1;mode=2")
ap.add_argument("--mesh-in-memory", action="store_true")
ap.add_argument("--tile-shard", type=int, default=0, help="time rank 0's shard of a W-way tile sharding (16x16 tiles) instead of the full frame")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
arrays = bench.build_arrays(cfg, via_file=not args.mesh_in_memory)
batch = cfg["width"] * cfg["height"]
seeds = bench.seeds_for(batch, 0)
dev = runtime.Device(0)
sc = scene.scene_from_arrays(arrays, dev)
import torch  # noqa: E402
stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", 0))
ref_state = None
ref_aggs = None
for setting in args.settings.split(";"):
    kv = dict(x.split("=") for x in setting.split(",") if x)
    dev.walk_mode = int(kv.pop("mode", 0))
    slices = int(kv.pop("slices", 0))
    dev.pass_slices = slices
    for k, v in kv.items():
        dev.knob(k, int(v))
    ids = None
    if args.tile_shard > 1:
        from clive2_b200 import sharding
        ids = sharding.tile_ids(cfg["width"], cfg["height"], 0, args.tile_shard, 16)
    r = Renderer(sc, seeds=seeds, sample_ids=ids)
    for _ in range(3):
        r.run_sample()
    dev.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.passes):
        r.run_sample()
    e1.record(stream)
    dev.synchronize()
    ms = e0.elapsed_time(e1) / args.passes
    r._fr.kernel_timing(True)
    for _ in range(args.passes):
        r.run_sample()
    km, n = r._fr.kernel_times(reset=True)
    r._fr.kernel_timing(False)
    n_local = batch if ids is None else len(ids)
    state = r._fr.read(runtime.BUF_SEEDS, np.uint32, 2 * n_local)
    same = True if ref_state is None else bool(np.array_equal(ref_state, state))
    ref_state = state if ref_state is None else ref_state
    # the last pass's aggregators (every connection weight and colour of every pixel) are order independent, too
    aggs = r._fr.read(runtime.BUF_AGGREGATORS, np.uint8, batch * T.WeightAggregator.itemsize) if ids is None else None
    same_aggs = None if aggs is None else (True if ref_aggs is None else bool(np.array_equal(ref_aggs, aggs)))
    ref_aggs = aggs if ref_aggs is None else ref_aggs
    print(json.dumps({"config": args.config, "setting": setting, "ms_per_pass": ms, "msamples_per_s": n_local / ms / 1e3, "samples_per_pass": n_local,
                      "kernel_ms": {k: round(v / max(n, 1), 4) for k, v in km.items()}, "same_rng_state_as_first": same, "same_aggregators_as_first": same_aggs}), flush=True)
    r.release()
