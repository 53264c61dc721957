#!/usr/bin/env python
"""One scene build, many context settings: ms per fused pass for each `VAR=val,VAR2=val` setting (context-level knobs
are read from the environment when a Device is created).

    python tools/sweep_env.py --config c4 --passes 5 "" CLV_PASS_SLICES=1 CLV_REFILL_BELOW=16 CLV_RAY_SCHEDULE=4,CLV_REFILL_BELOW=32
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from clive2_b200 import runtime, scene  # noqa: E402
from clive2_b200.renderer import Renderer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--passes", type=int, default=10)
ap.add_argument("settings", nargs="*", default=[""])
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
t0 = time.time()
arrays = bench.build_arrays(cfg)
print(f"scene built in {time.time() - t0:.1f} s", flush=True)
n = cfg["width"] * cfg["height"]
for setting in args.settings:
    pairs = [kv.split("=", 1) for kv in setting.split(",") if kv]
    for k, v in pairs:
        os.environ[k] = v
    dev = runtime.Device(0)
    sc = scene.scene_from_arrays(arrays, dev)
    r = Renderer(sc, seeds=bench.seeds_for(n, 0))
    for _ in range(3):
        r.run_sample()
    dev.synchronize()
    t = time.perf_counter()
    for _ in range(args.passes):
        r.run_sample()
    dev.synchronize()
    ms = (time.perf_counter() - t) * 1e3 / args.passes
    r._fr.kernel_timing(True)
    for _ in range(args.passes):
        r.run_sample()
    km, _ = r._fr.kernel_times(reset=True)
    print(f"{setting or 'default':50s} {ms:8.3f} ms/pass  {n / ms / 1e3:7.1f} Msamples/s  one-slice kernels: "
          + " ".join(f"{k[2:]}={v / args.passes:.3f}" for k, v in km.items()), flush=True)
    r.release()
    sc.release()
    dev.release()
    for k, _ in pairs:
        del os.environ[k]
