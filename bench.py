#!/usr/bin/env python
"""Benchmark of the BDPT hot path.  One "step" = one sample pass (Renderer.run_sample) over the FULL frame of the
configuration; `value` = samples of all steps / device time, whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4] [--impl ours|reference]

Default configuration: **c4** — BASELINE.json's north-star workload (1920x1080, ~10 M-triangle PLY mesh), the largest
configuration that fits one GPU.  c1-c3 are the other BASELINE configs (parity-test cases; selectable for profiling).

Prints ONE JSON line (rank 0):
  value            Msamples/s, scene + accumulators resident in HBM, CUDA events on the launch stream, max over ranks
  e2e              the same through the public call with pinned HOST buffers every step (seeds up, preview image down)
  roofline         dominant kernel of the pass: algorithmic bytes (device counters of an instrumented run of the same
                   passes) / its CUDA-event time, against the HBM peak of MEASURED_PEAKS.json when the traversal
                   records exceed the L2, against the L2 read peak measured in this run when they are L2 resident;
                   `traffic` = DRAM bytes of that kernel from ncu (profiles/dram_traffic.json, regenerated per round)
  rel_rmse         relative RMSE of the GPU image against the CPU reference (glibc build) at equal spp and against a
                   high-spp CPU image, on a reduced frame of the same scene (stated)
  cpu_baseline     the CPU reference on this box's host cores: oracle/_ref (the reference's own trace.metal compiled as
                   C++, kind "reference") when it was built, else oracle/trace_ref.c (kind "port")
`--impl reference` times that CPU code as the reference arm (same config / metric / unit).

N > 1 (torchrun, one rank per GPU): STRONG scaling of one frame.  The K passes of the frame are split over the GPUs by
sample pass and by image tile (clive2_b200.sharding.frame_plan: every GPU renders K // N full-frame passes from its own
RNG streams, the K % N passes left over are shared out as pass groups x 16x16-tile groups), replicated scene, and ONE NCCL
reduce of the packed float accumulators closes the timed region.
Extra keys carry the pure tile-sharded and the weak pass-sharded figures measured in the same run.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

L2_BYTES = 126 << 20
CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1": dict(preset="empty", width=256, height=256, mesh=None,
               workload="C1: Cornell box (preset 'empty', 16 triangles), 256x256, diffuse, max path length 6+6"),
    "c2": dict(preset="bench-spheres", width=512, height=512, mesh=None,
               workload="C2: Cornell box + two icospheres (2 x 20480 triangles), rough-dielectric (GGX transmission) "
                        "and coated-diffuse (GGX specular) materials, 512x512"),
    "c3": dict(preset="bench-terrain-1m", width=1024, height=1024, mesh=("obj", 708),
               workload="C3: Cornell enclosure + procedural terrain mesh (1 002 528 triangles, OBJ), 1024x1024"),
    # BASELINE.json configs[3]: the north-star configuration, quoted at N = 1, 2, 4, 8
    "c4": dict(preset="bench-terrain-10m", width=1920, height=1080, mesh=("ply", 2236),
               workload="C4: Cornell enclosure + procedural terrain mesh (9 999 392 triangles, PLY), 1920x1080"),
}
SEED = 20240928
RMSE_FRAME = {"c1": (128, 128), "c2": (128, 128), "c3": (128, 128), "c4": (160, 90)}  # reduced frames of rel_rmse
RMSE_SPP, RMSE_HI_SPP = 16, 128


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed regions run."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower() == "active":
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # median over the samples taken under load (the sampler also sees the idle gaps between timed regions)
            busy = [x for x in sm if x >= 0.6 * max(sm)]
            out.update(sm_mhz=statistics.median(busy or sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# scenes
# ---------------------------------------------------------------------------------------------------------------------
def build_arrays(cfg, width=None, height=None, via_file=True, bvh_builder="auto"):
    """`_build_arrays`, optionally memoised on /dev/shm across processes (CLV_SCENE_CACHE=1: measurement sweeps that start
    one process per build variant; never set by the driver's runs)."""
    if os.environ.get("CLV_SCENE_CACHE") != "1" or not os.path.isdir("/dev/shm"):
        return _build_arrays(cfg, width, height, via_file, bvh_builder)
    from clive2_b200.scene import SceneArrays
    w, h = width or cfg["width"], height or cfg["height"]
    base = f"/dev/shm/clv_scene_cache_{cfg['preset']}_{w}x{h}_{int(bool(via_file))}_"
    if all(os.path.exists(base + f + ".npy") for f in ARRAY_FIELDS):
        return SceneArrays(pixel_width=w, pixel_height=h, **{f: np.load(base + f + ".npy") for f in ARRAY_FIELDS})
    a = _build_arrays(cfg, width, height, via_file, bvh_builder)
    for f in ARRAY_FIELDS:
        np.save(base + f + ".tmp.npy", getattr(a, f))
        os.replace(base + f + ".tmp.npy", base + f + ".npy")
    return a


def _build_arrays(cfg, width=None, height=None, via_file=True, bvh_builder="auto"):
    """Scene arrays of a configuration.  C3 / C4 go through the file path BASELINE names (load.py:76-95 of the reference):
    the procedural mesh is written as OBJ / PLY and read back by the package's own readers (`via_file=False`: the
    in-memory mesh of the scene preset)."""
    from clive2_b200 import load, meshgen, scene
    w, h = width or cfg["width"], height or cfg["height"]
    if not (via_file and cfg.get("mesh")):
        return scene.preset_arrays(cfg["preset"], w, h, bvh_builder=bvh_builder)
    kind, quads = cfg["mesh"]
    p = scene.scene_presets[cfg["preset"]]
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(tmp, f"clv_bench_terrain_{quads}_{os.getuid()}.{kind}")
    if not os.path.exists(path):
        v, f = meshgen.terrain(quads, quads)
        part = path + f".{os.getpid()}.tmp." + kind
        (load.write_ply if kind == "ply" else load.write_obj)(part, v, f)
        os.replace(part, path)
    return scene.build_scene_arrays(w, h, p["cam_center"], p["cam_direction"],
                                    [{"file_path": path, "material": p["file_specs"][0]["material"]}], None, bvh_builder)


ARRAY_FIELDS = ("camera", "boxes", "triangles", "materials", "light_triangles", "light_counts", "light_surface_areas",
                "light_triangle_indices", "camera_triangle_indices")


def shared_arrays(cfg, rank, world, via_file):
    """Rank 0 builds the scene (mesh, normals, BVH: ~40 s at C4) and hands the arrays to the other ranks through
    /dev/shm instead of eight ranks building the same 10 M-triangle BVH on the same host cores."""
    if world == 1:
        return build_arrays(cfg, via_file=via_file)
    import torch.distributed as dist
    from clive2_b200.scene import SceneArrays
    base = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                        f"clv_bench_scene_{os.environ.get('MASTER_PORT', '0')}_")
    if rank == 0:
        a = build_arrays(cfg, via_file=via_file)
        for f in ARRAY_FIELDS:
            np.save(base + f + ".npy", getattr(a, f))
    dist.barrier()
    if rank != 0:
        a = SceneArrays(pixel_width=cfg["width"], pixel_height=cfg["height"],
                        **{f: np.load(base + f + ".npy", mmap_mode="r" if f in ("boxes", "triangles") else None)
                           for f in ARRAY_FIELDS})
        a.boxes, a.triangles = np.ascontiguousarray(a.boxes), np.ascontiguousarray(a.triangles)
    dist.barrier()
    if rank == 0:
        for f in ARRAY_FIELDS:
            try:
                os.unlink(base + f + ".npy")
            except OSError:
                pass
    return a


def seeds_for(batch, stream):
    s = np.random.RandomState(SEED + 7919 * stream).randint(0, 2 ** 32, size=(batch, 2), dtype=np.uint32)
    s[s == 0] = 1
    return s


def config_dict(cfg, arrays_or_none, world, plan=None):
    """The `config` object, identical in both arms (the driver compares them)."""
    if world == 1:
        par = "single GPU"
    else:
        par = (f"one frame on {world} GPUs: its sample passes dealt to the GPUs whole where they divide, the remainder split "
               f"by 16x16 image tile (clive2_b200.sharding.frame_plan), replicated scene, one NCCL reduce of the float accumulators")
    batch = cfg["width"] * cfg["height"]
    work = (768 + 200 + 128) * batch
    return {"workload": cfg["workload"], "width": cfg["width"], "height": cfg["height"], "samples_per_step": batch,
            "parallelism": par,
            "l2": (f"per-pass working set {work / 1e6:.0f} MB (path vertex records, connection scratch, aggregators, "
                   f"accumulators) is rewritten every step and exceeds the 126 MB L2; no explicit flush" if work > 126e6 else
                   "per-pass working set below the 126 MB L2; steps run back to back without flush")}


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference legs
# ---------------------------------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle import oracle as oracle_mod
    return ("reference", True) if oracle_mod.have_ref() else ("port", False)


def cpu_renderer(arrays, seeds):
    """The CPU reference, glibc build, all host threads: oracle/_ref (trace.metal compiled as C++) when built."""
    from oracle.oracle import OracleRenderer
    kind, use_ref = cpu_kind()
    r = OracleRenderer(arrays, seeds, shared_math=False, reference=use_ref)
    threads = r.oracle.use_all_cores()
    return r, threads, kind


CPU_WHAT = {"reference": "oracle/_ref: the reference's src/trace.metal compiled unmodified as C++ (g++ -O3 -fopenmp, glibc "
                         "libm) under the restated per-pass dispatcher of renderer.py (oracle/ref_glue.c)",
            "port": "oracle/trace_ref.c: C restatement of trace.metal + renderer.py's dispatcher (gcc -O3 -fopenmp, glibc libm)"}


def cpu_passes(arrays, n_passes, seed_stream=0):
    r, threads, kind = cpu_renderer(arrays, seeds_for(arrays.pixel_width * arrays.pixel_height, seed_stream))
    times = []
    for _ in range(n_passes):
        t0 = time.perf_counter()
        r.run_sample()
        times.append(time.perf_counter() - t0)
    est = np.nan_to_num(r.summed_image / r.summed_sample_weights)
    r.close()
    return times, threads, kind, est


def bounded_cpu_frame(cfg, arrays_full, budget_s, n_steps):
    """(arrays, width, height, note): the full frame when `n_steps` passes of it fit the budget, else the same scene at a
    frame reduced by a power of two per axis (per-sample cost does not depend on the frame size)."""
    width, height = cfg["width"], cfg["height"]
    probe_w, probe_h = (width // 8, height // 8) if width % 8 == 0 and height % 8 == 0 else (width, height)
    probe = build_arrays_like(arrays_full, cfg, probe_w, probe_h)
    t, _, _, _ = cpu_passes(probe, 2)
    per_sample = min(t) / (probe_w * probe_h)
    shrink = 1
    while (per_sample * (width // shrink) * (height // shrink) * n_steps > budget_s and min(width, height) // (2 * shrink) >= 32
           and width % (2 * shrink) == 0 and height % (2 * shrink) == 0):
        shrink *= 2
    if shrink == 1:
        return arrays_full, width, height, ""
    w, h = width // shrink, height // shrink
    return (build_arrays_like(arrays_full, cfg, w, h), w, h,
            f"frame reduced {shrink}x per axis from {width}x{height} to {w}x{h} to bound the run (same scene and BVH)")


def build_arrays_like(arrays_full, cfg, width, height):
    """The same scene at another frame size WITHOUT rebuilding mesh and BVH: only the camera record depends on the
    resolution... except that the two sensor triangles are part of the BVH and scale with the aspect ratio, so the
    aspect ratio is kept (all reduced frames are the full frame divided by an integer)."""
    import copy
    from clive2_b200.camera import Camera
    from clive2_b200 import scene
    p = scene.scene_presets[cfg["preset"]]
    assert abs(width / height - cfg["width"] / cfg["height"]) < 1e-9, "reduced frames keep the aspect ratio"
    cam = Camera(center=p["cam_center"], direction=p["cam_direction"], pixel_width=width, pixel_height=height,
                 phys_width=width / height, phys_height=1)
    a = copy.copy(arrays_full)
    a.pixel_width, a.pixel_height = width, height
    a.camera = np.array([cam.to_struct()]).reshape(1)
    return a


def cpu_baseline(cfg, arrays_full, budget_s=20.0):
    arrays, w, h, note = bounded_cpu_frame(cfg, arrays_full, budget_s, 3)
    t, threads, kind, _ = cpu_passes(arrays, 3)
    per_pass = min(t[1:])
    batch = w * h
    return {"value": batch / per_pass / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
            "sample": f"2 timed passes of {batch} samples each after 1 warm-up pass ({per_pass:.3f} s/pass)"
                      + (f"; {note}" if note else "; the full frame of the workload") + f"; {CPU_WHAT[kind]}"}


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    full = build_arrays(cfg, via_file=not args.mesh_in_memory)
    arrays, width, height, note = bounded_cpu_frame(cfg, full, 150.0, args.steps + args.warmup)
    batch = width * height
    r, threads, kind = cpu_renderer(arrays, seeds_for(batch, 0))
    for _ in range(args.warmup):
        r.run_sample()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.run_sample()
    dt = time.perf_counter() - t0
    r.close()
    value = batch * args.steps / dt / 1e6
    sample = (f"each step = one pass of {batch} samples" + (f"; {note}" if note else " = the full frame of the workload")
              + f"; {CPU_WHAT[kind]}")
    from clive2_b200 import sharding
    line = {
        "impl": "reference", "metric": "bdpt_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, full, world, sharding.shard_plan(world, args.steps)),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from clive2_b200 import runtime, scene, sharding
    from clive2_b200.renderer import Renderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; clive2_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    tdev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=tdev)

    width, height = cfg["width"], cfg["height"]
    batch = width * height
    K, W = args.steps, args.warmup
    t_scene = time.perf_counter()
    arrays = shared_arrays(cfg, rank, world, via_file=not args.mesh_in_memory)
    t_scene = time.perf_counter() - t_scene
    dev = runtime.Device(local_rank)
    t_up = time.perf_counter()
    sc = scene.scene_from_arrays(arrays, dev)

    # ---- how this rank takes part in the frame: one or two phases (sharding.frame_plan) ------------------------------------
    if world == 1:
        phases = [(K, 1, 1)]
    elif args.shard == "tile":
        phases = [(K, world, 1)]
    elif args.shard == "pass":
        if K % world:
            raise SystemExit(f"--shard pass needs --steps divisible by the number of GPUs ({K} % {world})")
        phases = [(K // world, 1, world)]
    else:
        phases = sharding.frame_plan(world, K)
    parts = []
    for i, (passes, tile_groups, pass_groups) in enumerate(phases):
        t, pgroup = rank % tile_groups, rank // tile_groups
        ids = sharding.tile_ids(width, height, t, tile_groups, 16) if tile_groups > 1 else None
        if world == 1 and args.tile:
            tw, th = (int(x) for x in args.tile.lower().split("x"))
            ids = sharding.tiled_order(width, height, tw, th)
        seeds_i = seeds_for(batch, 0 if world == 1 else 100 * i + pgroup)  # own RNG streams per pass group and phase
        parts.append(dict(passes=passes, ids=ids, seeds=seeds_i, n_local=batch if ids is None else len(ids),
                          r=Renderer(sc, seeds=seeds_i, sample_ids=ids)))
    dev.synchronize()
    t_up = time.perf_counter() - t_up
    r = parts[0]["r"]                      # phase 0 = the larger share of the work: per-kernel timing and e2e use it
    my_passes = sum(pt["passes"] for pt in parts)
    my_samples = sum(pt["passes"] * pt["n_local"] for pt in parts)
    n_local = parts[0]["n_local"]
    plan = (phases[-1][1], phases[-1][2])
    accel = dev.accel(sc.boxes, sc.triangles)
    accel_stats = accel.stats()
    extras_err = {}
    l2_read_gbs = hbm_read_gbs = None
    if rank == 0:
        try:
            l2_read_gbs = dev.read_bandwidth(32 << 20, 5)   # 32 MB: L2 resident
            hbm_read_gbs = dev.read_bandwidth(2 << 30, 5)   # 2 GB: streams from HBM
        except Exception as e:  # noqa: BLE001
            extras_err["read_bandwidth"] = repr(e)

    # ---- 1. instrumented run: work counters of exactly the passes that will be timed (same seeds => same work) ----
    dev.counters_enable(True)
    for pt in parts:
        for _ in range(W):
            pt["r"].run_sample()
    dev.counters(reset=True)
    for _ in range(parts[0]["passes"]):
        r.run_sample()
    counters = dev.counters(reset=True, per_class=True)  # phase 0 only: what the per-kernel roofline figures refer to
    for pt in parts[1:]:
        for _ in range(pt["passes"]):
            pt["r"].run_sample()
    counters_rest = dev.counters(reset=True, per_class=True)
    dev.counters_enable(False)
    for pt in parts:
        pt["r"]._fr.reset(pt["seeds"])

    stream = torch.cuda.ExternalStream(dev.stream, device=tdev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def timed_passes(work, with_reduce):
        """(ms of this rank's passes, ms of the reduce, launches): the passes of `work` = [(renderer, passes), ...] + one
        reduce between CUDA events on the launch stream, after W warm-up passes of every renderer."""
        for rend, _ in work:
            for _ in range(W):
                rend.run_sample()
        dev.synchronize()
        if world > 1 and with_reduce:
            with torch.cuda.stream(stream):  # the first reduce on a communicator sets up its channels: not timed
                warm = torch.zeros(height * width * 8, dtype=torch.float32, device=tdev)
                dist.reduce(warm, dst=0)
            torch.cuda.synchronize()
            del warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        n0 = dev.launch_count
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        for rend, n_passes in work:
            for _ in range(n_passes):
                rend.run_sample()
        e1.record(stream)
        if world > 1 and with_reduce:
            # issued with the renderer's stream as torch's current stream: NCCL's stream waits for the passes and the
            # renderer's stream for the reduce, all on the device (no host synchronisation inside the timed region)
            with torch.cuda.stream(stream):
                acc = torch.as_tensor(work[0][0].accumulator_buffer(), device=tdev)
                for rend, _ in work[1:]:  # a second phase has its own accumulators: one sum per GPU before the one reduce
                    acc += torch.as_tensor(rend.accumulator_buffer(), device=tdev)
                dist.reduce(acc, dst=0)
        e2.record(stream)
        dev.synchronize()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2), dev.launch_count - n0

    def over_ranks(x):
        """(max, mean) of a per-rank number."""
        if world == 1:
            return float(x), float(x)
        t = torch.tensor([float(x)], device=tdev, dtype=torch.float64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        v = [float(q.item()) for q in g]
        return max(v), sum(v) / len(v)

    # ---- 2. device-resident timed run -------------------------------------------------------------------------------
    pass_ms, reduce_ms, launches = timed_passes([(pt["r"], pt["passes"]) for pt in parts], True)
    elapsed_ms, _ = over_ranks(pass_ms + reduce_ms)
    pass_ms_max, pass_ms_mean = over_ranks(pass_ms)
    reduce_ms_max, _ = over_ranks(reduce_ms)

    # per-kernel durations for the roofline: more passes of the same work with CUDA events between the kernels on the
    # launch stream.  With events requested the pass runs as ONE slice on ONE stream (in the timed region above the
    # slices of a pass overlap on several streams, where "the duration of a kernel" is not defined).
    r._fr.kernel_timing(True)
    for _ in range(parts[0]["passes"]):
        r.run_sample()
    kernel_ms, timed = r._fr.kernel_times(reset=True)
    r._fr.kernel_timing(False)

    # ---- 2b. N > 1: the two pure shardings, for the record ----------------------------------------------------------------
    alt = {}
    if world > 1 and args.shard == "auto" and not args.no_alternatives:
        try:
            if phases != [(K, world, 1)]:
                rt = Renderer(sc, seeds=seeds_for(batch, 0), sample_ids=sharding.tile_ids(width, height, rank, world, 16))
                p_ms, r_ms, _ = timed_passes([(rt, K)], True)
                t_max, _ = over_ranks(p_ms + r_ms)
                pm, pa = over_ranks(p_ms)
                alt["tile_sharded"] = {"value": batch * K / (t_max * 1e-3) / 1e6, "unit": "Msamples/s", "scaling": "strong",
                                       "ms_per_step": t_max / K, "rank_pass_ms_max_over_mean": pm / pa,
                                       "what": f"every pass of the frame split over {world} GPUs by 16x16 tile "
                                               f"({batch // world} samples per GPU and pass), same RNG streams as 1 GPU"}
                rt.release()
            rp = Renderer(sc, seeds=seeds_for(batch, 500 + rank))
            p_ms, r_ms, _ = timed_passes([(rp, K)], True)
            t_max, _ = over_ranks(p_ms + r_ms)
            alt["pass_sharded_weak"] = {"value": batch * K * world / (t_max * 1e-3) / 1e6, "unit": "Msamples/s",
                                        "scaling": "weak", "ms_per_step": t_max / K,
                                        "what": f"every GPU renders {K} full-frame passes from its own RNG streams "
                                                f"({world}x the samples of the headline run), one reduce"}
            rp.release()
        except Exception as e:  # noqa: BLE001
            extras_err["alternative_shardings"] = repr(e)

    # ---- 3. end-to-end timed runs through the public API, HOST buffers in and out every step ---------------------------
    # (a) Renderer.run_sample_preview: seeds in, tone-mapped uint8 preview out = the loop body of the reference's
    #     render.py:31-37 (pass + `renderer.image`); (b) Renderer.run_sample_e2e: seeds in, raw float accumulators out.
    h_accum = runtime.PinnedArray((height, width, 8), np.float32)
    h_image = runtime.PinnedArray((height, width, 3), np.uint8)
    h_seeds = runtime.PinnedArray((batch, 2), np.uint32)
    h_seeds_b = runtime.PinnedArray((batch, 2), np.uint32)
    h_seeds.array[:] = parts[0]["seeds"]
    h_seeds_b.array[:] = seeds_for(batch, 2000 + rank)
    host_seed_buffers = (h_seeds, h_seeds_b)  # fresh host-side RNG state every step, no host arithmetic in the loop

    def timed_e2e(call):
        """Every pass of this rank's share of the frame through the public call `call(renderer, host seeds)`."""
        for pt in parts:
            pt["r"]._fr.reset(pt["seeds"])
            for k in range(W):
                call(pt["r"], host_seed_buffers[k & 1])
        dev.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for pt in parts:
            for k in range(pt["passes"]):
                call(pt["r"], host_seed_buffers[k & 1])
        dev.synchronize()
        return over_ranks(time.perf_counter() - t0)[0]

    e2e_s = timed_e2e(lambda rr, hs: rr.run_sample_preview(hs, h_image))
    checksum = int(h_image.array.astype(np.int64).sum())
    e2e_acc_s = timed_e2e(lambda rr, hs: rr.run_sample_e2e(hs, h_accum))
    clocks = sampler.stop() if rank == 0 else None

    # ---- 4. image error against the CPU reference (N = 1) -----------------------------------------------------------------
    rel_rmse = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rel_rmse = image_error(args, cfg, arrays, dev)
        except Exception as e:  # noqa: BLE001
            extras_err["rel_rmse"] = repr(e)

    if rank == 0:
        samples = batch * K
        value = samples / (elapsed_ms * 1e-3) / 1e6
        e2e_value = samples / e2e_s / 1e6
        tot = {k: counters["walk"][k] + counters["connect"][k] + counters_rest["walk"][k] + counters_rest["connect"][k]
               for k in counters["walk"]}
        peak_hbm, peak_src = measured_peaks()
        np_ = max(timed, 1)
        # the dominant kernel of the pass (by its own CUDA-event time) and its algorithmic bytes
        dom = max(kernel_ms, key=lambda k: kernel_ms[k])
        # per pass; every other kernel once.  12 bounce rounds + 4 short ones for the samples whose searches were suspended
        launches_of = {"k_walk_trace": 16, "k_walk_shade": 17 if "k_walk_trace" in kernel_ms else 1}
        cw, cc = ({k: v / max(parts[0]["passes"], 1) for k, v in counters[x].items()} for x in ("walk", "connect"))
        NBW = 128 if accel_stats["closest_uses_wide"] else 64   # bytes of the node record the walks fetch
        NBR = 128 if accel_stats["occlusion_uses_wide"] else 64  # ... and the connection rays (4-wide node = 128 B)
        NBT = 128 if accel_stats.get("trace_uses_wide") else 64  # ... and the wavefront trace kernel
        VB = 64  # bytes of one stored path vertex in the renderer's SoA records (four 16-byte parts)
        alg = {
            # node (64 B binary / 128 B 4-wide) + triangle (48 B) records fetched + the hit triangle's shading record
            # (64 B) + path vertices written (64 B) + seeds RMW (16 B) + accumulator RMW (64 B)
            "k_fused_walks": cw["node_fetches"] * NBW + cw["tri_fetches"] * 48 + cw["rays_ref"] * 64 + cw["vertices"] * VB
                             + n_local * (16 + 64),
            # node + triangle records fetched + per task: 4 B task word, origin + triangle parts of two vertices (4 x 16 B)
            "k_connect_rays": cc["node_fetches"] * NBR + cc["tri_fetches"] * 48 + cc["rays_ref"] * (4 + 64),
            # path vertices staged (64 B each), visibility bits + projection record (8 + 48 B), aggregator written (80 B)
            "k_connect_finish": cc["vertices"] * VB + n_local * (8 + 48 + 80),
            # origin / normal / material parts of every vertex (48 B), task words + projection record written
            "k_connect_enumerate": cc["vertices"] * 48 + cc["rays_ref"] * 4 + n_local * (48 + 8),
            # 9 aggregators read (80 B), light image RMW (32 B), accumulator RMW (64 B) per pixel
            "k_fused_finalize": batch * (9 * 80 + 32 + 64),
            # wavefront walks, all rounds of a pass.  trace: node + triangle records fetched + per ray origin/direction
            # parts read (32 B), hit written (16 B), queue entry (4 B).  (The continuation records of the < 1 % of the
            # searches that are suspended, <= 576 B written and read each, are not counted.)
            "k_walk_trace": cw["node_fetches"] * NBT + cw["tri_fetches"] * 48 + cw["rays_ref"] * (32 + 16 + 4),
            # shade (+ init): per ray hit (16 B) + walk state read and written (3-4 parts each way, 112 B) + seeds RMW
            # (16 B) + queue entry read + written (8 B) + the hit triangle's shading record (64 B); path vertices written
            # (64 B + 4 B patch); per sample: generators (seeds 16 B, jitter 8 B, state 48 B) + accumulator RMW (64 B)
            "k_walk_shade": cw["rays_ref"] * (16 + 112 + 16 + 8 + 64) + cw["vertices"] * 68 + n_local * (16 + 8 + 48 + 64),
        }
        # the pipelined form runs trace and shade in ONE kernel (plus the separately timed k_walk_init, counted under shade)
        alg["k_walk_pipeline"] = alg["k_walk_trace"] + alg["k_walk_shade"]
        for k in list(kernel_ms):
            alg.setdefault(k, 0.0)
        accel_bytes = accel_stats["node_bytes"] + accel_stats["wide_node_bytes"] + accel_stats["triangle_bytes"]
        resident = accel_bytes < L2_BYTES * 0.8
        dom_ms = kernel_ms[dom] / np_
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
        if resident and l2_read_gbs:
            bound, peak, src = "l2", l2_read_gbs, ("L2 read peak measured in this run (clv_measure_read_bandwidth, 32 MB "
                                                   "resident buffer); MEASURED_PEAKS.json has no L2 figure")
        else:
            bound, peak, src = "hbm", peak_hbm, peak_src
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(REPO, "profiles", "dram_traffic.json")))
            traffic = tj.get(args.config, {}).get(dom)
            traffic_note = tj.get("_note")
        except Exception:
            pass
        roofline = {"bound": bound, "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": src, "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": dom_ms,
                    "dram_frac": (traffic / (dom_ms * 1e-3) / 1e9 / peak_hbm) if traffic else None,
                    "traffic_source": traffic_note,
                    "launches_per_pass_of_this_kernel": launches_of.get(dom, 1),
                    "per_launch_definition": "bytes, time and traffic are per PASS: the sum over this kernel's launches of one "
                                             "pass (the wavefront walk kernels run once per bounce round: 12 rounds + 4 short ones "
                                             "for samples whose searches were suspended)"
                                             if dom in launches_of else "one launch per pass",
                    "traversal_records_bytes": accel_bytes, "traversal_records_l2_resident": bool(resident),
                    "l2_read_peak_gbs_measured_here": l2_read_gbs, "hbm_read_peak_gbs_measured_here": hbm_read_gbs,
                    "kernel_ms_per_pass": {k: v / np_ for k, v in kernel_ms.items()},
                    "algorithmic_gbs_per_kernel": {k: alg[k] / (kernel_ms[k] / np_ * 1e-3) / 1e9
                                                   for k in kernel_ms if kernel_ms[k] > 0 and alg.get(k)},
                    "per_ray": {"node_records_walk": cw["node_fetches"] / max(cw["rays_traced"], 1),
                                "triangle_records_walk": cw["tri_fetches"] / max(cw["rays_traced"], 1),
                                "node_records_connect": cc["node_fetches"] / max(cc["rays_traced"], 1),
                                "triangle_records_connect": cc["tri_fetches"] / max(cc["rays_traced"], 1)},
                    "note": "achieved = algorithmic bytes / CUDA-event time of the kernel; algorithmic bytes = node (64 B; "
                            "128 B 4-wide) and triangle (48 B) records the kernel fetches (device counters of an "
                            "instrumented run of the same passes) + the per-sample / per-ray records it reads and writes "
                            "(DESIGN.md §7).  dram_frac = ncu DRAM bytes of the kernel / its time / HBM peak: part of "
                            "the algorithmic bytes is served by L1/L2"}
        line = {
            "metric": "bdpt_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, arrays, world, plan),
            "run": {"triangles": int(len(arrays.triangles)), "boxes": int(len(arrays.boxes)), "accel_bytes": accel_bytes,
                    "bvh_levels": {"binary": accel_stats.get("binary_depth"), "four_wide": accel_stats.get("wide_depth")},
                    "phases": [{"passes_per_gpu": pa, "tile_groups": tg, "pass_groups": pg, "samples_per_pass_per_gpu": pt["n_local"]}
                               for (pa, tg, pg), pt in zip(phases, parts)],
                    "samples_per_pass_per_gpu": n_local, "passes_per_gpu": my_passes,
                    "launches_per_pass": launches // max(my_passes, 1),
                    "kernel_flavour": ("high-occupancy (traversal records exceed the caches)"
                                       if accel_stats["large_scene_tuning"] else "cache-resident"),
                    "walks": "wavefront (k_walk_init + (12 + 4) x (k_walk_trace, k_walk_shade))" if "k_walk_trace" in kernel_ms
                             else "wavefront, pipelined (k_walk_init + k_walk_pipeline)" if "k_walk_pipeline" in kernel_ms
                             else "megakernel (k_fused_walks)",
                    "mesh_source": "in-memory" if (args.mesh_in_memory or not cfg.get("mesh")) else
                                   f"{cfg['mesh'][0].upper()} file written and read back (load.write_{cfg['mesh'][0]} / read_{cfg['mesh'][0]})",
                    "scene_build_s": t_scene, "scene_upload_s": t_up, "excluded_from_timed_region": "scene build + upload",
                    "rank_pass_ms_max": pass_ms_max, "rank_pass_ms_mean": pass_ms_mean,
                    "rank_pass_ms_max_over_mean": pass_ms_max / max(pass_ms_mean, 1e-9),
                    "reduce_ms": reduce_ms_max, "reduce_bytes": batch * 32 if world > 1 else 0},
            "mrays_per_s": value * tot["rays_ref"] / max(my_samples, 1),
            "rays_per_sample": tot["rays_ref"] / max(my_samples, 1),
            "rays_definition": "a ray = one traverse_bvh invocation the reference makes for the same paths (SURVEY.md 8d: "
                               "closest hits of the walks, shadow rays, projections), counted by device counters of an "
                               "instrumented run of the same passes (rank 0's share at N > 1); mrays_per_s = value x "
                               "rays_per_sample",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": my_samples * 8 // max(my_passes, 1),
                    "d2h_bytes_per_step": batch * 3, "ms_per_step": e2e_s / K * 1e3, "checksum": checksum,
                    "api": "Renderer.run_sample_preview -> clv_renderer_run_sample_preview: pinned host seeds in, one "
                           "pass, device tone map, uint8 preview image out (render.py:31-37)"},
            "e2e_accumulators": {"value": samples / e2e_acc_s / 1e6, "unit": "Msamples/s",
                                 "h2d_bytes_per_step": my_samples * 8 // max(my_passes, 1), "d2h_bytes_per_step": batch * 32,
                                 "api": "Renderer.run_sample_e2e: seeds in, packed float accumulators [H][W][8] out"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if world > 1:
            line.update(alt)
        if rel_rmse is not None:
            line["rel_rmse"] = rel_rmse
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(cfg, arrays)
            except Exception as e:  # noqa: BLE001
                extras_err["cpu_baseline"] = repr(e)
        if extras_err:
            line["errors_in_extras"] = extras_err
        print(json.dumps(line), flush=True)

    for x in (h_seeds, h_seeds_b, h_accum, h_image):
        x.release()
    for pt in parts:
        pt["r"].release()
    sc.release()
    dev.release()
    if world > 1:
        dist.destroy_process_group()


def image_error(args, cfg, arrays_full, dev):
    """Relative RMSE (sqrt(mean((a-b)^2)) / mean(b)) of the estimate `summed_image / summed_sample_weights`
    (renderer.py:293-300) on a reduced frame of the same scene: GPU at RMSE_SPP vs CPU reference (glibc build) at the
    same spp (different RNG streams and elementary functions: this is Monte-Carlo noise of two independent renders), and
    each of them against a CPU image at RMSE_HI_SPP."""
    from clive2_b200 import scene
    from clive2_b200.renderer import Renderer
    w, h = RMSE_FRAME[args.config]
    a = build_arrays_like(arrays_full, cfg, w, h)

    def err(img, ref):
        img, ref = np.asarray(img, np.float64), np.asarray(ref, np.float64)
        return float(np.sqrt(np.mean((img - ref) ** 2)) / max(np.mean(ref), 1e-30))

    _, threads, kind, cpu_hi = cpu_passes(a, RMSE_HI_SPP, seed_stream=501)
    _, _, _, cpu_eq = cpu_passes(a, RMSE_SPP, seed_stream=502)
    sc = scene.scene_from_arrays(a, dev)
    r = Renderer(sc, seeds=seeds_for(w * h, 503))
    for _ in range(RMSE_SPP):
        r.run_sample()
    gpu_eq = r.estimate()
    for _ in range(RMSE_HI_SPP - RMSE_SPP):
        r.run_sample()
    gpu_hi = r.estimate()
    r.release()
    sc.release()
    return {"gpu_vs_cpu_equal_spp": err(gpu_eq, cpu_eq), "gpu_vs_hi_spp": err(gpu_eq, cpu_hi), "cpu_vs_hi_spp": err(cpu_eq, cpu_hi),
            "gpu_hi_vs_cpu_hi": err(gpu_hi, cpu_hi), "mean_ratio_gpu_over_cpu_hi": float(gpu_hi.mean() / cpu_hi.mean()),
            "spp": RMSE_SPP, "hi_spp": RMSE_HI_SPP, "frame": f"{w}x{h}", "cpu_kind": kind,
            "what": f"same scene and BVH as the workload at a {w}x{h} frame; independent RNG streams per image; the two "
                    f"equal-spp errors against the high-spp CPU image should agree (same estimator, same noise)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", choices=sorted(CONFIGS), default=os.environ.get("CLV_BENCH_CONFIG", "c4"))
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline and rel_rmse (profiling runs)")
    ap.add_argument("--no-alternatives", action="store_true", help="N > 1: skip the pure tile / weak pass measurements")
    ap.add_argument("--mesh-in-memory", action="store_true", help="c3/c4: do not round-trip the mesh through OBJ/PLY")
    ap.add_argument("--shard", choices=("auto", "pass", "tile"), default="auto",
                    help="N > 1: 'auto' = whole passes per GPU + the remainder as pass groups x tile groups (sharding.frame_plan, default), 'tile' = every pass split "
                         "by 16x16 tiles, 'pass' = passes dealt to the GPUs (needs steps % N == 0); all strong scaling")
    ap.add_argument("--tile", default=os.environ.get("CLV_BENCH_TILE", ""),
                    help="N = 1: trace samples in tiles of WxH pixels per warp-sized group (e.g. 8x4); default row-major")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    args.steps = max(args.steps, 1)
    cfg = CONFIGS[args.config]

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    run_ours(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
