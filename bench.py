#!/usr/bin/env python
"""Benchmark of the BDPT hot path (one "step" = one sample pass = Renderer.run_sample over a full frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is whole-job Msamples/s with everything resident in HBM (CUDA events on the
launch stream, max over ranks); `e2e` is the same metric through `Renderer.run_sample_preview` with pinned HOST buffers
(fresh seeds up, one pass, device tone map, uint8 preview image down, all inside the timed region; `e2e_accumulators`
downloads the raw float accumulators instead); `roofline` is for the dominant kernel of the pass;
`cpu_baseline` is the CPU transcription of the reference (oracle/, glibc build) timed on this box's host cores.
`--impl reference` times that CPU transcription as the reference arm (the reference's Metal kernels cannot run on
Linux and its Python needs `metalcompute`).

Multi-GPU (torchrun, one rank per GPU): every rank renders full-frame passes with its own RNG streams (the work
shards by sample pass, no data-path collective), and ONE NCCL reduce of the packed float accumulators closes the timed
region.  Per-GPU work is fixed => "scaling": "weak".
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

CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1": dict(preset="empty", width=256, height=256,
               workload="C1: Cornell box (preset 'empty', 16 triangles), 256x256, diffuse, max path length 6+6"),
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "c2": dict(preset="bench-spheres", width=512, height=512,
               workload="C2: Cornell box + two icospheres (2 x 20480 triangles), rough-dielectric (GGX transmission) "
                        "and coated-diffuse (GGX specular) materials, 512x512"),
    # BASELINE.json configs[2]
    "c3": dict(preset="bench-terrain-1m", width=1024, height=1024,
               workload="C3: Cornell enclosure + procedural terrain mesh (1 002 528 triangles), 1024x1024"),
}
CONFIGS["c4"] = dict(preset="bench-terrain-10m", width=1920, height=1080,
                     workload="C4: Cornell enclosure + procedural terrain mesh (9 999 392 triangles), 1920x1080")
SEED = 20240928


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
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


def build_arrays(cfg, width=None, height=None):
    from clive2_b200 import scene
    return scene.preset_arrays(cfg["preset"], width or cfg["width"], height or cfg["height"])


def seeds_for(batch, rank):
    s = np.random.RandomState(SEED + 7919 * rank).randint(0, 2 ** 32, size=(batch, 2), dtype=np.uint32)
    s[s == 0] = 1
    return s


# ---------------------------------------------------------------------------------------------------------------------
# CPU transcription (oracle) legs
# ---------------------------------------------------------------------------------------------------------------------
def oracle_passes(arrays, seeds, n_passes):
    """Seconds per pass of the C transcription (glibc build, all host threads) over `n_passes` passes."""
    from oracle.oracle import OracleRenderer
    r = OracleRenderer(arrays, seeds, shared_math=False)
    r.oracle.use_all_cores()
    times = []
    for _ in range(n_passes):
        t0 = time.perf_counter()
        r.run_sample()
        times.append(time.perf_counter() - t0)
    threads = r.oracle.num_threads
    r.close()
    return times, threads


def cpu_baseline(cfg, budget_s=20.0):
    arrays = build_arrays(cfg)
    batch = cfg["width"] * cfg["height"]
    seeds = seeds_for(batch, 0)
    t1, threads = oracle_passes(arrays, seeds, 1)
    n_more = int(max(1, min(6, (budget_s - t1[0]) // max(t1[0], 1e-3))))
    tn, _ = oracle_passes(arrays, seeds, n_more)
    per_pass = min(statistics.median(tn), t1[0])
    return {"value": batch / per_pass / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
            "sample": f"{n_more} full passes of the same workload ({batch} samples each) after 1 warm-up pass; "
                      f"oracle/trace_ref.c (gcc -O3 -fopenmp, glibc libm), {per_pass:.3f} s/pass"}


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    width, height = cfg["width"], cfg["height"]
    arrays = build_arrays(cfg)
    probe, threads = oracle_passes(arrays, seeds_for(width * height, 0), 1)
    shrink = 1
    while probe[0] / (shrink * shrink) * (args.steps + args.warmup) > 240.0 and min(width, height) // (2 * shrink) >= 64:
        shrink *= 2
    if shrink > 1:
        width, height = width // shrink, height // shrink
        arrays = build_arrays(cfg, width, height)
    batch = width * height
    from oracle.oracle import OracleRenderer
    r = OracleRenderer(arrays, seeds_for(batch, 0), shared_math=False)
    threads = r.oracle.use_all_cores()
    for _ in range(args.warmup):
        r.run_sample()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.run_sample()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt / 1e6
    sample = (f"each step = one full pass of the {width}x{height} frame ({batch} samples)"
              + (f", frame reduced {shrink}x per axis from {cfg['width']}x{cfg['height']} to bound the run" if shrink > 1 else ""))
    line = {
        "impl": "reference", "metric": "bdpt_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "width": width, "height": height, "samples_per_step": batch,
                   "what": "CPU transcription of the reference's trace.metal + renderer.py pass (oracle/trace_ref.c, "
                           "glibc build, OpenMP over all host threads); the reference's Metal kernels cannot run on Linux"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
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

    from clive2_b200 import runtime, scene
    from clive2_b200.renderer import Renderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; clive2_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    width, height = cfg["width"], cfg["height"]
    batch = width * height
    t_scene = time.perf_counter()
    arrays = build_arrays(cfg)
    t_scene = time.perf_counter() - t_scene
    dev = runtime.Device(local_rank)
    t_up = time.perf_counter()
    sc = scene.scene_from_arrays(arrays, dev)
    tile_sharded = args.shard == "tile" and world > 1
    seeds = seeds_for(batch, 0 if tile_sharded else rank)
    ids = None
    if tile_sharded:
        # strong scaling of ONE frame: this rank traces its 16x16 tiles with the global RNG rows (same streams as 1 GPU)
        from clive2_b200 import sharding
        ids = sharding.tile_ids(width, height, rank, world, 16)
    elif args.tile:
        from clive2_b200 import sharding
        tw, th = (int(x) for x in args.tile.lower().split("x"))
        ids = sharding.tiled_order(width, height, tw, th)
    r = Renderer(sc, seeds=seeds, sample_ids=ids)
    dev.synchronize()
    t_up = time.perf_counter() - t_up
    accel_stats = dev.accel(sc.boxes, sc.triangles).stats()
    l2_read_gbs = dev.read_bandwidth(32 << 20, 5) if rank == 0 else None     # 32 MB: L2 resident
    hbm_read_gbs = dev.read_bandwidth(2 << 30, 5) if rank == 0 else None      # 2 GB: streams from HBM
    K, W = args.steps, args.warmup

    # ---- 1. instrumented run: work counters of exactly the passes that will be timed (same seeds => same work) ----
    dev.counters_enable(True)
    for _ in range(W):
        r.run_sample()
    dev.counters(reset=True)
    for _ in range(K):
        r.run_sample()
    counters = dev.counters(reset=True, per_class=True)
    dev.counters_enable(False)
    r._fr.reset(seeds)

    stream = torch.cuda.ExternalStream(dev.stream, device=torch.device("cuda", local_rank))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- 2. device-resident timed run -------------------------------------------------------------------------------
    for _ in range(W):
        r.run_sample()
    dev.synchronize()
    if world > 1:
        # warm-up of the collective too: the first reduce on a communicator sets up its channels (the timed region's
        # reduce is then the steady-state one); a scratch tensor of the accumulators' size, not the accumulators
        with torch.cuda.stream(stream):
            warm = torch.zeros(height * width * 8, dtype=torch.float32, device=torch.device("cuda", local_rank))
            dist.reduce(warm, dst=0)
        torch.cuda.synchronize()
        del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = dev.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        r.run_sample()
    if world > 1:
        # the reduce is issued with the renderer's stream as torch's current stream: NCCL's stream waits for the passes
        # and the renderer's stream for the reduce, all on the device (no host synchronisation inside the timed region)
        with torch.cuda.stream(stream):
            acc = torch.as_tensor(r.accumulator_buffer(), device=torch.device("cuda", local_rank))
            dist.reduce(acc, dst=0)
    e1.record(stream)
    dev.synchronize()
    torch.cuda.synchronize()
    launches = dev.launch_count - launches0
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()
    # per-kernel durations for the roofline: K more passes of the same work with CUDA events between the kernels on the
    # launch stream.  With events requested the pass runs as ONE slice on ONE stream (in the timed region above the
    # slices of a pass overlap on several streams, where "the duration of a kernel" is not defined).
    r._fr.kernel_timing(True)
    for _ in range(K):
        r.run_sample()
    kernel_ms, timed_passes = r._fr.kernel_times(reset=True)
    r._fr.kernel_timing(False)

    # ---- 3. end-to-end timed runs through the public API, HOST buffers in and out every step ---------------------------
    # (a) Renderer.run_sample_preview: seeds in, tone-mapped uint8 preview out = the loop body of the reference's
    #     render.py:31-37 (pass + `renderer.image`); (b) Renderer.run_sample_e2e: seeds in, raw float accumulators out.
    h_seeds = runtime.PinnedArray((batch, 2), np.uint32)
    h_accum = runtime.PinnedArray((height, width, 8), np.float32)
    h_image = runtime.PinnedArray((height, width, 3), np.uint8)

    h_seeds_b = runtime.PinnedArray((batch, 2), np.uint32)
    h_seeds.array[:] = seeds
    h_seeds_b.array[:] = seeds_for(batch, rank + 1000)
    host_seed_buffers = (h_seeds, h_seeds_b)  # fresh host-side RNG state every step, no host arithmetic in the loop

    def timed_e2e(step):
        r._fr.reset(seeds)
        for k in range(W):
            step(host_seed_buffers[k & 1])
        dev.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(K):
            step(host_seed_buffers[k & 1])
        dev.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_s = timed_e2e(lambda hs: r.run_sample_preview(hs, h_image))
    checksum = int(h_image.array.astype(np.int64).sum())
    e2e_acc_s = timed_e2e(lambda hs: r.run_sample_e2e(hs, h_accum))
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        samples = batch * K * (1 if tile_sharded else world)
        value = samples / (elapsed_ms * 1e-3) / 1e6
        e2e_value = samples / e2e_s / 1e6
        tot = {k: counters["walk"][k] + counters["connect"][k] for k in counters["walk"]}
        rays_ref_per_pass = tot["rays_ref"] / K
        peak, peak_src = measured_peaks()
        # the dominant kernel of the pass (by its own CUDA-event time) and its algorithmic bytes
        dom = max(kernel_ms, key=lambda k: kernel_ms[k])
        cw, cc = ({k: v / K for k, v in counters[x].items()} for x in ("walk", "connect"))
        NBW = 128 if accel_stats["closest_uses_wide"] else 64   # bytes of the node record the walks fetch
        NBR = 128 if accel_stats["occlusion_uses_wide"] else 64  # ... and the connection rays (4-wide node = 128 B)
        VB = 64  # bytes of one stored path vertex in the renderer's SoA records (four 16-byte parts)
        alg = {
            # node (64 B binary / 128 B 4-wide) + triangle (48 B) records fetched + the hit triangle's shading record
            # (80 B of the reference Triangle) + path vertices written (64 B) + seeds RMW (16 B) + accumulator RMW (64 B)
            "k_fused_walks": cw["node_fetches"] * NBW + cw["tri_fetches"] * 48 + cw["rays_ref"] * 80 + cw["vertices"] * VB
                             + batch * (16 + 64),
            # node + triangle records fetched + per task: 4 B task word, origin + triangle parts of two vertices (4 x 16 B)
            "k_connect_rays": cc["node_fetches"] * NBR + cc["tri_fetches"] * 48 + cc["rays_ref"] * (4 + 64),
            # path vertices staged (64 B each), visibility bits + projection record (8 + 48 B), aggregator written (80 B)
            "k_connect_finish": cc["vertices"] * VB + batch * (8 + 48 + 80),
            # origin / normal / material parts of every vertex (48 B), task words + projection record written
            "k_connect_enumerate": cc["vertices"] * 48 + cc["rays_ref"] * 4 + batch * (48 + 8),
            # 9 aggregators read (80 B), light image RMW (32 B), accumulator RMW (64 B) per pixel
            "k_fused_finalize": batch * (9 * 80 + 32 + 64),
        }
        alg_bytes = alg[dom]
        dom_ms = kernel_ms[dom] / max(timed_passes, 1)
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": dom_ms,
                    "l2_read_peak_gbs_measured_here": l2_read_gbs, "hbm_read_peak_gbs_measured_here": hbm_read_gbs,
                    "frac_of_l2_read_peak": achieved / l2_read_gbs if l2_read_gbs else None,
                    "kernel_ms_per_pass": {k: v / max(timed_passes, 1) for k, v in kernel_ms.items()},
                    "algorithmic_gbs_per_kernel": {k: alg[k] / (kernel_ms[k] / max(timed_passes, 1) * 1e-3) / 1e9
                                                   for k in kernel_ms if kernel_ms[k] > 0},
                    "note": "algorithmic bytes = node (64 B) and triangle (48 B) records the kernel fetches (device "
                            "counters of an instrumented run of the same passes) + the per-sample / per-ray records it "
                            "reads and writes (DESIGN.md §7); BVH records of this configuration are cache resident, so "
                            "DRAM traffic (`traffic`, from ncu) is far below the algorithmic figure"}
        traffic_file = os.path.join(REPO, "profiles", "dram_traffic.json")
        if os.path.exists(traffic_file):
            try:
                roofline["traffic"] = json.load(open(traffic_file)).get(args.config, {}).get(roofline["kernel"])
            except Exception:
                pass
        line = {
            "metric": "bdpt_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "strong" if tile_sharded else "weak",
            "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "width": width, "height": height, "samples_per_step_per_gpu": batch,
                       "triangles": int(len(arrays.triangles)), "boxes": int(len(arrays.boxes)),
                       "parallelism": (f"tile-sharded x{world} (16x16 tiles round-robin, one NCCL reduce)" if tile_sharded else
                                       f"pass-sharded x{world} (one NCCL reduce)") if world > 1 else "single GPU",
                       "l2": f"per-pass working set {(768 + 200 + 128) * batch / 1e6:.0f} MB (path vertex records, connection "
                             f"scratch, aggregators + accumulators) is rewritten every step (> 126 MB L2); no explicit flush"
                             if (768 + 200 + 128) * batch > 126e6
                             else "per-pass working set below the 126 MB L2; steps run back to back without flush",
                       "accel_bytes": accel_stats["node_bytes"] + accel_stats["wide_node_bytes"] + accel_stats["triangle_bytes"],
                       "launches_per_step": launches // K,
                       "kernel_flavour": ("high-occupancy (walks 32 / rays 48 registers; traversal records exceed the caches)"
                                          if accel_stats["large_scene_tuning"] else
                                          "cache-resident (walks 64 / rays 63 registers)"),
                       "scene_build_s": t_scene, "scene_upload_s": t_up},
            "mrays_per_s": rays_ref_per_pass * K * world / (elapsed_ms * 1e-3) / 1e6,
            "mrays_traced_per_s": tot["rays_traced"] * world / (elapsed_ms * 1e-3) / 1e6,
            "rays_per_sample": rays_ref_per_pass / (len(ids) if ids is not None else batch),
            "rays_definition": "mrays_per_s counts traverse_bvh invocations the reference makes for the same paths "
                               "(SURVEY.md 8d); mrays_traced_per_s counts traversals this implementation starts "
                               "(connection rays whose target test fails first are skipped)",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": batch * 8,
                    "d2h_bytes_per_step": batch * 3, "ms_per_step": e2e_s / K * 1e3, "checksum": checksum,
                    "api": "Renderer.run_sample_preview -> clv_renderer_run_sample_preview: pinned host seeds in, one "
                           "pass, device tone map, uint8 preview image out (render.py:31-37)"},
            "e2e_accumulators": {"value": samples / e2e_acc_s / 1e6, "unit": "Msamples/s",
                                 "h2d_bytes_per_step": batch * 8, "d2h_bytes_per_step": batch * 32,
                                 "api": "Renderer.run_sample_e2e: seeds in, packed float accumulators [H][W][8] out"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line), flush=True)

    h_seeds.release()
    h_seeds_b.release()
    h_accum.release()
    h_image.release()
    r.release()
    sc.release()
    dev.release()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2")
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", choices=("pass", "tile"), default="pass",
                    help="N > 1: 'pass' = every rank renders full-frame passes from its own seeds (weak scaling, default); "
                         "'tile' = the ranks split every pass of ONE frame by 16x16 tiles (strong scaling)")
    ap.add_argument("--tile", default=os.environ.get("CLV_BENCH_TILE", ""),
                    help="WxH: trace samples in tiles of WxH pixels per warp-sized group (e.g. 8x4); default row-major")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    args.steps = max(args.steps, 1)
    cfg = CONFIGS[args.config]

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
