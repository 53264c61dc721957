"""Headless equivalents of the reference's CLIs: `render` (src/render.py:11-50) and the orbit `movie`
(src/movie.py:10-55), the latter frame-parallel over GPUs (BASELINE config 5).

    python -m clive2_b200.movie render --scene bench-spheres --width 512 --height 512 --samples 64 --out out.png
    torchrun --nproc-per-node 8 -m clive2_b200.movie movie --scene bench-spheres --movie-frames 120 --samples 32
    torchrun --nproc-per-node 8 -m clive2_b200.movie render --scene bench-terrain-10m --width 1920 --height 1080 --samples 256

Frames of a movie are independent (the reference rebuilds scene + BVH + renderer per frame, movie.py:29-55), so frame f
goes to rank f mod world_size and no collective is needed; within a rank the host work of neighbouring frames (scene
arrays + BVH of the next one, PNG encode of the previous one) overlaps the GPU's passes of the current one.  PNGs are written with cv2 when it is importable, else as
.npy arrays.
"""
import argparse
import os
import time

import numpy as np

from . import runtime, scene
from .camera import tone_map
from .renderer import TONE_MAP_EXPOSURE, Renderer


def frames_for_rank(start_frame, total_frames, rank, world):
    """Frame indices rendered by `rank`: round-robin over the frames still to do (movie.py --start-frame semantics)."""
    return [f for f in range(start_frame, total_frames) if (f - start_frame) % world == rank]


def write_image(path, image_bgr_u8):
    """PNG through cv2 (movie.py:52); only when cv2 itself is missing does the frame go out as .npy instead.  A failed
    write (unwritable path, unknown extension: cv2.imwrite returns False) is an error, not a silently missing frame."""
    try:
        import cv2
    except ImportError:
        np.save(path + ".npy", image_bgr_u8)
        return path + ".npy"
    if not cv2.imwrite(path, image_bgr_u8):
        raise OSError(f"cv2.imwrite could not write {path!r}")
    return path


def frame_arrays(preset, width, height, frame_idx=None, total_frames=None):
    """Host side of one frame: scene arrays + BVH (native builder; releases the GIL, so it overlaps a running render)."""
    if frame_idx is None:
        return scene.preset_arrays(preset, width, height)
    center, direction = scene.orbit_camera(frame_idx, total_frames)
    return scene.preset_arrays(preset, width, height, center, direction)


def render_arrays(device, arrays, samples, seed=None, raw=False):
    """Device side of one frame: upload -> `samples` fused passes -> tone-mapped BGR (`raw`: the packed float
    accumulators [H][W][8] instead, for `image_from_accumulators` on another thread)."""
    sc = scene.scene_from_arrays(arrays, device)
    r = Renderer(sc, seed=seed)
    for _ in range(samples):
        r.run_sample()
    result = r._accum() if raw else r.image
    r.release()
    sc.release()  # also drops the per-frame BVH cached for these buffers
    return result


def image_from_accumulators(a):
    """`Renderer.image` (renderer.py:293-300 of the reference) from the packed accumulators of `render_arrays(raw=True)`."""
    with np.errstate(invalid="ignore", divide="ignore"):
        est = a[:, :, 0:3] / a[:, :, 3:4]
    return tone_map(np.nan_to_num(est, neginf=0, posinf=0), exposure=TONE_MAP_EXPOSURE)


def render_frame(device, preset, width, height, samples, frame_idx=None, total_frames=None, seed=None):
    """One image: scene (+ orbit camera when frame_idx is given) -> BVH -> `samples` fused passes -> tone-mapped BGR."""
    return render_arrays(device, frame_arrays(preset, width, height, frame_idx, total_frames), samples, seed)


def passes_for_rank(samples, rank, world):
    """How many of the `samples` passes of one image `rank` renders (pass sharding: the first `samples % world` ranks
    take one more)."""
    return samples // world + (1 if rank < samples % world else 0)


def render_frame_sharded(device, preset, width, height, samples, rank, world, local_rank, seed=None):
    """One image on `world` GPUs (BASELINE config 4): every rank renders its share of the passes over the
    full frame from its own RNG streams (`sharding.rank_seeds`), ONE NCCL reduce sums the packed accumulators on rank 0,
    which tone-maps.  Returns the image on rank 0 and None elsewhere.  torch.distributed must be initialised (nccl)."""
    import torch
    import torch.distributed as dist

    from . import sharding
    arrays = frame_arrays(preset, width, height)
    sc = scene.scene_from_arrays(arrays, device)
    base = 20240928 if seed is None else int(seed)
    r = Renderer(sc, seeds=sharding.rank_seeds(width * height, base, rank))
    for _ in range(passes_for_rank(samples, rank, world)):
        r.run_sample()
    device.synchronize()
    acc = torch.as_tensor(r.accumulator_buffer(), device=torch.device("cuda", local_rank))
    sharding.reduce_accumulators(acc, dst=0)
    torch.cuda.synchronize(local_rank)
    image = r.image if rank == 0 else None  # accumulator_buffer() dropped the host cache: this reads the reduced sums
    r.release()
    sc.release()
    return image


def render_movie(device, preset, width, height, samples, frames, total_frames, seed=None, sink=None, overlap=True):
    """Frames `frames` of the orbit movie, in order; `sink(frame_idx, image)` receives each image (e.g. the PNG
    writer).  The reference does scene + BVH + render + imwrite serially per frame (movie.py:29-55); here the host
    work is taken off the GPU's critical path: while frame f renders, one worker thread prepares the arrays and BVH
    of the next frame and another tone-maps and encodes the previous image.  Images are identical to `render_frame`'s."""
    frames = list(frames)
    if not overlap:
        for f in frames:
            img = render_frame(device, preset, width, height, samples, f, total_frames, seed)
            if sink is not None:
                sink(f, img)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(1) as prep, ThreadPoolExecutor(1) as out:
        pending = prep.submit(frame_arrays, preset, width, height, frames[0], total_frames) if frames else None
        written = []
        for k, f in enumerate(frames):
            arrays = pending.result()
            if k + 1 < len(frames):
                pending = prep.submit(frame_arrays, preset, width, height, frames[k + 1], total_frames)
            acc = render_arrays(device, arrays, samples, seed, raw=True)
            written.append(out.submit(lambda f=f, acc=acc: sink and sink(f, image_from_accumulators(acc))))
        for w in written:
            w.result()  # surface writer errors


def main(argv=None):
    ap = argparse.ArgumentParser(prog="clive2_b200.movie")
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name in ("render", "movie"):
        p = sub.add_parser(name)
        p.add_argument("--samples", type=int, default=15)
        p.add_argument("--width", type=int, default=1280)
        p.add_argument("--height", type=int, default=720)
        p.add_argument("--scene", type=str, default="bench-spheres")
        p.add_argument("--seed", type=int, default=None)
        if name == "render":
            p.add_argument("--out", type=str, default="output/render.png")
        else:
            p.add_argument("--movie-name", type=str, default="test-movie")
            p.add_argument("--movie-frames", type=int, default=120)
            p.add_argument("--start-frame", type=int, default=0)
            p.add_argument("--out-dir", type=str, default="output")
            p.add_argument("--serial", action="store_true", help="no overlap of host scene prep / PNG encode with the GPU")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = runtime.Device(int(os.environ.get("LOCAL_RANK", 0)))
    if args.cmd == "render":
        t0 = time.time()
        if world > 1:  # torchrun: the passes of the one image are dealt to the ranks, one NCCL reduce at the end
            import torch
            import torch.distributed as dist
            local_rank = int(os.environ.get("LOCAL_RANK", 0))
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            img = render_frame_sharded(dev, args.scene, args.width, args.height, args.samples, rank, world, local_rank,
                                       seed=args.seed)
            dist.barrier()
            dist.destroy_process_group()
            if rank != 0:
                return
        else:
            img = render_frame(dev, args.scene, args.width, args.height, args.samples, seed=args.seed)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        print(f"{write_image(args.out, img)}: {args.samples} passes on {world} GPU(s) in {time.time() - t0:.2f} s")
        return
    out_dir = os.path.join(args.out_dir, args.movie_name)
    os.makedirs(out_dir, exist_ok=True)
    stamp = [time.time()]

    def sink(f, img):
        path = write_image(os.path.join(out_dir, f"frame_{f:04d}.png"), img)
        now = time.time()
        print(f"rank {rank}: {path} (+{now - stamp[0]:.2f} s)", flush=True)
        stamp[0] = now

    t0 = time.time()
    mine = frames_for_rank(args.start_frame, args.movie_frames, rank, world)
    render_movie(dev, args.scene, args.width, args.height, args.samples, mine, args.movie_frames, seed=args.seed,
                 sink=sink, overlap=not args.serial)
    if mine:
        print(f"rank {rank}: {len(mine)} frames in {time.time() - t0:.2f} s = {(time.time() - t0) / len(mine):.3f} s/frame",
              flush=True)


if __name__ == "__main__":
    main()
