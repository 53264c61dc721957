"""Headless equivalents of the reference's CLIs: `render` (src/render.py:11-50) and the orbit `movie`
(src/movie.py:10-55), the latter frame-parallel over GPUs (BASELINE config 5).

    python -m clive2_b200.movie render --scene bench-spheres --width 512 --height 512 --samples 64 --out out.png
    torchrun --nproc-per-node 8 -m clive2_b200.movie movie --scene bench-spheres --movie-frames 120 --samples 32

Frames of a movie are independent (the reference rebuilds scene + BVH + renderer per frame, movie.py:29-55), so frame f
goes to rank f mod world_size and no collective is needed.  PNGs are written with cv2 when it is importable, else as
.npy arrays.
"""
import argparse
import os
import time

import numpy as np

from . import runtime, scene
from .renderer import Renderer


def frames_for_rank(start_frame, total_frames, rank, world):
    """Frame indices rendered by `rank`: round-robin over the frames still to do (movie.py --start-frame semantics)."""
    return [f for f in range(start_frame, total_frames) if (f - start_frame) % world == rank]


def write_image(path, image_bgr_u8):
    try:
        import cv2
        cv2.imwrite(path, image_bgr_u8)
        return path
    except Exception:
        np.save(path + ".npy", image_bgr_u8)
        return path + ".npy"


def render_frame(device, preset, width, height, samples, frame_idx=None, total_frames=None, seed=None):
    """One image: scene (+ orbit camera when frame_idx is given) -> BVH -> `samples` fused passes -> tone-mapped BGR."""
    if frame_idx is None:
        arrays = scene.preset_arrays(preset, width, height)
    else:
        center, direction = scene.orbit_camera(frame_idx, total_frames)
        arrays = scene.preset_arrays(preset, width, height, center, direction)
    sc = scene.scene_from_arrays(arrays, device)
    r = Renderer(sc, seed=seed)
    for _ in range(samples):
        r.run_sample()
    image = r.image
    r.release()
    sc.release()  # also drops the per-frame BVH cached for these buffers
    return image


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
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = runtime.Device(int(os.environ.get("LOCAL_RANK", 0)))
    if args.cmd == "render":
        t0 = time.time()
        img = render_frame(dev, args.scene, args.width, args.height, args.samples, seed=args.seed)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        print(f"{write_image(args.out, img)}: {args.samples} passes in {time.time() - t0:.2f} s")
        return
    out_dir = os.path.join(args.out_dir, args.movie_name)
    os.makedirs(out_dir, exist_ok=True)
    for f in frames_for_rank(args.start_frame, args.movie_frames, rank, world):
        t0 = time.time()
        img = render_frame(dev, args.scene, args.width, args.height, args.samples, f, args.movie_frames, seed=args.seed)
        path = write_image(os.path.join(out_dir, f"frame_{f:04d}.png"), img)
        print(f"rank {rank}: {path} in {time.time() - t0:.2f} s", flush=True)


if __name__ == "__main__":
    main()
