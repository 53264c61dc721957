"""Multi-GPU partitioning of the per-pass work (new design: the reference is single-GPU).

Every sample id of a pass is independent and passes only meet in additive accumulators (renderer.py:271-278), so the
path shards without any data-path collective:
  * by image tile  — `tile_ids`: round-robin tiles of the frame; each rank traces its ids with the GLOBAL seed rows, so
    a G-GPU render consumes exactly the RNG streams of the 1-GPU render (strong scaling of one frame);
  * by sample pass — every rank renders full-frame passes from its own seed buffer (`rank_seeds`): statistically
    equivalent, not stream identical (weak scaling; what bench.py measures);
and ONE reduce(sum) of the packed float accumulators [H][W][8] at the end (`reduce_accumulators`, NCCL over NVLink
on GPUs; the same call runs on gloo/CPU tensors in the tests).
"""
import numpy as np


def tile_ids(width, height, rank, world, tile=16):
    """Sample ids (= pixel indices) of `rank`: tiles of `tile` x `tile` pixels dealt round-robin in row-major order."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    ys, xs = np.divmod(np.arange(width * height, dtype=np.int64), width)
    tiles_x = (width + tile - 1) // tile
    owner = ((ys // tile) * tiles_x + (xs // tile)) % world
    return np.flatnonzero(owner == rank).astype(np.uint32)


def rank_seeds(batch, base_seed, rank):
    """Independent uint32[batch][2] RNG state per rank for pass sharding (zero words replaced by 1)."""
    s = np.random.RandomState((int(base_seed) + 7919 * int(rank)) % (2 ** 32)).randint(0, 2 ** 32, size=(batch, 2), dtype=np.uint32)
    s[s == 0] = 1
    return s


def reduce_accumulators(tensor, dst=0, group=None):
    """Sum the packed accumulators of all ranks into rank `dst`, in place (torch.distributed; NCCL or gloo)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensor
    dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def tiled_order(width, height, tile_w=8, tile_h=4):
    """All sample ids of a frame ordered tile by tile (tile_w x tile_h pixels, row-major inside a tile), so that the 32
    lanes of a warp hold a compact block of pixels instead of a 32 x 1 strip.  Passing it as `sample_ids` changes
    which thread traces which sample, not any sample's result (ids keep their own seed rows and pixels)."""
    ys, xs = np.divmod(np.arange(width * height, dtype=np.int64), width)
    tiles_x = (width + tile_w - 1) // tile_w
    key = ((ys // tile_h) * tiles_x + xs // tile_w) * (tile_w * tile_h) + (ys % tile_h) * tile_w + xs % tile_w
    return np.argsort(key, kind="stable").astype(np.uint32)
