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


def _row_step(world):
    """Per-row rotation of the tile owners: the smallest odd step that is coprime with `world`."""
    for step in (3, 5, 7, 11, 13, 17):
        if np.gcd(step, world) == 1:
            return step
    return 1


def tile_owner(width, height, world, tile=16):
    """Owner rank of every pixel: tiles of `tile` x `tile` pixels dealt round-robin along a row, each row of tiles
    rotated by a step coprime with `world`.  (Plain row-major round-robin degenerates into vertical stripes whenever the
    number of tiles per row is a multiple of `world` — 1920 / 16 = 120 tiles on 8 ranks — and a column's cost then
    decides a rank's time.)  Every `world` consecutive tiles of a row, and every `world` consecutive rows of a column,
    go to `world` different ranks."""
    ys, xs = np.divmod(np.arange(width * height, dtype=np.int64), width)
    return ((xs // tile) + (ys // tile) * _row_step(world)) % world


def tile_ids(width, height, rank, world, tile=16):
    """Sample ids (= pixel indices) of `rank`, ascending: its tiles of `tile` x `tile` pixels (see `tile_owner`)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.flatnonzero(tile_owner(width, height, world, tile) == rank).astype(np.uint32)


def shard_plan(world, steps):
    """How `world` GPUs split `steps` sample passes of ONE frame (strong scaling): (tile_groups, pass_groups) with
    tile_groups * pass_groups == world.  A GPU in tile group t and pass group p traces the tiles of group t for the passes
    p, p + pass_groups, ... from pass group p's own RNG streams; one reduce sums everything.

    Passes are the cheaper axis to split — a GPU keeps whole frames, i.e. as many resident warps per pass as a single-GPU
    render — so pass_groups is the largest divisor of `world` that divides `steps` (every GPU then renders exactly
    steps / pass_groups passes); what is left is split by tile."""
    world, steps = int(world), int(steps)
    best = 1
    for p in range(1, world + 1):
        if world % p == 0 and steps % p == 0:
            best = p
    return world // best, best


def frame_plan(world, steps):
    """How `world` GPUs render `steps` sample passes of ONE frame: a list of phases (passes_per_gpu, tile_groups,
    pass_groups).  Whole-frame passes are the efficient unit (a GPU tracing half a frame pays the same per-round tails on
    half the rays: 64 instead of 72 Msamples/s per GPU at BASELINE config 4), so every GPU first renders steps // world
    full-frame passes from its own RNG streams, and only the steps % world passes that are left are shared out by
    `shard_plan` (pass groups x tile groups).  20 passes on 8 GPUs: 2 full passes each, then 4 passes on 2 GPUs each."""
    world, steps = int(world), int(steps)
    phases = []
    q, r = divmod(steps, world)
    if q:
        phases.append((q, 1, world))
    if r:
        tile_groups, pass_groups = shard_plan(world, r)
        phases.append((r // pass_groups, tile_groups, pass_groups))
    return phases


def hybrid_shard(width, height, rank, world, steps, tile=16):
    """(sample ids or None, pass group, passes of this rank, tile_groups, pass_groups) for `shard_plan(world, steps)`."""
    tile_groups, pass_groups = shard_plan(world, steps)
    t, p = rank % tile_groups, rank // tile_groups
    ids = None if tile_groups == 1 else tile_ids(width, height, t, tile_groups, tile)
    return ids, p, steps // pass_groups, tile_groups, pass_groups


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
