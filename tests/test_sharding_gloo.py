"""N > 1 host logic on CPU: world_size-2 gloo run of tile sharding + the accumulator reduce.
The per-rank sample passes are produced by the host build of the kernel source (tests/hostemu, test infrastructure);
on GPUs the same ids/seeds/reduce drive clive2_b200.Renderer and NCCL (bench.py, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clive2_b200 import sharding, struct_types as T
from tests.common import arrays, seeds_for


def test_tiles_partition_the_frame():
    for (w, h, world, tile) in ((64, 48, 2, 16), (100, 36, 8, 16), (33, 17, 4, 8)):
        ids = [sharding.tile_ids(w, h, r, world, tile) for r in range(world)]
        allids = np.sort(np.concatenate(ids))
        assert np.array_equal(allids, np.arange(w * h, dtype=np.uint32))
        sizes = [len(x) for x in ids]
        assert max(sizes) - min(sizes) <= tile * tile * 2
    with pytest.raises(ValueError):
        sharding.tile_ids(8, 8, 2, 2)


def test_frame_plan_covers_every_pass_of_the_frame_exactly_once():
    """Every phase (passes per GPU, tile groups, pass groups) uses all GPUs, and the passes add up: a GPU's passes count 1 /
    tile_groups of a frame each."""
    for world in (1, 2, 3, 4, 8):
        for steps in (1, 2, 3, 5, 8, 16, 20, 21, 24, 100):
            phases = sharding.frame_plan(world, steps)
            total = 0.0
            for passes, tile_groups, pass_groups in phases:
                assert tile_groups * pass_groups == world and passes >= 1
                total += passes * pass_groups  # whole-frame passes rendered by this phase (tile groups share one pass)
            assert total == steps, (world, steps, phases)
    assert sharding.frame_plan(8, 20) == [(2, 1, 8), (1, 2, 4)]
    assert sharding.frame_plan(4, 20) == [(5, 1, 4)]
    assert sharding.frame_plan(8, 3) == [(3, 8, 1)]


def test_rank_seeds_are_distinct_and_nonzero():
    a, b = sharding.rank_seeds(1000, 5, 0), sharding.rank_seeds(1000, 5, 1)
    assert (a != 0).all() and (b != 0).all() and (a != b).mean() > 0.99
    assert np.array_equal(a, sharding.rank_seeds(1000, 5, 0))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, passes, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.hostemu import HostEmu
        a = arrays("empty", w, h)
        b = w * h
        seeds = seeds_for(b)
        ids = sharding.tile_ids(w, h, rank, world, tile=8)
        E = HostEmu()
        acc = E.accel(a.boxes, a.triangles)
        s = np.ascontiguousarray(seeds[ids])
        lp, cp = np.zeros(len(ids), T.Path), np.zeros(len(ids), T.Path)
        aggs = np.zeros(b, T.WeightAggregator)
        li = np.zeros((b, 4), np.float32)
        accum = np.zeros((b, 8), np.float32)
        for _ in range(passes):
            E.fused_pass(a, acc, s, lp, cp, aggs, li, accum, sample_ids=ids)
        E.accel_destroy(acc)
        t = torch.from_numpy(accum)
        sharding.reduce_accumulators(t, dst=0)
        if rank == 0:
            np.save(out_path, accum)
    finally:
        dist.destroy_process_group()


def test_two_rank_tile_sharded_render_equals_single(tmp_path):
    w, h, passes = 32, 24, 2
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, passes, out), nprocs=2, join=True)
    reduced = np.load(out)
    from tests.hostemu import HostEmu
    a = arrays("empty", w, h)
    b = w * h
    E = HostEmu()
    acc = E.accel(a.boxes, a.triangles)
    s = seeds_for(b).copy()
    lp, cp = np.zeros(b, T.Path), np.zeros(b, T.Path)
    aggs = np.zeros(b, T.WeightAggregator)
    li = np.zeros((b, 4), np.float32)
    full = np.zeros((b, 8), np.float32)
    for _ in range(passes):
        E.fused_pass(a, acc, s, lp, cp, aggs, li, full)
    E.accel_destroy(acc)
    assert np.array_equal(reduced[:, 7], full[:, 7]) and (full[:, 7] == passes).all()
    assert np.array_equal(reduced[:, 4:7].view(np.uint32), full[:, 4:7].view(np.uint32))
    np.testing.assert_allclose(reduced[:, 0:4], full[:, 0:4], rtol=2e-4, atol=1e-6)


def test_tiled_order_is_a_permutation_of_compact_blocks():
    o = sharding.tiled_order(64, 32, 8, 4)
    assert np.array_equal(np.sort(o), np.arange(64 * 32, dtype=np.uint32))
    first = o[:32]
    ys, xs = np.divmod(first.astype(int), 64)
    assert xs.max() - xs.min() == 7 and ys.max() - ys.min() == 3  # one warp = one 8x4 pixel block
    o2 = sharding.tiled_order(70, 33, 8, 4)  # ragged edges
    assert np.array_equal(np.sort(o2), np.arange(70 * 33, dtype=np.uint32))


def test_more_ranks_than_tiles_leaves_empty_shards():
    """A 16x16 frame is ONE 16x16 tile: ranks 1.. own nothing (the renderer refuses an empty shard with a clear message;
    the caller picks a smaller tile)."""
    from clive2_b200 import sharding
    assert len(sharding.tile_ids(16, 16, 0, 4, 16)) == 256
    assert all(len(sharding.tile_ids(16, 16, r, 4, 16)) == 0 for r in (1, 2, 3))
    parts = [sharding.tile_ids(16, 16, r, 4, 4) for r in range(4)]
    assert all(len(p) == 64 for p in parts) and len(np.unique(np.concatenate(parts))) == 256
