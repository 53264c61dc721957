#!/bin/bash
# round 2, GPU call H: how a tile shard of the C4 frame (1/2, 1/8 of the samples) responds to slices and the trace grid
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
python tools/sweep_walk.py --config c4 --passes 8 --tile-shard 2 --settings "mode=0;mode=0,slices=2;mode=0,slices=2,walk_grid_divisor=1;mode=0,slices=1;mode=0,slices=3,walk_grid_divisor=1;mode=0,slices=4;mode=0,slices=3,walk_grid_divisor=3" > gpurun_out/r02h_shard2.jsonl 2> gpurun_out/r02h_shard2.err; echo "shard2 rc=$?"
python tools/sweep_walk.py --config c4 --passes 12 --tile-shard 8 --settings "mode=0;mode=0,slices=2;mode=0,slices=2,walk_grid_divisor=1;mode=0,slices=1;mode=1,slices=0;mode=1,slices=1" > gpurun_out/r02h_shard8.jsonl 2> gpurun_out/r02h_shard8.err; echo "shard8 rc=$?"
