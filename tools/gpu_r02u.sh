#!/bin/bash
# experiment: each round's queue sorted by a ray key (origin cell Morton code + direction octant); slices x suspension
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 600 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_sort_bits=18;mode=0,slices=1;mode=0,slices=2;mode=0,slices=1,walk_sort_bits=18;mode=0,slices=2,walk_grid_divisor=1" > gpurun_out/r02u_sweep_c4.jsonl 2> gpurun_out/r02u_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0;mode=0,walk_sort_bits=18" > gpurun_out/r02u_sweep_c3.jsonl 2> gpurun_out/r02u_sweep_c3.err; echo "c3 rc=$?"
