#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 400 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_tail_budget=64;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=16;mode=0,walk_tail_budget=128" > gpurun_out/r02s_sweep_c4.jsonl 2> gpurun_out/r02s_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c4 --passes 8 --tile-shard 2 --settings "mode=0;mode=0,walk_tail_budget=32" > gpurun_out/r02s_sweep_c4_shard2.jsonl 2> gpurun_out/r02s_sweep_c4_shard2.err; echo "c4 shard2 rc=$?"
