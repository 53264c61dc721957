#!/bin/bash
# continuation records (walk_tail_budget): parity first, then the A/B timing
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "suspend or c4_terrain or c3_terrain" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02t_pytest.log
timeout 500 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_tail_budget=64;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=16;mode=0,walk_tail_budget=128;mode=0,walk_tail_budget=256" > gpurun_out/r02t_sweep_c4.jsonl 2> gpurun_out/r02t_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c4 --passes 8 --tile-shard 2 --settings "mode=0;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=64" > gpurun_out/r02t_sweep_c4_shard2.jsonl 2> gpurun_out/r02t_sweep_c4_shard2.err; echo "c4 shard2 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0;mode=0,walk_tail_budget=32;mode=0,walk_tail_budget=64" > gpurun_out/r02t_sweep_c3.jsonl 2> gpurun_out/r02t_sweep_c3.err; echo "c3 rc=$?"
