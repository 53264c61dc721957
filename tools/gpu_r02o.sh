#!/bin/bash
mkdir -p gpurun_out
export CLV_SCENE_CACHE=1
CLV_WALK_SHARE_BELOW=16 timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest(share) rc=$?" | tee -a gpurun_out/r02o_pytest.log; tail -2 gpurun_out/r02o_pytest.log
timeout 400 python tools/sweep_walk.py --config c4 --passes 6 --settings "mode=0;mode=0,walk_share_below=8;mode=0,walk_share_below=16;mode=0,walk_share_below=24;mode=0,walk_share_below=32;mode=0,walk_share_below=16,walk_grid_divisor=1;mode=0,walk_share_below=16,slices=2,walk_grid_divisor=1;mode=0,walk_share_below=16,slices=1" > gpurun_out/r02o_sweep_c4.jsonl 2> gpurun_out/r02o_sweep_c4.err; echo "c4 rc=$?"
timeout 300 python tools/sweep_walk.py --config c4 --passes 8 --tile-shard 2 --settings "mode=0;mode=0,walk_share_below=16;mode=0,walk_share_below=32" > gpurun_out/r02o_sweep_c4_shard2.jsonl 2> gpurun_out/r02o_sweep_c4_shard2.err; echo "c4 shard2 rc=$?"
timeout 300 python tools/sweep_walk.py --config c4 --passes 12 --tile-shard 8 --settings "mode=1;mode=2;mode=2,walk_share_below=16;mode=2,walk_share_below=32" > gpurun_out/r02o_sweep_c4_shard8.jsonl 2> gpurun_out/r02o_sweep_c4_shard8.err; echo "c4 shard8 rc=$?"
timeout 300 python tools/sweep_walk.py --config c3 --passes 8 --settings "mode=0;mode=0,walk_share_below=16;mode=0,walk_share_below=32" > gpurun_out/r02o_sweep_c3.jsonl 2> gpurun_out/r02o_sweep_c3.err; echo "c3 rc=$?"
timeout 300 python tools/sweep_walk.py --config c2 --passes 20 --settings "mode=1;mode=2;mode=2,walk_share_below=16;mode=2,walk_share_below=32" > gpurun_out/r02o_sweep_c2.jsonl 2> gpurun_out/r02o_sweep_c2.err; echo "c2 rc=$?"
